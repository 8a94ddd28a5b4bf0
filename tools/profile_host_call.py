"""Where does the host time of Generator.forward go (bf16 plan)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib
from oracle import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = hv.Generator({"input_dim": 1, "ngf": 16}, True); g.load_state_dict(synth.synthetic_generator_state_dict()); g = g.cuda().eval(); g.precision = "bf16"
x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(n, seed=1))
with torch.no_grad():
    for _ in range(5): g(x, mask, cam, ratio)
torch.cuda.synchronize()
L = _lib.lib()
real = L.hv_generator_forward
acc = {"c": 0.0, "plan": 0.0}
def timed_forward(*a):
    t0 = time.perf_counter(); r = real(*a); acc["c"] += time.perf_counter() - t0; return r
class Proxy:
    def __getattr__(self, k): return timed_forward if k == "hv_generator_forward" else getattr(L, k)
orig_lib = _lib.lib
_lib.lib = lambda: Proxy()
orig_plan = g._ensure_plan
def timed_plan(*a):
    t0 = time.perf_counter(); r = orig_plan(*a); acc["plan"] += time.perf_counter() - t0; return r
g._ensure_plan = timed_plan
reps = 50
with torch.no_grad():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        g(x, mask, cam, ratio); torch.cuda.synchronize()
    tot = time.perf_counter() - t0
print(f"batch {n}: per forward incl. sync {tot / reps * 1e6:.0f} us; C call hv_generator_forward {acc['c'] / reps * 1e6:.0f} us; _ensure_plan (incl. its C calls) {acc['plan'] / reps * 1e6:.0f} us")
