"""Why is e2e slower than the device-resident forward?  PCIe copy rates alone / concurrent / under the forward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from oracle import synth

dev = torch.device("cuda", 0)
n = 16
g = hv.Generator({"input_dim": 1, "ngf": 16}, True); g.load_state_dict(synth.synthetic_generator_state_dict()); g = g.cuda().eval(); g.precision = "bf16"
x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(n, seed=1))
hin = torch.empty(12 << 20, dtype=torch.uint8).pin_memory(); din = torch.empty(12 << 20, dtype=torch.uint8, device=dev)
hout = torch.empty(16 << 20, dtype=torch.uint8).pin_memory(); dout = torch.empty(16 << 20, dtype=torch.uint8, device=dev)
s_in, s_out, s_main = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()

def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

def h2d():
    with torch.cuda.stream(s_in): din.copy_(hin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s_out): hout.copy_(dout, non_blocking=True)
def fwd():
    with torch.no_grad(): g(x, mask, cam, ratio)
def both(): h2d(); d2h()
def all3(): h2d(); d2h(); fwd()
for _ in range(3): fwd()
print(f"H2D 12 MiB alone      {timed(h2d):.3f} ms  ({12.58 / timed(h2d):.1f} GB/s)")
print(f"D2H 16 MiB alone      {timed(d2h):.3f} ms  ({16.78 / timed(d2h):.1f} GB/s)")
print(f"H2D + D2H concurrent  {timed(both):.3f} ms")
print(f"forward alone         {timed(fwd):.3f} ms")
print(f"forward + H2D + D2H (independent streams, no dependencies) {timed(all3):.3f} ms per step")
# many small copies like the bench (4 input tensors, 6 output tensors)
hs = [torch.empty(4 << 20, dtype=torch.uint8).pin_memory() for _ in range(4)]; ds = [torch.empty(4 << 20, dtype=torch.uint8, device=dev) for _ in range(4)]
def d2h4():
    with torch.cuda.stream(s_out):
        for h, d in zip(hs, ds): h.copy_(d, non_blocking=True)
print(f"D2H 4 x 4 MiB         {timed(d2h4):.3f} ms")
