"""Warm per-layer timing of the bf16 tensor-core conv plan (CUDA events, 20 back-to-back launches per layer)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib
from oracle import synth, generator_ref as gr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = hv.Generator({"input_dim": 1, "ngf": 16}, True); g.load_state_dict(synth.synthetic_generator_state_dict()); g = g.cuda().eval(); g.precision = "bf16"
x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(n, seed=1))
with torch.no_grad():
    g(x, mask, cam, ratio)
names = [f"{l[0].split('_')[0]}.{l[1]}" for l in gr.all_layers()]
specs = gr.all_layers()
tot = 0.0
for idx, name in enumerate(names):
    try:
        g.run_layer(idx, n)
    except Exception:
        continue
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.run_layer(idx, n)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    _, _, cin, cout, k, s, p, d, act = specs[idx]
    h = {"coarse.conv1": 256}.get(name, 0)
    tot += us
    print(f"{idx:2d} {name:28s} {cin:3d}->{cout:2d} k{k} s{s} d{d:2d}  {us:7.1f} us")
print("sum", tot)
