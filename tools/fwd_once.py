"""Two bf16 generator forwards at batch N (profiling target: ncu -k regex:conv_tc_kernel ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from oracle import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = hv.Generator({"input_dim": 1, "ngf": 16}, True); g.load_state_dict(synth.synthetic_generator_state_dict()); g = g.cuda().eval(); g.precision = "bf16"
x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(n, seed=1))
with torch.no_grad():
    for _ in range(reps):
        g(x, mask, cam, ratio)
torch.cuda.synchronize()
print("ok")
