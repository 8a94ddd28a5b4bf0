#!/usr/bin/env python
"""Does a second, independent pipeline (own generator plan, own streams) raise the end-to-end rate?  Kernels of two forwards can fill
each other's ramps and tails.  usage: python tools/diag_e2e_dual.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import healthivert_gan_b200 as hv
from oracle import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
BATCH = 16
sd = synth.synthetic_generator_state_dict()


def make(depth):
    g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    g.precision = "bf16"
    g.per_sample_mask, g.return_flow = True, False
    p = hv.SlicePipeline(g, batch=BATCH, depth=depth)
    rng = np.random.Generator(np.random.PCG64(7))
    for k in range(depth):
        s = p.slot(k)
        s.ct[:] = rng.integers(0, 256, size=s.ct.shape, dtype=np.uint8)
        s.cam[:] = rng.integers(0, 256, size=s.cam.shape, dtype=np.uint8)
        s.rows[:, 0] = 108
        s.rows[:, 1] = 149
        s.ct[:, 108:149] = 0
        s.ratio[:] = rng.random(BATCH).astype(np.float32)
    return g, p


def run(pipes, depth, nsteps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(nsteps):
        p = pipes[i % len(pipes)]
        k = (i // len(pipes)) % depth
        if i >= depth * len(pipes):
            p.wait(k)
        p.submit(k)
    for p in pipes:
        for k in range(depth):
            p.wait(k)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / nsteps * 1e3


for npipes, depth in ((1, 4), (2, 2), (2, 4), (3, 2)):
    objs = [make(depth) for _ in range(npipes)]
    pipes = [p for _, p in objs]
    run(pipes, depth, 40)
    ms = min(run(pipes, depth, steps) for _ in range(3))
    print(f"{npipes} pipeline(s) x depth {depth}: {ms:.4f} ms per batch-16 step = {BATCH / ms * 1e3:.0f} slices/s")
    for p in pipes:
        p.close()
    del objs, pipes
