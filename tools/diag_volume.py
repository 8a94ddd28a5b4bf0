#!/usr/bin/env python
"""Where does the time of VolumeSynthesizer.synthesize go?  Phase times of one 256^3 volume (device-synchronised between phases).
usage: python tools/diag_volume.py [depth]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import healthivert_gan_b200 as hv
from healthivert_gan_b200.volume import VolumeSynthesizer
from oracle import synth

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
g.load_state_dict(synth.synthetic_generator_state_dict())
g = g.cuda().eval()
g.precision = "bf16"
vs = VolumeSynthesizer(g, batch=64)
label, ct, cam = synth.synthetic_volume(seed=1, depth=depth)
acc = {}


def timed(name, fn):
    def wrap(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    return wrap


vs._to_u8_slices = timed("to_u8_slices (H2D float64 + convert)", vs._to_u8_slices)
vs._stage = timed("stages (prepare + forward + stitch + finish)", vs._stage)
axes = (2, 1) if depth == 256 else (2,)
for axis in axes:
    vs.synthesize(ct, label, cam, 20, axis=axis)
for rep in range(2):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for axis in axes:
        vs.synthesize(ct, label, cam, 20, axis=axis)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(f"total {total * 1e3:.1f} ms:", {k: round(v * 1e3, 1) for k, v in acc.items()}, "rest (counts, clone, permute, D2H)", round((total - sum(acc.values())) * 1e3, 1))
