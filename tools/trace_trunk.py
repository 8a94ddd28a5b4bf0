#!/usr/bin/env python
"""Per-role clock trace of one CTA of the dataflow trunk kernel (csrc/trunk_tc.cu) over the roofline chain (coarse conv5 .. conv11),
plus the mean time per layer of that chain (CUDA events, 20 chains, L2 flushed before each).
usage: python tools/trace_trunk.py [cta]      env HV_TRUNK_DEBUG=bits for the ablations (results are wrong with any bit set)"""
import ctypes
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib
from oracle import synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cta = int(args[0]) if args else 0
BATCH = int(args[1]) if len(args) > 1 else 16
torch.cuda.set_device(0)
g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
g.load_state_dict(synth.synthetic_generator_state_dict())
g = g.cuda().eval()
g.precision = "bf16"
x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(BATCH, seed=1))
with torch.no_grad():
    for _ in range(3):
        g(x, mask, cam, ratio)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
first, count = 4, 7
for _ in range(3):
    g.run_chain(first, count, BATCH)
evs = []
for _ in range(20):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.run_chain(first, count, BATCH)
    e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
us = sum(a.elapsed_time(b) for a, b in evs) / 20 * 1e3
print(f"batch {BATCH} HV_TRUNK_DEBUG={os.environ.get('HV_TRUNK_DEBUG', '0')} HV_TRUNK={os.environ.get('HV_TRUNK', '')}: chain of {count} layers {us:.1f} us = {us / count:.2f} us per layer")
if os.environ.get("HV_TRUNK") != "1" or "--no-trace" in sys.argv:
    sys.exit(0)
buf = torch.zeros(12000, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.hv_debug_trunk_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.hv_debug_trunk_trace(buf.data_ptr(), cta)
g.run_chain(first, count, BATCH)
torch.cuda.synchronize()
L.hv_debug_trunk_trace(None, 0)
t = buf.cpu().tolist()


def role(off):
    out = []
    for i in range(1300):
        tag, item, clk = t[off + 3 * i: off + 3 * i + 3]
        if tag == 0:
            break
        out.append((tag, item, clk))
    return out


prod, mma, epi = role(0), role(4000), role(8000)
t0 = min(r[0][2] for r in (prod, mma, epi) if r)
per = defaultdict(dict)
for tag, item, clk in prod + mma + epi:
    per[item][tag] = clk - t0
print("item  | producer: pull  +empty  +dep  +issue | issuer: full  +tempty  [+switch]  +issued | epilogue: tfull  +stored  +signalled   (cycles, 1.965 GHz)")
for item in sorted(per):
    d = per[item]
    g_ = lambda k: d.get(k)
    f = lambda a, b: f"{(d[a] - d[b]):6d}" if a in d and b in d else "     -"
    print(f"{item:5d} | {g_(10) if g_(10) is not None else -1:8d} {f(11, 10)} {f(12, 11)} {f(13, 12) if 12 in d else f(13, 11)} | {g_(20) if g_(20) is not None else -1:8d} {f(21, 20)} "
          f"{f(22, 21)} {f(23, 22) if 22 in d else f(23, 21)} | {g_(30) if g_(30) is not None else -1:8d} {f(31, 30)} {f(32, 31)}")
