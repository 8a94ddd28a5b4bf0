"""Per-CTA event timeline of one tcgen05 conv layer (CTA 0): where do the producer / MMA issuer / epilogue wait?
usage (GPU box): python tools/trace_conv.py <layer idx> [batch]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib
from oracle import synth


def main():
    layer = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    sd = synth.synthetic_generator_state_dict()
    g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    g.precision = "bf16"
    x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(n, seed=1))
    with torch.no_grad():
        g(x, mask, cam, ratio)
    L = _lib.lib()
    L.hv_debug_conv_trace.argtypes = [ctypes.c_void_p]
    for _ in range(2):
        g.run_layer(layer, n)
    torch.cuda.synchronize()
    buf = torch.zeros(12000 + 4 * 400, dtype=torch.int64, device="cuda")
    L.hv_debug_conv_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.run_layer(layer, n)            # predecessor (same layer) so that the traced launch is a dependent launch
    e0.record()
    g.run_layer(layer, n)
    e1.record()
    torch.cuda.synchronize()
    print(f"kernel (event to event, after a predecessor) {e0.elapsed_time(e1) * 1e3:.1f} us")
    L.hv_debug_conv_trace(None)
    b = buf.cpu().tolist()
    ev = []
    for base, who in ((0, "tma"), (4000, "mma"), (8000, "epi")):
        for i in range(2000):
            tag, clk = b[base + 2 * i], b[base + 2 * i + 1]
            if tag == 0:
                break
            ev.append((clk, who, tag))
    ev.sort()
    t0 = ev[0][0]
    names = {1: "tma issued", 10: "weights ready wait start", 11: "acc stage free", 12: "band full", 13: "band MMAs issued+commit",
             20: "epi tile start", 21: "acc full",
             2: "producer at slot wait", 3: "slot free", 14: "peeked next", 15: "MMAs issued"}
    print(f"layer {layer} batch {n}: {len(ev)} events, span {ev[-1][0] - t0} cycles")
    for clk, who, tag in ev[:140]:
        print(f"{clk - t0:8d} {who} {names.get(tag, tag)}")
    ctas = [(b[12000 + 4 * i], b[12000 + 4 * i + 1]) for i in range(400) if b[12000 + 4 * i]]
    if b[12000]:
        print(f"CTA 0: entry -> first event {t0 - b[12002]} cycles, last event -> exit {b[12003] - ev[-1][0]} cycles, entry -> exit {b[12003] - b[12002]} cycles")
    if ctas:
        g0 = min(c[0] for c in ctas)
        starts = sorted(c[0] - g0 for c in ctas)
        ends = sorted(c[1] - g0 for c in ctas)
        durs = sorted(c[1] - c[0] for c in ctas)
        q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
        print(f"CTAs {len(ctas)}: start ns p0/p50/p100 = {starts[0]}/{q(starts, .5)}/{starts[-1]}  end ns p0/p50/p100 = {ends[0]}/{q(ends, .5)}/{ends[-1]}"
              f"  duration ns p0/p50/p100 = {durs[0]}/{q(durs, .5)}/{durs[-1]}")
    if len(ev) > 140:
        print("...")
        for clk, who, tag in ev[-30:]:
            print(f"{clk - t0:8d} {who} {names.get(tag, tag)}")


if __name__ == "__main__":
    main()
