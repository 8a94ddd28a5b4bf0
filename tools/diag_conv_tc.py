"""Bring-up diagnostics for the tcgen05 conv kernel (run on the GPU box): prints error summaries per case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from healthivert_gan_b200 import _lib
from healthivert_gan_b200._lib import check, ptr

def bf(t): return t.to(torch.bfloat16).to(torch.float32)

def run(cin, cout, k, stride, dil, h, w, act="elu", n=1, up2=False, srcs=None, seed=0, heads=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    pad = (k - 1) // 2 * dil
    ref = F.conv2d(bf(x), bf(wt), b, stride=stride, padding=pad, dilation=dil)
    if heads:
        ref0, ref1 = ref[:, :1].clamp(-1, 1), torch.sigmoid(ref[:, 1:2])
    else:
        ref = {"elu": F.elu, "relu": F.relu, "none": lambda t: t, "sigmoid": torch.sigmoid}[act](ref)
        if up2: ref = ref.repeat_interleave(2, 2).repeat_interleave(2, 3)
    d = _lib.hv_conv_desc()
    xc = x.cuda()
    parts = [xc] if srcs is None else list(torch.split(xc, srcs, dim=1))
    parts = [p.contiguous() for p in parts]
    for i, p in enumerate(parts):
        d.src[i].ptr, d.src[i].channels, d.src[i].mode = p.data_ptr(), p.shape[1], 0
    d.n, d.cin, d.cout, d.hin, d.win, d.k, d.stride, d.pad, d.dil = n, cin, cout, h, w, k, stride, pad, dil
    d.act = 6 if heads else _lib.HV_ACT[act]
    d.nsrc = len(parts)
    ho, wo = h // stride, w // stride
    sc = 2 if up2 else 1
    y = torch.full((n, 1 if heads else cout, ho * sc, wo * sc), float("nan"), device="cuda")
    y2 = torch.full_like(y, float("nan")) if heads else None
    wc, bc = wt.cuda(), b.cuda()
    check(_lib.lib().hv_conv2d_bf16(d, ptr(wc), ptr(bc), ptr(y), ptr(y2), int(up2), None))
    torch.cuda.synchronize()
    outs = [(y.cpu(), ref0), (y2.cpu(), ref1)] if heads else [(y.cpu(), ref)]
    ok = True
    for yy, rr in outs:
        err = (yy - rr).abs()
        tol = 0.01 * rr.abs() + 0.01
        bad = (err > tol) | torch.isnan(yy)
        frac = bad.float().mean().item()
        ok &= frac == 0
        msg = f"max_err={err[~torch.isnan(err)].max().item() if (~torch.isnan(err)).any() else float('nan'):.4g} bad={frac:.4f} nan={torch.isnan(yy).float().mean().item():.4f}"
        if frac > 0:
            bc_ = bad.float().mean(dim=(0, 2, 3))
            by = bad.float().mean(dim=(0, 1, 3))
            bx = bad.float().mean(dim=(0, 1, 2))
            msg += f"\n    bad by channel: {[round(v,2) for v in bc_.tolist()][:16]}\n    bad by row(first 12): {[round(v,2) for v in by.tolist()][:12]} last: {[round(v,2) for v in by.tolist()][-4:]}\n    bad by col(first 12): {[round(v,2) for v in bx.tolist()][:12]} last: {[round(v,2) for v in bx.tolist()][-4:]}"
            msg += f"\n    sample got {yy.flatten()[:6].tolist()}\n    sample ref {rr.flatten()[:6].tolist()}"
        print(f"  {msg}")
    return ok

CASES = [
    ("64->64 3x3 d1 64x64", dict(cin=64, cout=64, k=3, stride=1, dil=1, h=64, w=64)),
    ("16->16 3x3 d1 32x32 none", dict(cin=16, cout=16, k=3, stride=1, dil=1, h=32, w=32, act="none")),
    ("64->64 3x3 d16 64x64 n2", dict(cin=64, cout=64, k=3, stride=1, dil=16, h=64, w=64, n=2)),
    ("64->64 3x3 d4 relu", dict(cin=64, cout=64, k=3, stride=1, dil=4, h=64, w=64, act="relu")),
    ("3->16 5x5 256x256", dict(cin=3, cout=16, k=5, stride=1, dil=1, h=256, w=256)),
    ("32->16 3x3 128x128 n3", dict(cin=32, cout=16, k=3, stride=1, dil=1, h=128, w=128, n=3)),
    ("64->32 3x3 up2 out", dict(cin=64, cout=32, k=3, stride=1, dil=1, h=64, w=64, up2=True)),
    ("128->64 concat 64+64", dict(cin=128, cout=64, k=3, stride=1, dil=1, h=64, w=64, srcs=[64, 64])),
    ("65->64 concat 64+1", dict(cin=65, cout=64, k=3, stride=1, dil=1, h=128, w=128, srcs=[64, 1])),
    ("4->16 concat 1+1+1+1 5x5", dict(cin=4, cout=16, k=5, stride=1, dil=1, h=64, w=64, srcs=[1, 1, 1, 1])),
    ("16->8 3x3", dict(cin=16, cout=8, k=3, stride=1, dil=1, h=64, w=64)),
    ("9->2 heads", dict(cin=9, cout=2, k=3, stride=1, dil=1, h=64, w=64, srcs=[8, 1], heads=True)),
    ("16->32 stride2 128->64", dict(cin=16, cout=32, k=3, stride=2, dil=1, h=128, w=128)),
    ("32->64 stride2 n2", dict(cin=32, cout=64, k=3, stride=2, dil=1, h=128, w=128, n=2)),
    ("64->64 batch16 64x64", dict(cin=64, cout=64, k=3, stride=1, dil=1, h=64, w=64, n=16)),
]

if __name__ == "__main__":
    only = sys.argv[1:] 
    allok = True
    for name, kw in CASES:
        if only and not any(o in name for o in only): continue
        print(name, flush=True)
        try:
            allok &= run(**kw)
        except Exception as e:
            print("  EXC", repr(e)[:300]); allok = False
            if "CUDA" in repr(e) or "launch" in repr(e): break
    print("ALL OK" if allok else "FAILURES")
