"""Host wrappers of the integer / mask kernels: threshold, height-adaptive stitch (A4) and
per-column vertebral heights + RHLV (A10)."""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


@torch.no_grad()
def threshold(p, value=1.0, as_u8=False):
    """torch.where(p > 0.5, value, 0) (reference pix2pix_model.py:201-202, eval:105)."""
    p = p.to(torch.float32).contiguous()
    out = torch.empty(p.shape, device=p.device, dtype=torch.uint8 if as_u8 else torch.float32)
    check(_lib.lib().hv_threshold(ptr(p), None if as_u8 else ptr(out), ptr(out) if as_u8 else None, float(value),
                                  p.numel(), _lib.stream()))
    return out


def _i32(t, dev):
    return torch.as_tensor(t).to(device=dev, dtype=torch.int32).contiguous()


@torch.no_grad()
def stitch(gen, real, pred_h, x1, x2, height, maxheight=40, return_rows=False):
    """Height-adaptive re-stitching without host syncs (reference pix2pix_model.py:206-252,
    eval_3d_sagittal_twostage.py:103-118).  gen/real: [N,1,H,W]; pred_h: [N] sigmoid outputs
    (NOT yet multiplied by maxheight); x1/x2/height: integer [N]."""
    n, c, h, w = gen.shape
    assert c == 1
    dev = gen.device
    gen = gen.to(torch.float32).contiguous()
    real = real.to(device=dev, dtype=torch.float32).contiguous()
    pred_h = pred_h.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(gen)
    rows = torch.empty(n, 4, device=dev, dtype=torch.int32)
    # keep every converted tensor alive until the launch is enqueued (the caching allocator would
    # otherwise hand a dead temporary's block to the next one)
    x1d, x2d, hd = _i32(x1, dev), _i32(x2, dev), _i32(height, dev)
    check(_lib.lib().hv_stitch(ptr(gen), ptr(real), ptr(pred_h), ptr(x1d), ptr(x2d), ptr(hd), int(maxheight),
                               ptr(out), ptr(rows), n, h, w, _lib.stream()))
    return (out, rows) if return_rows else out


@torch.no_grad()
def column_heights(vol_fake, vol_label, axis, z0, z1):
    """Device part of calculate_heights (reference evaluation/RHLV_quantification.py:41-73).
    vol_*: uint8 {0,1} CUDA tensors [d0,d1,d2].  Returns (counts [nz,8,ncols] int32, meta [nz,8] int32)."""
    assert vol_fake.dtype == torch.uint8 and vol_label.dtype == torch.uint8
    d0, d1, d2 = vol_label.shape
    ncols = d1 if axis == 2 else d2
    nz = z1 - z0
    counts = torch.empty(max(nz, 0), 8, ncols, device=vol_label.device, dtype=torch.int32)
    meta = torch.empty(max(nz, 0), 8, device=vol_label.device, dtype=torch.int32)
    if nz <= 0:
        if nz < 0:
            raise _lib.HvError(f"column_heights: empty window [{z0},{z1})")
        return counts, meta
    vf, vl = vol_fake.contiguous(), vol_label.contiguous()
    check(_lib.lib().hv_column_heights(ptr(vf), ptr(vl), d0, d1, d2, axis, z0, z1, ptr(counts), ptr(meta),
                                       _lib.stream()))
    return counts, meta


def calculate_heights(segmentation_fake, segmentation_label, height_threshold, axis=2, coronal=None):
    """reference evaluation/RHLV_quantification.py:41-118 (coronal twin: axis=1): the integer
    column scan runs on the GPU, the float64 ratio / threshold tail on the host."""
    coronal = (axis == 1) if coronal is None else coronal
    dev = torch.device("cuda", torch.cuda.current_device())
    f = torch.as_tensor(np.ascontiguousarray(segmentation_fake != 0)).to(dev).to(torch.uint8)
    l = torch.as_tensor(np.ascontiguousarray(segmentation_label != 0)).to(dev).to(torch.uint8)
    nz = l.shape[axis]
    counts, meta = column_heights(f, l, axis, 0, nz)
    counts = counts.cpu().numpy().astype(np.int64)
    meta = meta.cpu().numpy()
    names = ["all", "pre", "mid", "post"]
    acc = {k + s: [] for k in names for s in ("_fake", "_label")}
    eps = 0.0 if coronal else 1e-6
    for s in range(nz):
        valid, t1, t2, cenf, cenl = (int(v) for v in meta[s, :5])
        if not valid:
            continue
        ncols = counts.shape[2]
        seg = {"all": slice(0, ncols), "pre": slice(0, t1), "mid": slice(t1, t2), "post": slice(t2, ncols)}
        scale = {}
        for i, k in enumerate(names):
            fk, lb = counts[s, i, seg[k]], counts[s, 4 + i, seg[k]]
            r = 1
            if lb.size > 0 and fk.size > 0 and lb.max() > fk.max():
                r = lb.max() / (fk.max() + eps)
            scale[k] = r
        cf = cenf * scale["all"]
        for i, k in enumerate(names):
            fk = counts[s, i, seg[k]] * scale[k]
            lb = counts[s, 4 + i, seg[k]]
            acc[k + "_fake"].extend(fk[fk > cf * height_threshold])
            acc[k + "_label"].extend(lb[lb > cenl * height_threshold])
    return tuple(np.array(acc[k + s]) for k in names for s in ("_fake", "_label"))


def calculate_rhlv(segmentation_fake, segmentation_label, center_z, length, vertebra=None, height_threshold=0.64,
                   axis=2):
    """reference evaluation/RHLV_quantification.py:121-147 (same argument order)."""
    sl = [slice(None)] * 3
    sl[axis] = slice(center_z - length, center_z + length)
    hs = calculate_heights(segmentation_fake[tuple(sl)], segmentation_label[tuple(sl)], height_threshold, axis)
    m = [float(np.mean(h)) if h.size > 0 else 0 for h in hs]
    af, al, pf, pl, mf, ml, qf, ql = m
    rh = lambda f, l: (f - l) / (f + 1e-6)
    lo, hi = min(pl, ml, ql), max(pl, ml, ql)
    return rh(af, al), rh(pf, pl), rh(mf, ml), rh(qf, ql), lo / (hi + 1e-6)
