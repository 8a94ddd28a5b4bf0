"""Oracle restatement of the integer / mask rows of the path.

TEST INFRASTRUCTURE — never imported by the product package.  numpy only (byte/integer
arithmetic); small pure-Python loops where the reference loops.
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------- A9
def label_components_8(binary):
    """8-connected component labelling, scan order = first pixel in row-major order
    (equivalent to scipy.ndimage.label(x, ones((3,3))) up to label numbering).
    eval_3d_sagittal_twostage.py:20-21."""
    h, w = binary.shape
    lab = np.zeros((h, w), np.int32)
    nxt = 0
    for r in range(h):
        for c in range(w):
            if binary[r, c] and lab[r, c] == 0:
                nxt += 1
                stack = [(r, c)]
                lab[r, c] = nxt
                while stack:
                    y, x = stack.pop()
                    for dy in (-1, 0, 1):
                        for dx in (-1, 0, 1):
                            yy, xx = y + dy, x + dx
                            if 0 <= yy < h and 0 <= xx < w and binary[yy, xx] and lab[yy, xx] == 0:
                                lab[yy, xx] = nxt
                                stack.append((yy, xx))
    return lab, nxt


def remove_small_components(binary, min_size=50):
    """eval_3d_sagittal_twostage.py:16-30: zero every 8-connected component with fewer
    than ``min_size`` pixels."""
    out = np.array(binary, copy=True)
    lab, n = label_components_8(out != 0)
    if n:
        sizes = np.bincount(lab.ravel(), minlength=n + 1)
        small = sizes < min_size
        small[0] = False
        out[small[lab]] = 0
    return out


def slice_prep(cam, label, ct, vert_id, maxheight=40, eval_mask=True):
    """run_model's host-side slice preparation (eval_3d_sagittal_twostage.py:47-94).

    Returns None for an empty slice, else a dict with the three uint8 planes, the
    normalised fp32 planes the generator sees, and the integers (x1, x2, height, min_x,
    max_x).  ``eval_mask`` selects the 41-row inclusive mask of the eval driver (:75)
    vs the 40-row training mask (data/aligned_dataset.py:230)."""
    vert = np.zeros_like(label)
    vert[label == vert_id] = 1
    vert = remove_small_components(vert, 50)
    rows = np.nonzero(vert.any(axis=1))[0]
    if rows.size == 0:
        return None
    x1, x2 = int(rows.min()), int(rows.max())
    width = vert.shape[0]
    height = x2 - x1
    if height > maxheight:
        x_mean = int(np.mean(np.argwhere(vert)[:, 0]))
        x1 = x_mean - 20
        x2 = x1 + 40
    mask_x = (x1 + x2) // 2
    h2 = maxheight
    if mask_x <= h2 // 2:
        min_x, max_x = 0, h2
    elif width - mask_x <= h2 / 2:
        max_x = width
        min_x = max_x - h2
    else:
        min_x = mask_x - h2 // 2
        max_x = min_x + h2
    mask_u8 = np.zeros(vert.shape, np.uint8)
    if eval_mask:
        mask_u8[min_x:max_x + 1] = 255
    else:
        mask_u8[min_x:max_x] = 255
    ct_u8 = np.zeros(vert.shape, np.uint8)
    cam_u8 = np.zeros(vert.shape, np.uint8)
    # NumPy slice semantics are kept on purpose (a negative start wraps; the reference
    # raises a broadcast error in that case and so do we).
    ct_u8[:min_x, :] = ct[(x1 - min_x):x1, :]
    ct_u8[max_x:, :] = ct[x2:x2 + (width - max_x), :]
    cam_u8[:min_x, :] = cam[(x1 - min_x):x1, :]
    cam_u8[max_x:, :] = cam[x2:x2 + (width - max_x), :]
    ori_u8 = ct.astype(np.uint8)
    f32 = np.float32
    return {
        "ct_u8": ct_u8, "mask_u8": mask_u8, "cam_u8": cam_u8, "ori_u8": ori_u8,
        "ct": ((ct_u8.astype(f32) / f32(255.0)) - f32(0.5)) / f32(0.5),
        "ori_ct": ((ori_u8.astype(f32) / f32(255.0)) - f32(0.5)) / f32(0.5),
        "mask": mask_u8.astype(f32) / f32(255.0),
        "cam": cam_u8.astype(f32) / f32(255.0),
        "x1": x1, "x2": x2, "height": height, "min_x": min_x, "max_x": max_x,
    }


# ----------------------------------------------------------------------------- A4
def stitch_rows(pred_h_sigmoid, height, x1, maxheight=40):
    """pred_h = ceil(sigmoid*40); h = max(pred_h, height); rows of the generated band.
    eval_3d_sagittal_twostage.py:103-111 == pix2pix_model.py:208-218.
    ``pred_h_sigmoid`` is an fp32 value; the product is taken in fp32 like torch does."""
    pred_h = math.ceil(float(np.float32(pred_h_sigmoid) * np.float32(maxheight)))
    h = max(pred_h, int(height))
    d = h - int(height)
    x_upper = int(x1) - d // 2
    x_bottom = x_upper + h
    return h, d, x_upper, x_bottom


def stitch_plane(gen, real, x1, x2, height, pred_h_sigmoid, maxheight=40):
    """Height-adaptive re-stitching of ONE [H, W] plane: rows [x_up, x_bot) from the
    generator output, rows above from real[d//2 : x1], rows below from
    real[x2 : x2 + H - x_bot] (pix2pix_model.py:206-227, eval:108-118)."""
    hh = gen.shape[0]
    h, d, xu, xb = stitch_rows(pred_h_sigmoid, height, x1, maxheight)
    out = np.zeros_like(gen)
    out[xu:xb] = gen[xu:xb]
    out[:xu] = real[d // 2:x1]
    out[xb:] = real[x2:x2 + hh - xb]
    return out


def threshold_mask(p):
    """torch.where(p > 0.5, 1, 0) (pix2pix_model.py:201-202, eval:105)."""
    return (p > np.float32(0.5)).astype(np.float32)


def eval_postprocess(fine_seg, x_stage2, pred2_h, ori_ct, label, x1, x2, height, vert_id,
                     maxheight=40):
    """run_model after the forward (eval_3d_sagittal_twostage.py:103-130): returns
    (stitched label map float64 [H,W], stitched CT in 0..255 float32 [H,W])."""
    fake_ct = stitch_plane(x_stage2, ori_ct, x1, x2, height, pred2_h, maxheight)
    fake_ct = (fake_ct + np.float32(1)) * np.float32(127.5)
    h, d, xu, xb = stitch_rows(pred2_h, height, x1, maxheight)
    seg = threshold_mask(fine_seg)
    hh = seg.shape[0]
    mid = np.zeros_like(seg)
    mid[xu:xb] = seg[xu:xb] * vert_id
    up = np.zeros(seg.shape, np.float64)
    up[:xu] = label[d // 2:x1]
    bot = np.zeros(seg.shape, np.float64)
    bot[xb:] = label[x2:x2 + hh - xb]
    return mid + up + bot, fake_ct


# ----------------------------------------------------------------------------- A5
def sobel_edges(img):
    """Sobel.forward on [N,1,H,W] (models/edge_operator.py:41-48): replicate-pad 1,
    cross-correlate with Gx/Gy, magnitude, clamp to <= 1."""
    x = np.pad(img.astype(np.float32), ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    f32 = np.float32

    def sh(dy, dx):
        return x[:, :, 1 + dy:x.shape[2] - 1 + dy, 1 + dx:x.shape[3] - 1 + dx]

    gx = (-sh(-1, -1) + sh(-1, 1) - f32(2) * sh(0, -1) + f32(2) * sh(0, 1) - sh(1, -1) + sh(1, 1))
    gy = (sh(-1, -1) + f32(2) * sh(-1, 0) + sh(-1, 1) - sh(1, -1) - f32(2) * sh(1, 0) - sh(1, 1))
    e = np.sqrt(gx * gx + gy * gy, dtype=np.float32)
    return np.minimum(e, f32(1.0))


def edge_loss(fake_mask, real_mask):
    """800 * mse(sobel(fake), sobel(real)) (pix2pix_model.py:109,:263-264,:349).  On {0,1}
    masks this equals 800 * popcount(edge_fake XOR edge_real) / numel."""
    ef, er = sobel_edges(fake_mask), sobel_edges(real_mask)
    return float(800.0 * np.mean((ef.astype(np.float64) - er.astype(np.float64)) ** 2))


def edge_xor_count(fake_mask, real_mask):
    ef, er = sobel_edges(fake_mask), sobel_edges(real_mask)
    return int(np.count_nonzero((ef > 0) != (er > 0)))


# ----------------------------------------------------------------------------- A10
def column_heights(fake, label, axis=2, coronal=False):
    """Integer part of calculate_heights (evaluation/RHLV_quantification.py:41-73;
    coronal twin slices axis 1).  For every slice index where both masks are non-empty
    returns a dict of integer vectors / scalars:
    counts_{all,pre,mid,post}_{fake,label}, center_{fake,label}, t1, t2."""
    out = []
    n = label.shape[axis]
    for z in range(n):
        lab = np.take(label, z, axis=axis)
        fk = np.take(fake, z, axis=axis)
        if not (np.any(lab) and np.any(fk)):
            continue
        loc = np.where(fk)[1]
        y_min, y_max = int(loc.min()), int(loc.max())
        y_range = y_max - y_min
        t1 = int(y_min + y_range / 3)
        t2 = int(y_min + 2 * y_range / 3)
        cf = int(np.count_nonzero(fk[:, int(np.mean(loc))]))
        locl = np.where(lab)[1]
        cl = int(np.count_nonzero(lab[:, int(np.mean(locl))]))
        rec = {"z": z, "t1": t1, "t2": t2, "center_fake": cf, "center_label": cl}
        for nm, m in (("fake", fk), ("label", lab)):
            rec["all_" + nm] = np.count_nonzero(m, axis=0)
            rec["pre_" + nm] = np.count_nonzero(m[:, :t1], axis=0)
            rec["mid_" + nm] = np.count_nonzero(m[:, t1:t2], axis=0)
            rec["post_" + nm] = np.count_nonzero(m[:, t2:], axis=0)
        out.append(rec)
    return out


def calculate_heights(fake, label, height_threshold, axis=2, coronal=False):
    """calculate_heights (RHLV_quantification.py:41-118): the 8 kept-height vectors."""
    keys = ["all", "pre", "mid", "post"]
    acc = {k + "_" + s: [] for k in keys for s in ("fake", "label")}
    eps = 0.0 if coronal else 1e-6
    for rec in column_heights(fake, label, axis, coronal):
        cf = rec["center_fake"]
        cl = rec["center_label"]
        scale = {}
        for k in keys:
            f, l = rec[k + "_fake"], rec[k + "_label"]
            r = 1
            if l.size > 0 and f.size > 0 and l.max() > f.max():
                r = l.max() / (f.max() + eps)
            scale[k] = r
        cf = cf * scale["all"]
        for k in keys:
            f = rec[k + "_fake"] * scale[k]
            l = rec[k + "_label"]
            acc[k + "_fake"].extend(f[f > cf * height_threshold])
            acc[k + "_label"].extend(l[l > cl * height_threshold])
    return tuple(np.array(acc[k + "_" + s]) for k in keys for s in ("fake", "label"))


def calculate_rhlv(fake, label, center, length, height_threshold, axis=2, coronal=False):
    """calculate_rhlv (RHLV_quantification.py:121-147) -> (all, pre, mid, post RHLV,
    relative_height_label)."""
    sl = [slice(None)] * 3
    sl[axis] = slice(center - length, center + length)
    hs = calculate_heights(fake[tuple(sl)], label[tuple(sl)], height_threshold, axis, coronal)
    m = [float(np.mean(h)) if h.size > 0 else 0 for h in hs]
    af, al, pf, pl, mf, ml, qf, ql = m
    rh = lambda f, l: (f - l) / (f + 1e-6)
    lo, hi = min(pl, ml, ql), max(pl, ml, ql)
    return rh(af, al), rh(pf, pl), rh(mf, ml), rh(qf, ql), lo / (hi + 1e-6)
