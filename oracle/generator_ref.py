"""Oracle restatement of the two-stage generator forward.

TEST INFRASTRUCTURE — never imported by the product package.

Functional torch-CPU code driven by a reference-format ``state_dict`` (192 keys:
``<net>.<layer>.conv.{weight_orig,weight_u,weight_v,bias}`` + ``<net>.fc_height.*``).
The contextual attention is restated in its dense closed form (unfold -> matmul ->
flat-index diagonal fuse -> masked softmax -> matmul -> fold), which is *structurally
different* from the reference's per-sample conv formulation; ``tests/test_oracle_vs_reference.py``
checks it against the unmodified reference modules.

Reference: models/inpaint_networks.py, models/inpaint_tools.py,
torch/nn/utils/spectral_norm.py:92-114.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# (name, cin, cout, k, stride, pad, dilation, activation)
# models/inpaint_networks.py:41-63
COARSE_LAYERS = [
    ("conv1", 3, 16, 5, 1, 2, 1, "elu"),
    ("conv2_downsample", 16, 32, 3, 2, 1, 1, "elu"),
    ("conv3", 32, 32, 3, 1, 1, 1, "elu"),
    ("conv4_downsample", 32, 64, 3, 2, 1, 1, "elu"),
    ("conv5", 64, 64, 3, 1, 1, 1, "elu"),
    ("conv6", 64, 64, 3, 1, 1, 1, "elu"),
    ("conv7_atrous", 64, 64, 3, 1, 2, 2, "elu"),
    ("conv8_atrous", 64, 64, 3, 1, 4, 4, "elu"),
    ("conv9_atrous", 64, 64, 3, 1, 8, 8, "elu"),
    ("conv10_atrous", 64, 64, 3, 1, 16, 16, "elu"),
    ("conv11", 64, 64, 3, 1, 1, 1, "elu"),
    ("conv12", 64, 64, 3, 1, 1, 1, "elu"),
    ("conv20", 65, 64, 3, 1, 1, 1, "elu"),
    ("conv13", 64, 32, 3, 1, 1, 1, "elu"),
    ("conv14", 32, 32, 3, 1, 1, 1, "elu"),
    ("conv19", 33, 32, 3, 1, 1, 1, "elu"),
    ("conv15", 32, 16, 3, 1, 1, 1, "elu"),
    ("conv16", 16, 8, 3, 1, 1, 1, "elu"),
    ("conv17", 8, 1, 3, 1, 1, 1, "none"),
    ("conv18", 8, 1, 3, 1, 1, 1, "sigmoid"),
]
# models/inpaint_networks.py:126-165
FINE_LAYERS = [
    ("conv1", 4, 16, 5, 1, 2, 1, "elu"),
    ("conv2_downsample", 16, 16, 3, 2, 1, 1, "elu"),
    ("conv3", 16, 32, 3, 1, 1, 1, "elu"),
    ("conv4_downsample", 32, 32, 3, 2, 1, 1, "elu"),
    ("conv5", 32, 64, 3, 1, 1, 1, "elu"),
    ("conv6", 64, 64, 3, 1, 1, 1, "elu"),
    ("conv7_atrous", 64, 64, 3, 1, 2, 2, "elu"),
    ("conv8_atrous", 64, 64, 3, 1, 4, 4, "elu"),
    ("conv9_atrous", 64, 64, 3, 1, 8, 8, "elu"),
    ("conv10_atrous", 64, 64, 3, 1, 16, 16, "elu"),
    ("pmconv1", 4, 16, 5, 1, 2, 1, "elu"),
    ("pmconv2_downsample", 16, 16, 3, 2, 1, 1, "elu"),
    ("pmconv3", 16, 32, 3, 1, 1, 1, "elu"),
    ("pmconv4_downsample", 32, 64, 3, 2, 1, 1, "elu"),
    ("pmconv5", 64, 64, 3, 1, 1, 1, "elu"),
    ("pmconv6", 64, 64, 3, 1, 1, 1, "relu"),
    ("pmconv9", 64, 64, 3, 1, 1, 1, "elu"),
    ("pmconv10", 64, 64, 3, 1, 1, 1, "elu"),
    ("allconv11", 128, 64, 3, 1, 1, 1, "elu"),
    ("allconv19", 64, 64, 3, 1, 1, 1, "elu"),
    ("allconv12", 64, 64, 3, 1, 1, 1, "elu"),
    ("allconv13", 64, 32, 3, 1, 1, 1, "elu"),
    ("allconv14", 32, 32, 3, 1, 1, 1, "elu"),
    ("allconv15", 32, 16, 3, 1, 1, 1, "elu"),
    ("allconv16", 16, 8, 3, 1, 1, 1, "elu"),
    ("allconv17", 9, 1, 3, 1, 1, 1, "none"),
    ("allconv18", 9, 1, 3, 1, 1, 1, "sigmoid"),
]
_SPEC = {("coarse_generator", l[0]): l for l in COARSE_LAYERS}
_SPEC.update({("fine_generator", l[0]): l for l in FINE_LAYERS})


def all_layers():
    """[(net, name, cin, cout, k, stride, pad, dil, act)] in state_dict order."""
    out = [("coarse_generator",) + l for l in COARSE_LAYERS]
    out += [("fine_generator",) + l for l in FINE_LAYERS]
    return out


# ----------------------------------------------------------------------------- A1
def sn_power_iteration(w_orig, u, v, eps=1e-12):
    """One power iteration, in place on u and v (spectral_norm.py:92-110, train mode)."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
    u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))


def sn_sigma(w_orig, u, v):
    """sigma = u^T W v (spectral_norm.py:112)."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    return torch.dot(u, torch.mv(wm, v))


def _act(x, act):
    if act == "elu":
        return F.elu(x)
    if act == "relu":
        return F.relu(x)
    if act == "sigmoid":
        return torch.sigmoid(x)
    if act == "none":
        return x
    raise AssertionError(act)


def conv_block(sd, net, name, x, training=False, taps=None):
    """Conv2dBlock.forward = act(conv2d(x, W_orig/sigma, b)) (inpaint_networks.py:494-503)."""
    _, cin, cout, k, stride, pad, dil, act = _SPEC[(net, name)]
    p = f"{net}.{name}.conv."
    w, u, v, b = sd[p + "weight_orig"], sd[p + "weight_u"], sd[p + "weight_v"], sd[p + "bias"]
    if training:
        with torch.no_grad():
            sn_power_iteration(w, u, v)
    sigma = sn_sigma(w, u, v)
    y = F.conv2d(x, w / sigma, b, stride=stride, padding=pad, dilation=dil)
    y = _act(y, act)
    if taps is not None:
        taps[f"{net}.{name}"] = y
    return y


def _height_head(sd, net, x):
    """sigmoid(Linear(mean_HW(x))) (inpaint_networks.py:90-93, :211-214)."""
    g = x.mean(dim=(2, 3))
    return torch.sigmoid(F.linear(g, sd[f"{net}.fc_height.weight"], sd[f"{net}.fc_height.bias"]))


# ----------------------------------------------------------------------------- A2
def coarse_forward(sd, x, mask, cam, slice_ratio, training=False, taps=None):
    """CoarseGenerator.forward (inpaint_networks.py:68-117)."""
    net = "coarse_generator"
    n, _, h, w = x.shape
    ratio = slice_ratio.reshape(n, 1, 1, 1).expand(-1, -1, h, w).to(x.dtype)
    c = lambda name, t: conv_block(sd, net, name, t, training, taps)
    t = c("conv1", torch.cat([x, ratio, mask], dim=1))
    t = c("conv2_downsample", t)
    t = c("conv3", t)
    t = c("conv4_downsample", t)
    t = c("conv5", t)
    t = c("conv6", t)
    t = c("conv7_atrous", t)
    t = c("conv8_atrous", t)
    t = c("conv9_atrous", t)
    t = c("conv10_atrous", t)
    pred1_h = _height_head(sd, net, t)
    t = c("conv11", t)
    t = c("conv12", t)
    t = t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)  # nearest x2 (:97)
    cam128 = cam[:, :, ::2, ::2]                                   # nearest x0.5 (:98)
    t = c("conv20", torch.cat([t, cam128], dim=1))
    t = c("conv13", t)
    t = c("conv14", t)
    t = t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    t = c("conv19", torch.cat([t, cam], dim=1))
    t = c("conv15", t)
    t = c("conv16", t)
    x_stage1 = torch.clamp(c("conv17", t), -1.0, 1.0)
    coarse_seg = c("conv18", t)
    return coarse_seg, x_stage1, pred1_h


def fine_forward(sd, xin, x_stage1, mask, coarse_seg, slice_ratio, training=False, taps=None,
                 flow=True):
    """FineGenerator.forward (inpaint_networks.py:169-232)."""
    net = "fine_generator"
    n, _, h, w = xin.shape
    ratio = slice_ratio.reshape(n, 1, 1, 1).expand(-1, -1, h, w).to(xin.dtype)
    c = lambda name, t: conv_block(sd, net, name, t, training, taps)
    xnow = torch.cat([xin, coarse_seg, mask, ratio], dim=1)
    t = c("conv1", xnow)
    t = c("conv2_downsample", t)
    t = c("conv3", t)
    t = c("conv4_downsample", t)
    t = c("conv5", t)
    t = c("conv6", t)
    t = c("conv7_atrous", t)
    t = c("conv8_atrous", t)
    t = c("conv9_atrous", t)
    x_hallu = c("conv10_atrous", t)
    t = c("pmconv1", xnow)
    t = c("pmconv2_downsample", t)
    t = c("pmconv3", t)
    t = c("pmconv4_downsample", t)
    t = c("pmconv5", t)
    t = c("pmconv6", t)
    t, offsets = contextual_attention(t, mask)
    if taps is not None:
        taps["fine_generator.contextul_attention"] = t
        taps["fine_generator.contextul_attention.offsets"] = offsets
    offset_flow = flow_image(offsets) if flow else None
    t = c("pmconv9", t)
    pm = c("pmconv10", t)
    t = c("allconv11", torch.cat([x_hallu, pm], dim=1))
    pred2_h = _height_head(sd, net, t)
    t = c("allconv12", t)
    t = c("allconv19", t)
    t = t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    t = c("allconv13", t)
    t = c("allconv14", t)
    t = t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    t = c("allconv15", t)
    t = c("allconv16", t)
    t = torch.cat([t, x_stage1], dim=1)
    x_stage2 = torch.clamp(c("allconv17", t), -1.0, 1.0)
    fine_seg = c("allconv18", t)
    return fine_seg, x_stage2, offset_flow, pred2_h


def generator_forward(sd, x, mask, cam, slice_ratio, training=False, taps=None, flow=True):
    """Generator.forward (inpaint_networks.py:28-32) -> the reference 7-tuple."""
    coarse_seg, x_stage1, pred1_h = coarse_forward(sd, x, mask, cam, slice_ratio, training, taps)
    fine_seg, x_stage2, offset_flow, pred2_h = fine_forward(
        sd, x, x_stage1, mask, coarse_seg, slice_ratio, training, taps, flow)
    return coarse_seg, fine_seg, x_stage1, x_stage2, offset_flow, pred1_h, pred2_h


# ----------------------------------------------------------------------------- A3
def _cm(i, side=32):
    """row-major flat index (h*side+w) -> column-major flat index (w*side+h)."""
    return (i % side) * side + i // side


def ca_fuse(s, side=32):
    """The two identity-3x3 'fuse' convs on the [L_b, L_f] score map
    (inpaint_networks.py:350-361), as flat-index diagonal sums.  The +-1 shifts act on
    *flattened* indices (wrap across row / column ends); only flat indices outside
    [0, L) are zero."""
    L = side * side

    def diag(m):
        out = m.clone()
        out[1:, 1:] += m[:-1, :-1]
        out[:-1, :-1] += m[1:, 1:]
        return out

    t = diag(s)
    idx = torch.arange(L)
    perm = _cm(idx, side)            # perm[i] = cm(i)
    inv = torch.empty_like(perm)
    inv[perm] = idx                  # inv[cm(i)] = i
    tp = t[inv][:, inv]              # tp[i', j'] = t[cm^-1(i'), cm^-1(j')]
    up = diag(tp)
    return up[perm][:, perm]         # back to row-major


def ca_mask_valid(mask, rate=2):
    """mm[l] = 1 iff the zero-padded 3x3 neighbourhood of the 1/8-downsampled mask of
    SAMPLE 0 at l is all-zero (inpaint_networks.py:304-317)."""
    md = mask[0:1, :, :: 4 * rate, :: 4 * rate]
    m = F.unfold(F.pad(md, (1, 1, 1, 1)), kernel_size=3)  # [1, 9, L]
    return (m[0].mean(dim=0) == 0.0).to(torch.float32)     # [L]


def contextual_attention(f, mask, rate=2, ksize=3, softmax_scale=10.0, fuse=True,
                         return_intermediates=False):
    """ContextualAttention.forward(f, f, mask) with ksize=3, stride=1, rate=2, fuse_k=3
    (inpaint_networks.py:247-410).  Returns (y [N,C,H,W], offsets [N,2,H/2,W/2] int64 =
    (row, col) of argmax minus own position)."""
    n, c, h, w = f.shape
    side = h // rate
    L = side * side
    mm = ca_mask_valid(mask, rate)
    ys, offs, inter = [], [], []
    for i in range(n):
        fi = f[i:i + 1]
        # raw 4x4 stride-2 patches for pasting (:270-278); 'same' pad = 1 each side
        r = F.unfold(F.pad(fi, (1, 1, 1, 1)), kernel_size=2 * rate, stride=rate)[0]  # [C*16, L]
        fd = fi[:, :, ::rate, ::rate]
        p = F.unfold(F.pad(fd, (1, 1, 1, 1)), kernel_size=ksize)[0]                   # [C*9, L]
        nrm = torch.clamp(torch.sqrt((p * p).sum(dim=0)), min=1e-4)                   # (:341-345)
        s = (p / nrm).t() @ p                                                         # [L_b, L_f]
        if fuse:
            s = ca_fuse(s, side)
        logits = s * mm[:, None]
        a = torch.softmax(logits * softmax_scale, dim=0) * mm[:, None]                # (:364-366)
        am = torch.argmax(a, dim=0)                                                   # [L_f]
        cols = r @ a                                                                  # [C*16, L_f]
        y = F.fold(cols[None], output_size=(h, w), kernel_size=2 * rate, stride=rate,
                   padding=1) / 4.0                                                   # (:379)
        ys.append(y)
        ar = torch.arange(L)
        off = torch.stack([am // side - ar // side, am % side - ar % side], dim=0)
        offs.append(off.reshape(1, 2, side, side))
        if return_intermediates:
            inter.append({"P": p, "R": r, "norm": nrm, "S": s, "A": a, "argmax": am})
    y = torch.cat(ys, dim=0)
    offsets = torch.cat(offs, dim=0)
    if return_intermediates:
        return y, offsets, inter
    return y, offsets


# ------------------------------------------------------------------ flow colouring
def make_color_wheel():
    """models/inpaint_tools.py:244-273 (55 x 3 Middlebury wheel)."""
    ry, yg, gc, cb, bm, mr = 15, 6, 4, 11, 13, 6
    wheel = np.zeros([ry + yg + gc + cb + bm + mr, 3])
    col = 0
    wheel[0:ry, 0] = 255
    wheel[0:ry, 1] = np.floor(255 * np.arange(ry) / ry)
    col += ry
    wheel[col:col + yg, 0] = 255 - np.floor(255 * np.arange(yg) / yg)
    wheel[col:col + yg, 1] = 255
    col += yg
    wheel[col:col + gc, 1] = 255
    wheel[col:col + gc, 2] = np.floor(255 * np.arange(gc) / gc)
    col += gc
    wheel[col:col + cb, 1] = 255 - np.floor(255 * np.arange(cb) / cb)
    wheel[col:col + cb, 2] = 255
    col += cb
    wheel[col:col + bm, 2] = 255
    wheel[col:col + bm, 0] = np.floor(255 * np.arange(bm) / bm)
    col += bm
    wheel[col:col + mr, 2] = 255 - np.floor(255 * np.arange(mr) / mr)
    wheel[col:col + mr, 0] = 255
    return wheel


def flow_image(offsets, upscale=8):
    """flow_to_image + compute_color + x8 nearest upsample
    (inpaint_tools.py:73-100, :178-208; inpaint_networks.py:399-408).
    ``maxrad`` is a running maximum carried across the batch (reference quirk)."""
    off = offsets.permute(0, 2, 3, 1).cpu().numpy()
    wheel = make_color_wheel()
    ncols = wheel.shape[0]
    maxrad = -1.0
    out = []
    for i in range(off.shape[0]):
        u = off[i, :, :, 0].astype(np.float64)
        v = off[i, :, :, 1].astype(np.float64)
        rad = np.sqrt(u ** 2 + v ** 2)
        maxrad = max(maxrad, float(rad.max()))
        u = u / (maxrad + np.finfo(float).eps)
        v = v / (maxrad + np.finfo(float).eps)
        rad = np.sqrt(u ** 2 + v ** 2)
        a = np.arctan2(-v, -u) / np.pi
        fk = (a + 1) / 2 * (ncols - 1) + 1
        k0 = np.floor(fk).astype(int)
        k1 = k0 + 1
        k1[k1 == ncols + 1] = 1
        fr = fk - k0
        img = np.zeros(u.shape + (3,))
        for ch in range(3):
            tmp = wheel[:, ch]
            col0 = tmp[k0 - 1] / 255
            col1 = tmp[k1 - 1] / 255
            col = (1 - fr) * col0 + fr * col1
            idx = rad <= 1
            col[idx] = 1 - rad[idx] * (1 - col[idx])
            col[~idx] *= 0.75
            img[:, :, ch] = np.uint8(np.floor(255 * col))
        out.append(img)
    flow = torch.from_numpy(np.float32(np.uint8(out))) / 255.0
    flow = flow.permute(0, 3, 1, 2)
    return flow.repeat_interleave(upscale, dim=2).repeat_interleave(upscale, dim=3)
