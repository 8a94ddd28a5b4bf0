"""Deterministic synthetic weights, slices, train batches and straightened volumes.

TEST INFRASTRUCTURE — never imported by the product package (bench.py and tests use it
to build *inputs*; the product only ever sees tensors / state_dicts).

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` so that the build
container, the GPU box and the golden-vector script produce bit-identical inputs
without shipping a 4 MB checkpoint.  Recipes follow SURVEY.md §8(d).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from .generator_ref import all_layers

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEAD_BIAS_JSON = os.path.join(os.path.dirname(_HERE), "tests", "golden", "seg_head_bias.json")


def _normalize(x, eps=1e-12):
    return x / max(float(np.linalg.norm(x)), eps)


def synthetic_generator_state_dict(seed=0, power_iters=30, seg_head_margin=True):
    """Random-init generator weights in the reference ``state_dict`` format.

    * weight_orig, bias ~ U(+-1/sqrt(fan_in))  (torch Conv2d default init)
    * u, v ~ normalised N(0,1), then ``power_iters`` power iterations so that
      sigma = u^T W v is the converged spectral norm (un-warmed sigma makes
      activations explode).
    * seg heads (coarse conv18, fine allconv18): stored ``weight_u`` divided by 64 and
      bias re-centred (constants committed in tests/golden/seg_head_bias.json) so that
      thresholded masks are non-trivial and far from 0.5 (SURVEY §7 hard part 3).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for net, name, cin, cout, k, *_ in all_layers():
        fan_in = cin * k * k
        bound = 1.0 / np.sqrt(fan_in)
        w = rng.uniform(-bound, bound, size=(cout, cin, k, k))
        b = rng.uniform(-bound, bound, size=(cout,))
        u = _normalize(rng.standard_normal(cout))
        v = _normalize(rng.standard_normal(fan_in))
        wm = w.reshape(cout, -1)
        for _ in range(power_iters):
            v = _normalize(wm.T @ u)
            u = _normalize(wm @ v)
        p = f"{net}.{name}.conv."
        sd[p + "bias"] = torch.from_numpy(b.astype(np.float32))
        sd[p + "weight_orig"] = torch.from_numpy(w.astype(np.float32))
        sd[p + "weight_u"] = torch.from_numpy(u.astype(np.float32))
        sd[p + "weight_v"] = torch.from_numpy(v.astype(np.float32))
    for net in ("coarse_generator", "fine_generator"):
        bound = 1.0 / 8.0
        sd[f"{net}.fc_height.weight"] = torch.from_numpy(
            rng.uniform(-bound, bound, size=(1, 64)).astype(np.float32))
        sd[f"{net}.fc_height.bias"] = torch.from_numpy(
            rng.uniform(-bound, bound, size=(1,)).astype(np.float32))
    if seg_head_margin:
        for key in ("coarse_generator.conv18.conv.", "fine_generator.allconv18.conv."):
            sd[key + "weight_u"] = sd[key + "weight_u"] / 64.0
        if os.path.exists(_HEAD_BIAS_JSON):
            with open(_HEAD_BIAS_JSON) as fh:
                consts = json.load(fh)
            if consts.get("seed") == seed:
                for key, val in consts["bias"].items():
                    sd[key] = torch.tensor([val], dtype=torch.float32)
    return sd


def synthetic_slices(n=1, seed=123, mask_rows=(100, 141), per_sample_masks=False):
    """Config-1/2 inputs: x in [-1,1] zeroed on the mask rows, CAM in [0,1]
    (the caller passes 1-CAM like pix2pix_model.py:185), slice_ratio in [0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.random((n, 1, 256, 256), dtype=np.float32) * 2 - 1
    # low-pass so that the slices are not pure white noise
    x = 0.5 * x + 0.25 * np.roll(x, 1, axis=2) + 0.25 * np.roll(x, 1, axis=3)
    mask = np.zeros((n, 1, 256, 256), np.float32)
    for i in range(n):
        r0, r1 = mask_rows
        if per_sample_masks:
            sh = int(rng.integers(-24, 25))
            r0, r1 = r0 + sh, r1 + sh
        mask[i, :, r0:r1, :] = 1
    x = (x * (1 - mask)).astype(np.float32)
    cam = rng.random((n, 1, 256, 256), dtype=np.float32)
    ratio = rng.random((n,), dtype=np.float32)
    return (torch.from_numpy(x), torch.from_numpy(mask), torch.from_numpy(1 - cam),
            torch.from_numpy(ratio))


def synthetic_volume(seed=0, depth=64, target_id=20):
    """Config-3 straightened volume: label / CT / CAM of shape
    [256, 256, depth]; 7 vertebrae ids 17..23 stacked along axis 0 with elliptical
    cross-sections in (axis1, axis2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    label = np.zeros((256, 256, depth), np.float64)
    heights = rng.integers(22, 35, size=7)
    gaps = rng.integers(6, 9, size=7)
    while int(heights.sum() + gaps.sum()) > 248:   # seeds whose stack would not fit 256 rows (used to raise): trim the tallest
        heights[int(np.argmax(heights))] -= 1
    total = int(heights.sum() + gaps.sum())
    r = (256 - total) // 2
    yy, zz = np.meshgrid(np.arange(256), np.arange(depth), indexing="ij")
    centers = {}
    for i, vid in enumerate(range(17, 24)):
        ry = 25 + int(rng.integers(-3, 4))
        rz = 0.4 * depth
        ell = ((yy - 128) / ry) ** 2 + ((zz - depth / 2) / rz) ** 2 <= 1.0
        hgt = int(heights[i])
        # per-column height variation: shave 0..3 rows off top/bottom along axis1
        shave = (np.abs(yy - 128) / ry * 3).astype(int)
        for dr in range(hgt):
            keep = ell & (dr >= shave) & (dr < hgt - shave)
            label[r + dr][keep] = vid
        centers[vid] = r + hgt // 2
        r += hgt + int(gaps[i])
    ct = 20 + 10 * rng.standard_normal(label.shape) + 120 * (label > 0)
    sm = rng.standard_normal(label.shape)
    for ax in range(3):
        sm = (sm + np.roll(sm, 1, axis=ax) + np.roll(sm, -1, axis=ax)) / 3
    ct = np.clip(ct + 15 * sm, 0, 255)
    ct = np.floor(ct)
    xx = np.arange(256)[:, None, None]
    y2 = np.arange(256)[None, :, None]
    cam = np.exp(-((xx - centers[target_id]) ** 2 + (y2 - 128) ** 2) / (2 * 30.0 ** 2))
    cam = np.broadcast_to(cam, label.shape).copy()
    return label, ct, cam


def synthetic_train_batch(n=16, seed=7):
    """Config-4 batch dict in the dataset's collated format (SURVEY §8b/§8d,
    data/aligned_dataset.py:279-280)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x1 = np.array(([100, 98, 102, 96] * ((n + 3) // 4))[:n])
    height = np.array(([28, 30, 26, 33] * ((n + 3) // 4))[:n])
    x2 = x1 + height
    a = rng.random((n, 1, 256, 256), dtype=np.float32) * 2 - 1
    mask = np.zeros((n, 1, 256, 256), np.float32)
    a_mask = np.zeros_like(mask)
    normal = np.zeros_like(mask)
    for i in range(n):
        c = (x1[i] + x2[i]) // 2
        mask[i, :, c - 20:c + 20, :] = 1
        a_mask[i, :, x1[i]:x2[i], 90:140] = 1
        normal[i, :, 40:70, 90:140] = 1
    b = a * (1 - mask)
    cam = rng.random((n, 1, 256, 256), dtype=np.float32)
    ratio = rng.random((n,))
    t = torch.from_numpy
    return {
        "A": t(a), "B": t(b), "A_mask": t(a_mask), "mask": t(mask), "CAM": t(cam),
        "normal_vert": t(normal), "height": t(height), "x1": t(x1), "x2": t(x2),
        "h2": torch.full((n,), 40, dtype=torch.int64), "slice_ratio": t(ratio),
        "A_paths": ["synthetic"] * n, "B_paths": ["synthetic"] * n,
    }


def synthetic_discriminator_state_dict(seed=1, input_nc=1, ndf=64):
    """PatchGAN NLayerDiscriminator weights in the reference ``state_dict`` format
    (models/networks.py:555-602 after init_weights 'normal', gain 0.02: conv ~ N(0, 0.02), bias 0,
    BatchNorm gamma ~ N(1, 0.02), beta 0), drawn from numpy so both implementations share them."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    sd = {}
    chans = [(input_nc, ndf, True, False), (ndf, ndf * 2, False, True), (ndf * 2, ndf * 4, False, True),
             (ndf * 4, ndf * 8, False, True), (ndf * 8, 1, True, False)]
    idx = 0
    for cin, cout, bias, bn in chans:
        sd[f"model.{idx}.weight"] = torch.from_numpy((0.02 * rng.standard_normal((cout, cin, 4, 4))).astype(np.float32))
        if bias:
            sd[f"model.{idx}.bias"] = torch.zeros(cout)
        idx += 1
        if bn:
            sd[f"model.{idx}.weight"] = torch.from_numpy((1.0 + 0.02 * rng.standard_normal(cout)).astype(np.float32))
            sd[f"model.{idx}.bias"] = torch.zeros(cout)
            sd[f"model.{idx}.running_mean"] = torch.zeros(cout)
            sd[f"model.{idx}.running_var"] = torch.ones(cout)
            sd[f"model.{idx}.num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
            idx += 1
        idx += 1  # LeakyReLU slot (absent after the last conv)
    return sd


def train_options(**over):
    """argparse.Namespace with the effective defaults of `train.py --model pix2pix --direction BtoA`."""
    import argparse
    o = dict(gpu_ids=[], isTrain=True, checkpoints_dir="/tmp/hv_ckpt", name="synthetic", preprocess="none", input_nc=1,
             output_nc=1, ndf=64, netD="basic", n_layers_D=3, norm="batch", init_type="normal", init_gain=0.02,
             gan_mode="vanilla", lr=2e-4, beta1=0.5, direction="BtoA", lambda_L1=200.0, lr_policy="linear", epoch_count=1,
             n_epochs=200, n_epochs_decay=800, continue_train=False, verbose=False, load_iter=0, epoch="latest")
    o.update(over)
    return argparse.Namespace(**o)
