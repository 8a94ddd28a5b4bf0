"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Run from the repo root:  python -m oracle.make_golden
Needs /root/reference (read-only mount); the produced fixtures are committed so that
the GPU box (which has no reference) can check both the oracle and the CUDA path.

Fixtures
  generator_n1.npz   config-1 slice (synthetic_slices(1, seed=123)): the four image
                     outputs in full, both height heads, the 32x32 flow image, per-layer
                     activation statistics + 16 probe values for all 47 conv blocks and
                     the contextual-attention output.
  generator_n2.npz   two slices with DIFFERENT mask rows (proves the "sample-0 mask"
                     quirk of the contextual attention): outputs subsampled [::4, ::4].
  run_model.npz      eval driver slice prep + forward + stitch on 3 slices of
                     synthetic_volume(seed=0): stitched label map / CT per slice.
  rhlv_known.json    RHLV known answers on the reference's shipped label volumes and
                     on the packed copy in rhlv_label_0007_20.npz.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.simplefilter("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refshim, synth, nifti_min  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PROBES = np.random.Generator(np.random.PCG64(99)).random(16)


def _stats(t):
    t = t.detach().double().reshape(-1)
    idx = (PROBES * t.numel()).astype(np.int64)
    return np.concatenate([[t.mean().item(), t.std().item(), t.abs().max().item()],
                           t[torch.from_numpy(idx)].numpy()])


def generator_fixture(n, per_sample_masks, path, full):
    sd = synth.synthetic_generator_state_dict()
    g = refshim.reference_generator(sd)
    x, mask, cam, ratio = synth.synthetic_slices(n, seed=123, per_sample_masks=per_sample_masks)
    taps = {}
    hooks = []
    for net in ("coarse_generator", "fine_generator"):
        for name, mod in getattr(g, net).named_children():
            if type(mod).__name__ in ("Conv2dBlock",):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, key=f"{net}.{name}": taps.__setitem__(key, _stats(o))))
    hooks.append(g.fine_generator.contextul_attention.register_forward_hook(
        lambda m, i, o: taps.__setitem__("fine_generator.contextul_attention", _stats(o[0]))))
    with torch.no_grad():
        cs, fs, x1, x2, flow, p1, p2 = g(x, mask, cam, ratio)
    for h in hooks:
        h.remove()
    sub = (slice(None), slice(None), slice(None), slice(None)) if full else \
        (slice(None), slice(None), slice(None, None, 4), slice(None, None, 4))
    out = {
        "coarse_seg": cs[sub].numpy(), "fine_seg": fs[sub].numpy(),
        "x_stage1": x1[sub].numpy(), "x_stage2": x2[sub].numpy(),
        "pred1_h": p1.numpy(), "pred2_h": p2.numpy(),
        "flow32_u8": np.uint8(np.round(flow[:, :, ::8, ::8].numpy() * 255)),
        "tap_names": np.array(sorted(taps)),
        "tap_stats": np.stack([taps[k] for k in sorted(taps)]),
    }
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


def run_model_fixture(path):
    refshim.install()
    import eval_3d_sagittal_twostage as ev
    import torchvision.transforms as T
    a_t = T.Compose([T.Grayscale(1), T.ToTensor(), T.Normalize((0.5,), (0.5,))])
    m_t = T.Compose([T.ToTensor()])
    sd = synth.synthetic_generator_state_dict()
    g = refshim.reference_generator(sd)
    label, ct, cam = synth.synthetic_volume(seed=0, depth=64)
    cam255 = cam * 255
    out = {}
    for z, vid in ((32, 20), (20, 19), (40, 21)):
        ratio = torch.tensor([abs(z - 32) / 50 * 2])
        seg, fake, height = ev.run_model(g, cam255[:, :, z].copy(), label[:, :, z].copy(),
                                         ct[:, :, z].copy(), vid, ratio, a_t, m_t, "cpu", 40)
        out[f"seg_{z}_{vid}"] = seg.astype(np.uint8)
        out[f"ct_{z}_{vid}"] = fake.astype(np.float32)
        out[f"height_{z}_{vid}"] = np.int64(height)
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


def rhlv_fixture():
    refshim.install()
    import importlib
    sag = importlib.import_module("RHLV_quantification")
    cor = importlib.import_module("RHLV_quantification_coronal")
    known = {}
    base = "/root/reference/datasets/straightened/label"
    for vid, axis in ((20, 2), (20, 1), (19, 2), (21, 2)):
        v = nifti_min.load(f"{base}/0007_{vid}.nii.gz")
        lab = (v == vid).astype(np.float64)
        fk = np.maximum(lab, np.roll(lab, -3, axis=0))
        fk[-3:] = lab[-3:]
        loc = np.where(lab)[axis]
        c, ln = int(np.mean(loc)), int((loc.max() - loc.min()) // 5)
        mod = sag if axis == 2 else cor
        vals = mod.calculate_rhlv(fk, lab, c, ln, "x", 0.7)
        known[f"0007_{vid}_axis{axis}"] = {"center": c, "length": ln,
                                            "rhlv": [float(x) for x in vals]}
        if vid == 20 and axis == 2:
            np.savez_compressed(os.path.join(GOLD, "rhlv_label_0007_20.npz"),
                                label=v.astype(np.uint8))
    with open(os.path.join(GOLD, "rhlv_known.json"), "w") as fh:
        json.dump(known, fh, indent=1)
    print(known)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    generator_fixture(1, False, os.path.join(GOLD, "generator_n1.npz"), full=True)
    generator_fixture(2, True, os.path.join(GOLD, "generator_n2.npz"), full=False)
    run_model_fixture(os.path.join(GOLD, "run_model.npz"))
    rhlv_fixture()
