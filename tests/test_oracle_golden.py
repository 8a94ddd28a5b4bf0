"""The oracle restatement vs the committed golden vectors (outputs of the unmodified
reference, made by oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import generator_ref as gr
from oracle import mask_ops_ref as mo
from oracle import synth

PROBES = np.random.Generator(np.random.PCG64(99)).random(16)


def _stats(t):
    t = t.detach().double().reshape(-1)
    idx = (PROBES * t.numel()).astype(np.int64)
    return np.concatenate([[t.mean().item(), t.std().item(), t.abs().max().item()],
                           t[torch.from_numpy(idx)].numpy()])


def test_generator_config1_matches_reference_golden(golden_dir, synthetic_sd):
    g = np.load(os.path.join(golden_dir, "generator_n1.npz"))
    x, mask, cam, ratio = synth.synthetic_slices(1, seed=123)
    taps = {}
    with torch.no_grad():
        cs, fs, x1, x2, flow, p1, p2 = gr.generator_forward(synthetic_sd, x, mask, cam, ratio, taps=taps)
    for name, t in (("coarse_seg", cs), ("fine_seg", fs), ("x_stage1", x1), ("x_stage2", x2),
                    ("pred1_h", p1), ("pred2_h", p2)):
        assert np.abs(t.numpy() - g[name]).max() <= 2e-6, name
    assert np.array_equal(np.uint8(np.round(flow[:, :, ::8, ::8].numpy() * 255)), g["flow32_u8"])
    for name, ref in zip(g["tap_names"], g["tap_stats"]):
        got = _stats(taps[str(name)])
        assert np.abs(got - ref).max() <= 5e-6 * max(1.0, np.abs(ref).max()), name
    # thresholded masks: no flips outside the guard band
    for name, t in (("coarse_seg", cs), ("fine_seg", fs)):
        ref = g[name]
        guard = np.abs(ref - 0.5) > 1e-5
        assert np.array_equal((t.numpy() > 0.5)[guard], (ref > 0.5)[guard])


def test_generator_sample0_mask_quirk_matches_reference_golden(golden_dir, synthetic_sd):
    g = np.load(os.path.join(golden_dir, "generator_n2.npz"))
    x, mask, cam, ratio = synth.synthetic_slices(2, seed=123, per_sample_masks=True)
    assert not torch.equal(mask[0], mask[1])
    with torch.no_grad():
        cs, fs, x1, x2, flow, p1, p2 = gr.generator_forward(synthetic_sd, x, mask, cam, ratio)
    for name, t in (("coarse_seg", cs), ("fine_seg", fs), ("x_stage1", x1), ("x_stage2", x2)):
        assert np.abs(t[:, :, ::4, ::4].numpy() - g[name]).max() <= 2e-6, name
    assert np.array_equal(np.uint8(np.round(flow[:, :, ::8, ::8].numpy() * 255)), g["flow32_u8"])


def test_run_model_matches_reference_golden(golden_dir, synthetic_sd):
    g = np.load(os.path.join(golden_dir, "run_model.npz"))
    label, ct, cam = synth.synthetic_volume(seed=0, depth=64)
    for z, vid in ((32, 20), (20, 19), (40, 21)):
        prep = mo.slice_prep(cam[:, :, z] * 255, label[:, :, z], ct[:, :, z], vid)
        ratio = torch.tensor([abs(z - 32) / 50 * 2], dtype=torch.float32)
        t = lambda a: torch.from_numpy(a)[None, None]
        with torch.no_grad():
            out = gr.generator_forward(synthetic_sd, t(prep["ct"]), t(prep["mask"]),
                                       1 - t(prep["cam"]), ratio, flow=False)
        seg, fake = mo.eval_postprocess(out[1][0, 0].numpy(), out[3][0, 0].numpy(),
                                        float(out[6][0, 0]), prep["ori_ct"], label[:, :, z],
                                        prep["x1"], prep["x2"], prep["height"], vid)
        assert prep["height"] == int(g[f"height_{z}_{vid}"])
        assert np.array_equal(seg.astype(np.uint8), g[f"seg_{z}_{vid}"])
        assert np.abs(fake - g[f"ct_{z}_{vid}"]).max() <= 1e-3  # 0..255 scale


def test_rhlv_known_answers(golden_dir):
    known = json.load(open(os.path.join(golden_dir, "rhlv_known.json")))
    v = np.load(os.path.join(golden_dir, "rhlv_label_0007_20.npz"))["label"]
    lab = (v == 20).astype(np.float64)
    fk = np.maximum(lab, np.roll(lab, -3, axis=0))
    fk[-3:] = lab[-3:]
    for axis in (2, 1):
        k = known[f"0007_20_axis{axis}"]
        got = mo.calculate_rhlv(fk, lab, k["center"], k["length"], 0.7, axis=axis, coronal=(axis == 1))
        assert np.allclose(got, k["rhlv"], rtol=0, atol=1e-12)
    # SURVEY §8c known answers, sagittal vertebra 20
    hs = mo.calculate_heights(fk[:, :, 22:42], lab[:, :, 22:42], 0.7)
    assert (hs[0].sum(), hs[0].size, hs[1].sum(), hs[1].size) == (13235, 718, 10705, 692)
    # identity
    assert mo.calculate_rhlv(lab, lab, 32, 10, 0.7)[:4] == (0.0, 0.0, 0.0, 0.0)


def test_component_filter_matches_scipy():
    from scipy.ndimage import label as sp_label
    rng = np.random.Generator(np.random.PCG64(5))
    for _ in range(5):
        img = (rng.random((64, 64)) > 0.62).astype(np.float64)
        mine = mo.remove_small_components(img, 12)
        lab, n = sp_label(img, np.ones((3, 3), np.int32))
        ref = img.copy()
        for i in range(1, n + 1):
            if np.sum(lab == i) < 12:
                ref[lab == i] = 0
        assert np.array_equal(mine, ref)


def test_sobel_binary_is_integer_exact():
    rng = np.random.Generator(np.random.PCG64(3))
    a = (rng.random((2, 1, 64, 64)) > 0.5).astype(np.float32)
    b = (rng.random((2, 1, 64, 64)) > 0.5).astype(np.float32)
    ea, eb = mo.sobel_edges(a), mo.sobel_edges(b)
    assert set(np.unique(ea)) <= {0.0, 1.0}
    assert abs(mo.edge_loss(a, b) - 800.0 * mo.edge_xor_count(a, b) / a.size) < 1e-9
    import torch.nn.functional as F
    t = torch.from_numpy(a)
    w = torch.tensor([[[-1., 0, 1], [-2, 0, 2], [-1, 0, 1]], [[1., 2, 1], [0, 0, 0], [-1, -2, -1]]])[:, None]
    ref = F.conv2d(F.pad(t, (1, 1, 1, 1), mode="replicate"), w).pow(2).sum(1, keepdim=True).sqrt().clamp(max=1)
    assert np.array_equal(ref.numpy(), ea)
