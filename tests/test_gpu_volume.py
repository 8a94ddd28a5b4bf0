"""GPU parity of the device-side slice preparation (A9) and the batched volume driver against the oracle restatement of
run_model / process_nii_files and the golden vectors written by the unmodified reference (tests/golden/run_model.npz)."""
import os

import numpy as np
import pytest
import torch

import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib
from healthivert_gan_b200._lib import check, ptr
from healthivert_gan_b200.volume import VolumeSynthesizer
from oracle import generator_ref as gr
from oracle import mask_ops_ref as mo
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen(synthetic_sd):
    g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g.load_state_dict(synthetic_sd)
    return g.cuda().eval()


def _prepare(vs, label_u8, ct_u8, cam_u8, slices, vid):
    L = _lib.lib()
    S, h, w = label_u8.shape
    nb = len(slices)
    dev = label_u8.device
    idx = torch.tensor(slices, device=dev, dtype=torch.int32)
    v = torch.full((nb,), vid, device=dev, dtype=torch.int32)
    scratch = torch.empty(2 * nb * h * w, device=dev, dtype=torch.int32)
    keep = torch.empty(nb * h * w, device=dev, dtype=torch.uint8)
    meta = torch.empty(nb, 8, device=dev, dtype=torch.int32)
    planes = [torch.empty(nb, 1, h, w, device=dev) for _ in range(4)]
    ints = [torch.empty(nb, device=dev, dtype=torch.int32) for _ in range(3)]
    check(L.hv_slice_prepare(ptr(label_u8), ptr(ct_u8), ptr(cam_u8), ptr(idx), ptr(v), nb, h, w, 40, ptr(scratch), ptr(keep), ptr(meta),
                             *[ptr(p) for p in planes], *[ptr(i) for i in ints], None))
    torch.cuda.synchronize()
    return meta.cpu().numpy(), [p.cpu().numpy() for p in planes], keep.reshape(nb, h, w).cpu().numpy()


def test_slice_prepare_bit_exact_against_oracle(gen):
    label, ct, cam = synth.synthetic_volume(seed=3, depth=24)
    rng = np.random.Generator(np.random.PCG64(1))
    # specks (< 50 px, must be removed), a tall vertebra (height > 40 re-centring) and fractional CT values (uint8 truncation)
    label[5:9, 200:206, :] = 20
    label[60:130, 100:120, 5] = 21
    ct = ct + rng.random(ct.shape) * 0.999
    ct = np.clip(ct, 0, 255.999)
    vs = VolumeSynthesizer(gen)
    lab8, ct8, cam8 = vs._to_u8_slices(label, 2, 1.0), vs._to_u8_slices(ct, 2, 1.0), vs._to_u8_slices(cam, 2, 255.0)
    assert np.array_equal(ct8.cpu().numpy(), np.transpose(ct.astype(np.uint8), (2, 0, 1)))
    assert np.array_equal(cam8.cpu().numpy(), np.transpose((cam * 255).astype(np.uint8), (2, 0, 1)))
    for vid in (19, 20, 21, 7):
        slices = list(range(0, 24, 3)) + [5]
        meta, planes, keep = _prepare(vs, lab8, ct8, cam8, slices, vid)
        for b, z in enumerate(slices):
            ref = mo.slice_prep(cam[:, :, z] * 255, label[:, :, z], ct[:, :, z], vid)
            if ref is None:
                assert meta[b, 0] == 0, (vid, z)
                continue
            assert meta[b, 0] == 1
            assert tuple(meta[b, 1:6]) == (ref["x1"], ref["x2"], ref["height"], ref["min_x"], ref["max_x"]), (vid, z)
            vert = mo.remove_small_components((label[:, :, z] == vid).astype(np.uint8), 50)
            assert np.array_equal(keep[b], vert), (vid, z)
            assert np.array_equal(planes[0][b, 0], ref["ct"]), (vid, z)
            assert np.array_equal(planes[1][b, 0], ref["mask"]), (vid, z)
            assert np.array_equal(planes[2][b, 0], np.float32(1) - ref["cam"]), (vid, z)
            assert np.array_equal(planes[3][b, 0], ref["ori_ct"]), (vid, z)


def test_single_stage_against_reference_golden(gen, golden_dir):
    gold = np.load(os.path.join(golden_dir, "run_model.npz"))
    label, ct, cam = synth.synthetic_volume(seed=0, depth=64)
    vs = VolumeSynthesizer(gen)
    gen.per_sample_mask, gen.return_flow = True, False
    try:
        for z, vid in ((32, 20), (20, 19), (40, 21)):
            lab8, ct8, cam8 = vs._to_u8_slices(label, 2, 1.0), vs._to_u8_slices(ct, 2, 1.0), vs._to_u8_slices(cam, 2, 255.0)
            nxt = lab8.clone()
            ct_out = torch.zeros(64, 256, 256, device="cuda")
            lab_out = torch.zeros(64, 256, 256, device="cuda")
            vs._stage([z], vid, lab8, nxt, ct8, cam8, {z: abs(z - 32) / 50 * 2}, ct_out, lab_out)
            torch.cuda.synchronize()
            got_ct, got_seg = ct_out[z].cpu().numpy(), lab_out[z].cpu().numpy()
            assert np.abs(got_ct - gold[f"ct_{z}_{vid}"]).max() <= 0.05, (z, vid)     # 1e-3-level generator noise * 127.5
            assert (got_seg.astype(np.uint8) != gold[f"seg_{z}_{vid}"]).mean() <= 1e-4, (z, vid)
            assert int(vs.last_meta[0, 3]) == int(gold[f"height_{z}_{vid}"])
            assert np.array_equal(ct8[z].cpu().numpy(), got_ct.astype(np.uint8))      # hand-over plane of the next stage
    finally:
        gen.per_sample_mask, gen.return_flow = False, True


def _oracle_volume(sd, ct, label, cam, vert_id):
    """process_nii_files (eval:153-234) restated with the oracle pieces, batch 1, on the CPU."""
    out_ct, out_seg = np.zeros_like(ct), np.zeros_like(ct)
    loc = np.where(label == vert_id)
    z0, z1 = loc[2].min(), loc[2].max()
    rl = z1 - z0 + 1
    nl = int(rl * 4 / 5)
    nz0 = z0 + (rl - nl) // 2
    nz1 = nz0 + nl - 1
    center = (nz0 + nz1) // 2
    cam255 = cam * 255

    def run(lab2d, ct2d, z, vid, ratio):
        p = mo.slice_prep(cam255[:, :, z], lab2d, ct2d, vid)
        if p is None:
            return None
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None, None]
        with torch.no_grad():
            o = gr.generator_forward(sd, t(p["ct"]), t(p["mask"]), 1 - t(p["cam"]), torch.tensor([ratio], dtype=torch.float32), flow=False)
        seg, fake = mo.eval_postprocess(o[1][0, 0].numpy(), o[3][0, 0].numpy(), float(o[6][0, 0]), p["ori_ct"], lab2d, p["x1"], p["x2"],
                                        p["height"], vid)
        return seg, fake

    for z in range(nz0, nz1 + 1):
        ratio = abs(z - center) / rl * 2
        lab2d, ct2d = label[:, :, z], ct[:, :, z]
        if vert_id > 8 and np.sum(label[:, :, z] == vert_id - 1) > 200:
            lab2d, ct2d = run(lab2d, ct2d, z, vert_id - 1, ratio)
        if vert_id < 24 and np.sum(label[:, :, z] == vert_id + 1) > 200:
            lab2d, ct2d = run(lab2d, ct2d, z, vert_id + 1, ratio)
        o = run(lab2d, ct2d, z, vert_id, ratio)
        if o is not None:
            out_seg[:, :, z], out_ct[:, :, z] = o
    return out_ct, out_seg


def test_volume_driver_three_stage_loop_against_oracle(gen, synthetic_sd):
    label, ct, cam = synth.synthetic_volume(seed=5, depth=8)
    ref_ct, ref_seg = _oracle_volume(synthetic_sd, ct, label, cam, 20)
    vs = VolumeSynthesizer(gen, batch=4)
    got_ct, got_seg = vs.synthesize(ct, label, cam, 20)
    assert got_ct.shape == ct.shape
    done = np.nonzero(ref_seg.any(axis=(0, 1)))[0]
    assert done.size >= 3
    assert np.array_equal(np.nonzero(got_seg.any(axis=(0, 1)))[0], done)
    # three chained forwards: fp32 re-association noise can flip isolated threshold / truncation decisions
    assert (got_seg != ref_seg).mean() <= 2e-4
    assert np.abs(got_ct - ref_ct).max() <= 2.0 and np.abs(got_ct - ref_ct).mean() <= 0.01


def test_file_level_driver_matches_array_driver(gen, tmp_path):
    """synthesize_files (NIfTI in / out, process_nii_files loop body eval:153-241) == synthesize on the same volumes; the outputs
    carry the CT volume's affine and stay float64 like the reference's np.zeros_like(get_fdata()) volumes."""
    from healthivert_gan_b200 import nifti
    label, ct, cam = synth.synthetic_volume(seed=5, depth=8)
    aff = np.array([[0.0, 0.0, -1.0, 102.25], [0.0, -1.0, 0.0, 90.5], [-1.0, 0.0, 0.0, -1053.0], [0.0, 0.0, 0.0, 1.0]])   # float32-exact
    paths = {k: str(tmp_path / k / "case7_20.nii.gz") for k in ("CT", "label", "CAM", "CT_fake", "label_fake")}
    for k, vol in (("CT", ct), ("label", label), ("CAM", cam)):
        os.makedirs(os.path.dirname(paths[k]), exist_ok=True)
        nifti.save(paths[k], vol.astype(np.float64), aff)
    vs = VolumeSynthesizer(gen, batch=4)
    vid = vs.synthesize_files(paths["CT"], paths["label"], paths["CAM"], paths["CT_fake"], paths["label_fake"])
    assert vid == 20
    want_ct, want_seg = vs.synthesize(ct, label, cam, 20)
    got_ct, got_seg = nifti.load(paths["CT_fake"]), nifti.load(paths["label_fake"])
    assert got_ct.dataobj.dtype == np.float64 and got_ct.shape == ct.shape
    np.testing.assert_array_equal(got_ct.affine, aff)
    np.testing.assert_array_equal(got_seg.affine, aff)
    assert np.array_equal(got_ct.get_fdata(), want_ct.astype(np.float64))
    assert np.array_equal(got_seg.get_fdata(), want_seg.astype(np.float64))
    assert got_seg.get_fdata().any()


def test_coronal_driver_against_oracle_and_transposed_sagittal(gen, synthetic_sd):
    """axis = 1 (coronal slices vol[:, z, :], evaluation/RHLV_quantification_coronal.py:51-54) of a volume == the sagittal driver
    on the volume with its last two axes swapped (bit-identical: same slices through the same kernels) == the oracle restatement
    of process_nii_files on that swapped volume."""
    label, ct, cam = synth.synthetic_volume(seed=6, depth=8)
    sw = lambda v: np.ascontiguousarray(v.transpose(0, 2, 1))     # [256, 8, 256]: the slicing axis is axis 1
    vs = VolumeSynthesizer(gen, batch=4)
    cor_ct, cor_seg = vs.synthesize(sw(ct), sw(label), sw(cam), 20, axis=1)
    sag_ct, sag_seg = vs.synthesize(ct, label, cam, 20, axis=2)
    assert cor_ct.shape == (256, 8, 256)
    assert np.array_equal(cor_ct, sw(sag_ct)) and np.array_equal(cor_seg, sw(sag_seg))
    ref_ct, ref_seg = _oracle_volume(synthetic_sd, ct, label, cam, 20)
    assert (cor_seg != sw(ref_seg)).mean() <= 2e-4
    assert np.abs(cor_ct - sw(ref_ct)).max() <= 2.0 and np.abs(cor_ct - sw(ref_ct)).mean() <= 0.01


def test_config3_shape_sagittal_plus_coronal_with_rhlv(gen):
    """BASELINE.json config 3 at its stated slice shape through the driver: a 256 x 256 x 256 synthetic volume, sagittal AND coronal
    three-stage synthesis, RHLV features of both orientations (the 2.5D feature vector, SVM_grading_2.5d.py:14-27).  Size-independent
    properties at full size: the synthesis window is the middle 4/5 of the vertebra's extent along the slicing axis (eval:186-197),
    slices outside it stay zero, label_fake holds the target id plus (shifted) neighbour labels only, the CT stays in [0, 255],
    and RHLV(label_fake, label_fake) == 0."""
    from healthivert_gan_b200 import mask_ops
    label, ct, cam = synth.synthetic_volume(seed=2, depth=256)
    vs = VolumeSynthesizer(gen, batch=64)
    feats = {}
    for axis in (2, 1):
        ct_f, lab_f = vs.synthesize(ct, label, cam, 20, axis=axis)
        assert ct_f.shape == label.shape
        zs = np.where(label == 20)[axis]
        z0, z1 = int(zs.min()), int(zs.max())
        rl = z1 - z0 + 1
        nl = int(rl * 4 / 5)
        nz0 = z0 + (rl - nl) // 2
        done = np.nonzero(lab_f.any(axis=tuple(a for a in range(3) if a != axis)))[0]
        assert done.min() >= nz0 and done.max() <= nz0 + nl - 1 and done.size >= nl - 2
        ids = set(np.unique(lab_f))
        assert 20.0 in ids and ids <= {0.0} | {float(v) for v in range(17, 24)}
        assert ct_f.min() >= 0 and ct_f.max() <= 255          # (x + 1) * 127.5 of a clamped output (eval:121), stored as float
        fake = (lab_f == 20).astype(np.uint8)
        real = (label == 20).astype(np.uint8)
        c, ln = int(np.mean(np.where(real)[axis])), (z1 - z0) // 5
        feats[axis] = mask_ops.calculate_rhlv(fake, real, c, ln, "v20", 0.7, axis=axis)
        same = mask_ops.calculate_rhlv(fake, fake, c, ln, "v20", 0.7, axis=axis)
        assert all(abs(v) <= 1e-12 for v in same[:4])
    assert all(np.isfinite(v) for ax in feats for v in feats[ax])


def test_bf16_volume_loop_tracks_fp32_volume_loop(gen, synthetic_sd):
    """The three-stage loop feeds thresholded masks and uint8-truncated CT of one stage into the next, so bf16-mode decisions can
    differ from fp32 mode near 0.5 / near integer CT values.  Measured and bounded here: label flip rate, CT difference and the
    per-slice vertebral height (rows with any label) of label_fake."""
    label, ct, cam = synth.synthetic_volume(seed=5, depth=24)
    g16 = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g16.load_state_dict(synthetic_sd)
    g16 = g16.cuda().eval()
    g16.precision = "bf16"
    ct32, lab32 = VolumeSynthesizer(gen, batch=16).synthesize(ct, label, cam, 20)
    ct16, lab16 = VolumeSynthesizer(g16, batch=16).synthesize(ct, label, cam, 20)
    done = np.nonzero(lab32.any(axis=(0, 1)))[0]
    assert np.array_equal(np.nonzero(lab16.any(axis=(0, 1)))[0], done) and done.size >= 10
    flips = float((lab16[:, :, done] != lab32[:, :, done]).mean())
    dct = np.abs(ct16[:, :, done] - ct32[:, :, done])
    h32 = (lab32[:, :, done] != 0).any(axis=1).sum(axis=0)
    h16 = (lab16[:, :, done] != 0).any(axis=1).sum(axis=0)
    dh = np.abs(h32.astype(int) - h16.astype(int))
    print(f"bf16 vs fp32 volume loop: label flip rate {flips:.2e}, CT |diff| mean {dct.mean():.3f} max {dct.max():.1f} (of 255), "
          f"height |diff| mean {dh.mean():.2f} max {dh.max()} rows over {done.size} slices")
    assert flips <= 5e-3
    assert dct.mean() <= 0.5
    assert dh.mean() <= 1.0 and dh.max() <= 3


def test_straightening_against_reference_golden(golden_dir):
    """hv_resample_curve + straighten.straighten_case on the raw case the reference ships (multi-label mask + centroids, synthetic
    smooth CT) == the UNMODIFIED reference (vendored `straighten` package + straighten_mask_3d.py helpers, scipy map_coordinates):
    label maps bit-exact (SHA-256 of the straightened volume, before and after the posterior-element removal, and of the vertebra
    crop), trilinear CT within 1e-9."""
    import hashlib
    import json
    from healthivert_gan_b200 import nifti, straighten as st
    g = np.load(os.path.join(golden_dir, "straighten_0007.npz"))
    label = nifti.load(os.path.join(golden_dir, "raw_0007_msk.nii.gz")).get_fdata()
    entries = json.load(open(os.path.join(golden_dir, "raw_0007.json")))
    x, y, z = np.meshgrid(*(np.arange(s, dtype=np.float64) for s in label.shape), indexing="ij")
    ct = 500.0 * np.sin(x / 17.0) + 400.0 * np.cos(y / 23.0) + 2.5 * z - 150.0
    coords = [[e["X"], e["Y"], e["Z"]] for e in entries if isinstance(e, dict) and "X" in e]
    inter = st.Interpolator(st.extend_curve(np.array(coords), 20, (0, 0, 0), label.shape), step=1, get_local_basis=st.get_local_basis)
    raw = inter.interpolate_along(label, (128, 128), order=0)
    assert raw.shape == tuple(g["ct_shape"]) and raw.dtype == np.float64
    assert hashlib.sha256(raw.astype(np.uint8).tobytes()).hexdigest() == str(g["label_sha_before_split"])
    vids = [int(v) for v in g["vert_ids"]]
    sct, slab, crops = st.straighten_case(ct, label, entries, vids, outputsize=(128, 128, 128))
    assert hashlib.sha256(slab.astype(np.uint8).tobytes()).hexdigest() == str(g["label_sha"])
    assert [int((slab == i).sum()) for i in range(17, 25)] == [int(v) for v in g["label_counts"]]
    assert np.array_equal(slab.reshape(-1)[g["probe"]], g["label_probe"])
    np.testing.assert_allclose(sct.reshape(-1)[g["probe"]], g["ct_probe"], rtol=0, atol=1e-9)
    assert abs(sct.mean() - float(g["ct_mean"])) <= 1e-10
    assert sorted(crops) == vids
    for vid, want in zip(vids, g["centroids"]):
        np.testing.assert_allclose(crops[vid][2], want, rtol=0, atol=1e-8)
    crop_ct, crop_lab, _ = crops[20]
    assert crop_ct.shape == (128, 128, 128)
    assert hashlib.sha256(crop_lab.astype(np.uint8).tobytes()).hexdigest() == str(g["crop_label_sha"])
    np.testing.assert_allclose(crop_ct.reshape(-1)[g["probe"] % crop_ct.size], g["crop_ct_probe"], rtol=0, atol=1e-9)
    # the kernel against its own host grid + a scalar restatement of map_coordinates' rules on a handful of samples
    grid = inter.get_grid((128, 128))
    win = st.window(ct, -300, 800)
    for n, a, b in ((0, 0, 0), (100, 64, 64), (212, 127, 3), (57, 10, 120)):
        c = grid[:, n, a, b]
        inside = all(0 <= c[d] <= label.shape[d] - 1 for d in range(3))
        want = label[tuple(int(np.floor(v + 0.5)) for v in c)] if inside else 0.0
        assert raw[n, a, b] == want
    with pytest.raises(_lib.HvError):
        inter.interpolate_along(win, (128, 128), order=3)
