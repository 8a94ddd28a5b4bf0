"""Host geometry of the straightening step against the UNMODIFIED reference's values for the shipped raw case 0007
(tests/golden/straighten_0007.npz, oracle/make_golden_straighten.py).  The resampling kernel itself is covered in
tests/test_gpu_volume.py::test_straightening_against_reference_golden."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _module():
    """straighten.py imports the CUDA binding lazily enough (only interpolate_along touches it)."""
    from healthivert_gan_b200 import straighten
    return straighten


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "straighten_0007.npz"))


@pytest.fixture(scope="module")
def case():
    from oracle import nifti_min
    label = nifti_min.load(os.path.join(GOLD, "raw_0007_msk.nii.gz")).astype(np.float64)
    entries = json.load(open(os.path.join(GOLD, "raw_0007.json")))
    return label, entries


def _inter(case):
    st = _module()
    label, entries = case
    coords = [[e["X"], e["Y"], e["Z"]] for e in entries if isinstance(e, dict) and "X" in e]
    curve = st.extend_curve(np.array(coords), 20, (0, 0, 0), label.shape)
    return st, curve, st.Interpolator(curve, step=1, get_local_basis=st.get_local_basis)


def test_curve_knots_and_basis(gold, case):
    st, curve, inter = _inter(case)
    np.testing.assert_allclose(curve, gold["curve"], rtol=0, atol=1e-12)
    assert inter.knots.shape == gold["knots"].shape
    np.testing.assert_allclose(inter.knots, gold["knots"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(inter.basis, gold["basis"], rtol=0, atol=1e-10)
    # orthonormal, right-handed frames
    b = inter.basis
    np.testing.assert_allclose(np.einsum("nij,nik->njk", b, b), np.broadcast_to(np.eye(3), (len(b), 3, 3)), atol=1e-12)
    np.testing.assert_allclose(inter.get_grid((128, 128))[:, ::37, ::31, ::29], gold["grid_probe"], rtol=0, atol=1e-9)


def test_local_coordinates_of_the_centroids(gold, case):
    st, curve, inter = _inter(case)
    label, entries = case
    by_id = {int(e["label"]): (e["X"], e["Y"], e["Z"]) for e in entries if isinstance(e, dict) and e.get("label") is not None}
    for vid, want in zip(gold["vert_ids"], gold["centroids"]):
        got = inter.global_to_local(by_id[int(vid)], shape=(128, 128))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-8)
    back = inter.local_to_global(gold["centroids"][list(gold["vert_ids"]).index(20)], shape=(128, 128))
    np.testing.assert_allclose(back, gold["back20"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(back, by_id[20], atol=1e-3)          # round trip (piecewise-linear curve: approximate in the reference too)


def test_window_and_crop_helpers(gold):
    st = _module()
    sys.path.insert(0, os.path.dirname(GOLD))
    shape = (280, 180, 179)
    x, y, z = np.meshgrid(*(np.arange(s, dtype=np.float64) for s in shape), indexing="ij")
    ct = 500.0 * np.sin(x / 17.0) + 400.0 * np.cos(y / 23.0) + 2.5 * z - 150.0
    w = st.window(ct, -300, 800)
    np.testing.assert_allclose(w.reshape(-1)[gold["probe"] % w.size], gold["window_probe"], rtol=0, atol=1e-12)
    assert w.min() == 0.0 and w.max() == 255.0
    small = np.arange(4 * 5 * 6, dtype=np.float64).reshape(4, 5, 6)
    assert np.array_equal(st.window(small, -300, 800), small)       # already inside the window: unchanged (reference :176-177)
    vol = np.arange(10 * 12 * 8, dtype=np.float64).reshape(10, 12, 8)
    crop = st.extract_3d_volume(vol, (2.0, 11.0, 4.0), size=(6, 6, 4))
    # x: [0, 5) -> 5 planes at offset (6 - 5) // 2 = 0; y: [8, 12) -> 4 rows at offset 1; z: [2, 6) -> 4 columns at offset 0
    want = np.zeros((6, 6, 4))
    want[0:5, 1:5, 0:4] = vol[0:5, 8:12, 2:6]
    assert np.array_equal(crop, want)
    lab = np.zeros((3, 8, 5)); lab[:, 0:7, 2] = 4; lab[:, 6:, 0] = 4; lab[1, 7, 4] = 5
    out = st.remove_spine_labels_after_split(lab)
    assert (out[:, 7:, :] == 4).sum() == 0 and (out[:, :7, :] == 4).sum() == (lab[:, :7, :] == 4).sum()
    assert (out == 5).sum() == 0                                     # label 5 never touches the central column
