"""Generate tests/golden/straighten_0007.npz by running the UNMODIFIED reference straightening code (build container only).

TEST INFRASTRUCTURE.  Run from the repo root:  python -m oracle.make_golden_straighten
Reference code executed: the vendored `straighten` package (straighten/straighten/curve.py: Interpolator) and, from
straighten/straighten_mask_3d.py, extend_curve / get_local_basis / window / remove_spine_labels_after_split / extract_3d_volume
(its nibabel / skimage / matplotlib imports are stubbed, they are not used by these functions).  Inputs: the raw case the reference
ships (datasets/raw/0007: multi-label mask + centroid json; copied to tests/golden/ so that the GPU box has them) and a synthetic
smooth CT, because no raw CT is shipped.
"""
import hashlib
import importlib.util
import json
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import nifti_min  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def synthetic_ct(shape):
    x, y, z = np.meshgrid(*(np.arange(s, dtype=np.float64) for s in shape), indexing="ij")
    return 500.0 * np.sin(x / 17.0) + 400.0 * np.cos(y / 23.0) + 2.5 * z - 150.0


def main():
    for name in ("nibabel", "nibabel.orientations", "skimage", "skimage.morphology", "skimage.transform", "skimage.measure", "matplotlib",
                 "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.transform"].resize = None
    sys.modules["skimage"].morphology = sys.modules["skimage.morphology"]
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["nibabel"].orientations = sys.modules["nibabel.orientations"]
    sys.path.insert(0, os.path.join(REF, "straighten"))
    spec = importlib.util.spec_from_file_location("ref_straighten_mask_3d", os.path.join(REF, "straighten", "straighten_mask_3d.py"))
    ref = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(ref)
    except FileNotFoundError:
        pass   # the script's module-level driver opens a hard-coded json after every function has been defined
    from straighten import Interpolator

    for f in ("0007_msk.nii.gz", "0007.json"):
        shutil.copyfile(os.path.join(REF, "datasets", "raw", "0007", f), os.path.join(GOLD, "raw_" + f))
    label = nifti_min.load(os.path.join(GOLD, "raw_0007_msk.nii.gz")).astype(np.float64)
    entries = json.load(open(os.path.join(GOLD, "raw_0007.json")))
    ct = synthetic_ct(label.shape)

    coords = [[e["X"], e["Y"], e["Z"]] for e in entries if isinstance(e, dict) and "X" in e]
    curve = ref.extend_curve(np.array(coords), 20, (0, 0, 0), label.shape)
    ct_w = ref.window(ct.copy(), -300, 800)
    inter = Interpolator(curve, step=1, get_local_basis=ref.get_local_basis)
    shape = (128, 128)
    straight_ct = inter.interpolate_along(ct_w, shape, order=1)
    straight_label = inter.interpolate_along(label, shape, order=0)
    raw_label_sha = hashlib.sha256(straight_label.astype(np.uint8).tobytes()).hexdigest()
    straight_label = ref.remove_spine_labels_after_split(straight_label)
    rng = np.random.default_rng(7)
    probe = rng.integers(0, straight_ct.size, 8192)
    ids = [e["label"] for e in entries if isinstance(e, dict) and e.get("label") is not None]
    cents = {}
    for e in entries:
        if isinstance(e, dict) and e.get("label") is not None:
            cents[int(e["label"])] = inter.global_to_local((e["X"], e["Y"], e["Z"]), shape=shape)
    crop_ct = ref.extract_3d_volume(straight_ct, cents[20], size=(128, 128, 128))
    crop_lab = ref.extract_3d_volume(straight_label, cents[20], size=(128, 128, 128))
    back = inter.local_to_global(cents[20], shape=shape)
    np.savez_compressed(
        os.path.join(GOLD, "straighten_0007.npz"), curve=curve, knots=inter.knots, basis=inter.basis, grid_probe=inter.get_grid(shape)[:, ::37, ::31, ::29],
        ct_shape=np.array(straight_ct.shape), probe=probe, ct_probe=straight_ct.reshape(-1)[probe], ct_mean=straight_ct.mean(),
        label_probe=straight_label.reshape(-1)[probe], label_counts=np.array([(straight_label == i).sum() for i in range(17, 25)]),
        label_sha_before_split=np.array(raw_label_sha), label_sha=np.array(hashlib.sha256(straight_label.astype(np.uint8).tobytes()).hexdigest()),
        vert_ids=np.array(sorted(cents)), centroids=np.array([cents[k] for k in sorted(cents)]), back20=back,
        crop_ct_probe=crop_ct.reshape(-1)[probe % crop_ct.size], crop_label_sha=np.array(hashlib.sha256(crop_lab.astype(np.uint8).tobytes()).hexdigest()),
        window_probe=ct_w.reshape(-1)[probe % ct_w.size])
    print("points", inter.knots.shape, "straight", straight_ct.shape, "label counts", [(int(i), int((straight_label == i).sum())) for i in ids])


if __name__ == "__main__":
    main()
