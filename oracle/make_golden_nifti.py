"""Generate tests/golden/nifti_headers.json from the NIfTI volumes the reference ships (build container only).

TEST INFRASTRUCTURE.  Run from the repo root:  python -m oracle.make_golden_nifti
The reference's files were written by its own toolchain (nibabel, straighten/straighten_mask_3d.py); their headers are the
known-answer vectors for healthivert_gan_b200.nifti: the raw 352 header bytes (hex), the shape / dtype / affine the oracle's
minimal reader (oracle/nifti_min.py) and a by-hand parse give, and the SHA-256 of the voxel bytes.
"""
import gzip
import hashlib
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FILES = ["datasets/straightened/label/0007_20.nii.gz", "datasets/straightened/label/0007_23.nii.gz", "datasets/raw/0007/0007_msk.nii.gz"]


def main():
    out = []
    for rel in FILES:
        raw = gzip.open(os.path.join("/root/reference", rel)).read()
        dim = struct.unpack("<8h", raw[40:56])
        datatype, bitpix = struct.unpack("<2h", raw[70:74])
        off = int(struct.unpack("<f", raw[108:112])[0])
        shape = list(dim[1:1 + dim[0]])
        nbytes = int(np.prod(shape)) * bitpix // 8
        srow = struct.unpack("<12f", raw[280:328])
        vox = np.frombuffer(raw, dtype={64: np.float64, 512: np.uint16}[datatype], count=int(np.prod(shape)), offset=off)
        out.append({
            "file": rel, "header_hex": raw[:352].hex(), "shape": shape, "datatype": datatype, "bitpix": bitpix, "vox_offset": off,
            "sform_code": struct.unpack("<h", raw[254:256])[0], "qform_code": struct.unpack("<h", raw[252:254])[0],
            "affine": [list(srow[0:4]), list(srow[4:8]), list(srow[8:12]), [0.0, 0.0, 0.0, 1.0]],
            "voxel_sha256": hashlib.sha256(raw[off:off + nbytes]).hexdigest(),
            "voxel_sum": float(vox.astype(np.float64).sum()), "voxel_max": float(vox.max()),
            "first_nonzero_flat_index_F": int(np.flatnonzero(vox)[0]),
        })
    with open(os.path.join(ROOT, "tests", "golden", "nifti_headers.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
