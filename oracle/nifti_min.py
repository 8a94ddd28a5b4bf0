"""Minimal NIfTI-1 (.nii.gz) reader used by the oracle tests (no nibabel in the image).

TEST INFRASTRUCTURE.  Header facts: dim @40 (8 x int16), datatype @70 (int16),
vox_offset @108 (float32); voxel data is Fortran-ordered.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
           256: np.int8, 512: np.uint16, 768: np.uint32}


def load(path):
    with gzip.open(path, "rb") as fh:
        raw = fh.read()
    dim = struct.unpack("<8h", raw[40:56])
    datatype = struct.unpack("<h", raw[70:72])[0]
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    shape = tuple(dim[1:1 + dim[0]])
    count = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=_DTYPES[datatype], count=count, offset=vox_offset)
    return data.reshape(shape, order="F")
