"""NIfTI-1 single-file (.nii / .nii.gz) reader and writer for the volume driver.

The reference reads and writes its volumes with nibabel (`eval_3d_sagittal_twostage.py:163-181,:236-239`,
`evaluation/RHLV_quantification.py:159-167`): `nib.load(p).get_fdata()` (float64, scl_slope / scl_inter applied when they are
meaningful), `ct_nii.affine`, `nib.save(nib.Nifti1Image(data, affine), p)`.  nibabel is not a dependency of this package; this module
covers exactly that surface:

    img = nifti.load(path)          # NiftiImage: .dataobj (on-disk dtype), .affine, .header (dict), .get_fdata()
    nifti.save(path, data, affine)  # data dtype is kept (float64 volumes stay float64 like the reference's outputs)

Format facts used (NIfTI-1.1 header, 348 bytes + 4 extension bytes, little or big endian, voxel data Fortran-ordered):
sizeof_hdr @0, dim @40, datatype @70, bitpix @72, pixdim @76, vox_offset @108, scl_slope @112, scl_inter @116, qform_code @252,
sform_code @254, quatern_b..d / qoffset_x..z @256, srow_x..z @280, magic @344.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32,
           1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


class NiftiError(ValueError):
    pass


def parse_header(raw):
    """Decode the 348-byte header (bytes-like, at least 348 long).  Returns a dict; `endian` is '<' or '>'."""
    if len(raw) < 348:
        raise NiftiError("NIfTI header shorter than 348 bytes")
    endian = "<"
    if struct.unpack("<i", raw[0:4])[0] != 348:
        if struct.unpack(">i", raw[0:4])[0] != 348:
            raise NiftiError("not a NIfTI-1 file (sizeof_hdr != 348)")
        endian = ">"
    u = lambda fmt, a, b: struct.unpack(endian + fmt, raw[a:b])
    magic = bytes(raw[344:348])
    if magic not in (b"n+1\0", b"ni1\0"):
        raise NiftiError(f"unsupported NIfTI magic {magic!r}")
    if magic == b"ni1\0":
        raise NiftiError("two-file NIfTI (.hdr/.img) is not supported")
    dim = u("8h", 40, 56)
    if not 1 <= dim[0] <= 7:
        raise NiftiError(f"bad dim[0] = {dim[0]}")
    h = {
        "endian": endian, "dim": dim, "datatype": u("h", 70, 72)[0], "bitpix": u("h", 72, 74)[0], "pixdim": u("8f", 76, 108),
        "vox_offset": u("f", 108, 112)[0], "scl_slope": u("f", 112, 116)[0], "scl_inter": u("f", 116, 120)[0],
        "xyzt_units": raw[123], "descrip": bytes(raw[148:228]).rstrip(b"\0"), "qform_code": u("h", 252, 254)[0],
        "sform_code": u("h", 254, 256)[0], "quatern": u("3f", 256, 268), "qoffset": u("3f", 268, 280),
        "srow": np.array(u("12f", 280, 328), dtype=np.float64).reshape(3, 4), "magic": magic,
    }
    if h["datatype"] not in _DTYPES:
        raise NiftiError(f"unsupported NIfTI datatype code {h['datatype']}")
    return h


def _quatern_to_affine(h):
    b, c, d = (float(v) for v in h["quatern"])
    a2 = 1.0 - (b * b + c * c + d * d)
    if a2 < 1e-7:
        s = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d, a = b * s, c * s, d * s, 0.0
    else:
        a = np.sqrt(a2)
    r = np.array([[a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c],
                  [2 * b * c + 2 * a * d, a * a + c * c - b * b - d * d, 2 * c * d - 2 * a * b],
                  [2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a + d * d - c * c - b * b]])
    qfac = -1.0 if h["pixdim"][0] < 0 else 1.0
    zooms = np.array([h["pixdim"][1], h["pixdim"][2], h["pixdim"][3] * qfac], dtype=np.float64)
    aff = np.eye(4)
    aff[:3, :3] = r * zooms
    aff[:3, 3] = h["qoffset"]
    return aff


def header_affine(h):
    """nibabel's choice: sform if coded, else qform if coded, else the pixdim scaling."""
    if h["sform_code"] > 0:
        aff = np.eye(4)
        aff[:3, :] = h["srow"]
        return aff
    if h["qform_code"] > 0:
        return _quatern_to_affine(h)
    # neither code set: nibabel's get_base_affine() = shape_zoom_affine(shape, zooms, x_flip=True): voxel sizes on the diagonal
    # with the x axis flipped (radiological default of Nifti1Header) and the origin at the centre voxel of the first 3 dimensions
    ndim = int(h["dim"][0])
    shape = np.ones(3)
    zooms = np.ones(3)
    for i in range(min(ndim, 3)):
        shape[i] = float(h["dim"][1 + i])
        zooms[i] = float(h["pixdim"][1 + i])
    zooms[0] *= -1.0
    aff = np.eye(4)
    aff[:3, :3] = np.diag(zooms)
    aff[:3, 3] = -((shape - 1.0) / 2.0) * zooms
    return aff


class NiftiImage:
    def __init__(self, dataobj, affine, header):
        self.dataobj, self.affine, self.header = dataobj, affine, header

    @property
    def shape(self):
        return self.dataobj.shape

    def get_fdata(self):
        """float64 copy with scl_slope / scl_inter applied when the slope is finite and non-zero (nibabel's rule)."""
        out = np.array(self.dataobj, dtype=np.float64)
        slope, inter = float(self.header["scl_slope"]), float(self.header["scl_inter"])
        if np.isfinite(slope) and slope != 0.0:
            inter = inter if np.isfinite(inter) else 0.0
            if slope != 1.0 or inter != 0.0:
                out = out * slope + inter
        return out


def _read_all(path):
    with open(path, "rb") as fh:
        head = fh.read(2)
        fh.seek(0)
        if head == b"\x1f\x8b":
            with gzip.GzipFile(fileobj=fh) as gz:
                return gz.read()
        return fh.read()


def load(path):
    raw = _read_all(path)
    h = parse_header(raw)
    shape = tuple(int(v) for v in h["dim"][1:1 + h["dim"][0]])   # all dim[0] dimensions as stored, like nibabel's dataobj
    dt = np.dtype(_DTYPES[h["datatype"]]).newbyteorder(h["endian"])
    off = int(h["vox_offset"]) or 352
    count = int(np.prod(shape))
    if off + count * dt.itemsize > len(raw):
        raise NiftiError(f"{path}: file holds {len(raw) - off} voxel bytes, header promises {count * dt.itemsize}")
    data = np.frombuffer(raw, dtype=dt, count=count, offset=off).reshape(shape, order="F")
    if h["endian"] == ">":
        data = data.astype(dt.newbyteorder("<"))
    return NiftiImage(data, header_affine(h), h)


def _affine_to_quatern(aff):
    """Rotation part of `aff` as the NIfTI quaternion (b, c, d), the voxel sizes and qfac (nifti1_io mat44_to_quatern)."""
    r = np.array(aff[:3, :3], dtype=np.float64)
    zooms = np.sqrt((r * r).sum(axis=0))
    zooms[zooms == 0] = 1.0
    r = r / zooms
    qfac = 1.0
    if np.linalg.det(r) < 0:
        r[:, 2] = -r[:, 2]
        qfac = -1.0
    # nearest orthogonal matrix (polar decomposition) so that shear / rounding does not leak into the quaternion
    uu, _, vt = np.linalg.svd(r)
    r = uu @ vt
    a = r[0, 0] + r[1, 1] + r[2, 2] + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b = 0.25 * (r[2, 1] - r[1, 2]) / a
        c = 0.25 * (r[0, 2] - r[2, 0]) / a
        d = 0.25 * (r[1, 0] - r[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + r[0, 0] - (r[1, 1] + r[2, 2]), 1.0 + r[1, 1] - (r[0, 0] + r[2, 2]), 1.0 + r[2, 2] - (r[0, 0] + r[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd); c = 0.25 * (r[0, 1] + r[1, 0]) / b; d = 0.25 * (r[0, 2] + r[2, 0]) / b; a = 0.25 * (r[2, 1] - r[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd); b = 0.25 * (r[0, 1] + r[1, 0]) / c; d = 0.25 * (r[1, 2] + r[2, 1]) / c; a = 0.25 * (r[0, 2] - r[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd); b = 0.25 * (r[0, 2] + r[2, 0]) / d; c = 0.25 * (r[1, 2] + r[2, 1]) / d; a = 0.25 * (r[1, 0] - r[0, 1]) / d
        if a < 0:
            b, c, d = -b, -c, -d
    return (b, c, d), zooms, qfac


def build_header(shape, dtype, affine):
    """348 + 4 header bytes for a little-endian single-file NIfTI-1 image: sform (code 2, 'aligned') carries the affine,
    qform_code 0 with the quaternion filled in, no intensity scaling (slope 1, intercept 0) - what nibabel writes for
    `Nifti1Image(data, affine)`."""
    dt = np.dtype(dtype)
    key = dt.str[1:]
    if key not in _CODES:
        raise NiftiError(f"dtype {dt} has no NIfTI-1 datatype code")
    if not 1 <= len(shape) <= 7:
        raise NiftiError("NIfTI-1 stores 1 to 7 dimensions")
    aff = np.asarray(affine, dtype=np.float64)
    if aff.shape != (4, 4):
        raise NiftiError("affine must be 4 x 4")
    (qb, qc, qd), zooms, qfac = _affine_to_quatern(aff)
    dim = [len(shape)] + [int(v) for v in shape] + [1] * (7 - len(shape))
    pixdim = [qfac, zooms[0], zooms[1], zooms[2], 1.0, 1.0, 1.0, 1.0]
    h = bytearray(352)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, *dim)
    struct.pack_into("<2h", h, 70, _CODES[key], dt.itemsize * 8)
    struct.pack_into("<8f", h, 76, *pixdim)
    struct.pack_into("<3f", h, 108, 352.0, 1.0, 0.0)
    struct.pack_into("<2h", h, 252, 0, 2)
    struct.pack_into("<6f", h, 256, qb, qc, qd, aff[0, 3], aff[1, 3], aff[2, 3])
    struct.pack_into("<12f", h, 280, *aff[:3, :].reshape(-1))
    h[344:348] = b"n+1\0"
    return bytes(h)


def save(path, data, affine, compresslevel=1):
    """Write `data` (its dtype is kept) with `affine`; `.gz` paths are gzip-compressed (level 1 by default: the float64 volumes
    of the pipeline are mostly zeros and compress 100x at any level, the level only costs time)."""
    arr = np.asarray(data)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    head = build_header(arr.shape, arr.dtype, affine)
    body = np.asfortranarray(arr).tobytes(order="F")
    if str(path).endswith(".gz"):
        with open(path, "wb") as fh, gzip.GzipFile(filename="", mode="wb", fileobj=fh, compresslevel=compresslevel, mtime=0) as gz:
            gz.write(head)
            gz.write(body)
    else:
        with open(path, "wb") as fh:
            fh.write(head)
            fh.write(body)
