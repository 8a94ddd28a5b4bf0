"""Spine straightening: the step in front of the synthesis path.

Reference: `straighten/straighten/curve.py` (the `Interpolator` of the vendored `straighten` package) and the driver
`straighten/straighten_mask_3d.py` (`extend_curve :100-124`, `get_local_basis :155-171`, `window :173-186`,
`remove_spine_labels_after_split :126-142`, `extract_3d_volume :222-247`, `process_mask3d :463-562`).

The curve through the vertebra centroids is resampled at unit arc-length steps, every point gets a local orthonormal basis
(tangent, a vector in the sagittal plane, their cross product), and the CT (trilinear) and the label map (nearest) are gathered on
the planes spanned by the 2nd and 3rd basis vectors.  The host side (a few hundred curve points) stays numpy float64; the gather over
`n_points x 128 x 128` samples is one CUDA kernel (`hv_resample_curve`, csrc/resample.cu) with scipy.ndimage.map_coordinates'
`mode='constant'` semantics.  The per-slice 2-D box masks (`extract_mask_volume`, OpenCV minAreaRect) are not part of this module.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


# ------------------------------------------------------------------------------------------------ curve geometry (host)
def cumulative_length(curve):
    """curve.py:204-207."""
    seg = np.linalg.norm(np.diff(curve, axis=0), axis=1)
    return np.concatenate([[0.0], np.cumsum(seg)])


def _interp_rows(x_new, x, y):
    """Piecewise-linear interpolation of the rows of y (scipy.interpolate.interp1d(x, y, axis=0), its slope form)."""
    x_new = np.asarray(x_new, dtype=np.float64)
    hi = np.clip(np.searchsorted(x, x_new, side="left"), 1, len(x) - 1)
    lo = hi - 1
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])[:, None]
    return slope * (x_new - x[lo])[:, None] + y[lo]


def get_derivatives(curve, step):
    """curve.py:210-221: the curve resampled every `step` of arc length and its first `dim` numerical derivatives
    (numpy.gradient along the ORIGINAL points, then the same resampling)."""
    curve = np.asarray(curve, dtype=np.float64)
    lengths = cumulative_length(curve)
    xs = np.arange(0, lengths[-1], step)
    out = [_interp_rows(xs, lengths, curve)]
    grad = curve
    for _ in range(curve.shape[1]):
        grad = np.gradient(grad, axis=0)
        out.append(_interp_rows(xs, lengths, grad))
    return out


def frenet_serret(*gradients):
    """curve.py:11-23: Gram-Schmidt over the derivatives."""
    basis = []
    for grad in gradients:
        e = grad
        for v in basis:
            e = e - v * (v * grad).sum(axis=-1, keepdims=True)
        e = e / np.linalg.norm(e, axis=-1, keepdims=True)
        basis.append(e)
    return np.stack(basis, -1)


def get_local_basis(grad, *args):
    """straighten_mask_3d.py:155-171: tangent, a second vector in the sagittal (axis 0 / axis 2) plane, their cross product."""
    grad = grad / np.linalg.norm(grad, axis=1, keepdims=True)
    sagittal = grad[:, [0, 2]]
    second = sagittal[:, ::-1] * [1, -1]
    dets = np.linalg.det(np.stack([sagittal, second], -1))
    second = second * dets[:, None]
    second = second / np.linalg.norm(second, axis=1, keepdims=True)
    second = np.insert(second, 1, np.zeros_like(second[:, 0]), axis=1)
    third = np.cross(second, grad)
    return np.stack([grad, second, third], -1)


def extend_curve(curve, extension_length, min_bounds, max_bounds):
    """straighten_mask_3d.py:100-124: one extra point beyond each end, along the end segments, clamped to the volume."""
    curve = np.asarray(curve, dtype=np.float64)
    lo, hi = np.asarray(min_bounds, dtype=np.float64), np.asarray(max_bounds, dtype=np.float64)

    def beyond(p, q):
        d = p - q
        return np.minimum(np.maximum(p + d / np.linalg.norm(d) * extension_length, lo), hi)
    return np.vstack([beyond(curve[0], curve[1]), curve, beyond(curve[-1], curve[-2])])


def _interpolate_coords(coordinates, distance_to_origin, distance_to_plane):
    """curve.py:224-239: the curve point whose normal plane contains the query point (linear interpolation over the sign change
    of the signed plane distance nearest to the closest knot, extrapolating like interp1d(fill_value='extrapolate'))."""
    idx = int(distance_to_origin.argmin())
    candidates, = np.diff(np.sign(distance_to_plane)).nonzero()
    if len(candidates) > 0:
        idx = int(candidates[np.abs(candidates - idx).argmin()])
    sl = slice(max(0, idx - 2), idx + 2)
    x, y = distance_to_plane[sl], coordinates[sl]
    order = np.argsort(x, kind="mergesort")          # interp1d sorts its abscissae
    x, y = x[order], y[order]
    hi = int(np.clip(np.searchsorted(x, 0.0, side="left"), 1, len(x) - 1))
    lo = hi - 1
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    return slope * (0.0 - x[lo]) + y[lo]


class Interpolator:
    """curve.py:26-157 for unit spacing: `knots` (evenly spaced curve points), `basis` [n, 3, 3] (basis[n][:, j] = j-th local vector)."""

    def __init__(self, curve, step=1, get_local_basis=frenet_serret):
        curve = np.asarray(curve, dtype=np.float64)
        if curve.ndim != 2 or curve.shape[1] != 3:
            raise ValueError(f"The curve shape must be (n_points, 3), but {curve.shape} provided.")
        if not np.isfinite(curve).all():
            raise ValueError("The curve must contain only finite values.")
        even_curve, *grads = get_derivatives(curve, step)
        self.dim = 3
        self.knots = even_curve
        self.basis = get_local_basis(*grads)

    def get_grid(self, shape):
        """[3, n_points, s1, s0] sampling coordinates (host copy of what the kernel computes; for tests and small cases)."""
        s0, s1 = (int(v) for v in np.broadcast_to(shape, 2))
        g0, g1 = np.meshgrid(np.arange(s0) - s0 / 2, np.arange(s1) - s1 / 2)
        local = np.stack([np.zeros_like(g0), g0, g1])                             # [3, s1, s0]
        grid = np.einsum("Nij,j...->Ni...", self.basis, local)                    # [n, 3, s1, s0]
        return np.moveaxis(grid + self.knots[:, :, None, None], 1, 0)

    def interpolate_along(self, array, shape, fill_value=0, order=1, return_device=False):
        """map_coordinates(array, self.get_grid(shape), order=order, cval=fill_value) on the GPU -> [n_points, s1, s0] float64."""
        if callable(fill_value):
            fill_value = fill_value(array)
        s0, s1 = (int(v) for v in np.broadcast_to(shape, 2))
        dev = torch.device("cuda", torch.cuda.current_device())
        vol = array if isinstance(array, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(array, dtype=np.float64))
        vol = vol.to(dev, dtype=torch.float64).contiguous()
        if vol.dim() != 3:
            raise ValueError("interpolate_along expects a 3-D volume")
        knots = torch.as_tensor(np.ascontiguousarray(self.knots)).to(dev)
        basis = torch.as_tensor(np.ascontiguousarray(self.basis)).to(dev)
        n = knots.shape[0]
        out = torch.empty(n, s1, s0, device=dev, dtype=torch.float64)
        check(_lib.lib().hv_resample_curve(ptr(vol), vol.shape[0], vol.shape[1], vol.shape[2], ptr(knots), ptr(basis), n, s0, s1, int(order),
                                           float(fill_value), ptr(out), _lib.stream()))
        return out if return_device else out.cpu().numpy()

    def _centers(self, shape):
        centers = np.zeros_like(self.knots)
        centers[:, 0] = cumulative_length(self.knots)
        centers[:, 1:] = np.asarray(shape, dtype=np.float64) / 2
        return centers

    def global_to_local(self, point, shape):
        """curve.py:103-129: image coordinates -> (arc length, in-plane coordinates) of the straightened volume."""
        shape = np.broadcast_to(shape, 2)
        rel = np.asarray(point, dtype=np.float64) - self.knots
        to_origin = np.linalg.norm(rel, axis=-1)
        local = np.einsum("nji,nj->ni", self.basis, rel)
        return _interpolate_coords(local + self._centers(shape), to_origin, local[:, 0])

    def local_to_global(self, point, shape):
        """curve.py:109-137."""
        shape = np.broadcast_to(shape, 2)
        rel = np.asarray(point, dtype=np.float64) - self._centers(shape)
        glob = np.einsum("nij,nj->ni", self.basis, rel)
        return _interpolate_coords(glob + self.knots, np.linalg.norm(glob, axis=-1), rel[:, 0])


# ------------------------------------------------------------------------------------------------ volume helpers (host)
def window(img, win_min, win_max):
    """straighten_mask_3d.py:173-186 (bone window -300 .. 800 -> 0 .. 255); returns a new array."""
    img = np.asarray(img, dtype=np.float64)
    if img.max() < win_max and img.min() > win_min:
        return img.copy()
    return np.clip(255.0 * (img - win_min) / (win_max - win_min), 0, 255)


def remove_spine_labels_after_split(label_image):
    """straighten_mask_3d.py:126-142: behind the first in-plane row (from the centre outwards) where a vertebra's label no longer
    touches the central column, that label is removed (drops the posterior elements)."""
    out = np.array(label_image, copy=True)
    _, height, width = out.shape
    for lab in np.unique(out):
        if lab == 0:
            continue
        column_has = (out[:, height // 2:, width // 2] == lab).any(axis=0)
        missing = np.nonzero(~column_has)[0]
        if missing.size:
            h = height // 2 + int(missing[0])
            tail = out[:, h:, :]
            tail[tail == lab] = 0
    return out


def extract_3d_volume(data, center, size=(128, 128, 64)):
    """straighten_mask_3d.py:222-247: crop of `size` centred on `center`, zero padded where it leaves the volume."""
    x, y, z = center
    dx, dy, dz = size
    z_min, z_max = max(0, int(z - dz // 2)), min(data.shape[2], int(z + dz // 2))
    y_min, y_max = max(0, int(y - dy // 2)), min(data.shape[1], int(y + dy // 2))
    x_min, x_max = max(0, int(x - dx // 2)), min(data.shape[0], int(x + dx // 2))
    piece = data[x_min:x_max, y_min:y_max, z_min:z_max]
    out = np.zeros(size, dtype=data.dtype)
    sx, sy, sz = (dx - (x_max - x_min)) // 2, (dy - (y_max - y_min)) // 2, (dz - (z_max - z_min)) // 2
    if sz < 0:
        out[sx:sx + (x_max - x_min), sy:sy + (y_max - y_min), 0:size[2]] = piece[:, :, 0:size[2]]
    else:
        out[sx:sx + (x_max - x_min), sy:sy + (y_max - y_min), sz:sz + (z_max - z_min)] = piece
    return out


def straighten_case(ct_data, label_data, entries, vertebrae_ids, outputsize=(128, 128, 128), shape=(128, 128)):
    """The array-level body of process_mask3d (straighten_mask_3d.py:463-562): `entries` is the centroid list of the case's json
    ([{"label", "X", "Y", "Z"}, ...]).  Returns (straight_ct, straight_label, {vertebra id: (ct crop, label crop, local centroid)})."""
    coords = [[e["X"], e["Y"], e["Z"]] for e in entries if isinstance(e, dict) and "X" in e]
    ct = window(ct_data, -300, 800)
    inter = None
    if len(coords) > 1:
        curve = extend_curve(np.array(coords), 20, (0, 0, 0), label_data.shape)
        inter = Interpolator(curve, step=1, get_local_basis=get_local_basis)
        straight_ct = inter.interpolate_along(ct, shape, order=1)
        straight_label = inter.interpolate_along(label_data, shape, order=0)
    else:
        straight_ct, straight_label = ct, np.asarray(label_data, dtype=np.float64)
    straight_label = remove_spine_labels_after_split(straight_label)
    crops = {}
    for vid in vertebrae_ids:
        centroid = None
        for e in entries:
            if isinstance(e, dict) and e.get("label") == vid:
                centroid = (e["X"], e["Y"], e["Z"])
                if inter is not None:
                    centroid = inter.global_to_local(centroid, shape=shape)
        if centroid is None:
            continue
        crops[vid] = (extract_3d_volume(straight_ct, centroid, size=outputsize), extract_3d_volume(straight_label, centroid, size=outputsize),
                      np.asarray(centroid, dtype=np.float64))
    return straight_ct, straight_label, crops
