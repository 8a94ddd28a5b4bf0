// Resampling of a volume along a curve: the gather behind the reference's spine straightening
// (straighten/straighten/curve.py:54-101: Interpolator.get_grid + interpolate_along = scipy.ndimage.map_coordinates of the grid
//  knots[n] + basis[n] . (0, g0, g1), order 1 for the CT and order 0 for the label map, mode 'constant').
//
// One thread per output voxel out[n][a][b] (n: point on the curve, a in [0, s1), b in [0, s0)): local coordinates
// (0, b - s0 / 2, a - s1 / 2) - numpy.meshgrid's default 'xy' indexing puts the FIRST grid axis on the LAST output axis - are mapped
// through the local basis of point n, then the input volume is sampled.  map_coordinates' 'constant' mode: a sample with any
// coordinate outside [0, extent - 1] is the fill value (no interpolation beyond the edges); order 0 picks floor(c + 0.5); order 1 is
// the trilinear blend of the 8 neighbours.  All arithmetic in float64 like the reference.  HBM-bound gather: 8 x 8 B reads per output.
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

__global__ void __launch_bounds__(256) resample_curve_kernel(const double* __restrict__ vol, int d0, int d1, int d2,
                                                             const double* __restrict__ knots, const double* __restrict__ basis, int npts,
                                                             int s0, int s1, int order, double cval, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)npts * s0 * s1;
  if (i >= total) return;
  const int b = (int)(i % s0), a = (int)((i / s0) % s1), n = (int)(i / ((long long)s0 * s1));
  const double g1 = (double)b - (double)s0 / 2.0, g2 = (double)a - (double)s1 / 2.0;
  const double* B = basis + (size_t)n * 9;   // basis[n][i][j]: component i of basis vector j
  double c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) c[d] = (B[d * 3 + 1] * g1 + B[d * 3 + 2] * g2) + knots[(size_t)n * 3 + d];
  const int ext[3] = {d0, d1, d2};
  bool inside = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) inside = inside && c[d] >= 0.0 && c[d] <= (double)(ext[d] - 1);
  double v = cval;
  if (inside) {
    if (order == 0) {
      const long long x = (long long)floor(c[0] + 0.5), y = (long long)floor(c[1] + 0.5), z = (long long)floor(c[2] + 0.5);
      v = vol[((size_t)x * d1 + y) * d2 + z];
    } else {
      int lo[3], hi[3];
      double f[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double fl = floor(c[d]);
        lo[d] = (int)fl;
        hi[d] = min(lo[d] + 1, ext[d] - 1);
        f[d] = c[d] - fl;
      }
      v = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = (k & 4) ? hi[0] : lo[0], y = (k & 2) ? hi[1] : lo[1], z = (k & 1) ? hi[2] : lo[2];
        const double w = ((k & 4) ? f[0] : 1.0 - f[0]) * ((k & 2) ? f[1] : 1.0 - f[1]) * ((k & 1) ? f[2] : 1.0 - f[2]);
        v += w * vol[((size_t)x * d1 + y) * d2 + z];
      }
    }
  }
  out[i] = v;
}

}  // namespace hv

extern "C" int hv_resample_curve(const double* vol, int d0, int d1, int d2, const double* knots, const double* basis, int npts, int s0,
                                 int s1, int order, double cval, double* out, hv_stream_t stream) {
  using namespace hv;
  HV_CHECK_ARG(vol && knots && basis && out, "resample_curve: null argument");
  HV_CHECK_ARG(d0 >= 1 && d1 >= 1 && d2 >= 1 && npts >= 1 && s0 >= 1 && s1 >= 1, "resample_curve: bad extents");
  HV_CHECK_ARG(order == 0 || order == 1, "resample_curve: interpolation order %d is not supported (0 = nearest, 1 = linear)", order);
  const long long total = (long long)npts * s0 * s1;
  resample_curve_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(vol, d0, d1, d2, knots, basis, npts, s0, s1, order, cval, out);
  HV_LAUNCH_CHECK();
  return HV_OK;
}
