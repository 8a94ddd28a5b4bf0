// Bandwidth-bound mask / integer kernels of the path: SHRM height head, threshold,
// height-adaptive stitching, Sobel edges + XOR-count edge loss, per-column vertebral heights.
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

// ------------------------------------------------------------------ GAP -> Linear(c,1) -> sigmoid
// reference models/inpaint_networks.py:90-93,:211-214.  One CTA per sample.
__global__ void __launch_bounds__(256) gap_fc_sigmoid_kernel(const float* __restrict__ x, const float* __restrict__ fw,
                                                             const float* __restrict__ fb, float* __restrict__ out,
                                                             int c, int hw) {
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* xn = x + (size_t)n * c * hw;
  __shared__ float red[32];
  float acc = 0.f;
  for (int ch = warp; ch < c; ch += nw) {
    float s = 0.f;
    const float4* p = reinterpret_cast<const float4*>(xn + (size_t)ch * hw);
    for (int i = lane; i < hw / 4; i += 32) { float4 v = p[i]; s += (v.x + v.y) + (v.z + v.w); }
    for (int i = (hw / 4) * 4 + lane; i < hw; i += 32) s += xn[(size_t)ch * hw + i];
    s = warp_sum(s);
    acc += (s / (float)hw) * fw[ch];
  }
  float tot = 0.f;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nw; ++i) tot += red[i];
    out[n] = 1.f / (1.f + expf(-(tot + fb[0])));
  }
}

int gap_fc_sigmoid(const float* x, const float* fw, const float* fb, float* out, int n, int c, int hw, cudaStream_t st) {
  HV_CHECK_ARG(x && fw && fb && out && n > 0 && c > 0 && hw > 0, "gap_fc_sigmoid: bad argument");
  HV_CHECK_ARG(hw % 4 == 0, "gap_fc_sigmoid: hw must be a multiple of 4");
  gap_fc_sigmoid_kernel<<<n, 256, 0, st>>>(x, fw, fb, out, c, hw);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ threshold (pix2pix_model.py:201-202, eval:105)
__global__ void threshold_kernel(const float* __restrict__ p, float* __restrict__ of, uint8_t* __restrict__ ou,
                                 float value, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    const bool on = p[i] > 0.5f;
    if (of) of[i] = on ? value : 0.f;
    if (ou) ou[i] = on ? (uint8_t)value : (uint8_t)0;
  }
}

int threshold(const float* p, float* of, uint8_t* ou, float value, size_t count, cudaStream_t st) {
  HV_CHECK_ARG(p && (of || ou), "threshold: null argument");
  if (count == 0) return HV_OK;
  int blocks = (int)((count + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  threshold_kernel<<<blocks, 256, 0, st>>>(p, of, ou, value, count);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ height-adaptive stitch
// pix2pix_model.py:206-252 / eval_3d_sagittal_twostage.py:103-118.  One CTA row-block per
// (row, sample); the integer row arithmetic is done on the device (no .item() syncs).
__global__ void __launch_bounds__(256) stitch_kernel(const float* __restrict__ gen, const float* __restrict__ real,
                                                     const float* __restrict__ pred_h, const int32_t* __restrict__ x1,
                                                     const int32_t* __restrict__ x2, const int32_t* __restrict__ height,
                                                     int maxheight, float* __restrict__ out, int32_t* __restrict__ rows_out,
                                                     int h, int w) {
  const int n = blockIdx.y, r = blockIdx.x;
  const int pred = (int)ceilf(pred_h[n] * (float)maxheight);
  const int hh = height[n];
  const int hgt = pred < hh ? hh : pred;
  const int d = hgt - hh;
  const int xu = x1[n] - d / 2, xb = xu + hgt;
  if (rows_out && r == 0 && threadIdx.x == 0) {
    rows_out[n * 4 + 0] = hgt; rows_out[n * 4 + 1] = d; rows_out[n * 4 + 2] = xu; rows_out[n * 4 + 3] = xb;
  }
  const float* src;
  int sr;
  if (r >= xu && r < xb) { src = gen; sr = r; }
  else if (r < xu) { src = real; sr = d / 2 + r; }
  else { src = real; sr = x2[n] + (r - xb); }
  const size_t plane = (size_t)h * w;
  float* o = out + n * plane + (size_t)r * w;
  if (sr < 0 || sr >= h) {
    for (int x = threadIdx.x; x < w; x += blockDim.x) o[x] = 0.f;
    return;
  }
  const float* s = src + n * plane + (size_t)sr * w;
  for (int x = threadIdx.x; x < w; x += blockDim.x) o[x] = s[x];
}

int stitch(const float* gen, const float* real, const float* pred_h, const int32_t* x1, const int32_t* x2,
           const int32_t* height, int maxheight, float* out, int32_t* rows_out, int n, int h, int w, cudaStream_t st) {
  HV_CHECK_ARG(gen && real && pred_h && x1 && x2 && height && out, "stitch: null argument");
  HV_CHECK_ARG(n > 0 && n <= 65535 && h > 0 && w > 0, "stitch: bad extent");
  stitch_kernel<<<dim3(h, n), w >= 256 ? 256 : 64, 0, st>>>(gen, real, pred_h, x1, x2, height, maxheight, out, rows_out, h, w);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ Sobel (models/edge_operator.py:41-48)
__device__ __forceinline__ float sobel_at(const float* __restrict__ img, int y, int x, int h, int w) {
  const int ym = max(y - 1, 0), yp = min(y + 1, h - 1), xm = max(x - 1, 0), xp = min(x + 1, w - 1);
  const float a = img[ym * w + xm], b = img[ym * w + x], c = img[ym * w + xp];
  const float d = img[y * w + xm], f = img[y * w + xp];
  const float g = img[yp * w + xm], hh = img[yp * w + x], i = img[yp * w + xp];
  const float gx = (-a + c) + (-2.f * d + 2.f * f) + (-g + i);
  const float gy = (a + 2.f * b + c) - (g + 2.f * hh + i);
  return fminf(sqrtf(gx * gx + gy * gy), 1.f);
}

__global__ void __launch_bounds__(256) sobel_kernel(const float* __restrict__ img, float* __restrict__ edges, int h, int w) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  const size_t plane = (size_t)h * w;
  edges[n * plane + i] = sobel_at(img + n * plane, i / w, i % w, h, w);
}

int sobel(const float* img, float* edges, int n, int h, int w, cudaStream_t st) {
  HV_CHECK_ARG(img && edges && n > 0 && n <= 65535 && h > 0 && w > 0, "sobel: bad argument");
  sobel_kernel<<<dim3((h * w + 255) / 256, n), 256, 0, st>>>(img, edges, h, w);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// edge loss = 800 * mean((e_f - e_r)^2); on {0,1} masks e in {0,1} so it is an XOR count.
// General inputs are handled too: the squared difference is accumulated in fp32 per CTA, fp64 across CTAs.
__global__ void __launch_bounds__(256) edge_loss_kernel(const float* __restrict__ fake, const float* __restrict__ real,
                                                        unsigned long long* __restrict__ xor_count, double* __restrict__ sq_sum,
                                                        int h, int w) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)h * w;
  float sq = 0.f;
  int x = 0;
  if (i < h * w) {
    const float ef = sobel_at(fake + n * plane, i / w, i % w, h, w);
    const float er = sobel_at(real + n * plane, i / w, i % w, h, w);
    const float df = ef - er;
    sq = df * df;
    x = (ef > 0.f) != (er > 0.f);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, x);
  sq = warp_sum(sq);
  __shared__ float s_sq[8];
  __shared__ int s_x[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_sq[warp] = sq; s_x[warp] = __popc(bal); }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    int c = 0;
    for (int k = 0; k < 8; ++k) { t += s_sq[k]; c += s_x[k]; }
    if (c) atomicAdd(xor_count, (unsigned long long)c);
    if (t != 0.f) atomicAdd(sq_sum, (double)t);
  }
}

__global__ void edge_loss_finish_kernel(const double* sq_sum, float* loss, double inv_count) {
  *loss = (float)(800.0 * (*sq_sum) * inv_count);
}

int edge_xor_loss(const float* fake, const float* real, unsigned long long* xor_count, float* loss, int n, int h, int w,
                  cudaStream_t st) {
  HV_CHECK_ARG(fake && real && xor_count && loss && n > 0 && n <= 65535 && h > 0 && w > 0, "edge_xor_loss: bad argument");
  double* d_sq = static_cast<double*>(stream_scratch(st, sizeof(double)));  // per (device, stream): no shared accumulator
  HV_CHECK_ARG(d_sq, "edge_xor_loss: scratch allocation failed");
  HV_CUDA(cudaMemsetAsync(d_sq, 0, sizeof(double), st));
  HV_CUDA(cudaMemsetAsync(xor_count, 0, sizeof(unsigned long long), st));
  edge_loss_kernel<<<dim3((h * w + 255) / 256, n), 256, 0, st>>>(fake, real, xor_count, d_sq, h, w);
  HV_LAUNCH_CHECK();
  edge_loss_finish_kernel<<<1, 1, 0, st>>>(d_sq, loss, 1.0 / ((double)n * h * w));
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ the whole tail of Pix2PixModel.forward in ONE pass
// models/pix2pix_model.py:201-264 + the edge loss of :349: threshold of both segmentation heads, height-adaptive stitch of both CT
// heads, the masked centre crops, Sobel of the real and of the thresholded fake mask and the XOR-count edge loss.  Seven planes in,
// eight planes out, one launch: the separate kernels are launch-bound at this size (8 - 12 MB each in ~10 us = 0.1 - 0.15 of the HBM
// peak, profiles/r2_hbm_kernels.md); every output is bit-identical to theirs (tests/test_gpu_ops.py).
struct PostArgs {
  const float *fine, *coarse, *x2s, *x1s, *real, *real_mask, *mask, *pred2, *pred1;
  const int32_t *x1, *x2, *height;
  float *fake_mask, *coarse_bin, *fake_B, *fake_B_coarse, *fake_local, *real_local, *real_edges, *fake_edges;
  int32_t *rows_fine, *rows_coarse;
  double* sq_sum;                    // scratch: {double sq_sum; u64 xor; u32 ticket}
  unsigned long long* xor_out;
  float* loss_out;
  int maxheight, c0, c1, h, w, n;
};

__device__ __forceinline__ float thr01(float v) { return v > 0.5f ? 1.f : 0.f; }
// sobel_at on the thresholded plane
__device__ __forceinline__ float sobel_thr_at(const float* __restrict__ img, int y, int x, int h, int w) {
  const int ym = max(y - 1, 0), yp = min(y + 1, h - 1), xm = max(x - 1, 0), xp = min(x + 1, w - 1);
  const float a = thr01(img[ym * w + xm]), b = thr01(img[ym * w + x]), c = thr01(img[ym * w + xp]);
  const float d = thr01(img[y * w + xm]), f = thr01(img[y * w + xp]);
  const float g = thr01(img[yp * w + xm]), hh = thr01(img[yp * w + x]), i = thr01(img[yp * w + xp]);
  const float gx = (-a + c) + (-2.f * d + 2.f * f) + (-g + i);
  const float gy = (a + 2.f * b + c) - (g + 2.f * hh + i);
  return fminf(sqrtf(gx * gx + gy * gy), 1.f);
}

__device__ __forceinline__ float stitch_at(const float* __restrict__ gen, const float* __restrict__ real, int r, int x, int xu, int xb, int d,
                                           int x2, int h, int w) {
  int sr;
  const float* src;
  if (r >= xu && r < xb) { src = gen; sr = r; }
  else if (r < xu) { src = real; sr = d / 2 + r; }
  else { src = real; sr = x2 + (r - xb); }
  return (sr < 0 || sr >= h) ? 0.f : src[(size_t)sr * w + x];
}

__global__ void __launch_bounds__(256) post_forward_kernel(const PostArgs p) {
  const int n = blockIdx.y, h = p.h, w = p.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)h * w, o = (size_t)n * plane + i;
  // integer row arithmetic of the two stitches (the same expressions as stitch_kernel)
  const int hh = p.height[n];
  const int pf = (int)ceilf(p.pred2[n] * (float)p.maxheight), pc = (int)ceilf(p.pred1[n] * (float)p.maxheight);
  const int hf = pf < hh ? hh : pf, hc = pc < hh ? hh : pc;
  const int df = hf - hh, dc = hc - hh;
  const int xuf = p.x1[n] - df / 2, xbf = xuf + hf, xuc = p.x1[n] - dc / 2, xbc = xuc + hc;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.rows_fine[n * 4 + 0] = hf; p.rows_fine[n * 4 + 1] = df; p.rows_fine[n * 4 + 2] = xuf; p.rows_fine[n * 4 + 3] = xbf;
    p.rows_coarse[n * 4 + 0] = hc; p.rows_coarse[n * 4 + 1] = dc; p.rows_coarse[n * 4 + 2] = xuc; p.rows_coarse[n * 4 + 3] = xbc;
  }
  float sq = 0.f;
  int xr = 0;
  if (i < h * w) {
    const int r = i / w, x = i - r * w;
    const float* real = p.real + n * plane;
    p.fake_mask[o] = thr01(p.fine[o]);
    p.coarse_bin[o] = thr01(p.coarse[o]);
    const float fb = stitch_at(p.x2s + n * plane, real, r, x, xuf, xbf, df, p.x2[n], h, w);
    p.fake_B[o] = fb;
    p.fake_B_coarse[o] = stitch_at(p.x1s + n * plane, real, r, x, xuc, xbc, dc, p.x2[n], h, w);
    const bool centre = x >= p.c0 && x < p.c1;
    const float m = p.mask[o];
    p.fake_local[o] = centre ? fb * m : 0.f;
    p.real_local[o] = centre ? real[i] * m : 0.f;
    const float er = sobel_at(p.real_mask + n * plane, r, x, h, w);
    const float ef = sobel_thr_at(p.fine + n * plane, r, x, h, w);
    p.real_edges[o] = er;
    p.fake_edges[o] = ef;
    const float dd = ef - er;
    sq = dd * dd;
    xr = (ef > 0.f) != (er > 0.f);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, xr);
  sq = warp_sum(sq);
  __shared__ float s_sq[8];
  __shared__ int s_x[8];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_sq[warp] = sq; s_x[warp] = __popc(bal); }
  __syncthreads();
  unsigned long long* xor_acc = reinterpret_cast<unsigned long long*>(p.sq_sum + 1);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(p.sq_sum + 2);
  if (threadIdx.x == 0) {
    float t = 0.f;
    int c = 0;
    for (int k = 0; k < 8; ++k) { t += s_sq[k]; c += s_x[k]; }
    if (c) atomicAdd(xor_acc, (unsigned long long)c);
    if (t != 0.f) atomicAdd(p.sq_sum, (double)t);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    if (last) {   // every CTA has added its share: finish the loss (800 * mean squared difference, pix2pix_model.py:109, :349)
      __threadfence();
      *p.loss_out = (float)(800.0 * __ldcg(p.sq_sum) / ((double)p.n * h * w));
      *p.xor_out = __ldcg(xor_acc);
    }
  }
}

int post_forward(const PostArgs& a, cudaStream_t st) {
  HV_CHECK_ARG(a.fine && a.coarse && a.x2s && a.x1s && a.real && a.real_mask && a.mask && a.pred2 && a.pred1 && a.x1 && a.x2 && a.height &&
                   a.fake_mask && a.coarse_bin && a.fake_B && a.fake_B_coarse && a.fake_local && a.real_local && a.real_edges && a.fake_edges &&
                   a.rows_fine && a.rows_coarse && a.xor_out && a.loss_out,
               "post_forward: null argument");
  HV_CHECK_ARG(a.n > 0 && a.n <= 65535 && a.h > 0 && a.w > 0, "post_forward: bad extent");
  PostArgs q = a;
  q.sq_sum = static_cast<double*>(stream_scratch(st, 24));
  HV_CHECK_ARG(q.sq_sum, "post_forward: scratch allocation failed");
  HV_CUDA(cudaMemsetAsync(q.sq_sum, 0, 24, st));
  post_forward_kernel<<<dim3((a.h * a.w + 255) / 256, a.n), 256, 0, st>>>(q);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ per-column heights (RHLV_quantification.py:41-73)
// The slice z of a [d0,d1,d2] u8 volume is viewed as [rows = d0][cols]: sagittal (axis 2)
// cols = d1, coronal (axis 1) cols = d2.  Pass 1 counts non-zeros per (slice, column) with the
// thread index running along the contiguous volume axis (z for sagittal, column for coronal) so
// every warp reads whole 32 B sectors.  Pass 2 (one CTA per slice) derives the thirds / centre
// columns and splits the counts into all / pre / mid / post.
__global__ void __launch_bounds__(256) column_count_kernel(const uint8_t* __restrict__ fake, const uint8_t* __restrict__ label,
                                                           int d0, int d1, int d2, int axis, int z0, int nz,
                                                           int32_t* __restrict__ counts) {
  const uint8_t* vol = blockIdx.z ? label : fake;
  const int slot = blockIdx.z ? 4 : 0;
  const int ncols = axis == 2 ? d1 : d2;
  const long long row_stride = (long long)d1 * d2;
  if (axis == 2) {  // thread -> slice (contiguous), block -> column
    const int c = blockIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nz) return;
    const uint8_t* p = vol + (long long)c * d2 + z0 + s;
    int cnt = 0;
#pragma unroll 8
    for (int r = 0; r < d0; ++r) cnt += p[r * row_stride] != 0;
    counts[((size_t)s * 8 + slot) * ncols + c] = cnt;
  } else {          // thread -> column (contiguous), block -> slice
    const int s = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const uint8_t* p = vol + (long long)(z0 + s) * d2 + c;
    int cnt = 0;
#pragma unroll 8
    for (int r = 0; r < d0; ++r) cnt += p[r * row_stride] != 0;
    counts[((size_t)s * 8 + slot) * ncols + c] = cnt;
  }
}

__global__ void __launch_bounds__(256) column_split_kernel(int ncols, int32_t* __restrict__ counts, int32_t* __restrict__ meta) {
  const int s = blockIdx.x;
  int32_t* cnt = counts + (size_t)s * 8 * ncols;
  __shared__ int s_i[8];
  if (threadIdx.x == 0) {
    // np.where(slice)[1] statistics: min / max column and the mean column index (count-weighted)
    long long nf = 0, sf = 0, nl = 0, sl = 0;
    int ymin = -1, ymax = -1;
    for (int c = 0; c < ncols; ++c) {
      const int cf = cnt[c], cl = cnt[4 * ncols + c];
      if (cf) { if (ymin < 0) ymin = c; ymax = c; nf += cf; sf += (long long)cf * c; }
      nl += cl; sl += (long long)cl * c;
    }
    const int valid = nf > 0 && nl > 0;
    int t1 = 0, t2 = 0, cenf = 0, cenl = 0;
    if (valid) {
      const int yr = ymax - ymin;
      t1 = (int)((double)ymin + (double)yr / 3.0);
      t2 = (int)((double)ymin + (double)(2 * yr) / 3.0);
      cenf = cnt[(int)((double)sf / (double)nf)];
      cenl = cnt[4 * ncols + (int)((double)sl / (double)nl)];
    }
    s_i[0] = valid; s_i[1] = t1; s_i[2] = t2; s_i[3] = cenf; s_i[4] = cenl; s_i[5] = ymin; s_i[6] = ymax; s_i[7] = 0;
    for (int k = 0; k < 8; ++k) meta[s * 8 + k] = s_i[k];
  }
  __syncthreads();
  const int valid = s_i[0], t1 = s_i[1], t2 = s_i[2];
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    const int cf = valid ? cnt[c] : 0, cl = valid ? cnt[4 * ncols + c] : 0;
    const bool pre = c < t1, mid = c >= t1 && c < t2, post = c >= t2;
    cnt[0 * ncols + c] = cf;
    cnt[1 * ncols + c] = pre ? cf : 0;
    cnt[2 * ncols + c] = mid ? cf : 0;
    cnt[3 * ncols + c] = post ? cf : 0;
    cnt[4 * ncols + c] = cl;
    cnt[5 * ncols + c] = pre ? cl : 0;
    cnt[6 * ncols + c] = mid ? cl : 0;
    cnt[7 * ncols + c] = post ? cl : 0;
  }
}

int column_heights(const uint8_t* vf, const uint8_t* vl, int d0, int d1, int d2, int axis, int z0, int z1,
                   int32_t* counts, int32_t* meta, cudaStream_t st) {
  HV_CHECK_ARG(vf && vl && counts && meta, "column_heights: null argument");
  HV_CHECK_ARG(axis == 1 || axis == 2, "column_heights: axis must be 1 (coronal) or 2 (sagittal)");
  const int nzt = axis == 2 ? d2 : d1, ncols = axis == 2 ? d1 : d2;
  HV_CHECK_ARG(d0 > 0 && d1 > 0 && d2 > 0 && ncols <= 65535, "column_heights: bad extent");
  HV_CHECK_ARG(z0 >= 0 && z1 <= nzt && z0 <= z1, "column_heights: slice window [%d,%d) outside [0,%d)", z0, z1, nzt);
  const int nz = z1 - z0;
  if (nz == 0) return HV_OK;
  HV_CHECK_ARG(nz <= 65535, "column_heights: window too long");
  if (axis == 2) {
    const int bt = nz >= 256 ? 256 : (nz >= 128 ? 128 : (nz >= 64 ? 64 : 32));
    column_count_kernel<<<dim3((nz + bt - 1) / bt, ncols, 2), bt, 0, st>>>(vf, vl, d0, d1, d2, axis, z0, nz, counts);
  } else {
    column_count_kernel<<<dim3((ncols + 255) / 256, nz, 2), 256, 0, st>>>(vf, vl, d0, d1, d2, axis, z0, nz, counts);
  }
  HV_LAUNCH_CHECK();
  column_split_kernel<<<nz, 256, 0, st>>>(ncols, counts, meta);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv

extern "C" int hv_post_forward(const float* fine_seg, const float* coarse_seg, const float* x_stage2, const float* x_stage1, const float* real_B,
                               const float* real_B_mask, const float* mask, const float* pred2_h, const float* pred1_h, const int32_t* x1,
                               const int32_t* x2, const int32_t* height, int maxheight, int c0, int c1, float* fake_B_mask_raw,
                               float* coarse_seg_binary, float* fake_B, float* fake_B_coarse, float* fake_B_local, float* real_B_local,
                               float* real_edges, float* fake_edges, int32_t* rows_fine, int32_t* rows_coarse, unsigned long long* xor_count,
                               float* edge_loss, int n, int h, int w, hv_stream_t stream) {
  hv::PostArgs a;
  a.fine = fine_seg; a.coarse = coarse_seg; a.x2s = x_stage2; a.x1s = x_stage1; a.real = real_B; a.real_mask = real_B_mask; a.mask = mask;
  a.pred2 = pred2_h; a.pred1 = pred1_h; a.x1 = x1; a.x2 = x2; a.height = height;
  a.fake_mask = fake_B_mask_raw; a.coarse_bin = coarse_seg_binary; a.fake_B = fake_B; a.fake_B_coarse = fake_B_coarse;
  a.fake_local = fake_B_local; a.real_local = real_B_local; a.real_edges = real_edges; a.fake_edges = fake_edges;
  a.rows_fine = rows_fine; a.rows_coarse = rows_coarse; a.sq_sum = nullptr; a.xor_out = xor_count; a.loss_out = edge_loss;
  a.maxheight = maxheight; a.c0 = c0; a.c1 = c1; a.h = h; a.w = w; a.n = n;
  return hv::post_forward(a, hv::as_stream(stream));
}
