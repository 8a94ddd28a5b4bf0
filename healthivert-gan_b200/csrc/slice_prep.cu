// Device-side slice preparation and post-processing of the iterative eval loop:
// everything run_model does around the generator forward (reference eval_3d_sagittal_twostage.py:46-133) except the
// forward itself, batched over all slices of a stage, on uint8 planes:
//   * label == vert_id, 8-connected component labelling, components < 50 px removed (:16-30, :47-50)
//   * row bounds x1/x2, height, the height > 40 re-centring, the 40-row window (:51-72)
//   * mask (41 rows, inclusive, :75), shifted CT / CAM rows with uint8 truncation (:76-82), ToTensor + Normalize (:84-94)
//   * after the forward: (x+1)*127.5, uint8 truncation for the next stage, stitched label map (:119-130)
// Integer / byte outputs are bit-exact; the fp32 normalisation uses IEEE division / no FMA contraction so that the
// planes equal the reference's torchvision transforms bit for bit.
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

// ------------------------------------------------------------------ volume -> uint8 slices
// out[s][r][c] = (uint8) trunc(vol(r, c, s) * scale); the volume is [d0][d1][d2] C-contiguous float64, slices along `axis`
// (2: sagittal vol[:, :, s]; 1: coronal vol[:, s, :]); r = axis 0 index, c = the remaining axis.
__global__ void vol_to_u8_kernel(const double* __restrict__ vol, uint8_t* __restrict__ out, int d0, int d1, int d2, int axis, double scale,
                                 size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int ncol = axis == 2 ? d1 : d2;
  for (; i < total; i += stride) {
    const int c = i % ncol, r = (i / ncol) % d0, s = i / ((size_t)ncol * d0);
    const size_t src = axis == 2 ? ((size_t)r * d1 + c) * d2 + s : ((size_t)r * d1 + s) * d2 + c;
    const double v = vol[src] * scale;
    out[i] = (uint8_t)(long long)v;   // numpy astype(uint8) of an in-range float64: truncation toward zero
  }
}

int vol_to_u8(const double* vol, uint8_t* out, int d0, int d1, int d2, int axis, double scale, cudaStream_t st) {
  HV_CHECK_ARG(vol && out && (axis == 1 || axis == 2), "vol_to_u8: bad argument");
  const size_t total = (size_t)d0 * d1 * d2;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  vol_to_u8_kernel<<<(unsigned)blocks, 256, 0, st>>>(vol, out, d0, d1, d2, axis, scale, total);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// counts[s][j] = number of pixels of slice s equal to ids[j], j < 3  (eval:204,:213 "neighbour present" tests, z-range)
__global__ void __launch_bounds__(256) slice_id_counts_kernel(const uint8_t* __restrict__ label, int hw, int id0, int id1, int id2,
                                                              int32_t* __restrict__ counts) {
  __shared__ float red[32];
  const uint8_t* p = label + (size_t)blockIdx.x * hw;
  int c0 = 0, c1 = 0, c2 = 0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int v = p[i];
    c0 += v == id0; c1 += v == id1; c2 += v == id2;
  }
  const float s0 = block_sum((float)c0, red), s1 = block_sum((float)c1, red), s2 = block_sum((float)c2, red);
  if (threadIdx.x == 0) {
    counts[blockIdx.x * 3 + 0] = (int)s0; counts[blockIdx.x * 3 + 1] = (int)s1; counts[blockIdx.x * 3 + 2] = (int)s2;
  }
}

int slice_id_counts(const uint8_t* label, int nslices, int hw, int id0, int id1, int id2, int32_t* counts, cudaStream_t st) {
  HV_CHECK_ARG(label && counts && nslices > 0, "slice_id_counts: bad argument");
  slice_id_counts_kernel<<<nslices, 256, 0, st>>>(label, hw, id0, id1, id2, counts);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ CCL + bounds (one CTA of 1024 threads per slice)
// labels live in a per-slice int32 scratch plane (L2-resident); label = min pixel index of the component.
// meta[b][8] = (valid, x1, x2, height, min_x, max_x, kept pixels, 0)
__global__ void __launch_bounds__(1024) ccl_bounds_kernel(const uint8_t* __restrict__ label_planes, const int32_t* __restrict__ slice_idx,
                                                          const int32_t* __restrict__ vert_ids, int h, int w, int min_size, int maxheight,
                                                          int32_t* __restrict__ lab_scratch, uint8_t* __restrict__ keep_out,
                                                          int32_t* __restrict__ meta) {
  const int b = blockIdx.x, hw = h * w;
  const uint8_t* src = label_planes + (size_t)slice_idx[b] * hw;
  const int vid = vert_ids[b];
  int32_t* lab = lab_scratch + (size_t)b * hw;
  int32_t* size = lab_scratch + ((size_t)gridDim.x + b) * hw;   // second half of the scratch: component sizes at the root pixel
  __shared__ int s_changed;
  __shared__ int s_min, s_max, s_cnt;
  __shared__ unsigned long long s_rowsum;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) { lab[i] = src[i] == vid ? i : -1; size[i] = 0; }
  __syncthreads();
  // iterate: min over the 8-neighbourhood, then pointer jumping, until nothing changes
  for (int iter = 0; iter < 4096; ++iter) {
    if (threadIdx.x == 0) s_changed = 0;
    __syncthreads();
    int changed = 0;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      int l = lab[i];
      if (l < 0) continue;
      const int y = i / w, x = i - y * w;
      int m = l;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= w) continue;
          const int nl = lab[yy * w + xx];
          if (nl >= 0 && nl < m) m = nl;
        }
      }
      // pointer jumping: follow the chain of representatives
      int r = lab[m];
      while (r >= 0 && r < m) { m = r; r = lab[m]; }
      if (m < l) { lab[i] = m; changed = 1; }
    }
    if (changed) s_changed = 1;
    __syncthreads();
    if (!s_changed) break;
    __syncthreads();
  }
  // component sizes (histogram on the root pixel), then drop the small ones
  for (int i = threadIdx.x; i < hw; i += blockDim.x) if (lab[i] >= 0) atomicAdd(&size[lab[i]], 1);
  if (threadIdx.x == 0) { s_min = h; s_max = -1; s_cnt = 0; s_rowsum = 0ull; }
  __syncthreads();
  int mn = h, mx = -1, cnt = 0;
  unsigned long long rs = 0;
  uint8_t* keep = keep_out + (size_t)b * hw;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int l = lab[i];
    const bool k = l >= 0 && size[l] >= min_size;
    keep[i] = k ? 1 : 0;
    if (k) { const int y = i / w; mn = min(mn, y); mx = max(mx, y); ++cnt; rs += (unsigned long long)y; }
  }
  atomicMin(&s_min, mn); atomicMax(&s_max, mx); atomicAdd(&s_cnt, cnt); atomicAdd(&s_rowsum, rs);
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t* m = meta + b * 8;
    if (s_cnt == 0) { m[0] = 0; for (int j = 1; j < 8; ++j) m[j] = 0; return; }
    int x1 = s_min, x2 = s_max;
    const int height = x2 - x1;
    if (height > maxheight) {           // eval:57-60 (height itself is NOT updated, the reference returns the original)
      const int x_mean = (int)(s_rowsum / (unsigned long long)s_cnt);   // int(np.mean(rows)) of non-negative values
      x1 = x_mean - 20;
      x2 = x1 + 40;
    }
    const int mask_x = (x1 + x2) >> 1;  // floor division (non-negative)
    const int h2 = maxheight;
    int min_x, max_x;
    if (mask_x <= h2 / 2) { min_x = 0; max_x = h2; }
    else if (2 * (h - mask_x) <= h2) { max_x = h; min_x = max_x - h2; }   // width - mask_x <= h2 / 2 (true division)
    else { min_x = mask_x - h2 / 2; max_x = min_x + h2; }
    m[0] = 1; m[1] = x1; m[2] = x2; m[3] = height; m[4] = min_x; m[5] = max_x; m[6] = s_cnt; m[7] = 0;
  }
}

// ------------------------------------------------------------------ build the generator inputs of a batch of slices
// ct / mask / cam1m (= 1 - CAM) / ori_ct: [B][1][h][w] fp32;  ToTensor (u8 / 255) + Normalize((x - 0.5) / 0.5) in IEEE fp32.
__device__ __forceinline__ float norm_u8(int u) { return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), 0.5f), 0.5f); }

__global__ void __launch_bounds__(256) slice_build_kernel(const uint8_t* __restrict__ ct_planes, const uint8_t* __restrict__ cam_planes,
                                                          const int32_t* __restrict__ slice_idx, const int32_t* __restrict__ meta, int h, int w,
                                                          float* __restrict__ ct, float* __restrict__ mask, float* __restrict__ cam1m,
                                                          float* __restrict__ ori, int32_t* __restrict__ x1o, int32_t* __restrict__ x2o,
                                                          int32_t* __restrict__ ho) {
  const int b = blockIdx.y, hw = h * w;
  const int32_t* m = meta + b * 8;
  const int valid = m[0], x1 = m[1], x2 = m[2], min_x = m[4], max_x = m[5];
  if (blockIdx.x == 0 && threadIdx.x == 0) { x1o[b] = x1; x2o[b] = x2; ho[b] = m[3]; }
  const uint8_t* cs = ct_planes + (size_t)slice_idx[b] * hw;
  const uint8_t* ms = cam_planes + (size_t)slice_idx[b] * hw;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const int r = i / w, c = i - r * w;
  int ctv = 0, camv = 0, mk = 0;
  if (valid) {
    int srow = -1;
    if (r < min_x) srow = x1 - min_x + r;            // ct_slice[:min_x] = ct[(x1 - min_x):x1]
    else if (r >= max_x) srow = x2 + (r - max_x);    // ct_slice[max_x:] = ct[x2 : x2 + (width - max_x)]
    if (srow >= 0 && srow < h) { ctv = cs[srow * w + c]; camv = ms[srow * w + c]; }
    mk = (r >= min_x && r <= max_x) ? 255 : 0;       // mask_slice[min_x : max_x + 1] = 255 (41 rows, eval:75)
  }
  const size_t o = (size_t)b * hw + i;
  ct[o] = norm_u8(ctv);
  mask[o] = __fdiv_rn((float)mk, 255.f);
  cam1m[o] = __fsub_rn(1.f, __fdiv_rn((float)camv, 255.f));
  ori[o] = norm_u8(cs[i]);
}

// ------------------------------------------------------------------ after the forward + hv_stitch
// fake_ct: stitched CT in [-1, 1] (hv_stitch of x_stage2 with ori_ct); rows: hv_stitch rows_out (hgt, d, xu, xb).
// ct_out (may be null): fp32 0..255 plane written into the output volume slice; ct_u8_next / label_next: the uint8 planes the
// next stage of the same slice reads (numpy astype(uint8) truncation of the fp32 value, eval:49,:86).
__global__ void __launch_bounds__(256) slice_finish_kernel(const float* __restrict__ fake_ct, const float* __restrict__ fine_seg,
                                                           const int32_t* __restrict__ rows, const int32_t* __restrict__ meta,
                                                           const int32_t* __restrict__ slice_idx, const int32_t* __restrict__ vert_ids,
                                                           const uint8_t* __restrict__ label_in, int h, int w, float* __restrict__ ct_out,
                                                           float* __restrict__ label_out, uint8_t* __restrict__ ct_u8_next,
                                                           uint8_t* __restrict__ label_next) {
  const int b = blockIdx.y, hw = h * w;
  const int32_t* m = meta + b * 8;
  if (!m[0]) return;                                  // run_model returned None: the driver leaves the slice untouched
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const int r = i / w, c = i - r * w;
  const int x1 = m[1], x2 = m[2], d = rows[b * 4 + 1], xu = rows[b * 4 + 2], xb = rows[b * 4 + 3];
  const size_t plane = (size_t)slice_idx[b] * hw;
  const float v = __fmul_rn(__fadd_rn(fake_ct[(size_t)b * hw + i], 1.f), 127.5f);   // (fake_B + 1) * 127.5, eval:121
  if (ct_out) ct_out[plane + i] = v;
  if (ct_u8_next) ct_u8_next[plane + i] = (uint8_t)(int)v;
  int lab;
  if (r >= xu && r < xb) lab = fine_seg[(size_t)b * hw + i] > 0.5f ? vert_ids[b] : 0;      // eval:105,:124
  else {
    const int srow = r < xu ? d / 2 + r : x2 + (r - xb);                                    // eval:125-128
    lab = (srow >= 0 && srow < h && (r < xu ? srow < x1 : true)) ? label_in[plane + srow * w + c] : 0;
  }
  if (label_out) label_out[plane + i] = (float)lab;
  if (label_next) label_next[plane + i] = (uint8_t)lab;
}

int slice_prepare(const uint8_t* label_planes, const uint8_t* ct_planes, const uint8_t* cam_planes, const int32_t* slice_idx,
                  const int32_t* vert_ids, int nb, int h, int w, int maxheight, int32_t* lab_scratch, uint8_t* keep_scratch, int32_t* meta,
                  float* ct, float* mask, float* cam1m, float* ori, int32_t* x1, int32_t* x2, int32_t* height, cudaStream_t st) {
  HV_CHECK_ARG(label_planes && ct_planes && cam_planes && slice_idx && vert_ids && lab_scratch && keep_scratch && meta && ct && mask && cam1m &&
                   ori && x1 && x2 && height && nb > 0,
               "slice_prepare: bad argument");
  ccl_bounds_kernel<<<nb, 1024, 0, st>>>(label_planes, slice_idx, vert_ids, h, w, 50, maxheight, lab_scratch, keep_scratch, meta);
  HV_LAUNCH_CHECK();
  slice_build_kernel<<<dim3((h * w + 255) / 256, nb), 256, 0, st>>>(ct_planes, cam_planes, slice_idx, meta, h, w, ct, mask, cam1m, ori, x1, x2,
                                                                     height);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

int slice_finish(const float* fake_ct, const float* fine_seg, const int32_t* rows, const int32_t* meta, const int32_t* slice_idx,
                 const int32_t* vert_ids, const uint8_t* label_in, int nb, int h, int w, float* ct_out, float* label_out, uint8_t* ct_u8_next,
                 uint8_t* label_next, cudaStream_t st) {
  HV_CHECK_ARG(fake_ct && fine_seg && rows && meta && slice_idx && vert_ids && label_in && nb > 0, "slice_finish: bad argument");
  HV_CHECK_ARG(label_next != label_in, "slice_finish: the label planes are read with a row shift, they cannot be updated in place");
  slice_finish_kernel<<<dim3((h * w + 255) / 256, nb), 256, 0, st>>>(fake_ct, fine_seg, rows, meta, slice_idx, vert_ids, label_in, h, w, ct_out,
                                                                      label_out, ct_u8_next, label_next);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv

using namespace hv;
extern "C" {

int hv_vol_to_u8(const double* vol, uint8_t* out, int d0, int d1, int d2, int axis, double scale, hv_stream_t s) {
  return vol_to_u8(vol, out, d0, d1, d2, axis, scale, as_stream(s));
}
int hv_slice_id_counts(const uint8_t* label, int nslices, int hw, int id0, int id1, int id2, int32_t* counts, hv_stream_t s) {
  return slice_id_counts(label, nslices, hw, id0, id1, id2, counts, as_stream(s));
}
int hv_slice_prepare(const uint8_t* label_planes, const uint8_t* ct_planes, const uint8_t* cam_planes, const int32_t* slice_idx,
                     const int32_t* vert_ids, int nb, int h, int w, int maxheight, int32_t* lab_scratch, uint8_t* keep_scratch, int32_t* meta,
                     float* ct, float* mask, float* cam1m, float* ori, int32_t* x1, int32_t* x2, int32_t* height, hv_stream_t s) {
  return slice_prepare(label_planes, ct_planes, cam_planes, slice_idx, vert_ids, nb, h, w, maxheight, lab_scratch, keep_scratch, meta, ct, mask,
                       cam1m, ori, x1, x2, height, as_stream(s));
}
int hv_slice_finish(const float* fake_ct, const float* fine_seg, const int32_t* rows, const int32_t* meta, const int32_t* slice_idx,
                    const int32_t* vert_ids, const uint8_t* label_in, int nb, int h, int w, float* ct_out, float* label_out, uint8_t* ct_u8_next,
                    uint8_t* label_next, hv_stream_t s) {
  return slice_finish(fake_ct, fine_seg, rows, meta, slice_idx, vert_ids, label_in, nb, h, w, ct_out, label_out, ct_u8_next, label_next,
                      as_stream(s));
}

}  // extern "C"
