// Contextual attention, bf16 tensor-core path.
//
// Same closed form as ctx_attn.cu (reference models/inpaint_networks.py:247-410, ksize=3, stride=1, rate=2,
// fuse_k=3), with the two dense contractions on tcgen05 (gemm_tc.cu) and everything kept TRANSPOSED so that the
// softmax runs along contiguous rows and both GEMMs read K-major operands:
//   P   [f][t*64+c]   3x3 patches of the ::2-downsampled feature (bf16, exact copies of the bf16 activations)
//   Rt  [c*16+t][b]   4x4 stride-2 patches of the full feature, transposed
//   T   [f][b]  = sum_k P[f][k] P[b][k] / max(|P_b|, 1e-4)                     fp32   (GEMM 1, column scale)
//   A   [f][b]  = softmax_b(scale * fuse(T) * mm_b) * mm_b                      bf16   (fused fuse + softmax + argmax)
//   cols[ck][f] = sum_b Rt[ck][b] A[f][b]                                       bf16   (GEMM 2)
//   y = fold(cols) / 4 written straight into the chunked bf16 input buffer of pmconv9
#include <cuda_bf16.h>
#include "hv_common.cuh"
#include "kernels.h"
#include "conv_tc.cuh"

namespace hv {

int gemm_tc_nt(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, const float* colscale, int M, int N, int K, int batch,
               long long strideA, long long strideB, int out_bf16, cudaStream_t st);
int ca_mask_launch(const float* mask, float* mm, int n, int side, int mh, int mw, int per_sample, cudaStream_t st);
int ca_offsets_flow_launch(const int32_t* argmax, int32_t* offsets, float* flow, int n, int side, int up, int* scratch, cudaStream_t st);

constexpr int CA_C = 64, CA_H = 64, CA_SIDE = 32, CA_L = 1024, CA_KP = 576, CA_KR = 1024;

// ------------------------------------------------------------------ patches P + column norms (one warp per (n, l))
__global__ void __launch_bounds__(256) ca_tc_patches_kernel(TcBuf f, __nv_bfloat16* __restrict__ P, float* __restrict__ inv_norm) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n = gw / CA_L, l = gw - n * CA_L;
  if (n >= f.n) return;
  const int lh = l >> 5, lw = l & 31;
  uint4* dst = reinterpret_cast<uint4*>(P + ((size_t)n * CA_L + l) * CA_KP);
  float ss = 0.f;
  for (int i = lane; i < 72; i += 32) {  // 9 taps x 8 chunks of 8 channels
    const int t = i >> 3, ch = i & 7, ky = t / 3, kx = t - ky * 3;
    const int y = lh + ky - 1, x = lw + kx - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < CA_SIDE && x >= 0 && x < CA_SIDE)
      v = *reinterpret_cast<const uint4*>(f.ptr + f.chunk_base(n, ch) + f.pos(2 * y, 2 * x) * 8);
    dst[i] = v;
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 q = __bfloat1622float2(h[j]); ss = fmaf(q.x, q.x, ss); ss = fmaf(q.y, q.y, ss); }
  }
  ss = warp_sum(ss);
  if (lane == 0) inv_norm[(size_t)n * CA_L + l] = 1.f / fmaxf(sqrtf(ss), 1e-4f);
}

// ------------------------------------------------------------------ raw patches, transposed: Rt[c*16+ky*4+kx][b]
// one thread = (n, chunk, tap, bh, group of 8 bw): 8 loads of 8 channels, 8x8 transpose in registers, 8 stores of 16 B
__global__ void __launch_bounds__(256) ca_tc_raw_kernel(TcBuf f, __nv_bfloat16* __restrict__ Rt) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = i & 3, bh = (i >> 2) & 31, t = (i >> 7) & 15, ch = (i >> 11) & 7, n = i >> 14;
  if (n >= f.n) return;
  const int ky = t >> 2, kx = t & 3;
  const int y = 2 * bh - 1 + ky;
  __nv_bfloat16 v[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int x = 2 * (8 * g + j) - 1 + kx;
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < CA_H && x >= 0 && x < CA_H)
      raw = *reinterpret_cast<const uint4*>(f.ptr + f.chunk_base(n, ch) + f.pos(y, x) * 8);
    *reinterpret_cast<uint4*>(v[j]) = raw;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[j][c];
    *reinterpret_cast<uint4*>(Rt + ((size_t)n * CA_KR + (size_t)(ch * 8 + c) * 16 + t) * CA_L + bh * CA_SIDE + 8 * g) =
        *reinterpret_cast<const uint4*>(o);
  }
}

// ------------------------------------------------------------------ fuse (:350-361) + masked scaled softmax + argmax (:364-368)
// One warp per foreground row j; lane owns background columns i = lane + 32 k.  In the transposed domain
//   U[j][i] = sum_{cc,a in -1..1} T[cmi(cm(j)+cc)+a][cmi(cm(i)+cc)+a]   (flat indices outside [0,L) contribute 0)
__device__ __forceinline__ int ca_cm(int i) { return ((i & 31) << 5) | (i >> 5); }  // row-major <-> column-major (32x32)

__global__ void __launch_bounds__(256) ca_tc_fuse_softmax_kernel(const float* __restrict__ T, const float* __restrict__ mm,
                                                                 int mm_stride, __nv_bfloat16* __restrict__ A,
                                                                 int32_t* __restrict__ argmax_out, float scale, int fuse) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y, j = blockIdx.x * 8 + warp;
  const float* Tn = T + (size_t)n * CA_L * CA_L;
  const float* m = mm + (size_t)n * mm_stride;
  float u[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) u[k] = 0.f;
  if (fuse) {
    const int cj = ca_cm(j);
#pragma unroll
    for (int cc = -1; cc <= 1; ++cc) {
      const int pj = cj + cc;
      if (pj < 0 || pj >= CA_L) continue;
      const int q = ca_cm(pj);
#pragma unroll
      for (int a = -1; a <= 1; ++a) {
        const int s = q + a;
        if (s < 0 || s >= CA_L) continue;
        const float* row = Tn + (size_t)s * CA_L;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int pi = ca_cm(lane + 32 * k) + cc;
          if (pi < 0 || pi >= CA_L) continue;
          const int r = ca_cm(pi) + a;
          if (r < 0 || r >= CA_L) continue;
          u[k] += __ldg(row + r);
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) u[k] = __ldg(Tn + (size_t)j * CA_L + lane + 32 * k);
  }
  float mk[32];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    mk[k] = __ldg(m + lane + 32 * k);
    u[k] = u[k] * mk[k] * scale;
    mx = fmaxf(mx, u[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) { u[k] = __expf(u[k] - mx); sum += u[k]; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float best = -1.f;
  int bi = 0;
  __nv_bfloat16* out = A + ((size_t)n * CA_L + j) * CA_L;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float a = u[k] * inv * mk[k];
    out[lane + 32 * k] = __float2bfloat16(a);
    if (a > best) { best = a; bi = lane + 32 * k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) argmax_out[(size_t)n * CA_L + j] = bi;
}

// ------------------------------------------------------------------ overlap-add of the pasted patches (:377-379) -> chunked bf16
// one thread = (n, chunk, oy, ox): y[c][oy][ox] = 0.25 * sum over the (<= 4) patches covering the pixel
__global__ void __launch_bounds__(256) ca_tc_fold_kernel(const __nv_bfloat16* __restrict__ cols, TcBuf y) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ox = i & 63, oy = (i >> 6) & 63, ch = (i >> 12) & 7, n = i >> 15;
  if (n >= y.n) return;
  const __nv_bfloat16* cn = cols + (size_t)n * CA_KR * CA_L;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ky = ((oy + 1) & 1) + 2 * a, hf = (oy + 1 - ky) >> 1;
    if (oy + 1 - ky < 0 || hf >= CA_SIDE) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kx = ((ox + 1) & 1) + 2 * b, wf = (ox + 1 - kx) >> 1;
      if (ox + 1 - kx < 0 || wf >= CA_SIDE) continue;
      const __nv_bfloat16* src = cn + (size_t)(ch * 8 * 16 + ky * 4 + kx) * CA_L + hf * CA_SIDE + wf;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += __bfloat162float(src[(size_t)c * 16 * CA_L]);
    }
  }
  uint32_t pk[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * e] * 0.25f, acc[2 * e + 1] * 0.25f);
    pk[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(y.ptr + y.chunk_base(n, ch) + y.pos(oy, ox) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// ------------------------------------------------------------------ host orchestration
struct CaTcWorkspace {
  __nv_bfloat16 *P, *Rt, *A, *cols;
  float *inv_norm, *mm, *T;
  int32_t* argmax;
  int* scratch;
};

static size_t ca_tc_layout(int n, char* base, CaTcWorkspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += (bytes + 1023) & ~(size_t)1023; return p; };
  char* p;
  p = take(sizeof(__nv_bfloat16) * n * CA_L * CA_KP);  if (ws) ws->P = (__nv_bfloat16*)p;
  p = take(sizeof(__nv_bfloat16) * n * CA_KR * CA_L);  if (ws) ws->Rt = (__nv_bfloat16*)p;
  p = take(sizeof(float) * n * CA_L * CA_L);           if (ws) ws->T = (float*)p;
  p = take(sizeof(__nv_bfloat16) * n * CA_L * CA_L);   if (ws) ws->A = (__nv_bfloat16*)p;
  p = take(sizeof(__nv_bfloat16) * n * CA_KR * CA_L);  if (ws) ws->cols = (__nv_bfloat16*)p;
  p = take(sizeof(float) * n * CA_L);                  if (ws) ws->inv_norm = (float*)p;
  p = take(sizeof(float) * n * CA_L);                  if (ws) ws->mm = (float*)p;
  p = take(sizeof(int32_t) * n * CA_L);                if (ws) ws->argmax = (int32_t*)p;
  p = take(sizeof(int) * (n + 1));                     if (ws) ws->scratch = (int*)p;
  return off;
}

size_t ctx_attn_tc_workspace_bytes(int n) { return ca_tc_layout(n, nullptr, nullptr); }

// f: chunked bf16 [n][8 chunks][64x64, any border]; y: chunked bf16 output buffer of the same extent
// st_aux / evs (optional, 3 events): a second stream for the work that is off the critical path patches -> similarity -> softmax ->
// paste -> fold: the raw 4x4 patch matrix and the mask run beside the similarity GEMM, the offsets / flow outputs (which nothing
// downstream consumes) after the arg-max is known
int ctx_attn_fwd_tc(const TcBuf& f, const float* mask, const TcBuf& y, int32_t* offsets, float* flow, float scale, int fuse,
                    int per_sample_mask, void* workspace, cudaStream_t st, cudaStream_t st_aux, cudaEvent_t* evs) {
  HV_CHECK_ARG(f.ptr && y.ptr && mask && workspace, "ctx_attn_fwd_tc: null argument");
  HV_CHECK_ARG(f.chunks == 8 && f.h == CA_H && f.w == CA_H && !f.s2d && y.chunks == 8 && y.h == CA_H && y.w == CA_H && !y.s2d && y.n == f.n,
               "ctx_attn_fwd_tc: built for 64-channel 64x64 feature maps");
  const int n = f.n;
  const bool two = st_aux != nullptr && evs != nullptr;
  cudaStream_t sx = two ? st_aux : st;
  CaTcWorkspace ws;
  ca_tc_layout(n, (char*)workspace, &ws);
  if (two) {
    HV_CUDA(cudaEventRecord(evs[0], st));
    HV_CUDA(cudaStreamWaitEvent(sx, evs[0], 0));
  }
  HV_CUDA(launch_pdl(ca_tc_patches_kernel, dim3((n * CA_L + 7) / 8), dim3(256), 0, st, f, ws.P, ws.inv_norm));
  HV_LAUNCH_CHECK();
  int rc = ca_mask_launch(mask, ws.mm, n, CA_SIDE, 4 * CA_H, 4 * CA_H, per_sample_mask, sx);
  if (rc) return rc;
  HV_CUDA(launch_pdl(ca_tc_raw_kernel, dim3((n * 16384 + 255) / 256), dim3(256), 0, sx, f, ws.Rt));
  HV_LAUNCH_CHECK();
  if (two) HV_CUDA(cudaEventRecord(evs[1], sx));
  rc = gemm_tc_nt(ws.P, ws.P, ws.T, ws.inv_norm, CA_L, CA_L, CA_KP, n, (long long)CA_L * CA_KP, (long long)CA_L * CA_KP, 0, st);
  if (rc) return rc;
  if (two) HV_CUDA(cudaStreamWaitEvent(st, evs[1], 0));
  HV_CUDA(launch_pdl(ca_tc_fuse_softmax_kernel, dim3(CA_L / 8, n), dim3(256), 0, st, (const float*)ws.T, (const float*)ws.mm, (int)CA_L, ws.A, ws.argmax, scale, fuse));
  HV_LAUNCH_CHECK();
  if ((offsets || flow) && two) {
    HV_CUDA(cudaEventRecord(evs[2], st));
    HV_CUDA(cudaStreamWaitEvent(sx, evs[2], 0));
    rc = ca_offsets_flow_launch(ws.argmax, offsets, flow, n, CA_SIDE, 8, ws.scratch, sx);
    if (rc) return rc;
  }
  rc = gemm_tc_nt(ws.Rt, ws.A, ws.cols, nullptr, CA_KR, CA_L, CA_L, n, (long long)CA_KR * CA_L, (long long)CA_L * CA_L, 1, st);
  if (rc) return rc;
  HV_CUDA(launch_pdl(ca_tc_fold_kernel, dim3((n * 32768 + 255) / 256), dim3(256), 0, st, (const __nv_bfloat16*)ws.cols, y));
  HV_LAUNCH_CHECK();
  if ((offsets || flow) && !two) {
    rc = ca_offsets_flow_launch(ws.argmax, offsets, flow, n, CA_SIDE, 8, ws.scratch, st);
    if (rc) return rc;
  }
  return HV_OK;
}

}  // namespace hv
