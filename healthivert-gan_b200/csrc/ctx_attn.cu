// Contextual attention, fp32 parity path (dense closed form).
//
// Replaces ContextualAttention.forward(f, f, mask) (reference models/inpaint_networks.py:247-410,
// helpers models/inpaint_tools.py:7-70) for ksize=3, stride=1, rate=2, fuse_k=3:
//   P  = 3x3 patches of the ::2-downsampled feature      [L, 9c]
//   R  = 4x4 stride-2 patches of the full feature         [L, 16c]
//   S  = diag(1/max(|P_b|,1e-4)) P P^T                    [L_b, L_f]        (tensor contraction 1)
//   U  = two flat-index diagonal 3-tap sums of S (row-major, then column-major ordering)
//   A  = softmax_b(scale * U * mm_b) * mm_b ; argmax_b A
//   y  = fold(R^T A) / 4                                   (tensor contraction 2 + overlap-add)
// Per-sample intermediates live in a caller-provided workspace.
#include <cuda_bf16.h>
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

// ------------------------------------------------------------------ batched SGEMM (SIMT)
constexpr int GM = 128, GN = 128, GK = 8;

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ C, const float* __restrict__ rowscale,
                                                    int M, int N, int K, long long sA, long long sB,
                                                    long long sC, long long sS) {
  __shared__ __align__(16) float As[2][GK][GM];
  __shared__ __align__(16) float Bs[2][GK][GN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  A += blockIdx.z * sA; B += blockIdx.z * sB; C += blockIdx.z * sC;
  if (rowscale) rowscale += blockIdx.z * sS;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto gload = [&](int k0) {
    if (A_KMAJOR) { int m = tid >> 1, k = (tid & 1) * 4; ra = *reinterpret_cast<const float4*>(A + (size_t)(m0 + m) * K + k0 + k); }
    else { int k = tid >> 5, m = (tid & 31) * 4; ra = *reinterpret_cast<const float4*>(A + (size_t)(k0 + k) * M + m0 + m); }
    if (B_KMAJOR) { int n = tid >> 1, k = (tid & 1) * 4; rb = *reinterpret_cast<const float4*>(B + (size_t)(n0 + n) * K + k0 + k); }
    else { int k = tid >> 5, n = (tid & 31) * 4; rb = *reinterpret_cast<const float4*>(B + (size_t)(k0 + k) * N + n0 + n); }
  };
  auto sstore = [&](int buf) {
    if (A_KMAJOR) { int m = tid >> 1, k = (tid & 1) * 4; As[buf][k][m] = ra.x; As[buf][k + 1][m] = ra.y; As[buf][k + 2][m] = ra.z; As[buf][k + 3][m] = ra.w; }
    else { int k = tid >> 5, m = (tid & 31) * 4; *reinterpret_cast<float4*>(&As[buf][k][m]) = ra; }
    if (B_KMAJOR) { int n = tid >> 1, k = (tid & 1) * 4; Bs[buf][k][n] = rb.x; Bs[buf][k + 1][n] = rb.y; Bs[buf][k + 2][n] = rb.z; Bs[buf][k + 3][n] = rb.w; }
    else { int k = tid >> 5, n = (tid & 31) * 4; *reinterpret_cast<float4*>(&Bs[buf][k][n]) = rb; }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / GK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    const float sc = rowscale ? rowscale[m] : 1.f;
    float4 o0 = make_float4(acc[i][0] * sc, acc[i][1] * sc, acc[i][2] * sc, acc[i][3] * sc);
    float4 o1 = make_float4(acc[i][4] * sc, acc[i][5] * sc, acc[i][6] * sc, acc[i][7] * sc);
    *reinterpret_cast<float4*>(C + (size_t)m * N + n0 + tx * 4) = o0;
    *reinterpret_cast<float4*>(C + (size_t)m * N + n0 + 64 + tx * 4) = o1;
  }
}

int sgemm_batched(const float* A, const float* B, float* C, const float* rowscale, int M, int N, int K,
                  bool a_kmajor, bool b_kmajor, long long sA, long long sB, long long sC, long long sS,
                  int batch, cudaStream_t st) {
  HV_CHECK_ARG(M % GM == 0 && N % GN == 0 && K % GK == 0, "sgemm: M,N must be multiples of 128 and K of 8 (got %d,%d,%d)", M, N, K);
  dim3 grid(N / GN, M / GM, batch);
  if (a_kmajor && b_kmajor) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(A, B, C, rowscale, M, N, K, sA, sB, sC, sS);
  else if (!a_kmajor && !b_kmajor) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(A, B, C, rowscale, M, N, K, sA, sB, sC, sS);
  else if (a_kmajor) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(A, B, C, rowscale, M, N, K, sA, sB, sC, sS);
  else sgemm_kernel<false, true><<<grid, 256, 0, st>>>(A, B, C, rowscale, M, N, K, sA, sB, sC, sS);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ the same contractions on the tensor cores
// Training in the tensor-core mode: every dense contraction of the module (2 forward, 4 backward; 19 - 34 GFLOP each at batch 16) runs
// on gemm_tc_nt with operands ROUNDED TO BF16 and fp32 accumulation; inputs, outputs and the workspace tensors the backward reads stay
// fp32.  The operands are cast (and transposed where the contraction index is not the contiguous one) into the stream's scratch.
int gemm_tc_nt(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, const float* colscale, int M, int N, int K, int batch,
               long long strideA, long long strideB, int out_bf16, cudaStream_t st);

// dst bf16 [b][rows][cols] = src fp32 [b][rows][cols] * rowscale[b][row]; one thread = 8 consecutive columns
__global__ void __launch_bounds__(256) ca_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, const float* __restrict__ rowscale,
                                                      int rows, int cols, long long src_stride, long long scale_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = cols >> 3;
  if (i >= (long long)rows * c8) return;
  const int b = blockIdx.y, r = (int)(i / c8), c0 = (int)(i - (long long)r * c8) * 8;
  const float sc = rowscale ? rowscale[(size_t)b * scale_stride + r] : 1.f;
  const float4* p = reinterpret_cast<const float4*>(src + (size_t)b * src_stride + (size_t)r * cols + c0);
  const float4 a = __ldg(p), c = __ldg(p + 1);
  __align__(16) __nv_bfloat16 v[8] = {__float2bfloat16(a.x * sc), __float2bfloat16(a.y * sc), __float2bfloat16(a.z * sc), __float2bfloat16(a.w * sc),
                                      __float2bfloat16(c.x * sc), __float2bfloat16(c.y * sc), __float2bfloat16(c.z * sc), __float2bfloat16(c.w * sc)};
  *reinterpret_cast<uint4*>(dst + ((size_t)b * rows + r) * cols + c0) = *reinterpret_cast<const uint4*>(v);
}

// dst bf16 [b][cols][rows] = transpose of src fp32 [b][rows][cols]; 32 x 32 tiles through shared memory; grid (cols / 32, rows / 32, b)
__global__ void __launch_bounds__(256) ca_tcast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols, long long src_stride) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (size_t)b * src_stride;
#pragma unroll
  for (int r = ty; r < 32; r += 8) tile[r][tx] = s[(size_t)(r0 + r) * cols + c0 + tx];
  __syncthreads();
#pragma unroll
  for (int c = ty; c < 32; c += 8) dst[((size_t)b * cols + c0 + c) * rows + r0 + tx] = __float2bfloat16(tile[tx][c]);
}

// same contract as sgemm_batched (C[m][n] = rowscale[m] * sum_k A(m,k) B(n,k); a_kmajor: A stored [M][K], else [K][M]; same for B)
static int tc_contract(const float* A, const float* B, float* C, const float* rowscale, int M, int N, int K, bool a_kmajor, bool b_kmajor,
                       long long sA, long long sB, long long sC, long long sS, int batch, cudaStream_t st) {
  HV_CHECK_ARG(M % 128 == 0 && N % 128 == 0 && K % 64 == 0 && sC == (long long)M * N, "ctx_attn (tensor-core mode): contraction %d x %d x %d", M, N, K);
  const size_t a_bytes = ((size_t)batch * M * K * 2 + 255) & ~(size_t)255, b_bytes = (size_t)batch * N * K * 2;
  char* scratch = static_cast<char*>(stream_scratch(st, a_bytes + b_bytes));
  HV_CHECK_ARG(scratch, "ctx_attn (tensor-core mode): no memory for %zu bytes of operand scratch", a_bytes + b_bytes);
  __nv_bfloat16* Ab = reinterpret_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* Bb = reinterpret_cast<__nv_bfloat16*>(scratch + a_bytes);
  auto stage = [&](const float* src, __nv_bfloat16* dst, int R, bool kmajor, long long stride, const float* scale) -> int {
    if (kmajor) {
      ca_cast_kernel<<<dim3((unsigned)(((long long)R * (K / 8) + 255) / 256), batch), 256, 0, st>>>(src, dst, scale, R, K, stride, sS);
    } else {
      HV_CHECK_ARG(!scale, "ctx_attn (tensor-core mode): row scale on a transposed operand");
      ca_tcast_kernel<<<dim3(R / 32, K / 32, batch), 256, 0, st>>>(src, dst, K, R, stride);   // stored [K][R] -> [R][K]
    }
    HV_LAUNCH_CHECK();
    return HV_OK;
  };
  int rc = stage(A, Ab, M, a_kmajor, sA, rowscale);
  if (rc) return rc;
  if (B == A && a_kmajor == b_kmajor && !rowscale && M == N) Bb = Ab;
  else rc = stage(B, Bb, N, b_kmajor, sB, nullptr);
  if (rc) return rc;
  return gemm_tc_nt(Ab, Bb, C, nullptr, M, N, K, batch, (long long)M * K, (long long)N * K, 0, st);
}

static int contract(bool tc, const float* A, const float* B, float* C, const float* rowscale, int M, int N, int K, bool a_kmajor, bool b_kmajor,
                    long long sA, long long sB, long long sC, long long sS, int batch, cudaStream_t st) {
  return tc ? tc_contract(A, B, C, rowscale, M, N, K, a_kmajor, b_kmajor, sA, sB, sC, sS, batch, st)
            : sgemm_batched(A, B, C, rowscale, M, N, K, a_kmajor, b_kmajor, sA, sB, sC, sS, batch, st);
}

// ------------------------------------------------------------------ patch extraction
// P[n][l][c*9+ky*3+kx] = Fd[n][c][lh+ky-1][lw+kx-1], Fd = F[:, :, ::2, ::2]  (:282-295)
// R[n][l][c*16+ky*4+kx] = F[n][c][2lh-1+ky][2lw-1+kx]                        (:270-278)
// inv_norm[n][l] = 1 / max(sqrt(sum_k P^2), 1e-4)                             (:341-345)
__global__ void __launch_bounds__(256) ca_patches_kernel(const float* __restrict__ f, float* __restrict__ P,
                                                         float* __restrict__ R, float* __restrict__ inv_norm,
                                                         int c, int h, int w) {
  const int hs = h >> 1, ws = w >> 1, L = hs * ws;
  const int l = blockIdx.x, n = blockIdx.y;
  const int lh = l / ws, lw = l - lh * ws;
  const float* fn = f + (size_t)n * c * h * w;
  float* Pl = P + ((size_t)n * L + l) * (c * 9);
  float* Rl = R + ((size_t)n * L + l) * (c * 16);
  __shared__ float red[32];
  float ss = 0.f;
  for (int i = threadIdx.x; i < c * 9; i += blockDim.x) {
    int ch = i / 9, t = i - ch * 9, ky = t / 3, kx = t - ky * 3;
    int y = lh + ky - 1, x = lw + kx - 1;
    float v = (y >= 0 && y < hs && x >= 0 && x < ws) ? fn[((size_t)ch * h + 2 * y) * w + 2 * x] : 0.f;
    Pl[i] = v;
    ss = fmaf(v, v, ss);
  }
  for (int i = threadIdx.x; i < c * 16; i += blockDim.x) {
    int ch = i >> 4, ky = (i >> 2) & 3, kx = i & 3;
    int y = 2 * lh - 1 + ky, x = 2 * lw - 1 + kx;
    Rl[i] = (y >= 0 && y < h && x >= 0 && x < w) ? fn[((size_t)ch * h + y) * w + x] : 0.f;
  }
  float tot = block_sum(ss, red);
  if (threadIdx.x == 0) inv_norm[(size_t)n * L + l] = 1.f / fmaxf(sqrtf(tot), 1e-4f);
}

// mm[n][l] = 1 iff the zero-padded 3x3 neighbourhood of mask[src, 0, ::8, ::8] at l is all zero (:304-317)
__global__ void ca_mask_kernel(const float* __restrict__ mask, float* __restrict__ mm, int n, int hs, int ws,
                               int mh, int mw, int step, int per_sample) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * hs * ws) return;
  int s = i / (hs * ws), l = i - s * hs * ws, lh = l / ws, lw = l - lh * ws;
  const float* m = mask + (size_t)(per_sample ? s : 0) * mh * mw;
  float acc = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      int y = lh + dy, x = lw + dx;
      if (y >= 0 && y < hs && x >= 0 && x < ws) acc += m[(size_t)(y * step) * mw + x * step];
    }
  mm[i] = (acc / 9.f == 0.f) ? 1.f : 0.f;
}

// ------------------------------------------------------------------ fuse (:350-361)
// U[i][j] = sum_{cc in -1..1} T[cmi(cm(i)+cc)][cmi(cm(j)+cc)],  T[p][q] = sum_{a in -1..1} S[p+a][q+a]
// with flat indices outside [0,L) contributing zero; cm = row-major -> column-major index.
__global__ void __launch_bounds__(256) ca_fuse_kernel(const float* __restrict__ S, float* __restrict__ U, int side) {
  const int L = side * side;
  const float* Sn = S + (size_t)blockIdx.z * L * L;
  float* Un = U + (size_t)blockIdx.z * L * L;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= L) return;
  const int ci = (i % side) * side + i / side, cj = (j % side) * side + j / side;
  float acc = 0.f;
#pragma unroll
  for (int cc = -1; cc <= 1; ++cc) {
    const int pi = ci + cc, pj = cj + cc;
    if (pi < 0 || pi >= L || pj < 0 || pj >= L) continue;
    const int p = (pi % side) * side + pi / side, q = (pj % side) * side + pj / side;
#pragma unroll
    for (int a = -1; a <= 1; ++a) {
      const int r = p + a, s = q + a;
      if (r < 0 || r >= L || s < 0 || s >= L) continue;
      acc += Sn[(size_t)r * L + s];
    }
  }
  Un[(size_t)i * L + j] = acc;
}

// ------------------------------------------------------------------ masked scaled softmax + argmax (:364-368)
// one CTA per (32 foreground columns, sample); 8 row groups; in place on U -> A
__global__ void __launch_bounds__(256) ca_softmax_kernel(float* __restrict__ U, const float* __restrict__ mm,
                                                         int32_t* __restrict__ argmax_out, int L, float scale,
                                                         int mm_stride) {
  float* Un = U + (size_t)blockIdx.y * L * L;
  const float* m = mm + (size_t)blockIdx.y * mm_stride;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + lane;
  __shared__ float s_red[8][32];
  __shared__ int s_idx[8][32];
  float mx = -INFINITY;
  for (int b = g; b < L; b += 8) mx = fmaxf(mx, Un[(size_t)b * L + f] * m[b] * scale);
  s_red[g][lane] = mx;
  __syncthreads();
  mx = s_red[0][lane];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmaxf(mx, s_red[k][lane]);
  __syncthreads();
  float sum = 0.f;
  for (int b = g; b < L; b += 8) sum += expf(Un[(size_t)b * L + f] * m[b] * scale - mx);
  s_red[g][lane] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += s_red[k][lane];
  __syncthreads();
  float best = -1.f;
  int bi = 0;
  for (int b = g; b < L; b += 8) {
    const float mb = m[b];
    const float a = expf(Un[(size_t)b * L + f] * mb * scale - mx) / sum * mb;
    Un[(size_t)b * L + f] = a;
    if (a > best) { best = a; bi = b; }
  }
  s_red[g][lane] = best;
  s_idx[g][lane] = bi;
  __syncthreads();
  if (g == 0) {
    for (int k = 1; k < 8; ++k) {
      const float v = s_red[k][lane];
      const int id = s_idx[k][lane];
      if (v > best || (v == best && id < bi)) { best = v; bi = id; }
    }
    argmax_out[(size_t)blockIdx.y * L + f] = bi;
  }
}

// ------------------------------------------------------------------ overlap-add of the pasted patches (:377-379)
// y[n][c][oy][ox] = 0.25 * sum_{ky,kx,hf,wf : 2hf-1+ky=oy, 2wf-1+kx=ox} cols[n][c*16+ky*4+kx][hf*ws+wf]
__global__ void __launch_bounds__(256) ca_fold_kernel(const float* __restrict__ cols, float* __restrict__ y, int c,
                                                      int h, int w) {
  const int hs = h >> 1, ws = w >> 1, L = hs * ws;
  const size_t total = (size_t)c * h * w;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = blockIdx.y;
  const int ox = i % w, oy = (i / w) % h, ch = i / ((size_t)w * h);
  const float* cn = cols + (size_t)n * (c * 16) * L;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ky = ((oy + 1) & 1) + 2 * a, hf = (oy + 1 - ky) >> 1;
    if (oy + 1 - ky < 0 || hf >= hs) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kx = ((ox + 1) & 1) + 2 * b, wf = (ox + 1 - kx) >> 1;
      if (ox + 1 - kx < 0 || wf >= ws) continue;
      acc += cn[(size_t)(ch * 16 + ky * 4 + kx) * L + hf * ws + wf];
    }
  }
  y[(size_t)n * total + i] = acc * 0.25f;
}

// ------------------------------------------------------------------ offsets + flow colouring (:368-408)
__constant__ unsigned char c_wheel[55][3];
static bool g_wheel_ready = false;

static void make_wheel(unsigned char wheel[55][3]) {  // models/inpaint_tools.py:244-273
  const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
  int col = 0;
  for (int i = 0; i < 55; ++i) wheel[i][0] = wheel[i][1] = wheel[i][2] = 0;
  for (int i = 0; i < RY; ++i) { wheel[col + i][0] = 255; wheel[col + i][1] = (unsigned char)(255 * i / RY); }
  col += RY;
  for (int i = 0; i < YG; ++i) { wheel[col + i][0] = (unsigned char)(255 - 255 * i / YG); wheel[col + i][1] = 255; }
  col += YG;
  for (int i = 0; i < GC; ++i) { wheel[col + i][1] = 255; wheel[col + i][2] = (unsigned char)(255 * i / GC); }
  col += GC;
  for (int i = 0; i < CB; ++i) { wheel[col + i][1] = (unsigned char)(255 - 255 * i / CB); wheel[col + i][2] = 255; }
  col += CB;
  for (int i = 0; i < BM; ++i) { wheel[col + i][2] = 255; wheel[col + i][0] = (unsigned char)(255 * i / BM); }
  col += BM;
  for (int i = 0; i < MR; ++i) { wheel[col + i][2] = (unsigned char)(255 - 255 * i / MR); wheel[col + i][0] = 255; }
}

// offsets = argmax (row, col) - own (row, col) (:368-374, :389-397); smax[n] = max squared offset length of sample n
__global__ void __launch_bounds__(256) ca_offsets_kernel(const int32_t* __restrict__ argmax, int32_t* __restrict__ offsets,
                                                         int* __restrict__ smax, int side) {
  const int L = side * side, n = blockIdx.x;
  __shared__ int s_max[8];
  int mx = 0;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const int am = argmax[(size_t)n * L + l];
    const int du = am / side - l / side, dv = am % side - l % side;
    mx = max(mx, du * du + dv * dv);
    if (offsets) {
      offsets[((size_t)n * 2 + 0) * L + l] = du;
      offsets[((size_t)n * 2 + 1) * L + l] = dv;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = max(mx, s_max[i]);
    smax[n] = mx;
  }
}

// flow colouring (models/inpaint_tools.py:73-100,:178-208): the reference carries the maximum radius across the
// batch (maxrad of sample n = max over samples 0..n).  One thread = (cell, channel, pixel row of the up x up block).
__global__ void __launch_bounds__(256) ca_flow_kernel(const int32_t* __restrict__ argmax, const int* __restrict__ smax,
                                                      float* __restrict__ flow, int side, int up) {
  const int L = side * side, n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * 3 * up) return;
  const int yy = i % up, ch = (i / up) % 3, l = i / (3 * up);
  int m2 = 0;
  for (int s = 0; s <= n; ++s) m2 = max(m2, smax[s]);
  const double maxrad = sqrt((double)m2) + 2.220446049250313e-16;
  const int am = argmax[(size_t)n * L + l];
  const int du = am / side - l / side, dv = am % side - l % side;
  const double u = du / maxrad, v = dv / maxrad;
  const double rad = sqrt(u * u + v * v);
  const double a = atan2(-v, -u) / 3.141592653589793;
  const double fk = (a + 1.0) / 2.0 * 54.0 + 1.0;
  int k0 = (int)floor(fk), k1 = k0 + 1;
  if (k1 == 56) k1 = 1;
  const double fr = fk - k0;
  const double c0 = c_wheel[k0 - 1][ch] / 255.0, c1 = c_wheel[k1 - 1][ch] / 255.0;
  double col = (1.0 - fr) * c0 + fr * c1;
  if (rad <= 1.0) col = 1.0 - rad * (1.0 - col); else col *= 0.75;
  const float px = (float)(unsigned char)floor(255.0 * col) / 255.f;
  const int W = side * up;
  float* dst = flow + (((size_t)n * 3 + ch) * W + (size_t)(l / side) * up + yy) * W + (size_t)(l % side) * up;
  for (int xx = 0; xx < up; ++xx) dst[xx] = px;
}

static int ensure_wheel() {
  if (!g_wheel_ready) {
    unsigned char wheel[55][3];
    make_wheel(wheel);
    HV_CUDA(cudaMemcpyToSymbol(c_wheel, wheel, sizeof(wheel)));
    g_wheel_ready = true;
  }
  return HV_OK;
}

// scratch: n ints
int ca_offsets_flow_launch(const int32_t* argmax, int32_t* offsets, float* flow, int n, int side, int up, int* scratch,
                           cudaStream_t st) {
  int rc = ensure_wheel();
  if (rc) return rc;
  ca_offsets_kernel<<<n, 256, 0, st>>>(argmax, offsets, scratch, side);
  HV_LAUNCH_CHECK();
  if (flow) {
    ca_flow_kernel<<<dim3((side * side * 3 * up + 255) / 256, n), 256, 0, st>>>(argmax, scratch, flow, side, up);
    HV_LAUNCH_CHECK();
  }
  return HV_OK;
}

int ca_mask_launch(const float* mask, float* mm, int n, int side, int mh, int mw, int per_sample, cudaStream_t st) {
  ca_mask_kernel<<<(n * side * side + 255) / 256, 256, 0, st>>>(mask, mm, n, side, side, mh, mw, mh / side, per_sample);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

// ------------------------------------------------------------------ host orchestration
struct CaWorkspace {
  float *P, *R, *inv_norm, *mm, *S, *U, *cols;
  int32_t* argmax;
  int* scratch;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t ca_layout(int n, int c, int h, int w, char* base, CaWorkspace* ws) {
  const size_t L = (size_t)(h / 2) * (w / 2);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  char* p;
  p = take(sizeof(float) * n * L * c * 9);  if (ws) ws->P = (float*)p;
  p = take(sizeof(float) * n * L * c * 16); if (ws) ws->R = (float*)p;
  p = take(sizeof(float) * n * L);          if (ws) ws->inv_norm = (float*)p;
  p = take(sizeof(float) * n * L);          if (ws) ws->mm = (float*)p;
  p = take(sizeof(float) * n * L * L);      if (ws) ws->S = (float*)p;
  p = take(sizeof(float) * n * L * L);      if (ws) ws->U = (float*)p;
  p = take(sizeof(float) * n * L * c * 16); if (ws) ws->cols = (float*)p;
  p = take(sizeof(int32_t) * n * L);        if (ws) ws->argmax = (int32_t*)p;
  p = take(sizeof(int) * (n + 1));          if (ws) ws->scratch = (int*)p;
  return off;
}

size_t ctx_attn_workspace_bytes(int n, int c, int h, int w) { return ca_layout(n, c, h, w, nullptr, nullptr); }

int ctx_attn_fwd_fp32(const float* f, const float* mask, float* y, int32_t* offsets, float* flow, int n, int c,
                      int h, int w, float scale, int fuse, int per_sample_mask, void* workspace, cudaStream_t st, bool tc) {
  HV_CHECK_ARG(f && mask && y && workspace, "ctx_attn_fwd: null argument");
  HV_CHECK_ARG(n > 0 && n <= 65535 && c > 0 && h > 0 && w > 0, "ctx_attn_fwd: bad extent");
  HV_CHECK_ARG(h == w && (h % 2) == 0, "ctx_attn_fwd: square even feature maps only (h=%d w=%d)", h, w);
  const int side = h / 2, L = side * side;
  HV_CHECK_ARG(L % 128 == 0 && c % 8 == 0, "ctx_attn_fwd: needs (h/2)^2 %% 128 == 0 and c %% 8 == 0");
  CaWorkspace ws;
  ca_layout(n, c, h, w, (char*)workspace, &ws);
  ca_patches_kernel<<<dim3(L, n), 256, 0, st>>>(f, ws.P, ws.R, ws.inv_norm, c, h, w);
  HV_LAUNCH_CHECK();
  ca_mask_kernel<<<(n * L + 255) / 256, 256, 0, st>>>(mask, ws.mm, n, side, side, 4 * h, 4 * w, 8, per_sample_mask);
  HV_LAUNCH_CHECK();
  HV_CHECK_ARG(!tc || (c * 9) % 64 == 0, "ctx_attn_fwd (tensor-core mode): needs c * 9 %% 64 == 0");
  int rc = contract(tc, ws.P, ws.P, ws.S, ws.inv_norm, L, L, c * 9, true, true, (long long)L * c * 9,
                    (long long)L * c * 9, (long long)L * L, L, n, st);
  if (rc) return rc;
  float* A = ws.S;
  if (fuse) {
    ca_fuse_kernel<<<dim3((L + 255) / 256, L, n), 256, 0, st>>>(ws.S, ws.U, side);
    HV_LAUNCH_CHECK();
    A = ws.U;
  }
  ca_softmax_kernel<<<dim3(L / 32, n), 256, 0, st>>>(A, ws.mm, ws.argmax, L, scale, L);
  HV_LAUNCH_CHECK();
  // cols[ck][f] = sum_b R[b][ck] * A[b][f]
  rc = contract(tc, ws.R, A, ws.cols, nullptr, c * 16, L, L, false, false, (long long)L * c * 16,
                (long long)L * L, (long long)L * c * 16, 0, n, st);
  if (rc) return rc;
  const size_t per = (size_t)c * h * w;
  ca_fold_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>(ws.cols, y, c, h, w);
  HV_LAUNCH_CHECK();
  if (offsets || flow) {
    rc = ca_offsets_flow_launch(ws.argmax, offsets, flow, n, side, 8, ws.scratch, st);
    if (rc) return rc;
  }
  return HV_OK;
}

// =================================================================== backward (fp32)
// Adjoint of ctx_attn_fwd_fp32 w.r.t. the feature map f (it enters as foreground, background and raw patches alike).
// The FORWARD workspace must be passed unchanged: it still holds P, R, inv_norm, mm, S (raw scores) and A (in U).

// dcols[n][c*16+ky*4+kx][l] = 0.25 * dy[n][c][2lh-1+ky][2lw-1+kx]   (adjoint of the overlap-add, :377-379)
__global__ void __launch_bounds__(256) ca_unfold_dy_kernel(const float* __restrict__ dy, float* __restrict__ dcols, int c, int h, int w) {
  const int hs = h >> 1, ws = w >> 1, L = hs * ws;
  const size_t total = (size_t)c * 16 * L;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = blockIdx.y;
  const int l = i % L, t = (i / L) % 16, ch = i / ((size_t)L * 16);
  const int y = 2 * (l / ws) - 1 + (t >> 2), x = 2 * (l % ws) - 1 + (t & 3);
  float v = 0.f;
  if (y >= 0 && y < h && x >= 0 && x < w) v = 0.25f * dy[(((size_t)n * c + ch) * h + y) * w + x];
  dcols[(size_t)n * total + i] = v;
}

// softmax adjoint per foreground column: dU[b][f] = scale * mm[b] * A[b][f] * (dA[b][f] - sum_b' A[b'][f] dA[b'][f]); in place on dA
__global__ void __launch_bounds__(256) ca_softmax_bwd_kernel(const float* __restrict__ A, float* __restrict__ dA, const float* __restrict__ mm,
                                                             int L, float scale, int mm_stride) {
  const float* An = A + (size_t)blockIdx.y * L * L;
  float* dn = dA + (size_t)blockIdx.y * L * L;
  const float* m = mm + (size_t)blockIdx.y * mm_stride;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + lane;
  __shared__ float s_red[8][32];
  float dot = 0.f;
  for (int b = g; b < L; b += 8) dot = fmaf(An[(size_t)b * L + f], dn[(size_t)b * L + f], dot);
  s_red[g][lane] = dot;
  __syncthreads();
  dot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) dot += s_red[k][lane];
  for (int b = g; b < L; b += 8) {
    const size_t o = (size_t)b * L + f;
    dn[o] = scale * m[b] * An[o] * (dn[o] - dot);
  }
}

// adjoint of ca_fuse_kernel: dS[r][s] = sum_{a,cc} dU[cmi(cm(r+a)+cc)][cmi(cm(s+a)+cc)]  (the operator is symmetric up to the
// order of the two passes; a, cc in -1..1, flat indices outside [0, L) contribute nothing)
__global__ void __launch_bounds__(256) ca_fuse_bwd_kernel(const float* __restrict__ dU, float* __restrict__ dS, int side) {
  const int L = side * side;
  const float* Un = dU + (size_t)blockIdx.z * L * L;
  float* Sn = dS + (size_t)blockIdx.z * L * L;
  const int s = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (s >= L) return;
  float acc = 0.f;
#pragma unroll
  for (int a = -1; a <= 1; ++a) {
    const int p = r + a, q = s + a;
    if (p < 0 || p >= L || q < 0 || q >= L) continue;
    const int cp = (p % side) * side + p / side, cq = (q % side) * side + q / side;
#pragma unroll
    for (int cc = -1; cc <= 1; ++cc) {
      const int pi = cp + cc, pj = cq + cc;
      if (pi < 0 || pi >= L || pj < 0 || pj >= L) continue;
      const int i = (pi % side) * side + pi / side, j = (pj % side) * side + pj / side;
      acc += Un[(size_t)i * L + j];
    }
  }
  Sn[(size_t)r * L + s] = acc;
}

// per background row b: dinv[b] = sum_f dS[b][f] * S[b][f] / inv[b];  dG[b][:] = dS[b][:] * inv[b]  (in place)
__global__ void __launch_bounds__(256) ca_rowscale_bwd_kernel(float* __restrict__ dS, const float* __restrict__ S, const float* __restrict__ inv_norm,
                                                              float* __restrict__ dinv, int L) {
  __shared__ float red[32];
  const int b = blockIdx.x, n = blockIdx.y;
  float* row = dS + ((size_t)n * L + b) * L;
  const float* srow = S + ((size_t)n * L + b) * L;
  const float inv = inv_norm[(size_t)n * L + b];
  float acc = 0.f;
  for (int f = threadIdx.x; f < L; f += blockDim.x) { acc = fmaf(row[f], srow[f], acc); row[f] *= inv; }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) dinv[(size_t)n * L + b] = acc / inv;
}

// zero-padded copy of the patch matrix: Ppad[n][l][kpad]
__global__ void ca_pad_rows_kernel(const float* __restrict__ P, float* __restrict__ Ppad, int k, int kpad, size_t rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kpad) return;
  const int c = i % kpad;
  const size_t r = i / kpad;
  Ppad[i] = c < k ? P[r * k + c] : 0.f;
}

// df[n][c][Y][X] = sum of dR entries whose 4x4 stride-2 patch covers (Y, X)
//                + (Y, X both even) sum of dP entries whose 3x3 patch on the ::2 grid covers (Y/2, X/2), with
//                  dP[l] = C1[l] + C2[l] - inv[l]^3 dinv[l] P[l]   (the last term only where |P_l| > 1e-4)
__global__ void __launch_bounds__(256) ca_gather_df_kernel(const float* __restrict__ dR, const float* __restrict__ C1, const float* __restrict__ C2,
                                                           const float* __restrict__ P, const float* __restrict__ inv_norm,
                                                           const float* __restrict__ dinv, float* __restrict__ df, int c, int h, int w, int kpad) {
  const int hs = h >> 1, ws = w >> 1, L = hs * ws;
  const size_t total = (size_t)c * h * w;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = blockIdx.y;
  const int X = i % w, Y = (i / w) % h, ch = i / ((size_t)w * h);
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ky = ((Y + 1) & 1) + 2 * a, hf = (Y + 1 - ky) >> 1;
    if (Y + 1 - ky < 0 || hf >= hs) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kx = ((X + 1) & 1) + 2 * b, wf = (X + 1 - kx) >> 1;
      if (X + 1 - kx < 0 || wf >= ws) continue;
      acc += dR[((size_t)n * L + hf * ws + wf) * (c * 16) + ch * 16 + ky * 4 + kx];
    }
  }
  if (((Y | X) & 1) == 0) {
    const int y = Y >> 1, x = X >> 1;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int lh = y - ky + 1;
      if (lh < 0 || lh >= hs) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int lw = x - kx + 1;
        if (lw < 0 || lw >= ws) continue;
        const size_t l = (size_t)n * L + lh * ws + lw;
        const int col = ch * 9 + ky * 3 + kx;
        float v = C1[l * kpad + col] + C2[l * kpad + col];
        const float inv = inv_norm[l];
        if (inv < 1e4f) v -= inv * inv * inv * dinv[l] * P[l * (c * 9) + col];
        acc += v;
      }
    }
  }
  df[(size_t)n * total + i] = acc;
}

struct CaBwdWorkspace { float *X1, *X2, *dR, *Ppad, *C1, *C2, *dinv; };

static size_t ca_bwd_layout(int n, int c, int h, int w, char* base, CaBwdWorkspace* ws) {
  const size_t L = (size_t)(h / 2) * (w / 2);
  const size_t kpad = ((size_t)c * 9 + 127) / 128 * 128;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  char* p;
  p = take(sizeof(float) * n * L * L);        if (ws) ws->X1 = (float*)p;
  p = take(sizeof(float) * n * L * L);        if (ws) ws->X2 = (float*)p;
  p = take(sizeof(float) * n * L * c * 16);   if (ws) ws->dR = (float*)p;
  p = take(sizeof(float) * n * L * kpad);     if (ws) ws->Ppad = (float*)p;
  p = take(sizeof(float) * n * L * kpad);     if (ws) ws->C1 = (float*)p;
  p = take(sizeof(float) * n * L * kpad);     if (ws) ws->C2 = (float*)p;
  p = take(sizeof(float) * n * L);            if (ws) ws->dinv = (float*)p;
  return off;
}

size_t ctx_attn_bwd_workspace_bytes(int n, int c, int h, int w) { return ca_bwd_layout(n, c, h, w, nullptr, nullptr); }

int ctx_attn_bwd_fp32(const float* dy, float* df, int n, int c, int h, int w, float scale, int fuse, void* fwd_workspace,
                      void* bwd_workspace, cudaStream_t st, bool tc) {
  HV_CHECK_ARG(dy && df && fwd_workspace && bwd_workspace, "ctx_attn_bwd: null argument");
  HV_CHECK_ARG(h == w && (h % 2) == 0, "ctx_attn_bwd: square even feature maps only");
  const int side = h / 2, L = side * side, kp = c * 9, kr = c * 16;
  HV_CHECK_ARG(L % 128 == 0 && kr % 128 == 0 && c % 8 == 0, "ctx_attn_bwd: needs (h/2)^2 %% 128 == 0 and c %% 8 == 0");
  const int kpad = (kp + 127) / 128 * 128;
  CaWorkspace fw;
  ca_layout(n, c, h, w, (char*)fwd_workspace, &fw);
  CaBwdWorkspace bw;
  ca_bwd_layout(n, c, h, w, (char*)bwd_workspace, &bw);
  const float* A = fuse ? fw.U : fw.S;     // attention weights of the forward
  float* dcols = fw.cols;                   // the pasted columns are no longer needed
  const size_t per_cols = (size_t)kr * L;
  ca_unfold_dy_kernel<<<dim3((unsigned)((per_cols + 255) / 256), n), 256, 0, st>>>(dy, dcols, c, h, w);
  HV_LAUNCH_CHECK();
  // dA[b][f] = sum_ck R[b][ck] dcols[ck][f]
  int rc = contract(tc, fw.R, dcols, bw.X1, nullptr, L, L, kr, true, false, (long long)L * kr, (long long)kr * L, (long long)L * L, 0, n, st);
  if (rc) return rc;
  // dR[b][ck] = sum_f A[b][f] dcols[ck][f]
  rc = contract(tc, A, dcols, bw.dR, nullptr, L, kr, L, true, true, (long long)L * L, (long long)kr * L, (long long)L * kr, 0, n, st);
  if (rc) return rc;
  ca_softmax_bwd_kernel<<<dim3(L / 32, n), 256, 0, st>>>(A, bw.X1, fw.mm, L, scale, L);
  HV_LAUNCH_CHECK();
  float* dS = bw.X1;
  if (fuse) {
    ca_fuse_bwd_kernel<<<dim3((L + 255) / 256, L, n), 256, 0, st>>>(bw.X1, bw.X2, side);
    HV_LAUNCH_CHECK();
    dS = bw.X2;
    // the raw scores were overwritten by nothing: fw.S still holds inv * P P^T
  }
  HV_CHECK_ARG(fuse, "ctx_attn_bwd: the no-fuse variant overwrote its scores in the forward (not differentiable here)");
  ca_rowscale_bwd_kernel<<<dim3(L, n), 256, 0, st>>>(dS, fw.S, fw.inv_norm, bw.dinv, L);
  HV_LAUNCH_CHECK();
  const size_t rows = (size_t)n * L;
  ca_pad_rows_kernel<<<(unsigned)((rows * kpad + 255) / 256), 256, 0, st>>>(fw.P, bw.Ppad, kp, kpad, rows);
  HV_LAUNCH_CHECK();
  // C1[b][k] = sum_f dG[b][f] P[f][k];  C2[f][k] = sum_b dG[b][f] P[b][k]
  rc = contract(tc, dS, bw.Ppad, bw.C1, nullptr, L, kpad, L, true, false, (long long)L * L, (long long)L * kpad, (long long)L * kpad, 0, n, st);
  if (rc) return rc;
  rc = contract(tc, dS, bw.Ppad, bw.C2, nullptr, L, kpad, L, false, false, (long long)L * L, (long long)L * kpad, (long long)L * kpad, 0, n, st);
  if (rc) return rc;
  const size_t per = (size_t)c * h * w;
  ca_gather_df_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>(bw.dR, bw.C1, bw.C2, fw.P, fw.inv_norm, bw.dinv, df, c, h, w, kpad);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv
