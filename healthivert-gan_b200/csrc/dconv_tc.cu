// PatchGAN discriminator convolutions on the tensor cores (bf16 operands, fp32 accumulate): forward, data gradient and weight gradient
// of the 4x4 / padding 1 / stride 1|2 convolutions 64->128, 128->256, 256->512 (reference models/networks.py:575-598; 73 % of the
// training step's FLOPs) as batched "NT" GEMMs on the tcgen05 kernel of gemm_tc.cu over explicit im2col operands:
//
//   forward   y[b][co][p]     = sum_k  W[co][k]      * col[b][p][k]          M = Cout, N = pixels, K = Cin*16    (A broadcast over b)
//   wgrad     dW[b][co][k]    = sum_p  dy[b][co][p]  * colT[b][k][p]         M = Cout, N = Cin*16, K = pixels;   dW = sum_b dW[b]
//   dgrad     dcolT[b][k][p]  = sum_co Wt[k][co]     * dyT[b][p][co]         M = Cin*16, N = pixels, K = Cout;   dx = col2im(dcolT)
//
// Every GEMM output lands in the layout its consumer wants: y is NCHW (fp32), dW[b] is [Cout][Cin][4][4], dcolT is tap-major so that the
// col2im gather reads contiguous pixels.  Pixel counts are padded to a multiple of 256 (31*31 = 961 -> 1024) with zero operand rows.
// The fp32 SIMT kernels (conv_fp32.cu, train_ops.cu) remain the parity path; this is the bf16 training mode (Pix2PixModel.d_precision).
#include <cuda_bf16.h>
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

int gemm_tc_nt(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, const float* colscale, int M, int N, int K, int batch,
               long long strideA, long long strideB, int out_bf16, cudaStream_t st);

struct DcGeom {
  int n, cin, cout, h, w, stride, ho, wo, P, Pp, K;
};

static int dc_geom(DcGeom& g, int n, int cin, int cout, int h, int w, int stride) {
  HV_CHECK_ARG(n >= 1 && (stride == 1 || stride == 2) && cin >= 4 && (cin * 16) % 128 == 0 && cout % 128 == 0 && h >= 4 && w >= 4,
               "dconv_tc: needs k=4, pad=1, stride 1|2, Cin %% 8 == 0, Cout %% 128 == 0 (got cin %d cout %d stride %d)", cin, cout, stride);
  g.n = n; g.cin = cin; g.cout = cout; g.h = h; g.w = w; g.stride = stride;
  g.ho = (h + 2 - 4) / stride + 1; g.wo = (w + 2 - 4) / stride + 1;
  g.P = g.ho * g.wo; g.Pp = (g.P + 255) & ~255; g.K = cin * 16;
  return HV_OK;
}

// ---------------------------------------------------------------------------------------------------- operand builders
// col[b][p][ci*16 + ky*4 + kx] = x[b][ci][s*oy - 1 + ky][s*ox - 1 + kx]   (pixel-major: K contiguous; rows p >= P are zero)
// one thread = (b, ci, p): 16 loads, one 32-byte store (a full sector); p runs fastest so the loads of a warp are neighbours
__global__ void __launch_bounds__(256) dc_im2col_pm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, DcGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.n * g.cin * g.Pp;
  if (i >= total) return;
  const int p = (int)(i % g.Pp), ci = (int)((i / g.Pp) % g.cin), b = (int)(i / ((long long)g.Pp * g.cin));
  __align__(16) __nv_bfloat16 v[16];
  if (p < g.P) {
    const int oy = p / g.wo, ox = p - oy * g.wo;
    const float* src = x + ((size_t)b * g.cin + ci) * g.h * g.w;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int y = g.stride * oy - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int xx = g.stride * ox - 1 + kx;
        v[ky * 4 + kx] = __float2bfloat16((y >= 0 && y < g.h && xx >= 0 && xx < g.w) ? __ldg(src + (size_t)y * g.w + xx) : 0.f);
      }
    }
  } else {
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = __float2bfloat16(0.f);
  }
  uint4* dst = reinterpret_cast<uint4*>(col + ((size_t)b * g.Pp + p) * g.K + (size_t)ci * 16);
  dst[0] = reinterpret_cast<const uint4*>(v)[0];
  dst[1] = reinterpret_cast<const uint4*>(v)[1];
}

// colT[b][ci*16 + t][p]  (tap-major: pixels contiguous; columns p >= P are zero): one thread = (b, ci, p), 16 coalesced 2-byte stores
__global__ void __launch_bounds__(256) dc_im2col_km_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ colT, DcGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.n * g.cin * g.Pp;
  if (i >= total) return;
  const int p = (int)(i % g.Pp), ci = (int)((i / g.Pp) % g.cin), b = (int)(i / ((long long)g.Pp * g.cin));
  const int oy = p / g.wo, ox = p - oy * g.wo;
  const float* src = x + ((size_t)b * g.cin + ci) * g.h * g.w;
  __nv_bfloat16* dst = colT + ((size_t)b * g.K + (size_t)ci * 16) * g.Pp + p;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int y = g.stride * oy - 1 + ky;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int xx = g.stride * ox - 1 + kx;
      const bool ok = p < g.P && y >= 0 && y < g.h && xx >= 0 && xx < g.w;
      dst[(size_t)(ky * 4 + kx) * g.Pp] = __float2bfloat16(ok ? __ldg(src + (size_t)y * g.w + xx) : 0.f);
    }
  }
}

// weights: W fp32 [cout][K] -> Wb bf16 [cout][K] (forward A operand) and Wt bf16 [K][cout] (dgrad A operand)
__global__ void __launch_bounds__(256) dc_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb, __nv_bfloat16* __restrict__ wt,
                                                         int cout, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)cout * K) return;
  const int k = (int)(i % K), co = (int)(i / K);
  const __nv_bfloat16 v = __float2bfloat16(w[i]);
  if (wb) wb[i] = v;
  if (wt) wt[(size_t)k * cout + co] = v;
}

// dy fp32 [b][cout][P] -> dyb bf16 [b][cout][Pp] (wgrad A operand) and dyT bf16 [b][Pp][cout] (dgrad B operand); 32 x 32 tiles through
// shared memory so that both stores are row-contiguous; grid (Pp / 32, cout / 32, b)
__global__ void __launch_bounds__(256) dc_cast_dy_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dyb, __nv_bfloat16* __restrict__ dyT,
                                                         int cout, int P, int Pp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + tx;
    const float v = p < P ? dy[((size_t)b * cout + c0 + r) * P + p] : 0.f;
    tile[r][tx] = v;
    if (dyb) dyb[((size_t)b * cout + c0 + r) * Pp + p] = __float2bfloat16(v);
  }
  __syncthreads();
  if (dyT) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) dyT[((size_t)b * Pp + p0 + r) * cout + c0 + tx] = __float2bfloat16(tile[tx][r]);
  }
}

// y fp32 [b][cout][P] <- ypad [b][cout][Pp]
__global__ void __launch_bounds__(256) dc_compact_kernel(const float* __restrict__ ypad, float* __restrict__ y, long long rows, int P, int Pp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * P) return;
  const long long r = i / P;
  y[i] = ypad[r * Pp + (i - r * P)];
}

// dW[i] = sum_b part[b][i]
__global__ void __launch_bounds__(256) dc_reduce_b_kernel(const float* __restrict__ part, float* __restrict__ dw, long long count, int n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int b = 0; b < n; ++b) s += part[(size_t)b * count + i];
  dw[i] = s;
}

// dx[b][ci][y][x] = sum over the taps (ky, kx) whose output pixel exists: dcolT[b][ci*16 + ky*4 + kx][oy*wo + ox], oy = (y + 1 - ky) / s.
// One thread = 4 consecutive x (one 16-byte store).
__global__ void __launch_bounds__(256) dc_col2im_kernel(const __nv_bfloat16* __restrict__ dcolT, float* __restrict__ dx, DcGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W4 = (g.w + 3) >> 2;
  if (i >= (long long)g.n * g.cin * g.h * W4) return;
  const int x0 = (int)(i % W4) * 4, y = (int)((i / W4) % g.h);
  const long long bc = i / ((long long)W4 * g.h);     // b * cin + ci
  const int ci = (int)(bc % g.cin), b = (int)(bc / g.cin);
  const __nv_bfloat16* src = dcolT + ((size_t)b * g.K + (size_t)ci * 16) * g.Pp;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int ty = y + 1 - ky;
    if (ty < 0 || (g.stride == 2 && (ty & 1))) continue;
    const int oy = g.stride == 2 ? ty >> 1 : ty;
    if (oy >= g.ho) continue;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const __nv_bfloat16* tap = src + (size_t)(ky * 4 + kx) * g.Pp + (size_t)oy * g.wo;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tx = x0 + j + 1 - kx;
        if (tx < 0 || (g.stride == 2 && (tx & 1))) continue;
        const int ox = g.stride == 2 ? tx >> 1 : tx;
        if (ox < g.wo) acc[j] += __bfloat162float(tap[ox]);
      }
    }
  }
  float* out = dx + ((size_t)bc * g.h + y) * g.w + x0;
  if ((g.w & 3) == 0) *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else
    for (int j = 0; j < 4 && x0 + j < g.w; ++j) out[j] = acc[j];
}

static inline unsigned dc_blocks(long long n) { return (unsigned)((n + 255) / 256); }
static inline size_t dc_al(size_t b) { return (b + 255) & ~(size_t)255; }

// workspace layout (one buffer serves forward and backward of one layer call; sized by dconv_workspace_bytes)
struct DcWs {
  __nv_bfloat16 *col, *wb, *wt, *dyb, *dyT, *dcolT;
  float *ypad, *dwpart;
};
static size_t dc_layout(const DcGeom& g, char* base, DcWs* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += dc_al(bytes); return p; };
  const size_t colb = (size_t)g.n * g.Pp * g.K * 2;
  char* p;
  p = take(colb);                                   if (ws) ws->col = (__nv_bfloat16*)p;       // col (forward) / colT (wgrad)
  p = take((size_t)g.cout * g.K * 2);               if (ws) ws->wb = (__nv_bfloat16*)p;
  p = take((size_t)g.cout * g.K * 2);               if (ws) ws->wt = (__nv_bfloat16*)p;
  p = take((size_t)g.n * g.cout * g.Pp * 2);        if (ws) ws->dyb = (__nv_bfloat16*)p;
  p = take((size_t)g.n * g.cout * g.Pp * 2);        if (ws) ws->dyT = (__nv_bfloat16*)p;
  p = take(colb);                                   if (ws) ws->dcolT = (__nv_bfloat16*)p;
  const size_t ypad = (size_t)g.n * g.cout * g.Pp * 4, dwp = (size_t)g.n * g.cout * g.K * 4;
  p = take(ypad > dwp ? ypad : dwp);                if (ws) { ws->ypad = (float*)p; ws->dwpart = (float*)p; }
  return off;
}

size_t dconv_workspace_bytes(int n, int cin, int cout, int h, int w, int stride) {
  DcGeom g;
  if (dc_geom(g, n, cin, cout, h, w, stride)) return 0;
  return dc_layout(g, nullptr, nullptr);
}

int dconv_fwd_bf16(const float* x, const float* w, float* y, int n, int cin, int cout, int h, int wd, int stride, void* workspace, cudaStream_t st) {
  HV_CHECK_ARG(x && w && y && workspace, "dconv_fwd_bf16: null argument");
  DcGeom g;
  int rc = dc_geom(g, n, cin, cout, h, wd, stride);
  if (rc) return rc;
  DcWs ws;
  dc_layout(g, (char*)workspace, &ws);
  dc_weights_kernel<<<dc_blocks((long long)cout * g.K), 256, 0, st>>>(w, ws.wb, nullptr, cout, g.K);
  HV_LAUNCH_CHECK();
  dc_im2col_pm_kernel<<<dc_blocks((long long)n * cin * g.Pp), 256, 0, st>>>(x, ws.col, g);
  HV_LAUNCH_CHECK();
  float* out = g.Pp == g.P ? y : ws.ypad;
  rc = gemm_tc_nt(ws.wb, ws.col, out, nullptr, cout, g.Pp, g.K, n, 0, (long long)g.Pp * g.K, 0, st);
  if (rc) return rc;
  if (g.Pp != g.P) {
    dc_compact_kernel<<<dc_blocks((long long)n * cout * g.P), 256, 0, st>>>(ws.ypad, y, (long long)n * cout, g.P, g.Pp);
    HV_LAUNCH_CHECK();
  }
  return HV_OK;
}

// dy: gradient w.r.t. the conv output [n][cout][ho][wo]; dw [cout][cin][4][4] and / or dx [n][cin][h][w] (either may be NULL)
int dconv_bwd_bf16(const float* x, const float* w, const float* dy, float* dx, float* dw, int n, int cin, int cout, int h, int wd, int stride,
                   void* workspace, cudaStream_t st) {
  HV_CHECK_ARG(w && dy && workspace && (dx || dw) && (!dw || x), "dconv_bwd_bf16: null argument");
  DcGeom g;
  int rc = dc_geom(g, n, cin, cout, h, wd, stride);
  if (rc) return rc;
  DcWs ws;
  dc_layout(g, (char*)workspace, &ws);
  dc_cast_dy_kernel<<<dim3(g.Pp / 32, cout / 32, n), 256, 0, st>>>(dy, dw ? ws.dyb : nullptr, dx ? ws.dyT : nullptr, cout, g.P, g.Pp);
  HV_LAUNCH_CHECK();
  if (dw) {
    dc_im2col_km_kernel<<<dc_blocks((long long)n * cin * g.Pp), 256, 0, st>>>(x, ws.col, g);
    HV_LAUNCH_CHECK();
    rc = gemm_tc_nt(ws.dyb, ws.col, ws.dwpart, nullptr, cout, g.K, g.Pp, n, (long long)cout * g.Pp, (long long)g.K * g.Pp, 0, st);
    if (rc) return rc;
    dc_reduce_b_kernel<<<dc_blocks((long long)cout * g.K), 256, 0, st>>>(ws.dwpart, dw, (long long)cout * g.K, n);
    HV_LAUNCH_CHECK();
  }
  if (dx) {
    dc_weights_kernel<<<dc_blocks((long long)cout * g.K), 256, 0, st>>>(w, nullptr, ws.wt, cout, g.K);
    HV_LAUNCH_CHECK();
    rc = gemm_tc_nt(ws.wt, ws.dyT, ws.dcolT, nullptr, g.K, g.Pp, cout, n, 0, (long long)g.Pp * cout, 1, st);
    if (rc) return rc;
    dc_col2im_kernel<<<dc_blocks((long long)n * cin * h * ((wd + 3) >> 2)), 256, 0, st>>>(ws.dcolT, dx, g);
    HV_LAUNCH_CHECK();
  }
  return HV_OK;
}

}  // namespace hv

extern "C" {
size_t hv_dconv_workspace_bytes(int n, int cin, int cout, int h, int w, int stride) { return hv::dconv_workspace_bytes(n, cin, cout, h, w, stride); }
int hv_dconv_fwd_bf16(const float* x, const float* w, float* y, int n, int cin, int cout, int h, int wd, int stride, void* workspace, hv_stream_t s) {
  return hv::dconv_fwd_bf16(x, w, y, n, cin, cout, h, wd, stride, workspace, hv::as_stream(s));
}
int hv_dconv_bwd_bf16(const float* x, const float* w, const float* dy, float* dx, float* dw, int n, int cin, int cout, int h, int wd, int stride,
                      void* workspace, hv_stream_t s) {
  return hv::dconv_bwd_bf16(x, w, dy, dx, dw, n, cin, cout, h, wd, stride, workspace, hv::as_stream(s));
}
}
