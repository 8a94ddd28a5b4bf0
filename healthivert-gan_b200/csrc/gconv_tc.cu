// Data and weight gradients of the generator's convolutions on the tensor cores (bf16 operands, fp32 accumulate): the backward of
// Conv2dBlock (reference models/inpaint_networks.py:494-503 under torch.autograd, models/pix2pix_model.py:354) for ANY layer of the
// generator - k = 3 / 5, stride 1 / 2, dilation 1 .. 16, fused source gather (channel concat, nearest x2 / x0.5, scalar planes) - as
// batched "NT" GEMMs on the tcgen05 kernel of gemm_tc.cu over explicit im2col operands (same scheme as dconv_tc.cu):
//
//   wgrad   dWt[b,s][r][co] = sum_p colT[b,s][r][p] * dy[b,s][co][p]      r = ci*k*k + tap;  M = rows (padded to 128), N = Cout (padded to
//                                                                        32 / 64 / 128), K = a split of the output pixels;  dW = sum_{b,s}
//   dgrad   dcolT[b][r][p]  = sum_co Wt[r][co] * dyT[b][p][co]             M = rows, N = pixels, K = Cout (padded to 64);  dx = col2im(dcolT)
//
// dx is the gradient over the VIRTUAL concatenated input [n, cin, hin, win] of the forward descriptor, exactly like hv_conv2d_dgrad;
// the host routes channel ranges back to their source tensors.  fp32 SIMT kernels (train_ops.cu) remain the parity path.
#include <cuda_bf16.h>
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

int gemm_tc_nt(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, const float* colscale, int M, int N, int K, int batch,
               long long strideA, long long strideB, int out_bf16, cudaStream_t st);

struct GcGeom {
  hv_conv_src src[4];
  int nsrc;
  int n, cin, cout, hin, win, hout, wout, k, kk, stride, pad, dil;
  int R, Rp;        // im2col rows cin * k * k, padded to a multiple of 128
  int P, Pp;        // output pixels per image, padded to a multiple of 256
  int Np;           // Cout padded to 32 / 64 / a multiple of 128 (wgrad GEMM N)
  int Kc;           // Cout padded to a multiple of 64 (dgrad GEMM K)
  int nsplit, Pc;   // wgrad: the pixels of an image are split into nsplit chunks of Pc (multiple of 64) pixels
};

static int gc_geom(GcGeom& g, const hv_conv_desc* d) {
  HV_CHECK_ARG(d && d->nsrc >= 1 && d->nsrc <= 4 && d->cout >= 1 && d->cout <= 512, "gconv_tc: bad descriptor");
  int csum = 0;
  for (int i = 0; i < 4; ++i) {
    g.src[i] = d->src[i < d->nsrc ? i : 0];
    if (i < d->nsrc) csum += d->src[i].channels;
  }
  HV_CHECK_ARG(csum == d->cin, "gconv_tc: sources have %d channels, cin=%d", csum, d->cin);
  g.nsrc = d->nsrc; g.n = d->n; g.cin = d->cin; g.cout = d->cout; g.hin = d->hin; g.win = d->win;
  g.k = d->k; g.kk = d->k * d->k; g.stride = d->stride; g.pad = d->pad; g.dil = d->dil;
  const int eff = (d->k - 1) * d->dil + 1;
  g.hout = (d->hin + 2 * d->pad - eff) / d->stride + 1; g.wout = (d->win + 2 * d->pad - eff) / d->stride + 1;
  HV_CHECK_ARG(g.hout >= 1 && g.wout >= 1, "gconv_tc: empty output");
  g.R = d->cin * g.kk; g.Rp = (g.R + 127) & ~127;
  g.P = g.hout * g.wout; g.Pp = (g.P + 255) & ~255;
  g.Np = d->cout <= 32 ? 32 : (d->cout <= 64 ? 64 : ((d->cout + 127) & ~127));
  g.Kc = (d->cout + 63) & ~63;
  // wgrad split over the pixels: enough (image, chunk) batches to fill the GPU with Rp / 128 row tiles each; chunks of >= 1024 pixels
  g.Pc = g.Pp;
  g.nsplit = 1;
  while (g.Pc > 1024 && (g.Pc % 128) == 0 && (long long)g.n * g.nsplit * (g.Rp / 128) < 296) { g.Pc >>= 1; g.nsplit <<= 1; }
  return HV_OK;
}

// colT[b][s][r][pc]: row r = ci * kk + tap of image b, pixel s * Pc + pc.  One thread = (b, ci, 8 consecutive pixels): the source of
// the channel is resolved once, every tap gathers 8 values and writes ONE 16-byte store (the kernel is pure data movement: 9 - 25x
// the bytes of the input; 2-byte stores per thread ran at a tenth of the HBM rate).
struct GcSrc { const float* base; int mode, sh, sw; };   // plane of (image, channel) in its own tensor
__device__ __forceinline__ GcSrc gc_resolve(const GcGeom& g, int n, int ch) {
  int s = 0;
  while (s < g.nsrc - 1 && ch >= g.src[s].channels) { ch -= g.src[s].channels; ++s; }
  GcSrc r;
  r.mode = g.src[s].mode;
  r.sh = g.hin; r.sw = g.win;
  if (r.mode == HV_SRC_UP2) { r.sh >>= 1; r.sw >>= 1; }
  else if (r.mode == HV_SRC_SUB2) { r.sh <<= 1; r.sw <<= 1; }
  r.base = r.mode == HV_SRC_SCALAR ? g.src[s].ptr + n : g.src[s].ptr + ((size_t)n * g.src[s].channels + ch) * r.sh * r.sw;
  return r;
}
__device__ __forceinline__ float gc_fetch(const GcGeom& g, const GcSrc& r, int gy, int gx) {
  if (gy < 0 || gy >= g.hin || gx < 0 || gx >= g.win) return 0.f;
  if (r.mode == HV_SRC_SCALAR) return __ldg(r.base);
  if (r.mode == HV_SRC_UP2) { gy >>= 1; gx >>= 1; }
  else if (r.mode == HV_SRC_SUB2) { gy <<= 1; gx <<= 1; }
  return __ldg(r.base + (size_t)gy * r.sw + gx);
}

__global__ void __launch_bounds__(256) gc_im2col_kernel(GcGeom g, __nv_bfloat16* __restrict__ colT) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int P8 = g.Pp >> 3;
  if (i >= (long long)g.n * g.cin * P8) return;
  const int p0 = (int)(i % P8) * 8, ci = (int)((i / P8) % g.cin), b = (int)(i / ((long long)P8 * g.cin));
  const GcSrc src = gc_resolve(g, b, ci);
  const int s = p0 / g.Pc, pc = p0 - s * g.Pc;            // Pc is a multiple of 64: the 8 pixels stay in one chunk
  __nv_bfloat16* dst = colT + (((size_t)b * g.nsplit + s) * g.Rp + (size_t)ci * g.kk) * g.Pc + pc;
  int oy[8], ox[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int p = p0 + j; oy[j] = p / g.wout; ox[j] = p - oy[j] * g.wout; }
  for (int ky = 0; ky < g.k; ++ky)
    for (int kx = 0; kx < g.k; ++kx) {
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = __float2bfloat16(p0 + j < g.P ? gc_fetch(g, src, oy[j] * g.stride + ky * g.dil - g.pad, ox[j] * g.stride + kx * g.dil - g.pad) : 0.f);
      *reinterpret_cast<uint4*>(dst + (size_t)(ky * g.k + kx) * g.Pc) = *reinterpret_cast<const uint4*>(v);
    }
}

// dy fp32 [b][cout][P] -> dyb bf16 [b][s][Np][Pc] (wgrad B operand).  Only the rows of real filters are written: rows >= cout feed
// output columns >= cout of the GEMM, which nobody reads (a padding ROW or COLUMN of an operand may hold anything; only padding along
// K - pixels >= P here - must be zero, and is).  One thread = 8 consecutive pixels.
__global__ void __launch_bounds__(256) gc_cast_dy_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dyb, GcGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int P8 = g.Pp >> 3;
  if (i >= (long long)g.n * g.cout * P8) return;
  const int p0 = (int)(i % P8) * 8, co = (int)((i / P8) % g.cout), b = (int)(i / ((long long)P8 * g.cout));
  const int s = p0 / g.Pc, pc = p0 - s * g.Pc;
  const float* src = dy + ((size_t)b * g.cout + co) * g.P;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(p0 + j < g.P ? src[p0 + j] : 0.f);
  *reinterpret_cast<uint4*>(dyb + (((size_t)b * g.nsplit + s) * g.Np + co) * g.Pc + pc) = *reinterpret_cast<const uint4*>(v);
}

// dw[co][r] = sum over the (b, s) blocks of part[blk][r][co]   (part: fp32 [blocks][Rp][Np])
__global__ void __launch_bounds__(256) gc_reduce_dw_kernel(const float* __restrict__ part, float* __restrict__ dw, int blocks, int R, int Rp, int cout, int Np) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * cout) return;
  const int co = i % cout, r = i / cout;     // co fastest: coalesced reads of part
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += part[((size_t)b * Rp + r) * Np + co];
  dw[(size_t)co * R + r] = s;
}

// Wt bf16 [Rp][Kc]: Wt[ci * kk + t][co] = w[co][ci][t], zero padding
__global__ void __launch_bounds__(256) gc_wt_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, int cout, int R, int Rp, int Kc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rp * Kc) return;
  const int co = i % Kc, r = i / Kc;
  wt[i] = __float2bfloat16((co < cout && r < R) ? w[(size_t)co * R + r] : 0.f);
}

// dy fp32 [b][cout][P] -> dyT bf16 [b][Pp][Kc] through 32 x 32 shared-memory tiles; grid (Pp / 32, Kc / 32, b)
__global__ void __launch_bounds__(256) gc_dyT_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dyT, int cout, int Kc, int P, int Pp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + tx, co = c0 + r;
    tile[r][tx] = (p < P && co < cout) ? dy[((size_t)b * cout + co) * P + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) dyT[((size_t)b * Pp + p0 + r) * Kc + c0 + tx] = __float2bfloat16(tile[tx][r]);
}

// dx[b][ci][y][x] = sum over the taps whose output pixel exists: dcolT[b][ci * kk + t][oy * wout + ox], oy = (y + pad - ky * dil) / stride.
// One thread = 4 consecutive x of one input row (one 16-byte store; the row / tap arithmetic is shared by the four).
__global__ void __launch_bounds__(256) gc_col2im_kernel(const __nv_bfloat16* __restrict__ dcolT, float* __restrict__ dx, GcGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W4 = (g.win + 3) >> 2;
  if (i >= (long long)g.n * g.cin * g.hin * W4) return;
  const int x0 = (int)(i % W4) * 4, y = (int)((i / W4) % g.hin);
  const long long bc = i / ((long long)W4 * g.hin);
  const int ci = (int)(bc % g.cin), b = (int)(bc / g.cin);
  const __nv_bfloat16* src = dcolT + ((size_t)b * g.Rp + (size_t)ci * g.kk) * g.Pp;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < g.k; ++ky) {
    const int ty = y + g.pad - ky * g.dil;
    if (ty < 0 || (g.stride == 2 && (ty & 1))) continue;
    const int oy = g.stride == 2 ? ty >> 1 : ty;
    if (oy >= g.hout) continue;
    const __nv_bfloat16* row = src + (size_t)oy * g.wout;
    for (int kx = 0; kx < g.k; ++kx) {
      const __nv_bfloat16* tap = row + (size_t)(ky * g.k + kx) * g.Pp;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tx = x0 + j + g.pad - kx * g.dil;
        if (tx < 0 || (g.stride == 2 && (tx & 1))) continue;
        const int ox = g.stride == 2 ? tx >> 1 : tx;
        if (ox < g.wout) acc[j] += __bfloat162float(tap[ox]);
      }
    }
  }
  float* out = dx + ((size_t)bc * g.hin + y) * g.win + x0;
  if ((g.win & 3) == 0) *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else
    for (int j = 0; j < 4 && x0 + j < g.win; ++j) out[j] = acc[j];
}

static inline unsigned gc_blocks(long long n) { return (unsigned)((n + 255) / 256); }
static inline size_t gc_al(size_t b) { return (b + 255) & ~(size_t)255; }

size_t gconv_wgrad_workspace_bytes(const hv_conv_desc* d) {
  GcGeom g;
  if (gc_geom(g, d)) return 0;
  const size_t blocks = (size_t)g.n * g.nsplit;
  return gc_al(blocks * g.Rp * g.Pc * 2) + gc_al(blocks * g.Np * g.Pc * 2) + gc_al(blocks * g.Rp * g.Np * 4);
}
size_t gconv_dgrad_workspace_bytes(const hv_conv_desc* d) {
  GcGeom g;
  if (gc_geom(g, d)) return 0;
  return gc_al((size_t)g.Rp * g.Kc * 2) + gc_al((size_t)g.n * g.Pp * g.Kc * 2) + gc_al((size_t)g.n * g.Rp * g.Pp * 2);
}

int conv2d_wgrad_bf16(const hv_conv_desc* d, const float* dy, float* dw, float* db, void* workspace, cudaStream_t st) {
  HV_CHECK_ARG(d && dy && dw && workspace, "conv2d_wgrad_bf16: null argument");
  GcGeom g;
  int rc = gc_geom(g, d);
  if (rc) return rc;
  const int blocks = g.n * g.nsplit;
  char* base = (char*)workspace;
  __nv_bfloat16* colT = (__nv_bfloat16*)base;
  __nv_bfloat16* dyb = (__nv_bfloat16*)(base + gc_al((size_t)blocks * g.Rp * g.Pc * 2));
  float* part = (float*)((char*)dyb + gc_al((size_t)blocks * g.Np * g.Pc * 2));
  // rows R .. Rp of colT (padding of the GEMM's M) stay unwritten: they only produce output rows that gc_reduce_dw_kernel never reads
  gc_im2col_kernel<<<gc_blocks((long long)g.n * g.cin * (g.Pp >> 3)), 256, 0, st>>>(g, colT);
  HV_LAUNCH_CHECK();
  gc_cast_dy_kernel<<<gc_blocks((long long)g.n * g.cout * (g.Pp >> 3)), 256, 0, st>>>(dy, dyb, g);
  HV_LAUNCH_CHECK();
  rc = gemm_tc_nt(colT, dyb, part, nullptr, g.Rp, g.Np, g.Pc, blocks, (long long)g.Rp * g.Pc, (long long)g.Np * g.Pc, 0, st);
  if (rc) return rc;
  gc_reduce_dw_kernel<<<gc_blocks((long long)g.R * g.cout), 256, 0, st>>>(part, dw, blocks, g.R, g.Rp, g.cout, g.Np);
  HV_LAUNCH_CHECK();
  if (db) return channel_sum(dy, db, g.n, g.cout, g.P, st);   // bias gradient: fp32, deterministic (train_ops.cu)
  return HV_OK;
}

int conv2d_dgrad_bf16(const hv_conv_desc* d, const float* w, const float* dy, float* dx, void* workspace, cudaStream_t st) {
  HV_CHECK_ARG(d && w && dy && dx && workspace, "conv2d_dgrad_bf16: null argument");
  GcGeom g;
  int rc = gc_geom(g, d);
  if (rc) return rc;
  char* base = (char*)workspace;
  __nv_bfloat16* wt = (__nv_bfloat16*)base;
  __nv_bfloat16* dyT = (__nv_bfloat16*)(base + gc_al((size_t)g.Rp * g.Kc * 2));
  __nv_bfloat16* dcolT = (__nv_bfloat16*)((char*)dyT + gc_al((size_t)g.n * g.Pp * g.Kc * 2));
  gc_wt_kernel<<<gc_blocks((long long)g.Rp * g.Kc), 256, 0, st>>>(w, wt, g.cout, g.R, g.Rp, g.Kc);
  HV_LAUNCH_CHECK();
  gc_dyT_kernel<<<dim3(g.Pp / 32, g.Kc / 32, g.n), 256, 0, st>>>(dy, dyT, g.cout, g.Kc, g.P, g.Pp);
  HV_LAUNCH_CHECK();
  rc = gemm_tc_nt(wt, dyT, dcolT, nullptr, g.Rp, g.Pp, g.Kc, g.n, 0, (long long)g.Pp * g.Kc, 1, st);
  if (rc) return rc;
  gc_col2im_kernel<<<gc_blocks((long long)g.n * g.cin * g.hin * ((g.win + 3) >> 2)), 256, 0, st>>>(dcolT, dx, g);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv

extern "C" {
size_t hv_conv2d_wgrad_bf16_workspace_bytes(const hv_conv_desc* d) { return hv::gconv_wgrad_workspace_bytes(d); }
size_t hv_conv2d_dgrad_bf16_workspace_bytes(const hv_conv_desc* d) { return hv::gconv_dgrad_workspace_bytes(d); }
int hv_conv2d_wgrad_bf16(const hv_conv_desc* d, const float* dy, float* dw, float* db, void* workspace, hv_stream_t s) {
  return hv::conv2d_wgrad_bf16(d, dy, dw, db, workspace, hv::as_stream(s));
}
int hv_conv2d_dgrad_bf16(const hv_conv_desc* d, const float* w, const float* dy, float* dx, void* workspace, hv_stream_t s) {
  return hv::conv2d_dgrad_bf16(d, w, dy, dx, workspace, hv::as_stream(s));
}
}
