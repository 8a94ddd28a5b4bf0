// Internal (non-ABI) declarations shared between the hv_b200 translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/hv_b200.h"

namespace hv {

struct SnJob {
  const float* w;   // [cout, kdim] weight_orig
  float* u;         // [cout]
  float* v;         // [kdim]
  int cout, kdim;
  float* w_eff;     // [cout, kdim] = w / sigma (may be null)
  float* sigma;     // scalar (may be null)
};

int sn_prepare_batched(const SnJob* d_jobs, int njobs, int training, cudaStream_t st);
int sn_prepare_single(const SnJob& job, int training, cudaStream_t st);

int conv2d_fwd_fp32(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2,
                    cudaStream_t st);
int conv2d_fwd_fp32_ex(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2, int hout, int wout,
                       cudaStream_t st);
// sub-pixel output of conv2d_fwd_fp32_sub: row / column padding of their own, outputs at (oy * os + ooy, ox * os + oox) of a yH x yW plane
struct ConvSubpixel { int pad_y, pad_x, os, ooy, oox, yH, yW; };
int conv2d_fwd_fp32_sub(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2, int hout, int wout,
                        const ConvSubpixel* sub, cudaStream_t st);
// internal source mode: [N,ch,H/2,W/2] read through zero insertion (value at even (y,x) only) - stride-2 data gradient
#define HV_SRC_ZEROINS2 4

// out[ch] = sum over (n, hw) of x[n][ch][:] (deterministic two-pass sum; bias gradients)
int channel_sum(const float* x, float* out, int n, int c, int hw, cudaStream_t st);
// stand-alone tcgen05 convolution on fp32 NCHW tensors (conv_tc_api.cu); buffers come from the stream's scratch
int conv2d_bf16(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2, int flags, cudaStream_t st);

int gap_fc_sigmoid(const float* x, const float* fc_w, const float* fc_b, float* out, int n, int c,
                   int hw, cudaStream_t st);

size_t ctx_attn_workspace_bytes(int n, int c, int h, int w);
int ctx_attn_fwd_fp32(const float* f, const float* mask, float* y, int32_t* offsets, float* flow,
                      int n, int c, int h, int w, float scale, int fuse, int per_sample_mask,
                      void* workspace, cudaStream_t st, bool tc = false);   // tc: the contractions on tcgen05 (bf16 operands)

// C[m][n] = rowscale[m] * sum_k A(m,k) * B(n,k); batched over blockIdx.z.
// a_kmajor: A stored [M][K] (lda = K) else [K][M];  b_kmajor: B stored [N][K] else [K][N].
int sgemm_batched(const float* A, const float* B, float* C, const float* rowscale, int M, int N, int K,
                  bool a_kmajor, bool b_kmajor, long long strideA, long long strideB, long long strideC,
                  long long strideScale, int batch, cudaStream_t st);

size_t ctx_attn_bwd_workspace_bytes(int n, int c, int h, int w);
int ctx_attn_bwd_fp32(const float* dy, float* df, int n, int c, int h, int w, float scale, int fuse, void* fwd_workspace,
                      void* bwd_workspace, cudaStream_t st, bool tc = false);

int stitch(const float* gen, const float* real, const float* pred_h, const int32_t* x1, const int32_t* x2,
           const int32_t* height, int maxheight, float* out, int32_t* rows_out, int n, int h, int w,
           cudaStream_t st);
int threshold(const float* p, float* out_f32, uint8_t* out_u8, float value, size_t count, cudaStream_t st);
int sobel(const float* img, float* edges, int n, int h, int w, cudaStream_t st);
int edge_xor_loss(const float* fake_mask, const float* real_mask, unsigned long long* xor_count, float* loss,
                  int n, int h, int w, cudaStream_t st);
int column_heights(const uint8_t* vol_fake, const uint8_t* vol_label, int d0, int d1, int d2, int axis,
                   int z0, int z1, int32_t* counts, int32_t* meta, cudaStream_t st);

}  // namespace hv
