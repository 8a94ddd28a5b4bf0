// Stand-alone bf16 tensor-core convolution op on fp32 NCHW tensors (C ABI: hv_conv2d_bf16).
// Packs the sources into chunked bf16 buffers, runs the tcgen05 kernel, unpacks the result; it is
// the module-level / parity-test entry to the kernel the generator plan launches directly.
#include <vector>
#include "hv_common.cuh"
#include "conv_tc.cuh"
#include "kernels.h"

namespace hv {

static int alloc_buf(TcBuf& b, int n, int channels, int h, int w, int border, bool s2d, cudaStream_t st) {
  b.n = n; b.chunks = ((channels + 15) / 16) * 2; b.h = h; b.w = w; b.border = border; b.s2d = s2d;
  HV_CUDA(cudaMallocAsync((void**)&b.ptr, b.bytes() + TcBuf::kSlackBytes, st));
  HV_CUDA(cudaMemsetAsync(b.ptr, 0, b.bytes() + TcBuf::kSlackBytes, st));
  return HV_OK;
}

int conv2d_bf16(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2, int flags,
                cudaStream_t st) {
  const int up2_out = flags & 1;
  HV_CHECK_ARG(d && w && y, "conv2d_bf16: null argument");
  HV_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= 4, "conv2d_bf16: nsrc out of range");
  HV_CHECK_ARG(d->k == 3 || d->k == 5, "conv2d_bf16: kernel %d not built (3 or 5)", d->k);
  HV_CHECK_ARG(d->pad == (d->k - 1) / 2 * d->dil, "conv2d_bf16: only 'same' padding (pad = dil*(k-1)/2)");
  HV_CHECK_ARG(d->stride == 1 || (d->stride == 2 && d->k == 3 && d->dil == 1 && (d->hin % 2) == 0 && (d->win % 2) == 0),
               "conv2d_bf16: stride-2 needs k=3, dil=1, even extent");
  HV_CHECK_ARG(d->cout <= 64, "conv2d_bf16: cout <= 64");
  HV_CHECK_ARG(d->act != HV_ACT_HEADS || (d->cout == 2 && y2), "conv2d_bf16: HEADS needs cout=2 and y2");
  const int border = d->pad > 0 ? d->pad : 1;
  TcSource srcs[2];
  int nts = d->nsrc == 1 ? 1 : 2;
  int ch0 = d->src[0].channels, ch1 = 0;
  for (int i = 1; i < d->nsrc; ++i) ch1 += d->src[i].channels;
  HV_CHECK_ARG(ch0 + ch1 == d->cin, "conv2d_bf16: sources have %d channels, cin=%d", ch0 + ch1, d->cin);
  int rc = alloc_buf(srcs[0].buf, d->n, ch0, d->hin, d->win, border, d->stride == 2, st);
  if (rc) return rc;
  srcs[0].real_channels = ch0;
  rc = tc_pack_nchw(d->src[0].ptr, ch0, d->src[0].mode, srcs[0].buf, 0, st);
  if (rc) return rc;
  if (nts == 2) {
    rc = alloc_buf(srcs[1].buf, d->n, ch1, d->hin, d->win, border, d->stride == 2, st);
    if (rc) return rc;
    srcs[1].real_channels = ch1;
    int off = 0;
    for (int i = 1; i < d->nsrc; ++i) {
      rc = tc_pack_nchw(d->src[i].ptr, d->src[i].channels, d->src[i].mode, srcs[1].buf, off, st);
      if (rc) return rc;
      off += d->src[i].channels;
    }
  }
  TcConv c;
  rc = tc_conv_setup(c, srcs, nts, d->k, d->stride, d->dil, d->cout, d->n);
  if (rc) return rc;
  c.force_generic = (flags & 2) != 0;
  const int ho = d->hin / d->stride, wo = d->win / d->stride;
  TcBuf out;
  if (d->act == HV_ACT_HEADS) {
    rc = tc_conv_pack_weights(c, w, bias, 1, w + (size_t)d->cin * d->k * d->k, bias ? bias + 1 : nullptr, 1, st);
    if (rc) return rc;
    tc_conv_set_output_heads(c, y, y2, nullptr, nullptr);
  } else {
    rc = tc_conv_pack_weights(c, w, bias, d->cout, nullptr, nullptr, 0, st);
    if (rc) return rc;
    const int sc = up2_out ? 2 : 1;
    rc = alloc_buf(out, d->n, d->cout, ho * sc, wo * sc, 1, false, st);
    if (rc) return rc;
    tc_conv_set_output_chunked(c, out, 0, c.n_pad / 8 < out.chunks ? c.n_pad / 8 : out.chunks, up2_out != 0, d->act);   // never past the buffer's own chunks
  }
  rc = tc_conv_launch(c, st);
  if (rc) return rc;
  if (d->act != HV_ACT_HEADS) {
    rc = tc_unpack_nchw(out, 0, d->cout, y, st);
    if (rc) return rc;
    HV_CUDA(cudaFreeAsync(out.ptr, st));
  }
  for (int i = 0; i < nts; ++i) HV_CUDA(cudaFreeAsync(srcs[i].buf.ptr, st));
  HV_CUDA(cudaFreeAsync(c.w_packed, st));
  HV_CUDA(cudaFreeAsync(c.bias_pad, st));
  return HV_OK;
}

}  // namespace hv

extern "C" int hv_conv2d_bf16(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2, int flags,
                              hv_stream_t stream) {
  return hv::conv2d_bf16(d, w, bias, y, y2, flags, hv::as_stream(stream));
}

// bf16 tensor-core variant of hv_ctx_attn_fwd on fp32 NCHW tensors: pack -> ctx_attn_fwd_tc -> unpack
extern "C" int hv_ctx_attn_fwd_bf16(const float* f, const float* mask, float* y, int32_t* offsets, float* flow, int n, int c,
                                    int h, int w, float softmax_scale, int fuse, int per_sample_mask, hv_stream_t stream) {
  using namespace hv;
  cudaStream_t st = as_stream(stream);
  HV_CHECK_ARG(f && mask && y, "ctx_attn_fwd_bf16: null argument");
  HV_CHECK_ARG(c == 64 && h == 64 && w == 64 && n >= 1, "ctx_attn_fwd_bf16: built for [n,64,64,64] features (got c=%d h=%d w=%d)", c, h, w);
  TcBuf fb, yb;
  int rc = alloc_buf(fb, n, 64, 64, 64, 1, false, st);
  if (rc) return rc;
  rc = alloc_buf(yb, n, 64, 64, 64, 1, false, st);
  if (rc) return rc;
  rc = tc_pack_nchw(f, 64, HV_SRC_DIRECT, fb, 0, st);
  if (rc) return rc;
  void* ws = nullptr;
  HV_CUDA(cudaMallocAsync(&ws, ctx_attn_tc_workspace_bytes(n), st));
  rc = ctx_attn_fwd_tc(fb, mask, yb, offsets, flow, softmax_scale, fuse, per_sample_mask, ws, st);
  if (rc) return rc;
  rc = tc_unpack_nchw(yb, 0, 64, y, st);
  if (rc) return rc;
  HV_CUDA(cudaFreeAsync(ws, st));
  HV_CUDA(cudaFreeAsync(fb.ptr, st));
  HV_CUDA(cudaFreeAsync(yb.ptr, st));
  return HV_OK;
}

// debug hook (not part of the drop-in surface): dev_buf = 12000 int64 on the device, or NULL to switch tracing off
// debug hook: conv launch i (in host launch order) records {first CTA entry, last CTA exit} (globaltimer ns) at dev_buf[4 i], [4 i + 1];
// the caller initialises the entry slots to a large value and the exit slots to 0
extern "C" int hv_debug_conv_timeline(void* dev_buf) {
  hv::tc_set_timeline(reinterpret_cast<long long*>(dev_buf));
  return HV_OK;
}
extern "C" int hv_debug_conv_trace(void* dev_buf) {
  hv::tc_set_trace(reinterpret_cast<long long*>(dev_buf));
  return HV_OK;
}
