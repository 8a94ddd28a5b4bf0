// Stand-alone bf16 tensor-core convolution op on fp32 NCHW tensors (C ABI: hv_conv2d_bf16).
// Packs the sources into chunked bf16 buffers, runs the tcgen05 kernel, unpacks the result; it is
// the module-level / parity-test entry to the kernel the generator plan launches directly.
#include <vector>
#include "hv_common.cuh"
#include "conv_tc.cuh"
#include "kernels.h"

namespace hv {

// The stand-alone ops carve their chunked buffers out of the per-stream grow-only scratch (hv_api.cu): no allocation calls inside a
// training step (cudaMalloc synchronises the device; the stream-ordered pool gives its blocks back at every synchronisation).
struct Carver {
  char* base = nullptr;
  size_t off = 0;
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 1023) & ~(size_t)1023;
    return p;
  }
};

static void shape_buf(TcBuf& b, int n, int channels, int h, int w, int border, bool s2d) {
  b.n = n; b.chunks = ((channels + 15) / 16) * 2; b.h = h; b.w = w; b.border = border; b.s2d = s2d;
}

static void place_buf(TcBuf& b, Carver& cv) { b.ptr = static_cast<__nv_bfloat16*>(cv.take(b.bytes() + TcBuf::kSlackBytes)); }

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
  const bool heads = d->act == HV_ACT_HEADS;
  const int ho = d->hin / d->stride, wo = d->win / d->stride;
  const int sc = up2_out ? 2 : 1;
  TcBuf out;
  shape_buf(srcs[0].buf, d->n, ch0, d->hin, d->win, border, d->stride == 2);
  if (nts == 2) shape_buf(srcs[1].buf, d->n, ch1, d->hin, d->win, border, d->stride == 2);
  if (!heads) shape_buf(out, d->n, d->cout, ho * sc, wo * sc, 1, false);
  Carver cv;                                   // first pass: sizes; second pass: addresses
  for (int pass = 0; pass < 2; ++pass) {
    cv.off = 0;
    void* arena = cv.take(TcConv::kArenaBytes);
    for (int i = 0; i < nts; ++i) place_buf(srcs[i].buf, cv);
    if (!heads) place_buf(out, cv);
    if (pass == 0) {
      cv.base = static_cast<char*>(stream_scratch(st, cv.off));
      HV_CHECK_ARG(cv.base, "conv2d_bf16: no memory for %zu bytes of scratch", cv.off);
    } else {
      HV_CUDA(cudaMemsetAsync(static_cast<char*>(arena) + TcConv::kArenaBytes, 0, cv.off - TcConv::kArenaBytes, st));   // zero borders / padding channels
    }
  }
  TcConv c;
  c.arena = cv.base;
  srcs[0].real_channels = ch0;
  int rc = tc_pack_nchw(d->src[0].ptr, ch0, d->src[0].mode, srcs[0].buf, 0, st);
  if (rc) return rc;
  if (nts == 2) {
    srcs[1].real_channels = ch1;
    int off = 0;
    for (int i = 1; i < d->nsrc; ++i) {
      rc = tc_pack_nchw(d->src[i].ptr, d->src[i].channels, d->src[i].mode, srcs[1].buf, off, st);
      if (rc) return rc;
      off += d->src[i].channels;
    }
  }
  rc = tc_conv_setup(c, srcs, nts, d->k, d->stride, d->dil, d->cout, d->n);
  if (rc) return rc;
  c.force_generic = (flags & 2) != 0;
  if (heads) {
    rc = tc_conv_pack_weights(c, w, bias, 1, w + (size_t)d->cin * d->k * d->k, bias ? bias + 1 : nullptr, 1, st);
    if (rc) return rc;
    tc_conv_set_output_heads(c, y, y2, nullptr, nullptr);
  } else {
    rc = tc_conv_pack_weights(c, w, bias, d->cout, nullptr, nullptr, 0, st);
    if (rc) return rc;
    tc_conv_set_output_chunked(c, out, 0, c.n_pad / 8 < out.chunks ? c.n_pad / 8 : out.chunks, up2_out != 0, d->act);   // never past the buffer's own chunks
  }
  rc = tc_conv_launch(c, st);
  if (rc) return rc;
  if (!heads) {
    rc = tc_unpack_nchw(out, 0, d->cout, y, st);
    if (rc) return rc;
  }
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
  shape_buf(fb, n, 64, 64, 64, 1, false);
  shape_buf(yb, n, 64, 64, 64, 1, false);
  Carver cv;
  void* ws = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    cv.off = 0;
    place_buf(fb, cv);
    place_buf(yb, cv);
    ws = cv.take(ctx_attn_tc_workspace_bytes(n));
    if (pass == 0) {
      cv.base = static_cast<char*>(stream_scratch(st, cv.off));
      HV_CHECK_ARG(cv.base, "ctx_attn_fwd_bf16: no memory for %zu bytes of scratch", cv.off);
    }
  }
  HV_CUDA(cudaMemsetAsync(fb.ptr, 0, (size_t)(reinterpret_cast<char*>(ws) - reinterpret_cast<char*>(fb.ptr)), st));
  int rc = tc_pack_nchw(f, 64, HV_SRC_DIRECT, fb, 0, st);
  if (rc) return rc;
  rc = ctx_attn_fwd_tc(fb, mask, yb, offsets, flow, softmax_scale, fuse, per_sample_mask, ws, st);
  if (rc) return rc;
  return tc_unpack_nchw(yb, 0, 64, y, st);
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
