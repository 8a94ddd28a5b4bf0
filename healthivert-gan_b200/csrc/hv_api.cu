// extern "C" surface of libhv_b200.so (declared in include/hv_b200.h) for the stand-alone ops,
// plus the thread-local error string and the launch counter.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Small internal scratch buffers keyed by (device, stream): kernels enqueued on one stream are ordered, so one grow-only
// buffer per stream is race-free; different streams / devices / host threads get different buffers (the C ABI stays re-entrant).
void* stream_scratch(cudaStream_t st, size_t bytes) {
  static std::mutex mu;
  struct Buf { void* ptr = nullptr; size_t cap = 0; };
  static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  Buf& b = bufs[std::make_pair(dev, st)];
  if (bytes > b.cap) {
    if (b.ptr) {   // work already enqueued on this stream may still read the old buffer
      if (cudaStreamSynchronize(st) != cudaSuccess) return nullptr;
      cudaFree(b.ptr);
      b.ptr = nullptr; b.cap = 0;
    }
    const size_t grow = bytes > (1u << 20) ? bytes : (1u << 20);
    if (cudaMalloc(&b.ptr, grow) != cudaSuccess) { b.ptr = nullptr; return nullptr; }
    b.cap = grow;
  }
  return b.ptr;
}

}  // namespace hv

using namespace hv;

extern "C" {

const char* hv_last_error(void) { return g_err; }
int hv_version(void) { return 100; }
uint64_t hv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int hv_sn_prepare(const float* w_orig, float* u, float* v, int cout, int kdim, int training, float* w_eff,
                  float* sigma_out, hv_stream_t stream) {
  SnJob j;
  j.w = w_orig; j.u = u; j.v = v; j.cout = cout; j.kdim = kdim; j.w_eff = w_eff; j.sigma = sigma_out;
  return sn_prepare_single(j, training, as_stream(stream));
}

int hv_conv2d_fwd(const hv_conv_desc* d, const float* w, const float* bias, float* y, float* y2, hv_stream_t stream) {
  return conv2d_fwd_fp32(d, w, bias, y, y2, as_stream(stream));
}

int hv_gap_fc_sigmoid(const float* x, const float* fc_w, const float* fc_b, float* out, int n, int c, int hw,
                      hv_stream_t stream) {
  return gap_fc_sigmoid(x, fc_w, fc_b, out, n, c, hw, as_stream(stream));
}

size_t hv_ctx_attn_workspace_bytes(int n, int c, int h, int w) {
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return 0;
  return ctx_attn_workspace_bytes(n, c, h, w);
}

int hv_ctx_attn_fwd(const float* f, const float* mask, float* y, int32_t* offsets, float* flow, int n, int c, int h,
                    int w, float softmax_scale, int fuse, int per_sample_mask, void* workspace, hv_stream_t stream) {
  return ctx_attn_fwd_fp32(f, mask, y, offsets, flow, n, c, h, w, softmax_scale, fuse, per_sample_mask, workspace,
                           as_stream(stream));
}

size_t hv_ctx_attn_bwd_workspace_bytes(int n, int c, int h, int w) {
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return 0;
  return ctx_attn_bwd_workspace_bytes(n, c, h, w);
}

int hv_ctx_attn_bwd(const float* dy, float* df, int n, int c, int h, int w, float softmax_scale, int fuse, void* fwd_workspace,
                    void* bwd_workspace, hv_stream_t stream) {
  return ctx_attn_bwd_fp32(dy, df, n, c, h, w, softmax_scale, fuse, fwd_workspace, bwd_workspace, as_stream(stream));
}

int hv_ctx_attn_fwd_tc(const float* f, const float* mask, float* y, int32_t* offsets, float* flow, int n, int c, int h,
                       int w, float softmax_scale, int fuse, int per_sample_mask, void* workspace, hv_stream_t stream) {
  return ctx_attn_fwd_fp32(f, mask, y, offsets, flow, n, c, h, w, softmax_scale, fuse, per_sample_mask, workspace,
                           as_stream(stream), true);
}

int hv_ctx_attn_bwd_tc(const float* dy, float* df, int n, int c, int h, int w, float softmax_scale, int fuse, void* fwd_workspace,
                       void* bwd_workspace, hv_stream_t stream) {
  return ctx_attn_bwd_fp32(dy, df, n, c, h, w, softmax_scale, fuse, fwd_workspace, bwd_workspace, as_stream(stream), true);
}

int hv_stitch(const float* gen, const float* real, const float* pred_h, const int32_t* x1, const int32_t* x2,
              const int32_t* height, int maxheight, float* out, int32_t* rows_out, int n, int h, int w,
              hv_stream_t stream) {
  return stitch(gen, real, pred_h, x1, x2, height, maxheight, out, rows_out, n, h, w, as_stream(stream));
}

int hv_threshold(const float* p, float* out_f32, uint8_t* out_u8, float value, size_t count, hv_stream_t stream) {
  return threshold(p, out_f32, out_u8, value, count, as_stream(stream));
}

int hv_sobel(const float* img, float* edges, int n, int h, int w, hv_stream_t stream) {
  return sobel(img, edges, n, h, w, as_stream(stream));
}

int hv_edge_xor_loss(const float* fake_mask, const float* real_mask, unsigned long long* xor_count, float* loss, int n,
                     int h, int w, hv_stream_t stream) {
  return edge_xor_loss(fake_mask, real_mask, xor_count, loss, n, h, w, as_stream(stream));
}

int hv_column_heights(const uint8_t* vol_fake, const uint8_t* vol_label, int d0, int d1, int d2, int axis, int z0,
                      int z1, int32_t* counts, int32_t* meta, hv_stream_t stream) {
  return column_heights(vol_fake, vol_label, d0, d1, d2, axis, z0, z1, counts, meta, as_stream(stream));
}

}  // extern "C"
