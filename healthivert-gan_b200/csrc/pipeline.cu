// uint8 host interface of the generator with a captured CUDA graph and a ring of pinned host slots.
//
// The reference's inference driver hands the generator uint8 planes and keeps uint8 results
// (eval_3d_sagittal_twostage.py:84-98: ct / CAM planes go through numpy_to_pil -> ToTensor -> Normalize, the mask is a row range
// painted 255; :103-130: thresholded mask, (x + 1) * 127.5 later truncated by astype(uint8), pred_h).  This file is that boundary
// for HOST buffers: per slot one contiguous pinned input block (ct u8 | cam u8 | mask row ranges | slice ratios) and one
// pinned output block (ct u8 | fine mask u8 | coarse mask u8 | pred1_h, pred2_h), i.e. 1/4 of the bytes of the fp32 tensors
// each way.  submit() = one H2D copy on the input stream, ONE cudaGraphLaunch (u8 -> fp32 unpack, the ~60 kernels of the
// two-stage forward on their three streams, fp32 -> u8 finish) on the compute stream, one D2H copy on the output stream;
// consecutive slots overlap (copy of slot i+1 / i-1 under the forward of slot i).  The host does ~10 driver calls per batch.
#include <string.h>
#include <utility>
#include <vector>
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

constexpr int PL_H = 256, PL_W = 256, PL_HW = PL_H * PL_W;

// ToTensor (u8 / 255) + Normalize((x - 0.5) / 0.5) in IEEE fp32, as slice_prep.cu
__device__ __forceinline__ float pl_norm_u8(int u) { return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), 0.5f), 0.5f); }

// one thread = 4 consecutive pixels: a 4-byte load per u8 plane, three 16-byte stores
__global__ void __launch_bounds__(256) pl_unpack_kernel(const uint8_t* __restrict__ ct, const uint8_t* __restrict__ cam,
                                                        const int32_t* __restrict__ rows, float* __restrict__ x,
                                                        float* __restrict__ mask, float* __restrict__ cam1m, int n) {
  const int i4 = blockIdx.x * blockDim.x + threadIdx.x;   // index of a group of 4 pixels
  if (i4 >= n * (PL_HW / 4)) return;
  const int b = i4 / (PL_HW / 4), r = (i4 - b * (PL_HW / 4)) / (PL_W / 4);
  const uchar4 c = reinterpret_cast<const uchar4*>(ct)[i4], m = reinterpret_cast<const uchar4*>(cam)[i4];
  const float mk = (r >= rows[2 * b] && r < rows[2 * b + 1]) ? 1.f : 0.f;   // mask_slice[r0:r1] = 255 -> ToTensor -> 1.0
  reinterpret_cast<float4*>(x)[i4] = make_float4(pl_norm_u8(c.x), pl_norm_u8(c.y), pl_norm_u8(c.z), pl_norm_u8(c.w));
  reinterpret_cast<float4*>(mask)[i4] = make_float4(mk, mk, mk, mk);
  reinterpret_cast<float4*>(cam1m)[i4] = make_float4(__fsub_rn(1.f, __fdiv_rn((float)m.x, 255.f)), __fsub_rn(1.f, __fdiv_rn((float)m.y, 255.f)),
                                                     __fsub_rn(1.f, __fdiv_rn((float)m.z, 255.f)), __fsub_rn(1.f, __fdiv_rn((float)m.w, 255.f)));
}

__device__ __forceinline__ uint8_t pl_ct_u8(float v) {   // (fake_B + 1) * 127.5 (eval:121), then numpy astype(uint8) truncation
  return (uint8_t)(int)__fmul_rn(__fadd_rn(v, 1.f), 127.5f);
}

__global__ void __launch_bounds__(256) pl_finish_kernel(const float* __restrict__ x2, const float* __restrict__ fine, const float* __restrict__ coarse,
                                                        const float* __restrict__ p1, const float* __restrict__ p2, uint8_t* __restrict__ ct,
                                                        uint8_t* __restrict__ fine_u8, uint8_t* __restrict__ coarse_u8, float* __restrict__ heights, int n, int batch) {
  const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 < n) { heights[i4] = p1[i4]; heights[batch + i4] = p2[i4]; }   // [2][batch] like the host view
  if (i4 >= n * (PL_HW / 4)) return;
  const float4 v = reinterpret_cast<const float4*>(x2)[i4], f = reinterpret_cast<const float4*>(fine)[i4], c = reinterpret_cast<const float4*>(coarse)[i4];
  reinterpret_cast<uchar4*>(ct)[i4] = make_uchar4(pl_ct_u8(v.x), pl_ct_u8(v.y), pl_ct_u8(v.z), pl_ct_u8(v.w));
  reinterpret_cast<uchar4*>(fine_u8)[i4] = make_uchar4(f.x > 0.5f, f.y > 0.5f, f.z > 0.5f, f.w > 0.5f);      // torch.where(seg > 0.5, 1, 0), eval:105
  reinterpret_cast<uchar4*>(coarse_u8)[i4] = make_uchar4(c.x > 0.5f, c.y > 0.5f, c.z > 0.5f, c.w > 0.5f);
}

struct Slot {
  uint8_t* h_in = nullptr;    // pinned: ct | cam | rows | ratio
  uint8_t* h_out = nullptr;   // pinned: ct | fine | coarse | heights
  uint8_t* d_in = nullptr;
  uint8_t* d_out = nullptr;
  cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_out = nullptr;
  std::vector<std::pair<int, cudaGraphExec_t>> graphs;   // (batch, executable graph)
  std::vector<int> graph_kernels;
  bool in_flight = false;
};

}  // namespace hv

using namespace hv;

struct hv_pipeline {
  hv_generator* gen = nullptr;
  int batch = 0, depth = 0, per_sample_mask = 0, use_graph = 1;
  size_t in_bytes = 0, out_bytes = 0;
  cudaStream_t s_in = nullptr, s_main = nullptr, s_out = nullptr;
  // fp32 staging of the forward (shared by the slots: forwards are serialised on s_main)
  float *x = nullptr, *mask = nullptr, *cam = nullptr, *ratio_unused = nullptr;
  float *coarse = nullptr, *fine = nullptr, *x1 = nullptr, *x2 = nullptr, *p1 = nullptr, *p2 = nullptr;
  std::vector<Slot> slots;
};

namespace hv {

static size_t pl_in_bytes(int n) { return (size_t)n * PL_HW * 2 + (size_t)n * 2 * sizeof(int32_t) + (size_t)n * sizeof(float); }
static size_t pl_out_bytes(int n) { return (size_t)n * PL_HW * 3 + (size_t)n * 2 * sizeof(float); }

// the work of one batch on the compute stream (captured into a graph, or launched directly when graphs are switched off)
static int pl_enqueue(hv_pipeline* p, Slot& s, int n) {
  const int B = p->batch;
  const uint8_t* d_ct = s.d_in;
  const uint8_t* d_cam = s.d_in + (size_t)B * PL_HW;
  const int32_t* d_rows = reinterpret_cast<const int32_t*>(s.d_in + (size_t)B * PL_HW * 2);
  const float* d_ratio = reinterpret_cast<const float*>(s.d_in + (size_t)B * PL_HW * 2 + (size_t)B * 2 * sizeof(int32_t));
  const int blocks = (n * (PL_HW / 4) + 255) / 256;
  pl_unpack_kernel<<<blocks, 256, 0, p->s_main>>>(d_ct, d_cam, d_rows, p->x, p->mask, p->cam, n);
  HV_LAUNCH_CHECK();
  int rc = hv_generator_forward(p->gen, p->x, p->mask, p->cam, d_ratio, n, p->coarse, p->fine, p->x1, p->x2, nullptr, p->p1, p->p2, nullptr,
                                p->per_sample_mask, (hv_stream_t)p->s_main);
  if (rc) return rc;
  uint8_t* o_ct = s.d_out;
  pl_finish_kernel<<<blocks, 256, 0, p->s_main>>>(p->x2, p->fine, p->coarse, p->p1, p->p2, o_ct, o_ct + (size_t)B * PL_HW, o_ct + (size_t)B * PL_HW * 2,
                                                   reinterpret_cast<float*>(o_ct + (size_t)B * PL_HW * 3), n, B);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv

extern "C" {

int hv_pipeline_destroy(hv_pipeline* p) {
  if (!p) return HV_OK;
  if (p->s_main) cudaStreamSynchronize(p->s_main);
  if (p->s_in) cudaStreamSynchronize(p->s_in);
  if (p->s_out) cudaStreamSynchronize(p->s_out);
  for (Slot& s : p->slots) {
    for (auto& g : s.graphs) cudaGraphExecDestroy(g.second);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    cudaFree(s.d_in); cudaFree(s.d_out);
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_comp) cudaEventDestroy(s.ev_comp);
    if (s.ev_out) cudaEventDestroy(s.ev_out);
  }
  cudaFree(p->x);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_main) cudaStreamDestroy(p->s_main);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  return HV_OK;
}

int hv_pipeline_create(hv_pipeline** out, hv_generator* gen, int batch, int depth, int per_sample_mask, int use_graph) {
  HV_CHECK_ARG(out && gen, "pipeline_create: null argument");
  HV_CHECK_ARG(batch >= 1 && batch <= 4096 && depth >= 1 && depth <= 64, "pipeline_create: batch %d / depth %d out of range", batch, depth);
  hv_pipeline* p = new hv_pipeline();
  p->gen = gen; p->batch = batch; p->depth = depth; p->per_sample_mask = per_sample_mask; p->use_graph = use_graph;
  p->in_bytes = pl_in_bytes(batch); p->out_bytes = pl_out_bytes(batch);
#define PL_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { hv::set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); hv_pipeline_destroy(p); return HV_ERR_CUDA; } } while (0)
  PL_TRY(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  PL_TRY(cudaStreamCreateWithFlags(&p->s_main, cudaStreamNonBlocking));
  PL_TRY(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  const size_t plane = (size_t)batch * PL_HW;
  PL_TRY(cudaMalloc((void**)&p->x, sizeof(float) * (7 * plane + 2 * (size_t)batch + 64)));
  p->mask = p->x + plane; p->cam = p->mask + plane; p->coarse = p->cam + plane; p->fine = p->coarse + plane;
  p->x1 = p->fine + plane; p->x2 = p->x1 + plane; p->p1 = p->x2 + plane; p->p2 = p->p1 + batch;
  p->slots.resize(depth);
  for (Slot& s : p->slots) {
    PL_TRY(cudaHostAlloc((void**)&s.h_in, p->in_bytes, cudaHostAllocDefault));
    PL_TRY(cudaHostAlloc((void**)&s.h_out, p->out_bytes, cudaHostAllocDefault));
    PL_TRY(cudaMalloc((void**)&s.d_in, p->in_bytes));
    PL_TRY(cudaMalloc((void**)&s.d_out, p->out_bytes));
    PL_TRY(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
    PL_TRY(cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
    PL_TRY(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    memset(s.h_in, 0, p->in_bytes);
  }
#undef PL_TRY
  *out = p;
  return HV_OK;
}

/* host pointers into slot `slot` (all inside its two pinned blocks, sized for the pipeline's batch) */
int hv_pipeline_slot(hv_pipeline* p, int slot, uint8_t** ct_in, uint8_t** cam_in, int32_t** rows_in, float** ratio_in, uint8_t** ct_out,
                     uint8_t** fine_mask_out, uint8_t** coarse_mask_out, float** heights_out) {
  HV_CHECK_ARG(p && slot >= 0 && slot < p->depth, "pipeline_slot: bad handle or slot %d", slot);
  Slot& s = p->slots[slot];
  const size_t plane = (size_t)p->batch * PL_HW;
  if (ct_in) *ct_in = s.h_in;
  if (cam_in) *cam_in = s.h_in + plane;
  if (rows_in) *rows_in = reinterpret_cast<int32_t*>(s.h_in + 2 * plane);
  if (ratio_in) *ratio_in = reinterpret_cast<float*>(s.h_in + 2 * plane + (size_t)p->batch * 2 * sizeof(int32_t));
  if (ct_out) *ct_out = s.h_out;
  if (fine_mask_out) *fine_mask_out = s.h_out + plane;
  if (coarse_mask_out) *coarse_mask_out = s.h_out + 2 * plane;
  if (heights_out) *heights_out = reinterpret_cast<float*>(s.h_out + 3 * plane);
  return HV_OK;
}

size_t hv_pipeline_bytes(hv_pipeline* p, int which) { return p ? (which == 0 ? p->in_bytes : p->out_bytes) : 0; }

void* hv_pipeline_stream(hv_pipeline* p, int which) {
  if (!p) return nullptr;
  return which == 0 ? (void*)p->s_in : (which == 1 ? (void*)p->s_main : (void*)p->s_out);
}

int hv_pipeline_wait(hv_pipeline* p, int slot) {
  HV_CHECK_ARG(p && slot >= 0 && slot < p->depth, "pipeline_wait: bad handle or slot %d", slot);
  Slot& s = p->slots[slot];
  if (!s.in_flight) return HV_OK;
  HV_CUDA(cudaEventSynchronize(s.ev_out));
  s.in_flight = false;
  return HV_OK;
}

int hv_pipeline_submit(hv_pipeline* p, int slot, int n) {
  HV_CHECK_ARG(p && slot >= 0 && slot < p->depth, "pipeline_submit: bad handle or slot %d", slot);
  HV_CHECK_ARG(n >= 1 && n <= p->batch, "pipeline_submit: batch %d outside 1..%d", n, p->batch);
  Slot& s = p->slots[slot];
  if (s.in_flight) { set_error("pipeline_submit: slot %d is still in flight (call hv_pipeline_wait first)", slot); return HV_ERR_STATE; }
  // inputs: one copy of the whole block (the row ranges / ratios sit behind the planes)
  HV_CUDA(cudaMemcpyAsync(s.d_in, s.h_in, p->in_bytes, cudaMemcpyHostToDevice, p->s_in));
  HV_CUDA(cudaEventRecord(s.ev_in, p->s_in));
  HV_CUDA(cudaStreamWaitEvent(p->s_main, s.ev_in, 0));
  if (p->use_graph) {
    cudaGraphExec_t exec = nullptr;
    int nk = 0;
    for (size_t i = 0; i < s.graphs.size(); ++i)
      if (s.graphs[i].first == n) { exec = s.graphs[i].second; nk = s.graph_kernels[i]; }
    if (!exec) {
      // first use of (slot, n): warm up once un-captured (lazy one-off work of the library: kernel attributes, tensor-map cache),
      // then capture the same enqueue and keep the executable graph
      int rc = pl_enqueue(p, s, n);
      if (rc) return rc;
      HV_CUDA(cudaStreamSynchronize(p->s_main));
      cudaGraph_t graph = nullptr;
      HV_CUDA(cudaStreamBeginCapture(p->s_main, cudaStreamCaptureModeThreadLocal));
      rc = pl_enqueue(p, s, n);
      cudaError_t e = cudaStreamEndCapture(p->s_main, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) { set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); return HV_ERR_CUDA; }
      size_t nodes = 0;
      HV_CUDA(cudaGraphGetNodes(graph, nullptr, &nodes));
      std::vector<cudaGraphNode_t> list(nodes);
      HV_CUDA(cudaGraphGetNodes(graph, list.data(), &nodes));
      for (size_t i = 0; i < nodes; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(list[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++nk;
      }
      e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return HV_ERR_CUDA; }
      s.graphs.push_back(std::make_pair(n, exec));
      s.graph_kernels.push_back(nk);
    }
    HV_CUDA(cudaGraphLaunch(exec, p->s_main));
    count_launch(nk);
  } else {
    int rc = pl_enqueue(p, s, n);
    if (rc) return rc;
  }
  HV_CUDA(cudaEventRecord(s.ev_comp, p->s_main));
  HV_CUDA(cudaStreamWaitEvent(p->s_out, s.ev_comp, 0));
  HV_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, p->out_bytes, cudaMemcpyDeviceToHost, p->s_out));
  HV_CUDA(cudaEventRecord(s.ev_out, p->s_out));
  s.in_flight = true;
  return HV_OK;
}

}  // extern "C"
