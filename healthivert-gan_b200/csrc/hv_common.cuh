// Shared helpers for the hv_b200 kernels (error plumbing, launch counting, math).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/hv_b200.h"

namespace hv {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// grow-only device scratch private to (current device, stream); nullptr on allocation failure
void* stream_scratch(cudaStream_t st, size_t bytes);

#define HV_CHECK_ARG(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      hv::set_error(__VA_ARGS__);                    \
      return HV_ERR_INVALID;                         \
    }                                                \
  } while (0)

#define HV_CUDA(expr)                                                            \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      hv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                    __FILE__, __LINE__);                                         \
      return HV_ERR_CUDA;                                                        \
    }                                                                            \
  } while (0)

// after a kernel launch: catches launch-configuration errors without synchronising
#define HV_LAUNCH_CHECK()                 \
  do {                                    \
    hv::count_launch();                   \
    HV_CUDA(cudaGetLastError());          \
  } while (0)

static inline cudaStream_t as_stream(hv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch for the chains of short kernels (attention, input packing): the launch of kernel i + 1 overlaps
// the execution of kernel i.  A kernel launched this way calls pdl_prologue() before it reads anything its predecessor wrote.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case HV_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case HV_ACT_RELU: return fmaxf(x, 0.f);
    case HV_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case HV_ACT_LRELU02: return x > 0.f ? x : 0.2f * x;
    case HV_ACT_CLAMP1: return fminf(fmaxf(x, -1.f), 1.f);
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? smem32[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

}  // namespace hv
