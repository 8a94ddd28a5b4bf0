// Spectral-norm weight preparation: optional power iteration (train mode), sigma = u^T W v,
// w_eff = W / sigma.  One CTA per layer; a device-side job table batches all 47 layers of
// the generator into a single launch.
//
// Replaces SpectralNorm.compute_weight (torch/nn/utils/spectral_norm.py:92-114) as hooked
// onto every generator conv by the reference (models/inpaint_networks.py:491-492).
#include "hv_common.cuh"
#include "kernels.h"

namespace hv {

__device__ void sn_prepare_block(const SnJob& j, int training, float* red) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const float* W = j.w;
  const int cout = j.cout, kd = j.kdim;
  __shared__ float s_wv[1024];
  if (training) {
    // v <- normalize(W^T u)
    float ss = 0.f;
    for (int k = tid; k < kd; k += nt) {
      float a = 0.f;
      for (int co = 0; co < cout; ++co) a = fmaf(W[(size_t)co * kd + k], j.u[co], a);
      j.v[k] = a;
      ss = fmaf(a, a, ss);
    }
    float nrm = fmaxf(sqrtf(block_sum(ss, red)), 1e-12f);
    for (int k = tid; k < kd; k += nt) j.v[k] = j.v[k] / nrm;
    __syncthreads();
  }
  // wv = W v  (one warp per output row)
  for (int co = warp; co < cout; co += nw) {
    float a = 0.f;
    for (int k = lane; k < kd; k += 32) a = fmaf(W[(size_t)co * kd + k], j.v[k], a);
    a = warp_sum(a);
    if (lane == 0) s_wv[co] = a;
  }
  __syncthreads();
  if (training) {
    float ss = 0.f;
    for (int co = tid; co < cout; co += nt) ss = fmaf(s_wv[co], s_wv[co], ss);
    float nrm = fmaxf(sqrtf(block_sum(ss, red)), 1e-12f);
    for (int co = tid; co < cout; co += nt) j.u[co] = s_wv[co] / nrm;
    __syncthreads();
  }
  float d = 0.f;
  for (int co = tid; co < cout; co += nt) d = fmaf(j.u[co], s_wv[co], d);
  const float sigma = block_sum(d, red);
  if (tid == 0 && j.sigma) *j.sigma = sigma;
  if (j.w_eff) {
    const size_t total = (size_t)cout * kd;
    for (size_t i = tid; i < total; i += nt) j.w_eff[i] = W[i] / sigma;
  }
}

__global__ void __launch_bounds__(512) sn_prepare_kernel(const SnJob* jobs, int training) {
  __shared__ float red[32];
  sn_prepare_block(jobs[blockIdx.x], training, red);
}

__global__ void __launch_bounds__(512) sn_prepare_single_kernel(SnJob job, int training) {
  __shared__ float red[32];
  sn_prepare_block(job, training, red);
}

int sn_prepare_batched(const SnJob* d_jobs, int njobs, int training, cudaStream_t st) {
  sn_prepare_kernel<<<njobs, 512, 0, st>>>(d_jobs, training);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

int sn_prepare_single(const SnJob& job, int training, cudaStream_t st) {
  HV_CHECK_ARG(job.w && job.u && job.v, "sn_prepare: null argument");
  HV_CHECK_ARG(job.cout > 0 && job.cout <= 1024 && job.kdim > 0, "sn_prepare: cout must be in 1..1024");
  sn_prepare_single_kernel<<<1, 512, 0, st>>>(job, training);
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv
