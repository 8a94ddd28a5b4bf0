// L2 -> shared-memory ingest probe (run on a B200 under gpurun): what does one SM / the whole chip sustain for the
// band loads of the tcgen05 conv kernel, by copy flavour and by where the data lives?
//   flavour T : one tensor-map box {256 x 8 B, rows, chunks} per load (what conv_tc issues)
//   flavour B : the same bytes as `rows * chunks` 1-D cp.async.bulk copies of 2 KB on one barrier
//   working set: 16 MB (L2 resident, re-read many times) or 1 GB (streams from HBM)
// Every CTA (1 thread) keeps `inflight` loads of `bytes` outstanding; output = bytes / cycle / SM on CTA 0 and the chip total.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_l2_probe tma_l2_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000ll) __trap(); }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Cfg {
  int flavour;        // 0 tensor box, 1 bulk rows
  int rows;           // 2 KB rows per load (box: chunks dimension)
  int inflight, loads;
  long long ws_rows;  // working set in 2 KB rows (loads wrap inside it)
  long long plane_b;  // bytes between box rows in global memory
};

__global__ void __launch_bounds__(32, 1) probe_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* g, Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)c.rows * 2048u;
    const long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t phase[16] = {0};
    // positions per plane available to slide over: ws_rows / rows 2-KB units
    const long long units = c.ws_rows / c.rows;
    while (waited < c.loads) {
      while (issued < c.loads && issued - waited < c.inflight) {
        const int s = issued % c.inflight;
        const uint32_t bar = smem_u32(&bars[s]);
        const uint32_t dst = smem_u32(smem + (size_t)s * bytes);
        const long long unit = ((long long)issued * gridDim.x + blockIdx.x) % units;
        mbar_expect_tx(bar, bytes);
        if (c.flavour == 0) {
          const int c0 = (int)(unit * 256 + 2 * (blockIdx.x & 7));   // 16 B granular start like a conv band
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(&map), "r"(bar), "r"(c0), "r"(0), "r"(0) : "memory");
        } else {
          const uint8_t* src = g + unit * 2048 + 16 * (blockIdx.x & 7);
          for (int r = 0; r < c.rows; ++r)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (uint32_t)r * 2048u), "l"(src + (long long)r * c.plane_b), "r"(2048u), "r"(bar) : "memory");
        }
        ++issued;
      }
      const int s = waited % c.inflight;
      mbar_wait(smem_u32(&bars[s]), phase[s]);
      phase[s] ^= 1u;
      ++waited;
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 148 * 8));
  long long h_out[148];
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)fnp;
  const size_t total = (size_t)1 << 30;
  uint8_t* g; CK(cudaMalloc(&g, total + (1 << 20))); CK(cudaMemset(g, 1, total + (1 << 20)));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int rows : {8, 24}) {                       // 16 KB (one kernel row of a 64-channel band) and 48 KB (3 rows x 8 chunks)
    for (long long ws_mb : {16ll, 1024ll}) {
      const long long ws_rows = ws_mb * 1024 * 1024 / 2048;
      // `rows` planes of ws_rows/rows 2-KB units each: box row r lives plane_b bytes after row r-1
      const long long plane_b = ws_rows / rows * 2048;
      CUtensorMap map;
      cuuint64_t dims[3] = {(cuuint64_t)plane_b / 8, (cuuint64_t)rows, 1};
      cuuint64_t strides[2] = {(cuuint64_t)plane_b, (cuuint64_t)plane_b * rows};
      cuuint32_t box[3] = {256, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      for (int flavour : {0, 1})
        for (int nctas : {148, 8})
          for (int inflight : {1, 2, 4, 8}) {
            if ((size_t)inflight * rows * 2048 > 190 * 1024) continue;
            Cfg c{flavour, rows, inflight, 256, ws_rows, plane_b};
            double best = 1e30, worst = 0;
            for (int rep = 0; rep < 3; ++rep) {   // rep 0 warms L2 for the 16 MB set
              probe_kernel<<<nctas, 32, 200 * 1024>>>(map, g, c, d_out);
              CK(cudaDeviceSynchronize());
              CK(cudaMemcpy(h_out, d_out, nctas * 8, cudaMemcpyDeviceToHost));
              if (rep == 0) continue;
              double mx = 0;
              for (int i = 0; i < nctas; ++i) mx = h_out[i] > mx ? h_out[i] : mx;
              best = mx < best ? mx : best; worst = mx > worst ? mx : worst;
            }
            const double bpc = (double)rows * 2048 * c.loads / best;
            printf("%s rows=%2d (%2d KB) ws=%4lld MB ctas=%3d inflight=%d : %6.0f cyc/load  %5.1f B/cyc/SM  chip %.2f TB/s\n", flavour ? "bulk1d" : "tensor", rows,
                   rows * 2, ws_mb, nctas, inflight, best / c.loads, bpc, bpc * nctas * clk_khz * 1e3 / 1e12);
          }
    }
  }
  return 0;
}
