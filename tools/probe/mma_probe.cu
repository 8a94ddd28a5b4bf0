// Micro-probe for design constants of the tcgen05 conv / GEMM kernels (run on a B200 under gpurun):
//  (A) cycles per tcgen05.mma for operand layouts (no-swizzle interleaved vs 128B swizzle), N, and shifted A views
//  (B) numeric check of a 128B-swizzled A operand whose start address is shifted by whole rows (base_offset field)
//  (C) TMA load latency / concurrency for several box shapes
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe mma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <math.h>

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// layout_type: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// ------------------------------------------------------------------------------------------- (A) MMA rate
// mode 0: no-swizzle interleaved ([kchunk][rows][8]) LBO = rows*16, SBO = 128; mode 1: SW128 rows of 128 B
struct RateCfg { int mode, n, shift_bytes, reps, distinct; };
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(RateCfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 96 * 1024;
    const uint32_t idesc = make_idesc(c.n);
    uint64_t adv[9], bdv[9];
#pragma unroll
    for (int v = 0; v < 9; ++v) {
      if (c.mode == 0) {
        const uint32_t a0 = a_base + (uint32_t)(v % 3) * c.shift_bytes + (uint32_t)(v / 3) * 4096u;
        adv[v] = make_desc(a0, 2048 + 0, 128, 0, 0);
        bdv[v] = make_desc(b_base + (uint32_t)v * 2u * c.n * 16u, c.n * 16u, 128, 0, 0);
      } else {
        const uint32_t a0 = a_base + (uint32_t)(v % 3) * c.shift_bytes + (uint32_t)((v / 3) % 4) * 32u;
        adv[v] = make_desc(a0, 16, 1024, 2, 0);
        bdv[v] = make_desc(b_base + (uint32_t)((v / 3) % 4) * 32u, 16, 1024, 2, 0);
      }
    }
    long long t0 = clock64();
    for (int r = 0; r < c.reps; r += 9) {
#pragma unroll
      for (int v = 0; v < 9; ++v) { if (elected) umma_bf16(tmem, c.distinct == 1 ? adv[0] : adv[v], c.distinct == 1 ? bdv[0] : bdv[v], idesc, (r + v) > 0); }
    }
    if (elected) umma_commit(smem_u32(&bar));
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && elected) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory"); }
}

// ------------------------------------------------------------------------------------------- (B) SW128 shifted view check
// A band: rows x 64 bf16 (128 B per row) stored with the TMA 128B swizzle relative to a 1024-aligned base.
// D[128][n] = sum_k A[shift + m][k] * B[j][k]
__global__ void __launch_bounds__(128, 1) sw128_check_kernel(const __nv_bfloat16* a_lin, const __nv_bfloat16* b_lin, int band_rows, int n,
                                                             int shift, int use_base_off, float* d_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sa = smem; uint8_t* sb = smem + 64 * 1024;
  for (int i = threadIdx.x; i < band_rows * 8; i += blockDim.x) {   // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(a_lin)[i];
    *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < n * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(b_lin)[i];
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a0 = smem_u32(sa) + (uint32_t)shift * 128u + ks * 32u;
      const uint32_t b0 = smem_u32(sb) + ks * 32u;
      umma_bf16(tmem, make_desc(a0, 16, 1024, 2, use_base_off ? ((a0 >> 7) & 7) : 0), make_desc(b0, 16, 1024, 2, 0), idesc, ks > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < n; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * n + c0 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory"); }
}

// ------------------------------------------------------------------------------------------- (C) TMA probe
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  return (PFN_encodeTiled)ptr;
}
struct TmaCfg { int ndim; int c[4]; int step_dim; int step; int inflight; int loads; uint32_t bytes; };
__global__ void __launch_bounds__(32, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap map, TmaCfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t phase[16] = {0};
    while (waited < c.loads) {
      while (issued < c.loads && issued - waited < c.inflight) {
        const int s = issued % c.inflight;
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, c.bytes);
        int cc[4] = {c.c[0], c.c[1], c.c[2], c.c[3]};
        cc[c.step_dim] += (issued * (int)gridDim.x + (int)blockIdx.x) * c.step;
        const uint32_t dst = smem_u32(smem + (size_t)s * 32768);
        if (c.ndim == 2) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(&map), "r"(bar), "r"(cc[0]), "r"(cc[1]) : "memory");
        else if (c.ndim == 3) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(&map), "r"(bar), "r"(cc[0]), "r"(cc[1]), "r"(cc[2]) : "memory");
        ++issued;
      }
      const int s = waited % c.inflight;
      mbar_wait(smem_u32(&bars[s]), phase[s]);
      phase[s] ^= 1u;
      ++waited;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}


__global__ void __launch_bounds__(32, 1) tma_burst_kernel(const __grid_constant__ CUtensorMap map, TmaCfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long tot = 0;
    uint32_t ph = 0;
    for (int it = 0; it < c.loads; ++it) {
      long long t0 = clock64();
      mbar_expect_tx(smem_u32(&bar), c.bytes * c.inflight);
      for (int k = 0; k < c.inflight; ++k) {
        int cc[4] = {c.c[0], c.c[1], c.c[2], c.c[3]};
        cc[c.step_dim] += ((it * c.inflight + k) * (int)gridDim.x + (int)blockIdx.x) * c.step;
        const uint32_t dst = smem_u32(smem + (size_t)k * c.bytes);
        if (c.ndim == 2) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(&map), "r"(smem_u32(&bar)), "r"(cc[0]), "r"(cc[1]) : "memory");
        else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(&map), "r"(smem_u32(&bar)), "r"(cc[0]), "r"(cc[1]), "r"(cc[2]) : "memory");
      }
      mbar_wait(smem_u32(&bar), ph);
      ph ^= 1u;
      tot += clock64() - t0;
    }
    if (blockIdx.x == 0) out[0] = tot;
  }
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 64));
  long long h_out[2];
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("== (A) tcgen05.mma rate, 148 CTAs, M=128 K=16; cycles per MMA (issue / complete); floor = N/2\n");
  const int ns[] = {16, 32, 64, 128, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 16, 128, 2048})
      for (int n : ns) {
        if (mode == 0 && n > 64) continue;
        if (mode == 1 && (shift == 16 || shift == 2048)) continue;
        RateCfg c{mode, n, shift, 2048 / 9 * 9, 9};
        mma_rate_kernel<<<148, 128, 200 * 1024>>>(c, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h_out, d_out, 16, cudaMemcpyDeviceToHost));
        printf("mode=%s shift=%4dB N=%3d : issue %.1f  complete %.1f cyc/MMA (floor %d)\n", mode ? "sw128" : "noswz", shift, n, h_out[0] / 2043.0, h_out[1] / 2043.0, n / 2);
      }
  // same-address repeats (distinct=1) to separate address effects
  for (int mode = 0; mode < 2; ++mode) {
    RateCfg c{mode, 64, 0, 2048 / 9 * 9, 1};
    mma_rate_kernel<<<148, 128, 200 * 1024>>>(c, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, 16, cudaMemcpyDeviceToHost));
    printf("mode=%s same operands N=64 : issue %.1f complete %.1f\n", mode ? "sw128" : "noswz", h_out[0] / 2043.0, h_out[1] / 2043.0);
  }

  printf("== (B) SW128 shifted A view numeric check\n");
  {
    const int band = 160, n = 64;
    std::vector<__nv_bfloat16> ha(band * 64), hb(n * 64);
    std::vector<float> fa(band * 64), fb(n * 64);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = v; }
    for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(v); fb[i] = v; }
    __nv_bfloat16 *da, *db; float* dd;
    CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dd, 128 * n * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(sw128_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    std::vector<float> hd(128 * n);
    for (int use_bo = 0; use_bo < 2; ++use_bo)
      for (int shift : {0, 1, 2, 3, 5, 8, 9, 17}) {
        CK(cudaMemset(dd, 0, 128 * n * 4));
        sw128_check_kernel<<<1, 128, 100 * 1024>>>(da, db, band, n, shift, use_bo, dd);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d base_off %d: CUDA error %s\n", shift, use_bo, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(hd.data(), dd, 128 * n * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int j = 0; j < n; ++j) {
            float ref = 0;
            for (int k = 0; k < 64; ++k) ref += fa[(shift + m) * 64 + k] * fb[j * 64 + k];
            double err = fabs(ref - hd[m * n + j]);
            if (err > 1e-3) ++bad;
            if (err > maxerr) maxerr = err;
          }
        printf("shift=%2d rows base_off=%s : max err %.4g, bad %d / %d\n", shift, use_bo ? "(addr>>7)&7" : "0", maxerr, bad, 128 * n);
      }
  }

  printf("== (C) TMA loads: cycles per load vs loads in flight (148 CTAs, 64 loads each)\n");
  {
    PFN_encodeTiled enc = get_encode();
    const size_t bytes = (size_t)1 << 30;
    void* g; CK(cudaMalloc(&g, bytes)); CK(cudaMemset(g, 1, bytes));
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(tma_burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Shape { const char* name; int ndim; CUtensorMapDataType dt; cuuint64_t dims[3]; cuuint64_t strides[2]; cuuint32_t box[3]; CUtensorMapSwizzle sw; uint32_t bytes; int step_dim; int step; };
    const cuuint64_t plane = 1300000;  // positions per chunk plane
    Shape shapes[] = {
        {"3D u64 {256x8B, 2 chunks} 4KB (conv1 band)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT64, {plane * 2, 2, 1}, {plane * 16, plane * 32}, {256, 2, 1}, CU_TENSOR_MAP_SWIZZLE_NONE, 4096, 0, 256},
        {"3D u64 {256x8B, 8 chunks} 16KB (64ch band)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT64, {plane * 2, 8, 1}, {plane * 16, plane * 128}, {256, 8, 1}, CU_TENSOR_MAP_SWIZZLE_NONE, 16384, 0, 256},
        {"2D bf16 SW128 {64, 128 rows} 16KB", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, {64, 2500000, 1}, {128, 0}, {64, 128, 1}, CU_TENSOR_MAP_SWIZZLE_128B, 16384, 1, 128},
        {"2D bf16 SW128 {64, 256 rows} 32KB", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, {64, 2500000, 1}, {128, 0}, {64, 256, 1}, CU_TENSOR_MAP_SWIZZLE_128B, 32768, 1, 256},
        {"2D bf16 SW32 {16, 256 rows} 8KB", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, {16, 2500000, 1}, {32, 0}, {16, 256, 1}, CU_TENSOR_MAP_SWIZZLE_32B, 8192, 1, 256},
        {"2D bf16 SW64 {32, 256 rows} 16KB", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, {32, 2500000, 1}, {64, 0}, {32, 256, 1}, CU_TENSOR_MAP_SWIZZLE_64B, 16384, 1, 256},
        {"2D bf16 noswz {8(16B), 256 rows} 4KB", 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, {8, 2500000, 1}, {16, 0}, {8, 256, 1}, CU_TENSOR_MAP_SWIZZLE_NONE, 4096, 1, 256},
    };
    for (auto& s : shapes) {
      CUtensorMap map;
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&map, s.dt, s.ndim, g, s.dims, s.strides, s.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, s.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", s.name, (int)r); continue; }
      for (int inflight : {1, 2, 4, 6}) {
        TmaCfg c{s.ndim, {0, 0, 0, 0}, s.step_dim, s.step, inflight, 64, s.bytes};
        for (int rep = 0; rep < 2; ++rep) {   // rep 0: cold (HBM), rep 1: L2-warm
          tma_probe_kernel<<<148, 32, 200 * 1024>>>(map, c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s: CUDA error %s\n", s.name, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost));
          printf("%-48s inflight=%d %s: %.0f cyc/load  (%.1f B/cyc/SM)\n", s.name, inflight, rep ? "warm" : "cold", h_out[0] / 64.0, s.bytes * 64.0 / h_out[0]);
        }
      }
      for (int nctas : {148, 8})
      for (int burst : {1, 2, 4, 8}) {
        if ((size_t)burst * s.bytes > 190 * 1024) continue;
        TmaCfg c{s.ndim, {0, 0, 0, 0}, s.step_dim, s.step, burst, 16, s.bytes};
        for (int rep = 0; rep < 2; ++rep) {
          tma_burst_kernel<<<nctas, 32, 200 * 1024>>>(map, c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s: CUDA error %s\n", s.name, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost));
          printf("%-48s ctas=%3d burst=%d %s: %.0f cyc/burst  (%.1f B/cyc/SM)\n", s.name, nctas, burst, rep ? "warm" : "cold", h_out[0] / 16.0, s.bytes * burst * 16.0 / h_out[0]);
        }
      }
    }
  }
  return 0;
}
