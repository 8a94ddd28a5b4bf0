// Batched bf16 "NT" GEMM on the 5th-generation tensor cores:  C[b][m][n] = sum_k A[b][m][k] * B[b][n][k]
// (both operands K-major = row-major [rows][K]), fp32 accumulation in TMEM.
//
// It carries the two dense contractions of the contextual-attention module (reference
// models/inpaint_networks.py:347-348 foreground-background similarity, :377-379 patch paste) in bf16 mode:
//   scores  T[f][b]   = sum_k P[f][k] P[b][k] * inv_norm[b]          (M = N = 1024, K = 576, fp32 out, column scale)
//   paste   cols[ck][f] = sum_b Rt[ck][b] A[f][b]                     (M = N = K = 1024, bf16 out)
//
// Kernel structure (persistent, one CTA per SM, 192 threads):
//   warp 0    : TMA producer, 128B-swizzled [rows][64] boxes of A and B into a 6-stage ring
//   warp 1    : TMEM allocator + tcgen05.mma issuer (M = 128, N = BN, K = 16; 4 MMAs per stage), peeks the next
//               stage's mbarrier before issuing so that the wait latency hides under the queued MMAs
//   warps 2-5 : epilogue, one TMEM lane quadrant each; accumulators are ring-buffered in all 512 TMEM columns
#include <cuda_bf16.h>
#include <stdlib.h>
#include "hv_common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace hv {

constexpr int G_BM = 128, G_BK = 64, G_TMEM_COLS = 512;
// operand ring depth: 6 x 32 KB (BN = 128) or 4 x 48 KB (BN = 256).  BN = 256 is the default: an SS-mode MMA re-reads A (4 KB)
// and B (BN x 32 B) from shared memory, TMA writes the same bytes once, and at 128 B/clk of shared-memory bandwidth a 128 x 128
// tile is bound at 50 % of the tensor peak by that traffic, a 128 x 256 tile at 67 %
__host__ __device__ constexpr int gemm_stages(int bn) { return bn == 256 ? 4 : 6; }
constexpr int G_STAGE_PITCH = 36;   // floats per row of the epilogue transpose tile (32 + 4: conflict-free float4 rows)

struct GemmParams {
  CUtensorMap map_a, map_b;
  void* c;                  // fp32 or bf16 [batch][M][N]
  const float* colscale;    // [batch][N] or null
  int M, N, K, batch;
  int tiles_m, tiles_n, total_tiles, kblocks;
  int a_bcast;              // A is shared by every batch entry (batch stride 0): its tensor map has a batch extent of 1
  // tap mode (weight gradient of a stride-1 convolution without an im2col operand, gemm_tc_taps below): a tile is
  // (image, K chunk, tap group); its A rows are (tap, filter): one TMA box of co8 filter rows per tap, each read at its own SHIFT
  // along K (K = positions of the zero-bordered plane; a tap is a constant shift of the flattened position); B = the input planes
  int tap_mode, ntaps, tpm, co8, ntg, nkc, kb0;   // kb0: first K block (the planes start with all-zero border rows)
  int tap_off[25];           // multiples of 8 positions: the innermost TMA coordinate must be 16-byte aligned (an odd element offset
  int tap_rep[25];           // raises "illegal instruction"); the sub-8 part of a shift selects a pre-shifted REPLICA of the A planes
  int nrep;
};

template <int BN, bool OUT_BF16>
__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  constexpr int ACC_STAGES = G_TMEM_COLS / BN > 4 ? 4 : G_TMEM_COLS / BN;   // BN = 32 / 64 (narrow outputs): four stages are plenty
  constexpr int G_STAGES = gemm_stages(BN);
  constexpr uint32_t A_BYTES = G_BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)G_STAGES * STAGE_BYTES);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8u * G_STAGES;
  const uint32_t bar_tfull = bar_empty + 8u * G_STAGES;
  const uint32_t bar_tempty = bar_tfull + 8u * ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 2 * ACC_STAGES);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_b) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < G_STAGES; ++i) { mbar_init(bar_full + 8u * i, 1); mbar_init(bar_empty + 8u * i, 1); }
      for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, 128); }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(G_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();   // barriers / TMEM / tensor-map prefetch above overlap the predecessor's tail
  const int tiles_per_batch = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ===================================================================== TMA producer
    const bool leader = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t ring = smem_u32(smem);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      if (p.tap_mode) {
        const int tg = tile % p.ntg, blk = tile / p.ntg;
        const int kc = blk % p.nkc, img = blk / p.nkc;
        const int t0 = tg * p.tpm, nt_here = min(p.tpm, p.ntaps - t0);
        const uint32_t bytes = (uint32_t)(nt_here * p.co8 + BN) * 128u;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_empty + 8u * slot, phase ^ 1u);
          if (leader) {
            const uint32_t fb = bar_full + 8u * slot, dst = ring + (uint32_t)slot * STAGE_BYTES;
            const int k0 = (p.kb0 + kc * p.kblocks + kb) * G_BK;
            mbar_expect_tx(fb, bytes);
            for (int tl = 0; tl < nt_here; ++tl)   // out-of-range positions (before / after the plane) are zero-filled by TMA
              tma_load_3d(dst + (uint32_t)(tl * p.co8) * 128u, &p.map_a, fb, k0 - p.tap_off[t0 + tl], 0, img * p.nrep + p.tap_rep[t0 + tl]);
            tma_load_3d(dst + A_BYTES, &p.map_b, fb, k0, 0, img);
          }
          if (++slot == G_STAGES) { slot = 0; phase ^= 1u; }
        }
        continue;
      }
      const int b = tile / tiles_per_batch, r = tile - b * tiles_per_batch;
      const int mt = r / p.tiles_n, nt = r - mt * p.tiles_n;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(bar_empty + 8u * slot, phase ^ 1u);
        if (leader) {
          const uint32_t fb = bar_full + 8u * slot, dst = ring + (uint32_t)slot * STAGE_BYTES;
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_3d(dst, &p.map_a, fb, kb * G_BK, mt * G_BM, p.a_bcast ? 0 : b);
          tma_load_3d(dst + A_BYTES, &p.map_b, fb, kb * G_BK, nt * BN, b);
        }
        if (++slot == G_STAGES) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ring_lo = kDescLoSw128 + (smem_u32(smem) >> 4);
    int slot = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    bool ready = mbar_peek(bar_full, 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait_peeked(ready, bar_full + 8u * slot, phase);
        tc_fence_after();
        const uint32_t a_lo = ring_lo + (uint32_t)slot * (STAGE_BYTES >> 4);
        const uint32_t b_lo = a_lo + (A_BYTES >> 4);
        const uint32_t cur_empty = bar_empty + 8u * slot;
        if (++slot == G_STAGES) { slot = 0; phase ^= 1u; }
        ready = mbar_peek(bar_full + 8u * slot, phase);  // next stage (possibly of the next tile)
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < G_BK / 16; ++ks)  // 32 B per K step inside the 128 B swizzle row
            umma_bf16(d_tmem, ((uint64_t)kDescHiSw128 << 32) | (a_lo + 2u * ks), ((uint64_t)kDescHiSw128 << 32) | (b_lo + 2u * ks),
                      idesc, (kb | ks) != 0 ? 1u : 0u);
          umma_commit(cur_empty);
        }
      }
      if (leader) umma_commit(bar_tfull + 8u * acc);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int quad = warp & 3;
    float* stage_base = reinterpret_cast<float*>(smem + (size_t)G_STAGES * STAGE_BYTES + 256);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_batch, r = tile - b * tiles_per_batch;
      const int mt = r / p.tiles_n, nt = r - mt * p.tiles_n;
      const size_t out_row0 = ((size_t)b * p.M + (size_t)mt * G_BM) * p.N + (size_t)nt * BN;   // row 0 of the tile
      const float* cs = p.colscale ? p.colscale + (size_t)b * p.N + (size_t)nt * BN : nullptr;
      mbar_wait(bar_tfull + 8u * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      // 32 x 32 blocks go through a per-warp staging tile so that global stores are row-contiguous: tcgen05.ld hands every lane
      // its own ROW (32 lanes = 32 rows N*4 bytes apart); after the transpose 8 lanes cover 32 consecutive columns of one row
      float* stage = stage_base + (warp - 2) * (32 * G_STAGE_PITCH);
      const int tr = lane >> 3, tc = (lane & 7) * 4;   // read-back: rows tr + 4 i, columns tc .. tc + 3
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld<32>(t_addr + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == BN) {  // all columns of this accumulator stage are in registers
          tc_fence_before();
          mbar_arrive(bar_tempty + 8u * acc);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stage + lane * G_STAGE_PITCH + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (cs) sc = __ldg(reinterpret_cast<const float4*>(cs + c0 + tc));
        const size_t blk = out_row0 + (size_t)(quad * 32) * p.N + c0 + tc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = tr + 4 * i;
          float4 t = *reinterpret_cast<const float4*>(stage + r * G_STAGE_PITCH + tc);
          t.x *= sc.x; t.y *= sc.y; t.z *= sc.z; t.w *= sc.w;
          if (OUT_BF16) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(t.x, t.y), h1 = __floats2bfloat162_rn(t.z, t.w);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.c) + blk + (size_t)r * p.N) = pk;
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.c) + blk + (size_t)r * p.N) = t;
          }
        }
        __syncwarp();
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled gemm_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

static int gemm_map(CUtensorMap* map, const void* base, int rows, int K, int batch, long long batch_stride_elems, int box_rows,
                    long long row_stride_elems = 0) {
  PFN_encodeTiled enc = gemm_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return HV_ERR_CUDA; }
  if (row_stride_elems == 0) row_stride_elems = K;
  if (batch_stride_elems == 0) { batch = 1; batch_stride_elems = (long long)rows * K; }   // broadcast operand
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 2, (cuuint64_t)batch_stride_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (gemm operand) failed with CUresult %d", (int)r); return HV_ERR_CUDA; }
  return HV_OK;
}

template <int BN, bool OUT_BF16>
static int gemm_launch(const GemmParams& p, int grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)gemm_stages(BN) * (G_BM * 128 + BN * 128) + 256 + 4 * 32 * G_STAGE_PITCH * sizeof(float);
  static bool configured = false;
  if (!configured) {
    HV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, OUT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  HV_CUDA(launch_pdl(gemm_tc_kernel<BN, OUT_BF16>, dim3(grid), dim3(192), smem, st, p));
  HV_LAUNCH_CHECK();
  return HV_OK;
}

int gemm_tc_nt(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, const float* colscale, int M, int N, int K, int batch,
               long long strideA, long long strideB, int out_bf16, cudaStream_t st) {
  HV_CHECK_ARG(A && B && C, "gemm_tc: null argument");
  HV_CHECK_ARG(M % G_BM == 0 && (N % 128 == 0 || N == 32 || N == 64) && K % G_BK == 0 && batch >= 1,
               "gemm_tc: M %% 128, N %% 128 (or N = 32 / 64), K %% 64 must be 0 (got %d,%d,%d)", M, N, K);
  // the attention module calls with the same workspace operands every forward: keep the encoded tensor maps of the last few
  // distinct problems (encoding costs several microseconds of host time per map)
  struct Cached { const void *a, *b; int M, N, K, batch, bn; long long sa, sb; CUtensorMap ma, mb; };
  static thread_local Cached cache[4];
  static thread_local int next = 0;
  static const bool bn128 = getenv("HV_GEMM_BN128") != nullptr;
  GemmParams p;
  const int bn_pick = N < 128 ? N : ((N % 256 == 0 && out_bf16 && !bn128) ? 256 : 128);   // N = 32 / 64: one narrow tile column
  const Cached* hit = nullptr;
  for (const Cached& e : cache)
    if (e.a == A && e.b == B && e.M == M && e.N == N && e.K == K && e.batch == batch && e.bn == bn_pick && e.sa == strideA && e.sb == strideB) hit = &e;
  int rc = HV_OK;
  if (hit) { p.map_a = hit->ma; p.map_b = hit->mb; }
  else rc = gemm_map(&p.map_a, A, M, K, batch, strideA, G_BM);
  if (rc) return rc;
  // measured on B200 (batch 16, M = N = 1024): fp32 output K = 576: 28.5 us with BN = 128 vs 31.5 us with BN = 256 (512 tiles are
  // only 3.5 waves); bf16 output K = 1024: 34.6 us vs 31.9 us
  const int bn = bn_pick;
  if (!hit) {
    rc = gemm_map(&p.map_b, B, N, K, batch, strideB, bn);
    if (rc) return rc;
    Cached& e = cache[next];
    next = (next + 1) % 4;
    e = Cached{A, B, M, N, K, batch, bn, strideA, strideB, p.map_a, p.map_b};
  }
  p.tap_mode = 0;
  p.c = C; p.colscale = colscale; p.M = M; p.N = N; p.K = K; p.batch = batch; p.a_bcast = strideA == 0 ? 1 : 0;
  p.tiles_m = M / G_BM; p.tiles_n = N / bn; p.total_tiles = p.tiles_m * p.tiles_n * batch; p.kblocks = K / G_BK;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (bn == 256) return out_bf16 ? gemm_launch<256, true>(p, grid, st) : gemm_launch<256, false>(p, grid, st);
  if (bn == 64) { HV_CHECK_ARG(!out_bf16, "gemm_tc: narrow tiles write fp32"); return gemm_launch<64, false>(p, grid, st); }
  if (bn == 32) { HV_CHECK_ARG(!out_bf16, "gemm_tc: narrow tiles write fp32"); return gemm_launch<32, false>(p, grid, st); }
  return out_bf16 ? gemm_launch<128, true>(p, grid, st) : gemm_launch<128, false>(p, grid, st);
}

// Weight gradient of a stride-1 convolution as shifted-operand GEMMs (no im2col):
//   part[(img, kc, tg)][tl * co8 + co][ci] = sum over the positions k of K chunk kc of  DY[img][rep[t]][co][k - off[t]] * X[img][ci][k],   t = tg * tpm + tl
// DY: bf16 [n][nrep][cout][plane], X: bf16 [n][cin][plane], zero-bordered planes of `plane` positions (a multiple of 8); off[t] are
// multiples of 8; part: fp32 [n * nkc * ntg][128][bn].  Rows of filters >= cout and of inputs >= cin are zero-filled by TMA (the boxes
// are taller than the tensors).
int gemm_tc_taps(const __nv_bfloat16* DY, const __nv_bfloat16* X, float* part, int n, int cout, int cin, int plane, const int* tap_off,
                 const int* tap_rep, int nrep, int ntaps, int tpm, int co8, int bn, int nkc, int kb0, int kblocks, cudaStream_t st) {
  HV_CHECK_ARG(DY && X && part && tap_off && tap_rep && nrep >= 1, "gemm_tc_taps: null argument");
  for (int i = 0; i < ntaps && i < 25; ++i)
    HV_CHECK_ARG((tap_off[i] & 7) == 0 && tap_rep[i] >= 0 && tap_rep[i] < nrep, "gemm_tc_taps: tap %d: offset %d / replica %d", i, tap_off[i], tap_rep[i]);
  HV_CHECK_ARG(ntaps >= 1 && ntaps <= 25 && tpm >= 1 && tpm * co8 <= G_BM && (co8 & 7) == 0 && (plane & 7) == 0 && cout <= co8,
               "gemm_tc_taps: bad tap tiling (taps %d, per tile %d, filter rows %d, plane %d)", ntaps, tpm, co8, plane);
  HV_CHECK_ARG(bn == 32 || bn == 64 || bn == 128 || bn == 256, "gemm_tc_taps: bn = %d", bn);
  GemmParams p;
  int rc = gemm_map(&p.map_a, DY, cout, plane, n * nrep, (long long)cout * plane, co8);
  if (rc) return rc;
  rc = gemm_map(&p.map_b, X, cin, plane, n, (long long)cin * plane, bn);
  if (rc) return rc;
  p.tap_mode = 1; p.ntaps = ntaps; p.tpm = tpm; p.co8 = co8; p.ntg = (ntaps + tpm - 1) / tpm; p.nkc = nkc; p.kb0 = kb0;
  for (int i = 0; i < 25; ++i) p.tap_off[i] = i < ntaps ? tap_off[i] : 0;
  for (int i = 0; i < 25; ++i) p.tap_rep[i] = i < ntaps ? tap_rep[i] : 0;
  p.nrep = nrep;
  p.c = part; p.colscale = nullptr; p.M = G_BM; p.N = bn; p.K = kblocks * G_BK; p.a_bcast = 0;
  p.tiles_m = p.tiles_n = 1; p.batch = p.total_tiles = n * nkc * p.ntg; p.kblocks = kblocks;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  switch (bn) {
    case 32: return gemm_launch<32, false>(p, grid, st);
    case 64: return gemm_launch<64, false>(p, grid, st);
    case 128: return gemm_launch<128, false>(p, grid, st);
    default: return gemm_launch<256, false>(p, grid, st);
  }
}

}  // namespace hv
