// Dataflow kernel for CHAINS of 64 -> 64 3x3 convolution layers on the 64x64 trunk of the generator
// (coarse conv5 .. conv12, fine conv6 .. conv10_atrous, pmconv5/6, pmconv9/10, allconv12/19: reference
// models/inpaint_networks.py:45-54, :133-137, :148-149, :152-158): ONE persistent launch runs every layer of a chain.
//
// Why: at batch 16 a 64x64 layer is only 3.5 tiles per SM; as one launch per layer, 56 % of the forward was per-launch
// ramp (prologue, first band, last epilogue, launch gap).  Here the (layer, tile) items of the whole chain form ONE queue:
//   * a CTA pulls the next item with an atomic counter (dynamic: correctness never depends on which CTAs are resident, so
//     two chains on two streams may share the GPU);
//   * a tile of layer l reads image i of layer l-1's output: the producer spins on a per-(layer, image) completion counter
//     that the epilogue warps bump after their stores (release / acquire at GPU scope, then fence.proxy.async before the
//     TMA load): images complete in queue order, so in the steady state the counter is already full and layers overlap -
//     no grid barrier, no drain between layers;
//   * when a CTA crosses into the next layer its MMA warp waits for its own in-flight MMAs (tcgen05.commit -> mbarrier),
//     reloads the 72 KB filter bank with one bulk copy and goes on: ~1 us per CTA per layer instead of a ~4 us ramp.
// The tile pipeline is that of conv_tc_kernel<64, ELU, 0x11334> (csrc/conv_tc.cu): one 4-D TMA box brings the three
// kernel-row bands of a tile, 36 UMMAs (M = 128 positions, N = 64 filters, K = 16) read them through shifted descriptors,
// 16 epilogue warps drain 4 TMEM accumulator stages.
#include <stdlib.h>
#include <string.h>
#include "hv_common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace hv {

constexpr int TR_MAX_LAYERS = 8;
constexpr int TR_SLOTS = 3;                       // band ring: 3 x 48 KB
constexpr int TR_ACC = 4;                         // TMEM accumulator stages (4 x 64 columns)
constexpr int TR_RING = 16;                       // item-id ring (producer -> issuer / epilogue)
constexpr int TR_EPI_WARPS = 16;
constexpr int TR_THREADS = 32 * (TR_EPI_WARPS + 2);
constexpr uint32_t TR_W_BYTES = 9u * 8u * 64u * 16u;          // 73,728: [tap][chunk][64 filters][8 ch] bf16
constexpr uint32_t TR_BAND_BYTES = 3u * 8u * TC_TILE_M * 16u;  // 49,152: [chunk][row][128 positions][8 ch] bf16
constexpr uint32_t TR_SMEM = TR_W_BYTES + TR_SLOTS * TR_BAND_BYTES + 1024 /* over-read pad */ + 1024 /* barriers, ring */ +
                             TR_MAX_LAYERS * 64 * 4 /* bias */;

struct TrunkLayer {
  const void* w_packed;
  const float* bias;
  __nv_bfloat16* out;
  int in_pitch, in_border, q_first, tile_adv, tiles_per_image, dil;
  unsigned long long pitch_magic, tiles_magic;
  int out_pitch, out_border, out_plane, out_chunks_total;
  int up2, relu;
  int item_base;     // first queue item of this layer
  int dep;           // layer of the chain whose output this layer reads, or -1 (input produced before the launch)
  int dep_target;    // completion count of one image of layer `dep` (tiles_per_image * epilogue warps)
  int signal;        // a later layer of the chain reads this layer's output
};

struct TrunkParams {
  CUtensorMap maps[TR_MAX_LAYERS];
  TrunkLayer layers[TR_MAX_LAYERS];
  int nlayers, total_items, n_images;
  int* counters;     // [0] queue head, [1] exit ticket, [2 + l * n_images + i] completion of image i of layer l
  int debug;         // HV_TRUNK_DEBUG ablation bits (results are WRONG with any of them): 1 no dependency wait, 2 no proxy fence,
                     // 4 no completion signal (needs 1), 8 no filter-bank reload after the first
  long long* trace;  // per-role clock stamps of CTA `trace_cta` (hv_debug_trunk_trace), or null
  int trace_cta;
};

// trace record: (tag, item, clock64); tags 1x producer, 2x issuer, 3x epilogue warp 0
__device__ __forceinline__ void tr_ev(long long* tr, int& n, int tag, int item) {
  if (tr && n < 1300) { tr[3 * n] = tag; tr[3 * n + 1] = item; tr[3 * n + 2] = clock64(); ++n; }
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float exp_neg_fast_tr(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

__global__ void __launch_bounds__(TR_THREADS, 1) trunk_tc_kernel(const __grid_constant__ TrunkParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_PRODUCER = TR_EPI_WARPS, W_MMA = TR_EPI_WARPS + 1;
  uint8_t* s_slots = smem + TR_W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_slots + TR_SLOTS * TR_BAND_BYTES + 1024);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8u * TR_SLOTS;
  const uint32_t bar_w = bar_empty + 8u * TR_SLOTS;        // filter bank of the current layer has landed
  const uint32_t bar_wfree = bar_w + 8u;                   // every MMA that read the previous filter bank has completed
  const uint32_t bar_tfull = bar_wfree + 8u;
  const uint32_t bar_tempty = bar_tfull + 8u * TR_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TR_SLOTS + 2 + 2 * TR_ACC);
  volatile int* s_item = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [TR_RING]
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);                // [nlayers][64]

  if (warp == W_PRODUCER && lane == 0)
    for (int l = 0; l < p.nlayers; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.maps[l]) : "memory");
  if (warp == W_MMA) {
    if (lane == 0) {
      for (int i = 0; i < TR_SLOTS; ++i) { mbar_init(bar_full + 8u * i, 1); mbar_init(bar_empty + 8u * i, 1); }
      mbar_init(bar_w, 1);
      mbar_init(bar_wfree, 1);
      for (int i = 0; i < TR_ACC; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, 32 * TR_EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.nlayers * 64; i += TR_THREADS) s_bias[i] = __ldg(p.layers[i >> 6].bias + (i & 63));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == W_PRODUCER) {
    // ===================================================================== work queue + TMA producer (one lane)
    if (lane == 0) {
      // the queue / completion counters were reset by the previous launch on them, the activations come from the predecessor
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int slot = 0, round = 0, ntr = 0;
      uint32_t phase = 0;
      const uint32_t slots_base = smem_u32(s_slots);
      long long* tr = (p.trace && (int)blockIdx.x == p.trace_cta) ? p.trace : nullptr;
      // the queue pull of the NEXT round is issued before this round's work: its L2 round trip (~0.4 us) stays off the
      // producer's critical path (one thread runs this loop; a tile lasts ~1.3 us)
      const bool static_q = (p.debug & 32) != 0;   // ablation: static round-robin assignment instead of the atomic queue
      int item = static_q ? (int)blockIdx.x : atomicAdd(&p.counters[0], 1);
      while (true) {
        const int next = static_q ? item + (int)gridDim.x : (item < p.total_items ? atomicAdd(&p.counters[0], 1) : item);
        tr_ev(tr, ntr, 10, item);
        mbar_wait(bar_empty + 8u * slot, phase ^ 1u);
        tr_ev(tr, ntr, 11, item);
        const uint32_t fb = bar_full + 8u * slot;
        if (item >= p.total_items) {          // end of the queue: pass the sentinel down the pipeline
          s_item[round & (TR_RING - 1)] = -1;
          mbar_arrive(fb);
          break;
        }
        int l = 0;
#pragma unroll 1
        while (l + 1 < p.nlayers && item >= p.layers[l + 1].item_base) ++l;
        const TrunkLayer& L = p.layers[l];
        const int tile = item - L.item_base;
        const int img = (int)(((unsigned long long)tile * L.tiles_magic) >> 40);
        s_item[round & (TR_RING - 1)] = item;
        mbar_expect_tx(fb, TR_BAND_BYTES);    // arms the barrier (release: the item id above is visible to whoever sees the phase complete)
        if (L.dep >= 0 && !(p.debug & 1)) {   // image `img` of the producing layer must be complete (all its tiles stored)
          const int* flag = p.counters + 2 + L.dep * p.n_images + img;
          if (ld_acquire_gpu(flag) < L.dep_target) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flag) < L.dep_target) {
              __nanosleep(32);
              if (clock64() - t0 > 4000000000ll) __trap();   // a protocol bug fails the launch instead of hanging the GPU
            }
          }
          tr_ev(tr, ntr, 12, item);
          // generic-proxy stores of other SMs (observed through the acquire above) -> this thread's async-proxy (TMA) reads of global memory
          if (!(p.debug & 2)) asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        const int c_tile = 2 * ((tile - img * L.tiles_per_image) * L.tile_adv + L.q_first);   // tensor-map inner unit = 8 B
        const int rel2 = 2 * (-L.dil * L.in_pitch - L.dil);
        tma_load_4d(slots_base + (uint32_t)slot * TR_BAND_BYTES, &p.maps[l], fb, c_tile + rel2, 0, 0, img);
        tr_ev(tr, ntr, 13, item);
        ++round;
        if (++slot == TR_SLOTS) { slot = 0; phase ^= 1u; }
        item = next;
      }
    }
  } else if (warp == W_MMA) {
    // ===================================================================== MMA issuer (warp-uniform, one lane issues)
    const bool leader = elect_one();
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
    constexpr uint32_t a_lbo = 3u * TC_TILE_M;                 // K chunks of the band image are 3 rows apart (16 B units)
    const uint32_t b_lo_base = ((uint32_t)64 << 16) + (smem_u32(smem) >> 4);
    const uint32_t a_lo_base = (smem_u32(s_slots) >> 4) + (a_lbo << 16);
    int slot = 0, acc = 0, round = 0, cur_layer = -1;
    uint32_t phase = 0, acc_phase = 0, w_phase = 0, wfree_phase = 0;
    uint32_t a_step = 1;
    int ntr = 0;
    long long* tr = (p.trace && (int)blockIdx.x == p.trace_cta && leader) ? p.trace + 4000 : nullptr;
    while (true) {
      mbar_wait(bar_full + 8u * slot, phase);
      tc_fence_after();
      const int item = s_item[round & (TR_RING - 1)];
      tr_ev(tr, ntr, 20, item);
      mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u);
      tc_fence_after();
      tr_ev(tr, ntr, 21, item);
      if (item < 0) {                         // sentinel: wake the epilogue with it and stop
        if (leader) mbar_arrive(bar_tfull + 8u * acc);
        break;
      }
      int l = 0;
#pragma unroll 1
      while (l + 1 < p.nlayers && item >= p.layers[l + 1].item_base) ++l;
      if (l != cur_layer && !((p.debug & 8) && cur_layer >= 0)) {
        // layer switch of this CTA: wait for the MMAs still reading the old filter bank, then bring the new one (72 KB bulk copy)
        if (cur_layer >= 0) {
          if (leader) umma_commit(bar_wfree);
          mbar_wait(bar_wfree, wfree_phase);
          wfree_phase ^= 1u;
        }
        if (leader) {
          mbar_expect_tx(bar_w, TR_W_BYTES);
          bulk_load(smem_u32(smem), p.layers[l].w_packed, TR_W_BYTES, bar_w);
        }
        mbar_wait(bar_w, w_phase);
        tc_fence_after();
        w_phase ^= 1u;
        cur_layer = l;
        a_step = (uint32_t)p.layers[l].dil;
        tr_ev(tr, ntr, 22, item);
      }
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
      const uint32_t a_row = a_lo_base + (uint32_t)slot * (TR_BAND_BYTES >> 4);
      if (leader) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t a_lo = a_row + (uint32_t)r * TC_TILE_M + (uint32_t)i * a_step + (uint32_t)ks * 2u * a_lbo;
              const uint32_t b_lo = b_lo_base + (uint32_t)(r * 3 + i) * 512u + (uint32_t)ks * 2u * 64u;
              umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, (r | i | ks) == 0 ? 0u : 1u);
            }
        umma_commit(bar_empty + 8u * slot);   // the band slot is free once these MMAs have read it
        umma_commit(bar_tfull + 8u * acc);    // accumulator complete -> epilogue
      }
      tr_ev(tr, ntr, 23, item);
      __syncwarp();
      ++round;
      if (++slot == TR_SLOTS) { slot = 0; phase ^= 1u; }
      if (++acc == TR_ACC) { acc = 0; acc_phase ^= 1u; }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (warps 0 .. 15: lane quadrant x 16-column slice)
    const int quad = warp & 3;
    const int col0 = (warp >> 2) * 16;
    const int m = quad * 32 + lane;
    int acc = 0, round = 0, ntr = 0;
    uint32_t acc_phase = 0;
    long long* tr = (p.trace && (int)blockIdx.x == p.trace_cta && threadIdx.x == 0) ? p.trace + 8000 : nullptr;
    // Completion signal of a tile = fence (the warp's stores are performed at GPU scope) + one RED on the image's counter.  The fence
    // stalls the warp until its outstanding stores are acknowledged (~1 us right after they were issued), which would double the
    // epilogue's time per tile; it is therefore issued one round LATER, just before the next tile's stores (by then the old stores
    // have landed and the fence is cheap) - or at once when the warp would otherwise sit idle waiting for the next accumulator
    // (which also makes the delay deadlock-free: a warp never blocks with an unpublished tile).
    int* pend = nullptr;
    auto flush = [&]() {
      if (pend) {
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(pend, 1);
        pend = nullptr;
      }
    };
    while (true) {
      if (pend && !__all_sync(0xffffffffu, mbar_peek(bar_tfull + 8u * acc, acc_phase))) flush();
      mbar_wait(bar_tfull + 8u * acc, acc_phase);
      tc_fence_after();
      const int item = s_item[round & (TR_RING - 1)];
      if (item < 0) break;
      tr_ev(tr, ntr, 30, item);
      int l = 0;
#pragma unroll 1
      while (l + 1 < p.nlayers && item >= p.layers[l + 1].item_base) ++l;
      const TrunkLayer& L = p.layers[l];
      const int tile = item - L.item_base;
      const int img = (int)(((unsigned long long)tile * L.tiles_magic) >> 40);
      const int q = (tile - img * L.tiles_per_image) * L.tile_adv + m + L.q_first;
      const int qrow = (int)(((unsigned long long)q * L.pitch_magic) >> 40);
      const int yy = qrow - L.in_border, xx = q - qrow * L.in_pitch - L.in_border;
      const bool valid = m < L.tile_adv && yy >= 0 && yy < 64 && xx >= 0 && xx < 64;
      float v[16];
      tmem_ld<16>(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 64) + (uint32_t)col0, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_tempty + 8u * acc);     // this warp's slice is in registers
      uint4 val[2];
      {
        const float* bias = s_bias + l * 64 + col0;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f0 = v[jj * 8 + 2 * e] + bias[jj * 8 + 2 * e], f1 = v[jj * 8 + 2 * e + 1] + bias[jj * 8 + 2 * e + 1];
            if (L.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
            else { f0 = f0 > 0.f ? f0 : exp_neg_fast_tr(f0) - 1.f; f1 = f1 > 0.f ? f1 : exp_neg_fast_tr(f1) - 1.f; }
            __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          val[jj] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      flush();                                 // the previous tile of this warp (its stores were issued a round ago)
      if (valid) {
        __nv_bfloat16* out_img = L.out + (size_t)img * L.out_chunks_total * L.out_plane * 8;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          __nv_bfloat16* plane = out_img + (size_t)((col0 >> 3) + jj) * L.out_plane * 8;
          if (L.up2) {   // nearest x2 upsample fused into the store (inpaint_networks.py:97, :219)
            const size_t o = (size_t)(2 * yy + L.out_border) * L.out_pitch + 2 * xx + L.out_border;
            *reinterpret_cast<uint4*>(plane + o * 8) = val[jj];
            *reinterpret_cast<uint4*>(plane + (o + 1) * 8) = val[jj];
            *reinterpret_cast<uint4*>(plane + (o + L.out_pitch) * 8) = val[jj];
            *reinterpret_cast<uint4*>(plane + (o + L.out_pitch + 1) * 8) = val[jj];
          } else {
            *reinterpret_cast<uint4*>(plane + ((size_t)(yy + L.out_border) * L.out_pitch + xx + L.out_border) * 8) = val[jj];
          }
        }
      }
      tr_ev(tr, ntr, 31, item);
      if (L.signal && !(p.debug & 4)) pend = p.counters + 2 + l * p.n_images + img;
      if (p.debug & 16) flush();               // ablation: publish at once (the fence then waits for the stores just issued)
      tr_ev(tr, ntr, 32, item);
      ++round;
      if (++acc == TR_ACC) { acc = 0; acc_phase ^= 1u; }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
  // the last CTA out resets the queue and the completion counters for the next launch on them
  if (threadIdx.x == 0) {
    __threadfence();
    const int t = atomicAdd(&p.counters[1], 1);
    if (t == (int)gridDim.x - 1) {
      const int total = 2 + p.nlayers * p.n_images;
      for (int i = 0; i < total; ++i) p.counters[i] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------- host
static long long* g_trunk_trace = nullptr;
static int g_trunk_trace_cta = 0;
void tc_trunk_set_trace(long long* dev_buf, int cta) { g_trunk_trace = dev_buf; g_trunk_trace_cta = cta; }

bool tc_trunk_eligible(const TcConv& c) {
  const TcParams& p = c.p;
  return c.n_pad == 64 && c.k == 3 && c.stride == 1 && c.nsrc == 1 && !c.src[0].kxpack && c.src[0].buf.chunks == 8 && c.src[0].buf.xp == 1 &&
         !c.src[0].buf.s2d && c.src[0].buf.h == 64 && c.src[0].buf.w == 64 && p.nseg == 1 && p.segs[0].nrows == 3 && p.segs[0].ntaps == 3 &&
         p.w_bytes == TR_W_BYTES && p.slot_bytes == TR_BAND_BYTES && p.out_xp == 1 && p.pair == 0 && p.out_chunk_off == 0 && p.out_nchunks == 8 &&
         (p.out_mode == TC_OUT_CHUNKED || p.out_mode == TC_OUT_CHUNKED_UP2) && (p.act == HV_ACT_ELU || p.act == HV_ACT_RELU) &&
         c.dil * 2 + 64 <= 128;
}

size_t tc_trunk_counter_ints(int max_images) { return 2 + (size_t)TR_MAX_LAYERS * max_images; }

// convs[0 .. count): consecutive layers, layer i + 1 reading layer i's output buffer.  counters: tc_trunk_counter_ints(n_images)
// zero-initialised ints private to this chain (left zero by every launch).
int tc_trunk_launch(const TcConv* const* convs, int count, int n_images, int* counters, cudaStream_t st) {
  HV_CHECK_ARG(convs && count >= 1 && count <= TR_MAX_LAYERS && counters && n_images >= 1, "tc_trunk_launch: bad argument");
  TrunkParams q;
  memset(&q, 0, sizeof(q));
  int base = 0;
  for (int l = 0; l < count; ++l) {
    const TcConv& c = *convs[l];
    HV_CHECK_ARG(tc_trunk_eligible(c), "tc_trunk_launch: layer %d of the chain is not a 64 -> 64 3x3 trunk layer", l);
    const TcParams& p = c.p;
    TrunkLayer& L = q.layers[l];
    q.maps[l] = p.maps[0];
    L.w_packed = p.w_packed; L.bias = p.bias; L.out = p.out;
    L.in_pitch = p.in_pitch; L.in_border = p.in_border; L.q_first = p.q_first; L.tile_adv = p.tile_adv;
    L.tiles_per_image = p.tiles_per_image; L.dil = c.dil; L.pitch_magic = p.pitch_magic; L.tiles_magic = p.tiles_magic;
    L.out_pitch = p.out_pitch; L.out_border = p.out_border; L.out_plane = p.out_plane; L.out_chunks_total = p.out_chunks_total;
    L.up2 = p.out_mode == TC_OUT_CHUNKED_UP2; L.relu = p.act == HV_ACT_RELU;
    L.item_base = base;
    base += p.tiles_per_image * n_images;
    L.dep = l > 0 ? l - 1 : -1;
    L.dep_target = l > 0 ? convs[l - 1]->p.tiles_per_image * TR_EPI_WARPS : 0;
    L.signal = l + 1 < count;
    if (l > 0) {
      HV_CHECK_ARG(convs[l - 1]->p.out == c.src[0].buf.ptr && convs[l - 1]->p.out_mode == TC_OUT_CHUNKED,
                   "tc_trunk_launch: layer %d does not read layer %d's output", l, l - 1);
    }
  }
  q.nlayers = count; q.total_items = base; q.n_images = n_images; q.counters = counters;
  static const int debug = getenv("HV_TRUNK_DEBUG") ? atoi(getenv("HV_TRUNK_DEBUG")) : 0;
  q.debug = debug; q.trace = g_trunk_trace; q.trace_cta = g_trunk_trace_cta;
  static bool configured = false;
  if (!configured) {
    HV_CUDA(cudaFuncSetAttribute(trunk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR_SMEM));
    configured = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  static const bool pdl = getenv("HV_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(base < sms ? base : sms);
  cfg.blockDim = dim3(TR_THREADS);
  cfg.dynamicSmemBytes = TR_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  HV_CUDA(cudaLaunchKernelEx(&cfg, trunk_tc_kernel, q));
  HV_LAUNCH_CHECK();
  return HV_OK;
}

}  // namespace hv

// debug hook (not part of the drop-in surface): dev_buf = 12000 int64 on the device (3 roles x 1300 records of (tag, item, clock)),
// or NULL to switch tracing off; the CTA with index `cta` of every subsequent trunk launch records
extern "C" int hv_debug_trunk_trace(void* dev_buf, int cta) {
  hv::tc_trunk_set_trace(reinterpret_cast<long long*>(dev_buf), cta);
  return HV_OK;
}
