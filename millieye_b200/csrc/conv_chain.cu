// A run of consecutive conv layers as ONE persistent kernel of CTA pairs with tile-level dependencies.
//
// Why (profiles/round1/trace_r1j.txt): launched one kernel per layer, the 75 convolutions of
// Darknet-53 at batch 32 spent ~6.7 us per layer outside any CTA's lifetime plus ~4 us inside it on setup, pipeline
// fill and the exposed last epilogue - a third of the conv time - and the 13^2 / 26^2 layers lost another 20-37 %
// to wave quantisation (88 or 170 tiles on 74 SM pairs).  Here the 74 CTA pairs stay resident for the whole run of
// layers: every pair walks its own list of (layer, tile) work items; a tile of layer l starts as soon as the
// tiles of layer l-1 that cover its input rows (and its residual rows) have been written, which it learns from
// per-m-tile completion counters in global memory (release/acquire at gpu scope around the TMA stores / loads).
// No grid-wide barrier, no relaunch, no TMEM / mbarrier / tensor-map setup between layers, and the next layer's
// tiles fill the SM pairs a partial last wave used to leave idle.
//
// Tiles are the 256 x BN (BN = 128 | 256) cta_group::2 tiles of conv_gemm_pair.cu, same operand path (im2col-mode
// TMA for 3x3, tiled TMA for 1x1, 128B swizzle, 5 stages of 32 KB) and the same fused epilogue (bias + LeakyReLU
// (+ residual) -> fp16 -> swizzled staging -> TMA store).  Per CTA:
//   warp 0      TMA producer; before a tile's first A load it waits for the producer layer's counters
//   warp 1      (leader CTA) tcgen05.mma.cta_group::2 issuer, accumulators double-buffered in TMEM
//   warps 2-9   epilogue: tcgen05.ld -> math -> staging tile
//   warps 10-11 store warps, alternating tiles: TMA store, hand the staging tile to the next tile (residual
//               prefetch), then wait for the store to COMPLETE and publish the tile's counter - off the
//               epilogue's critical path.
// Deadlock freedom: work lists are cut from one global order (layer-major); every pair takes its items in that
// order and a tile only waits for tiles that precede it in the order; all 74 pairs are co-resident (1 CTA / SM).
//
// Replaces, for the layers it covers, the per-layer launches of Darknet.forward's module loop
// (reference yolov3/models.py:247-262).
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <vector>

namespace me {

unsigned long long* conv_debug_word();   // conv_gemm.cu
unsigned long long* conv_trace_buffer();
int conv_ensure_debug_word();
bool conv_pdl_enabled();

namespace {

constexpr int kBM = 128;            // rows per CTA (256 per pair)
constexpr int kBK = 64;
constexpr int kStoreWarps = 2;
constexpr int kEpiThreads = 256;
constexpr int kThreads = 64 + kEpiThreads + 32 * kStoreWarps;
constexpr uint32_t kEpiBarrierId = 1;
constexpr int kMaxStages = 8;
constexpr int kABytes = kBM * kBK * 2;           // 16 KB
constexpr int kBBytesMax = 128 * kBK * 2;        // 16 KB (BN = 256: 128 weight rows per CTA)
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kSubBytes = kBM * 128;             // 128 rows x 64 fp16 columns
constexpr int kStagingBytes = 4 * kSubBytes;     // BN = 256
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                  // TMEM columns between the two accumulators
constexpr int kTailBytes = 256 * 4 + 32 * 8 + 16;
constexpr int kItemShift = 20;                   // work item = layer << 20 | tile

constexpr int smem_bytes(int stages) { return 1024 + stages * kStageBytes + kStagingBytes + kTailBytes; }

// One layer of the chain, in global memory (tensor maps need 64-byte alignment).
struct alignas(128) ChainLayer {
  CUtensorMap tmA, tmB, tmC, tmR;
  const float* bias;
  int M, Ho, Wo, stride, pad;
  int kb_per_tap, num_kb, tiles_n, bn;
  int act, has_res, im2col;
  int out_f32;       // 1: fp32 output (YOLO head logits): BN = 128, four 32-column staging sub-tiles
  int dep_kind;      // -1: input complete before the launch; 0: same rows (1x1); 1: 3x3 stride 1; 2: 3x3 stride 2
  int dep_base;      // first counter of the layer that produces the input
  int dep_target;    // arrivals that complete one of its m tiles (2 CTAs x its n tiles)
  int dep_M;         // its row count
  int Hin, Win;      // its spatial size
  int res_base;      // first counter of the layer that produces the residual, -1: none / complete before the launch
  int res_target;
  int ctr_base;      // first counter of this layer (one per 256-row m tile)
};

struct ChainHeader {
  unsigned long long magic;
  int n_layers, npairs, work_stride, stages, smem, n_counters;
  long long layers_off, work_off, counters_off, total_bytes;
};
constexpr unsigned long long kMagic = 0x4d45434841494e31ull;  // "MECHAIN1"
constexpr unsigned long long kMagicPlan = 0x4d45434841494e50ull;  // "MECHAINP": schedule only, no tensor maps
constexpr size_t kHeaderBytes = 256;

struct ChainParams {
  const ChainLayer* layers;
  const int* work;
  int work_stride;
  int stages;
  int* counters;
  unsigned long long* debug;
  unsigned long long* trace;   // me_conv_set_trace: 16 words per CTA (tools/chain_trace.py), else nullptr
};

// wait with the cycles spent in it added to `acc` when tracing
#define ME_CHAIN_TRACED(acc, stmt)        \
  do {                                    \
    if (tr) {                             \
      const long long t_ = clock64();     \
      stmt;                               \
      (acc) += clock64() - t_;            \
    } else {                              \
      stmt;                               \
    }                                     \
  } while (0)

__device__ __forceinline__ void watchdog_trap(unsigned long long* dbg, uint32_t tag, uint32_t aux) {
  if (dbg) {
    *reinterpret_cast<volatile unsigned long long*>(dbg) =
        (static_cast<unsigned long long>(tag | 0xC000u) << 32) | (static_cast<unsigned long long>(blockIdx.x) << 8) | (aux & 0x7fu) | 0x80u;
    __threadfence_system();
  }
  __trap();
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long* dbg, uint32_t tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) watchdog_trap(dbg, tag, parity);
  }
}

// Blocks until *ctr >= target (acquire, gpu scope), then orders the async proxy (TMA loads) after it.
__device__ __forceinline__ void wait_counter(const int* ctr, int target, unsigned long long* dbg, uint32_t tag) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  if (v < target) {
    const long long t0 = clock64();
    uint32_t spins = 0;
    do {
      __nanosleep(32);
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((++spins & 255u) == 0 && clock64() - t0 > 4000000000LL) watchdog_trap(dbg, tag, static_cast<uint32_t>(v));
    } while (v < target);
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void publish_counter(int* ctr) {
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
}

// Rows [lo, hi] of the producer layer that the consumer rows [m0, m1] read.
__host__ __device__ __forceinline__ void dep_rows(const ChainLayer& L, int m0, int m1, int* lo, int* hi) {
  if (L.dep_kind == 0) {
    *lo = m0;
    *hi = m1;
  } else if (L.dep_kind == 1) {
    *lo = m0 - L.Win - 1;
    *hi = m1 + L.Win + 1;
  } else {
    const int hw = L.Ho * L.Wo;
    const int n0 = m0 / hw, y0 = (m0 - n0 * hw) / L.Wo;
    const int n1 = m1 / hw, y1 = (m1 - n1 * hw) / L.Wo;
    int r0 = 2 * y0 - 1, r1 = 2 * y1 + 1;
    if (r0 < 0) r0 = 0;
    if (r1 > L.Hin - 1) r1 = L.Hin - 1;
    *lo = (n0 * L.Hin + r0) * L.Win;
    *hi = (n1 * L.Hin + r1) * L.Win + L.Win - 1;
  }
  if (*lo < 0) *lo = 0;
  if (*hi > L.dep_M - 1) *hi = L.dep_M - 1;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) conv_chain_kernel(const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + p.stages * kStageBytes;
  float* s_bias = reinterpret_cast<float*>(staging + kStagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* full_bar = bars;                    // leader CTA only
  uint64_t* empty_bar = bars + kMaxStages;      // per CTA, multicast-arrived by the leader's commits
  uint64_t* tmem_full = bars + 2 * kMaxStages;  // per CTA
  uint64_t* tmem_empty = tmem_full + 2;         // leader CTA only, 16 arrivals
  // Staging: 64 KB.  Tiles of <= 128 fp16 columns (<= 32 KB) alternate between its two halves, so the epilogue of tile
  // i+1 writes one half while tile i's store still reads the other (the short-K 1x1 layers were bound by that hand-over);
  // 256-column and fp32 tiles take the whole area.  Work item i uses barrier / store warp i & 1:
  uint64_t* stg_ready = tmem_empty + 2;         // [2] the tile's staging region is free (and its residual, if any, landed)
  uint64_t* stg_full = stg_ready + 2;           // [2] the region has been written by the 8 epilogue warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stg_full + 2);
  volatile int* released = reinterpret_cast<volatile int*>(tmem_ptr + 2);   // [2] last item whose staging reads are done

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = static_cast<int>(ptx::cluster_id_x());
  const int* wl = p.work + static_cast<size_t>(pair) * p.work_stride;
  unsigned long long* tr = p.trace ? p.trace + 16ull * blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  ptx::pdl_launch_dependents();
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < p.stages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(&tmem_full[a], 1);
        ptx::mbar_init(&tmem_empty[a], 16);  // 8 epilogue warps x 2 CTAs
      }
      ptx::mbar_init(&stg_ready[0], 1);
      ptx::mbar_init(&stg_ready[1], 1);
      ptx::mbar_init(&stg_full[0], kEpiThreads / 32);
      ptx::mbar_init(&stg_full[1], kEpiThreads / 32);
      released[0] = -1;
      released[1] = -1;
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  ptx::cluster_sync();  // both CTAs' barriers exist before any remote arrive / multicast commit
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::pdl_wait();
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      long long w_dep = 0, w_empty = 0;
      int cur = -1;
      const ChainLayer* L = nullptr;
      for (int i = 0;; ++i) {
        const int item = __ldg(wl + i);
        if (item < 0) break;
        const int l = item >> kItemShift, tile = item & ((1 << kItemShift) - 1);
        if (l != cur) {
          cur = l;
          L = p.layers + l;
          ptx::prefetch_tmap(&L->tmA);
          ptx::prefetch_tmap(&L->tmB);
        }
        const int tiles_n = L->tiles_n, bn = L->bn, num_kb = L->num_kb, kb_per_tap = L->kb_per_tap;
        const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
        int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM;   // this CTA's 128 rows
        const bool live = m0 < L->M;
        if (!live) m0 = 0;  // ragged last pair: rows are discarded by the store warp, load something valid
        const int nb = tn * bn + static_cast<int>(rank) * (bn / 2);   // this CTA's half of the weight rows
        const uint32_t stage_tx = 2u * static_cast<uint32_t>(kABytes + (bn / 2) * kBK * 2);
        int cw = 0, ch = 0, cn = 0;
        if (L->im2col) {
          const int q0 = m0 % L->Wo;
          const int t = m0 / L->Wo;
          cw = q0 * L->stride - L->pad;
          ch = (t % L->Ho) * L->stride - L->pad;
          cn = t / L->Ho;
        }
        bool deps_ok = L->dep_kind < 0;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          ME_CHAIN_TRACED(w_empty, mbar_wait(&empty_bar[stage], phase ^ 1, p.debug, 0x100u + stage));
          uint8_t* sa = stage_base + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          ptx::tma_load_2d_pair(&L->tmB, full_leader, sb, kb * kBK, nb);   // weights never wait
          if (!deps_ok) {
            // the input rows of this tile: wait until the producer layer has stored them
            int m1 = m0 + kBM - 1;
            if (m1 > L->M - 1) m1 = L->M - 1;
            int lo, hi;
            dep_rows(*L, m0, m1, &lo, &hi);
            for (int t = lo / (2 * kBM); t <= hi / (2 * kBM); ++t)
              ME_CHAIN_TRACED(w_dep, wait_counter(p.counters + L->dep_base + t, L->dep_target, p.debug, 0x700u));
            fence_proxy_async_all();
            deps_ok = true;
          }
          if (L->im2col) {
            const int r = tap / 3, s = tap - r * 3;
            ptx::tma_load_im2col_4d_pair(&L->tmA, full_leader, sa, cb * kBK, cw, ch, cn, (uint16_t)s, (uint16_t)r);
          } else {
            ptx::tma_load_2d_pair(&L->tmA, full_leader, sa, cb * kBK, m0);
          }
          if (++cb == kb_per_tap) { cb = 0; ++tap; }
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (tr) { tr[2] = w_dep; tr[3] = w_empty; tr[4] = clock64(); }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      int cur = -1, num_kb = 0;
      uint32_t idesc = 0;
      long long w_full = 0, w_acc = 0, t_first = 0;
      int n_items = 0;
      for (int it = 0;; ++it) {
        const int item = __ldg(wl + it);
        if (item < 0) break;
        const int l = item >> kItemShift;
        if (l != cur) {
          cur = l;
          num_kb = p.layers[l].num_kb;
          idesc = ptx::make_idesc_f16(2 * kBM, p.layers[l].bn);
        }
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ME_CHAIN_TRACED(w_acc, mbar_wait(&tmem_empty[acc], acc_phase ^ 1, p.debug, 0x200u + acc));
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        ++n_items;
        for (int kb = 0; kb < num_kb; ++kb) {
          ME_CHAIN_TRACED(w_full, mbar_wait(&full_bar[stage], phase, p.debug, 0x300u + stage));
          if (tr && t_first == 0) { t_first = clock64(); w_full = 0; }
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stage_base + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = ptx::make_kmajor_desc(a_addr + k * 32, kBK * 2);
            const uint64_t bdesc = ptx::make_kmajor_desc(b_addr + k * 32, kBK * 2);
            ptx::umma_f16_ss_pair(d_tmem, adesc, bdesc, idesc, (kb != 0 || k != 0) ? 1u : 0u);
          }
          ptx::umma_commit_pair(&empty_bar[stage], 0b11);   // frees the stage in both CTAs
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pair(&tmem_full[acc], 0b11);       // accumulator ready in both CTAs
      }
      if (tr) { tr[5] = w_full; tr[6] = w_acc; tr[7] = t_first; tr[8] = clock64(); tr[14] = n_items; }
    }
    __syncwarp();
  } else if (warp < 2 + kEpiThreads / 32) {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = (warp - 2) >> 2;   // which BN/2 columns this warp converts
    const int etid = threadIdx.x - 64;
    int cur = -1, bn = 0, tiles_n = 1, act = 0, has_res = 0, out_f32 = 0;
    const float* bias = nullptr;
    const bool eleader = threadIdx.x == 64;
    int bias_layer = -1, bias_tn = -1;
    long long w_tfull = 0, w_stg = 0;
    if (!eleader) tr = nullptr;
    for (int it = 0;; ++it) {
      const int item = __ldg(wl + it);
      if (item < 0) break;
      const int l = item >> kItemShift, tile = item & ((1 << kItemShift) - 1);
      if (l != cur) {
        cur = l;
        const ChainLayer* L = p.layers + l;
        bn = L->bn;
        tiles_n = L->tiles_n;
        act = L->act;
        has_res = L->has_res;
        out_f32 = L->out_f32;
        bias = L->bias;
      }
      const int tn = tile % tiles_n;
      const int n0 = tn * bn;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (item >> kItemShift != bias_layer || tn != bias_tn) {   // same layer and n tile as the last one: bias is in place
        bias_layer = item >> kItemShift;
        bias_tn = tn;
        ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);   // everybody is done with the previous tile's bias
        for (int k = etid; k < bn; k += kEpiThreads) s_bias[k] = __ldg(bias + n0 + k);
        ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);
      }
      ME_CHAIN_TRACED(w_tfull, mbar_wait(&tmem_full[acc], acc_phase, p.debug, 0x400u + acc));
      ptx::tc_fence_after();
      ME_CHAIN_TRACED(w_stg, mbar_wait(&stg_ready[it & 1], (it >> 1) & 1, p.debug, 0x500u + (it & 1)));
      // this tile's staging region (see the layout comment at the top of the kernel)
      uint8_t* stg = staging + ((bn <= 128 && !out_f32) ? (it & 1) * (kStagingBytes / 2) : 0);

      const int c_base = half * (bn / 2);
      const int nch = bn / 64;   // 32-column chunks per warp: 1, 2 or 4
      // bias + activation (+ residual from the staging tile) -> fp16 -> swizzled staging tile, 32 columns of this row;
      // fp32 layers (head logits): 32 fp32 columns = one 128-byte row of staging sub-tile c / 32
      auto convert_chunk = [&](int c, const uint32_t (&r)[32]) {
        float v[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c + 4 * j4);
          v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + b4.x;
          v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + b4.y;
          v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + b4.z;
          v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + b4.w;
        }
        if (act == ME_ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
        } else if (act == ME_ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        }
        if (out_f32) {
          uint8_t* sub = stg + (c >> 5) * kSubBytes;
          const uint32_t rbase = row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t off = rbase + j * 16;
            off ^= ((off >> 7) & 7u) << 4;
            *reinterpret_cast<float4*>(sub + off) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          return;
        }
        uint8_t* sub = stg + (c >> 6) * kSubBytes;
        const uint32_t rbase = row * 128 + (c & 63) * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t off = rbase + j * 16;
          off ^= ((off >> 7) & 7u) << 4;
          uint4* dst = reinterpret_cast<uint4*>(sub + off);
          float* vv = v + 8 * j;
          if (has_res) {
            const uint4 rr = *dst;
            const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(rh[e]);
              vv[2 * e] += f.x;
              vv[2 * e + 1] += f.y;
            }
          }
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(vv[2 * e], vv[2 * e + 1]);
          *dst = o;
        }
      };
      const uint32_t t_row = tmem_base + acc * kAccStride + c_base + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32b_x32(t_row, r0);
      for (int ci = 0; ci < nch; ci += 2) {
        ptx::tmem_ld_wait_regs(r0);
        if (ci + 1 < nch) ptx::tmem_ld_32x32b_x32(t_row + (ci + 1) * 32, r1);
        convert_chunk(c_base + ci * 32, r0);
        if (ci + 1 < nch) {
          ptx::tmem_ld_wait_regs(r1);
          if (ci + 2 < nch) ptx::tmem_ld_32x32b_x32(t_row + (ci + 2) * 32, r0);
          convert_chunk(c_base + (ci + 1) * 32, r1);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));
      ptx::fence_proxy_async_smem();   // staging writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&stg_full[it & 1]);
    }
    if (tr) { tr[9] = w_tfull; tr[10] = w_stg; tr[12] = clock64(); }
  } else {
    // ------------------------------------------------------------------ store warps (alternating tiles)
    // Warp w owns the work items i with (i & 1) == w: it hands item i its staging region (stg_ready[w]; the residual
    // tile is TMA-loaded into the region first), waits for the epilogue (stg_full[w]), issues the TMA store, marks the
    // region's reads done (released[w] = i), then waits for the store to COMPLETE and publishes the tile's counter.
    const int w = warp - (2 + kEpiThreads / 32);
    if (ptx::elect_one()) {
      auto half_tile = [&](int item) {   // <= 32 KB of staging: alternates between the two halves
        const ChainLayer* N = p.layers + (item >> kItemShift);
        return N->bn <= 128 && !N->out_f32;
      };
      auto region = [&](int i, int item) { return staging + (half_tile(item) ? (i & 1) * (kStagingBytes / 2) : 0); };
      // Gives work item i its staging region: the residual tile is loaded into it (after the layer that produces the
      // residual has stored those rows), or the barrier is simply arrived on.
      auto hand_over = [&](int i, int item) {
        const ChainLayer* N = p.layers + (item >> kItemShift);
        const int tile = item & ((1 << kItemShift) - 1);
        if (!N->has_res) {
          ptx::mbar_arrive(&stg_ready[w]);
          return;
        }
        const int tm = tile / N->tiles_n, tn = tile - tm * N->tiles_n;
        int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM;
        if (m0 >= N->M) m0 = 0;
        if (N->res_base >= 0) {
          wait_counter(p.counters + N->res_base + tm, N->res_target, p.debug, 0x800u);
          fence_proxy_async_all();
        }
        const int nsub = N->bn / 64;   // (residual layers are fp16)
        uint8_t* dst = region(i, item);
        ptx::mbar_arrive_expect_tx(&stg_ready[w], nsub * kSubBytes);
        for (int sub = 0; sub < nsub; ++sub)
          ptx::tma_load_2d(&N->tmR, &stg_ready[w], dst + sub * kSubBytes, tn * N->bn + sub * 64, m0);
      };
      int cur = -1;
      const ChainLayer* L = nullptr;
      int prev_item = -1;        // item i - 1 (the other warp's)
      bool handed = false;       // item i was already handed over at the end of item i - 2
      for (int i = 0;; ++i) {
        const int item = __ldg(wl + i);
        if (item < 0) break;
        const int l = item >> kItemShift, tile = item & ((1 << kItemShift) - 1);
        if (l != cur) {
          cur = l;
          L = p.layers + l;
          ptx::prefetch_tmap(&L->tmC);
          if (L->has_res) ptx::prefetch_tmap(&L->tmR);
        }
        if ((i & 1) != w) {
          prev_item = item;
          continue;
        }
        if (!handed) {
          // the region overlaps item i - 1's when either takes the whole area: wait until the other warp's store has
          // read it (item i - 2's region was released by this warp before it got here)
          if (i >= 1 && (!half_tile(item) || !half_tile(prev_item))) {
            const long long t0 = clock64();
            uint32_t spins = 0;
            while (released[1 - w] < i - 1) {
              if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) watchdog_trap(p.debug, 0x900u + w, 0);
            }
            __threadfence_block();
          }
          hand_over(i, item);
        }
        handed = false;
        const int tm = tile / L->tiles_n, tn = tile - tm * L->tiles_n;
        const int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM, n0 = tn * L->bn;
        mbar_wait(&stg_full[w], (i >> 1) & 1, p.debug, 0x600u + w);
        if (m0 < L->M) {
          const int sub_cols = L->out_f32 ? 32 : 64;
          const int nsub = L->bn / sub_cols;
          const uint8_t* src = region(i, item);
          for (int sub = 0; sub < nsub; ++sub)
            ptx::tma_store_2d(&L->tmC, src + sub * kSubBytes, n0 + sub * sub_cols, m0);
          ptx::tma_store_commit();
          ptx::tma_store_wait_read0();    // the staging region has been read
        }
        __threadfence_block();
        released[w] = i;
        // Item i + 2 reuses this very half when items i, i+1, i+2 are all half tiles: hand it over right away, unless
        // its residual may depend on the tile just stored (then this thread must publish first, or it would wait for
        // itself inside hand_over()).
        const int next1 = __ldg(wl + i + 1);
        const int next2 = next1 >= 0 ? __ldg(wl + i + 2) : -1;
        if (next2 >= 0 && half_tile(item) && half_tile(next1) && half_tile(next2)) {
          const ChainLayer* N2 = p.layers + (next2 >> kItemShift);
          if (!(N2->has_res && N2->res_base >= 0)) {
            hand_over(i + 2, next2);
            handed = true;
          }
        }
        ptx::tma_store_wait_all0();       // the tile is in memory
        fence_proxy_async_all();
        publish_counter(p.counters + L->ctr_base + tm);
      }
    }
    __syncwarp();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();  // the peer may still be reading our barriers / issuing MMAs on our TMEM
  ptx::tc_fence_after();
  if (p.trace && threadIdx.x == 0) p.trace[16ull * blockIdx.x + 13] = clock64();
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
}

// Tile width of a layer: the widest of 256 / 128 / 64 that divides cout (fp32 layers: 128, their staging rows are 32 columns)
inline int chain_bn(const me_conv_desc& d) {
  if (d.out_f32) return 128;
  return d.cout % 256 == 0 ? 256 : (d.cout % 128 == 0 ? 128 : 64);
}

struct HostLayer {
  int tiles_m, tiles_n, num_kb, dep, res, dep_kind;
  int M, Ho, Wo, Hin, Win, dep_M;
};

}  // namespace
}  // namespace me

extern "C" {

int me_conv_chain_eligible(const me_conv_desc* d) {
  if (!d) return 0;
  if (d->ksize != 1 && d->ksize != 3) return 0;
  if (!(d->stride == 1 || (d->stride == 2 && d->ksize == 3))) return 0;
  if (d->cin <= 0 || d->cin % me::kBK != 0) return 0;
  if (d->out_f32) {   // YOLO head logits: 128-column tiles of fp32, no residual
    if (d->cout <= 0 || d->cout % 128 != 0 || d->res_pitch > 0) return 0;
    if (d->out_pitch < d->cout || d->out_pitch % 4 != 0) return 0;
  } else {
    if (d->cout <= 0 || d->cout % 64 != 0) return 0;
    if (d->out_pitch < d->cout || d->out_pitch % 8 != 0) return 0;
  }
  if (d->in_pitch < d->cin || d->in_pitch % 8 != 0) return 0;
  return 1;
}

size_t me_conv_chain_blob_bytes(const me_chain_layer* layers, int n_layers) {
  using namespace me;
  if (!layers || n_layers <= 0) return 0;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int npairs = sms / 2;
  long long tiles = 0, mtiles = 0;
  for (int l = 0; l < n_layers; ++l) {
    const me_conv_desc& d = layers[l].d;
    const int pad = (d.ksize - 1) / 2;
    const long long Ho = (d.h + 2 * pad - d.ksize) / d.stride + 1, Wo = (d.w + 2 * pad - d.ksize) / d.stride + 1;
    const long long M = d.n * Ho * Wo;
    const long long tm = (M + 2 * kBM - 1) / (2 * kBM);
    const int bn = chain_bn(d);
    tiles += tm * (d.cout / bn);
    mtiles += tm;
  }
  // every pair's list can in principle hold every tile; lists are bounded by tiles/npairs * 4 + slack in build
  const long long stride = tiles + 1;
  size_t bytes = kHeaderBytes + static_cast<size_t>(n_layers) * sizeof(ChainLayer);
  bytes += static_cast<size_t>(npairs) * stride * sizeof(int);
  bytes = (bytes + 255) & ~size_t(255);
  bytes += static_cast<size_t>(mtiles) * sizeof(int);
  return (bytes + 255) & ~size_t(255);
}

int me_conv_chain_verify(const void* host_blob);

static int chain_build_impl(const me_chain_layer* layers, int n_layers, void* host_blob, size_t blob_bytes, bool encode) {
  using namespace me;
  ME_REQUIRE(layers && host_blob && n_layers > 0, "conv_chain: null argument");
  ME_REQUIRE(n_layers < (1 << (31 - kItemShift)), "conv_chain: too many layers");
  const size_t need = me_conv_chain_blob_bytes(layers, n_layers);
  ME_REQUIRE(blob_bytes >= need, "conv_chain: blob of %zu bytes given, %zu needed", blob_bytes, need);
  ME_REQUIRE((reinterpret_cast<uintptr_t>(host_blob) & 127) == 0, "conv_chain: host blob must be 128-byte aligned");
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int npairs = sms / 2;
  memset(host_blob, 0, need);
  unsigned char* base = static_cast<unsigned char*>(host_blob);
  ChainHeader* H = reinterpret_cast<ChainHeader*>(base);
  ChainLayer* CL = reinterpret_cast<ChainLayer*>(base + kHeaderBytes);

  std::vector<HostLayer> hl(n_layers);
  long long tiles = 0;
  int ctr = 0;
  for (int l = 0; l < n_layers; ++l) {
    const me_chain_layer& a = layers[l];
    const me_conv_desc& d = a.d;
    ME_REQUIRE(me_conv_chain_eligible(&d), "conv_chain: layer %d is not eligible (k=%d s=%d cin=%d cout=%d f32=%d)", l,
               d.ksize, d.stride, d.cin, d.cout, d.out_f32);
    ME_REQUIRE(a.x && a.w_packed && a.bias && a.y, "conv_chain: layer %d has a null pointer", l);
    ME_REQUIRE(a.dep_layer < l && a.res_layer < l, "conv_chain: layer %d depends on a later layer", l);
    const int pad = (d.ksize - 1) / 2;
    const int Ho = (d.h + 2 * pad - d.ksize) / d.stride + 1;
    const int Wo = (d.w + 2 * pad - d.ksize) / d.stride + 1;
    const long long M64 = 1LL * d.n * Ho * Wo;
    ME_REQUIRE(M64 > 0 && M64 < (1LL << 31), "conv_chain: pixel count out of range");
    const int M = static_cast<int>(M64);
    const int taps = d.ksize * d.ksize;
    const int cin_pad = round_up(d.cin, kBK);
    const int ktot = taps * cin_pad;
    const int bn = chain_bn(d);
    ChainLayer& L = CL[l];
    L.bias = a.bias;
    L.M = M;
    L.Ho = Ho;
    L.Wo = Wo;
    L.stride = d.stride;
    L.pad = pad;
    L.kb_per_tap = cin_pad / kBK;
    L.num_kb = taps * L.kb_per_tap;
    L.tiles_n = d.cout / bn;
    L.bn = bn;
    L.act = d.act;
    L.has_res = (d.res_pitch > 0 && a.residual != nullptr) ? 1 : 0;
    L.im2col = d.ksize == 3 ? 1 : 0;
    L.out_f32 = d.out_f32 ? 1 : 0;
    const int tiles_m = ceil_div(M, 2 * kBM);
    ME_REQUIRE(tiles_m * L.tiles_n < (1 << kItemShift), "conv_chain: too many tiles in layer %d", l);
    L.ctr_base = ctr;
    ctr += tiles_m;
    hl[l] = HostLayer{tiles_m, L.tiles_n, L.num_kb, a.dep_layer, L.has_res ? a.res_layer : -1, -1, M, Ho, Wo, d.h, d.w, 0};
    tiles += 1LL * tiles_m * L.tiles_n;
    L.dep_kind = -1;
    L.dep_base = 0;
    L.dep_target = 0;
    L.dep_M = 0;
    L.Hin = d.h;
    L.Win = d.w;
    if (a.dep_layer >= 0) {
      const HostLayer& P = hl[a.dep_layer];
      ME_REQUIRE(P.M == d.n * d.h * d.w, "conv_chain: layer %d reads %d rows but layer %d writes %d", l, d.n * d.h * d.w,
                 a.dep_layer, P.M);
      L.dep_kind = d.ksize == 1 ? 0 : (d.stride == 1 ? 1 : 2);
      L.dep_base = CL[a.dep_layer].ctr_base;
      L.dep_target = 2 * P.tiles_n;
      L.dep_M = P.M;
      hl[l].dep_kind = L.dep_kind;
      hl[l].dep_M = P.M;
    }
    L.res_base = -1;
    L.res_target = 0;
    if (L.has_res && a.res_layer >= 0) {
      const HostLayer& P = hl[a.res_layer];
      ME_REQUIRE(P.M == M, "conv_chain: residual of layer %d has %d rows, the layer %d", l, P.M, M);
      L.res_base = CL[a.res_layer].ctr_base;
      L.res_target = 2 * P.tiles_n;
    }
    if (!encode) continue;     // schedule only (me_conv_chain_plan): no driver needed, the blob cannot be run
    int rc;
    if (L.im2col) {
      rc = encode_im2col_nhwc(&L.tmA, a.x, d.n, d.h, d.w, d.cin, d.in_pitch, d.ksize, pad, d.stride, kBK, kBM,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      rc = encode_tiled_2d(&L.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.x, d.cin, M, d.in_pitch, kBK, kBM,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc != ME_OK) return rc;
    rc = encode_tiled_2d(&L.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.w_packed, ktot, d.cout, ktot, kBK, bn / 2,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != ME_OK) return rc;
    if (d.out_f32)
      rc = encode_tiled_2d(&L.tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.y, d.cout, M, d.out_pitch, 32, kBM,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    else
      rc = encode_tiled_2d(&L.tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.y, d.cout, M, d.out_pitch, 64, kBM,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != ME_OK) return rc;
    if (L.has_res) {
      rc = encode_tiled_2d(&L.tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.residual, d.cout, M, d.res_pitch, 64, kBM,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != ME_OK) return rc;
    } else {
      L.tmR = L.tmC;
    }
  }

  // ---- work lists: the tiles in global (layer-major) order, each given to the pair that can start it first in a
  // small timing model (K blocks of 512 tensor clocks, a fixed per-tile cost, a tile starts once the m tiles it
  // reads are complete plus a hand-over latency).  Lists stay in global order, which is what makes the waits safe.
  const long long stride = tiles + 1;
  H->magic = encode ? kMagic : kMagicPlan;
  H->n_layers = n_layers;
  H->npairs = npairs;
  H->work_stride = static_cast<int>(stride);
  int stages = (227 * 1024 - 1024 - kStagingBytes - kTailBytes) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  {
    const char* e = getenv("ME_PAIR_STAGES");
    const int cap = e ? atoi(e) : 0;
    if (cap >= 2 && cap < stages) stages = cap;
  }
  ME_REQUIRE(stages >= 2, "conv_chain: not enough shared memory");
  H->stages = stages;
  H->smem = smem_bytes(stages);
  H->n_counters = ctr;
  H->layers_off = kHeaderBytes;
  H->work_off = kHeaderBytes + static_cast<long long>(n_layers) * sizeof(ChainLayer);
  long long off = H->work_off + static_cast<long long>(npairs) * stride * sizeof(int);
  off = (off + 255) & ~255LL;
  H->counters_off = off;
  H->total_bytes = static_cast<long long>(need);
  int* work = reinterpret_cast<int*>(base + H->work_off);

  // Rows [lo, hi] of the producer that m tile `tm` of layer h reads, as producer m tiles [a, b] (same rule as dep_rows()).
  auto dep_tiles = [&](const HostLayer& h, int tm, int* a, int* b) {
    int m0 = tm * 2 * kBM, m1 = m0 + 2 * kBM - 1;
    if (m1 > h.M - 1) m1 = h.M - 1;
    int lo, hi;
    if (h.dep_kind == 0) {
      lo = m0;
      hi = m1;
    } else if (h.dep_kind == 1) {
      lo = m0 - h.Win - 1;
      hi = m1 + h.Win + 1;
    } else {
      const int hw = h.Ho * h.Wo;
      const int n0 = m0 / hw, y0 = (m0 - n0 * hw) / h.Wo, n1 = m1 / hw, y1 = (m1 - n1 * hw) / h.Wo;
      int r0 = 2 * y0 - 1, r1 = 2 * y1 + 1;
      if (r0 < 0) r0 = 0;
      if (r1 > h.Hin - 1) r1 = h.Hin - 1;
      lo = (n0 * h.Hin + r0) * h.Win;
      hi = (n1 * h.Hin + r1) * h.Win + h.Win - 1;
    }
    if (lo < 0) lo = 0;
    if (hi > h.dep_M - 1) hi = h.dep_M - 1;
    *a = lo / (2 * kBM);
    *b = hi / (2 * kBM);
  };
  // Work lists: the tiles in strict layer-major order, each given to the pair that can start it first in a small timing
  // model (clocks): K blocks of 512 tensor clocks, a fixed cost per tile, a tile can start kHand after the m tiles it reads
  // were stored.  (An event-driven list scheduler that lets tiles of layer l+1 overtake the tail of layer l was measured
  // and dropped: conv stack 2.46 vs 2.42 ms on Darknet-53, and the kernel relies on layer-major lists.)
  const double kKb = 512.0, kTile = 1500.0, kHand = 3000.0;
  std::vector<double> pair_t(npairs, 0.0);
  std::vector<int> pair_n(npairs, 0);
  std::vector<std::vector<double>> done(n_layers);   // completion time of every m tile
  for (int l = 0; l < n_layers; ++l) done[l].assign(hl[l].tiles_m, 0.0);
  {
    for (int l = 0; l < n_layers; ++l) {
      const HostLayer& h = hl[l];
      const double cost = h.num_kb * kKb + kTile;
      for (int tm = 0; tm < h.tiles_m; ++tm) {
        double ready = 0.0;
        if (h.dep >= 0) {
          int a, b;
          dep_tiles(h, tm, &a, &b);
          for (int t = a; t <= b; ++t)
            if (done[h.dep][t] > ready) ready = done[h.dep][t];
          ready += kHand;
        }
        if (h.res >= 0 && done[h.res][tm] + kHand > ready) ready = done[h.res][tm] + kHand;
        for (int tn = 0; tn < h.tiles_n; ++tn) {
          int best = 0;
          double best_key = 1e300;
          for (int q = 0; q < npairs; ++q) {
            const double key = pair_t[q] >= ready ? pair_t[q] : ready + (ready - pair_t[q]) * 1e-6;
            if (key < best_key) {
              best_key = key;
              best = q;
            }
          }
          const double start = pair_t[best] > ready ? pair_t[best] : ready;
          pair_t[best] = start + cost;
          if (pair_t[best] > done[l][tm]) done[l][tm] = pair_t[best];
          work[static_cast<long long>(best) * stride + pair_n[best]++] = (l << kItemShift) | (tm * h.tiles_n + tn);
        }
      }
    }
  }
  for (int q = 0; q < npairs; ++q) work[static_cast<long long>(q) * stride + pair_n[q]] = -1;
  // never hand out a schedule that cannot complete: replay it against the rules the kernel waits by
  return me_conv_chain_verify(host_blob);
}

int me_conv_chain_build(const me_chain_layer* layers, int n_layers, void* host_blob, size_t blob_bytes) {
  return chain_build_impl(layers, n_layers, host_blob, blob_bytes, true);
}

// The same layer table and work lists WITHOUT tensor maps: needs no CUDA driver (tests, inspection); me_conv_chain_run
// refuses such a blob.
int me_conv_chain_plan(const me_chain_layer* layers, int n_layers, void* host_blob, size_t blob_bytes) {
  return chain_build_impl(layers, n_layers, host_blob, blob_bytes, false);
}

// Replays the work lists of a blob the way the kernel executes them - every pair takes its items in list order; an item
// can run once the counters of the m tiles it reads (dep_rows) and of its residual tile have reached their targets; a
// finished item adds 2 to its m tile's counter - and checks that every tile of every layer is listed exactly once, that
// all lists run to their end (no deadlock whatever the timing, all pairs being co-resident) and that every counter ends
// at its target.  Pairs are modelled as strictly sequential, which is stricter than the kernel's pipelining.
int me_conv_chain_verify(const void* host_blob) {
  using namespace me;
  ME_REQUIRE(host_blob, "conv_chain_verify: null argument");
  const unsigned char* base = static_cast<const unsigned char*>(host_blob);
  const ChainHeader* H = reinterpret_cast<const ChainHeader*>(base);
  ME_REQUIRE(H->magic == kMagic || H->magic == kMagicPlan, "conv_chain_verify: not a chain blob");
  const ChainLayer* CL = reinterpret_cast<const ChainLayer*>(base + H->layers_off);
  const int* work = reinterpret_cast<const int*>(base + H->work_off);
  const int nl = H->n_layers, np = H->npairs;
  std::vector<int> counters(H->n_counters, 0);
  std::vector<std::vector<unsigned char>> seen(nl);
  std::vector<int> tiles_m(nl);
  long long total = 0;
  for (int l = 0; l < nl; ++l) {
    tiles_m[l] = ceil_div(CL[l].M, 2 * kBM);
    seen[l].assign(static_cast<size_t>(tiles_m[l]) * CL[l].tiles_n, 0);
    total += static_cast<long long>(tiles_m[l]) * CL[l].tiles_n;
  }
  std::vector<long long> pos(np, 0);
  long long done = 0;
  bool progress = true;
  auto ready = [&](int item, int* why_ctr, int* why_val, int* why_target) {
    const int l = item >> kItemShift, tile = item & ((1 << kItemShift) - 1);
    const ChainLayer& L = CL[l];
    const int tm = tile / L.tiles_n;
    if (L.dep_kind >= 0) {
      const int m0 = tm * 2 * kBM;
      int m1 = m0 + 2 * kBM - 1;
      if (m1 > L.M - 1) m1 = L.M - 1;
      int lo, hi;
      dep_rows(L, m0, m1, &lo, &hi);
      for (int t = lo / (2 * kBM); t <= hi / (2 * kBM); ++t)
        if (counters[L.dep_base + t] < L.dep_target) {
          *why_ctr = L.dep_base + t; *why_val = counters[L.dep_base + t]; *why_target = L.dep_target;
          return false;
        }
    }
    if (L.has_res && L.res_base >= 0 && counters[L.res_base + tm] < L.res_target) {
      *why_ctr = L.res_base + tm; *why_val = counters[L.res_base + tm]; *why_target = L.res_target;
      return false;
    }
    return true;
  };
  while (progress) {
    progress = false;
    for (int q = 0; q < np; ++q) {
      for (;;) {
        const int item = work[static_cast<long long>(q) * H->work_stride + pos[q]];
        if (item < 0) break;
        const int l = item >> kItemShift, tile = item & ((1 << kItemShift) - 1);
        ME_REQUIRE(l < nl && tile < static_cast<int>(seen[l].size()), "conv_chain_verify: pair %d lists item %d of layer %d out of range", q, tile, l);
        int c, v, t;
        if (!ready(item, &c, &v, &t)) break;
        ME_REQUIRE(!seen[l][tile], "conv_chain_verify: tile %d of layer %d is listed twice", tile, l);
        seen[l][tile] = 1;
        counters[CL[l].ctr_base + tile / CL[l].tiles_n] += 2;
        ++pos[q];
        ++done;
        progress = true;
      }
    }
  }
  if (done != total) {
    for (int q = 0; q < np; ++q) {
      const int item = work[static_cast<long long>(q) * H->work_stride + pos[q]];
      if (item < 0) continue;
      int c = -1, v = 0, t = 0;
      ready(item, &c, &v, &t);
      return fail(ME_ERR_ARG, "conv_chain_verify: schedule cannot complete (%lld of %lld tiles ran): pair %d is blocked at "
                  "layer %d tile %d waiting for counter %d = %d < %d", done, total, q, item >> kItemShift,
                  item & ((1 << kItemShift) - 1), c, v, t);
    }
    return fail(ME_ERR_ARG, "conv_chain_verify: %lld of %lld tiles are listed", done, total);
  }
  for (int l = 0; l < nl; ++l)
    for (int tm = 0; tm < tiles_m[l]; ++tm)
      ME_REQUIRE(counters[CL[l].ctr_base + tm] == 2 * CL[l].tiles_n, "conv_chain_verify: counter of layer %d m tile %d ends at %d, "
                 "target %d", l, tm, counters[CL[l].ctr_base + tm], 2 * CL[l].tiles_n);
  return ME_OK;
}

int me_conv_chain_run(const void* host_blob, void* dev_blob, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(host_blob && dev_blob, "conv_chain: null argument");
  const ChainHeader* H = static_cast<const ChainHeader*>(host_blob);
  ME_REQUIRE(H->magic != kMagicPlan, "conv_chain: the blob holds a schedule only (me_conv_chain_plan); build it with me_conv_chain_build");
  ME_REQUIRE(H->magic == kMagic, "conv_chain: the host blob was not written by me_conv_chain_build");
  ME_REQUIRE((reinterpret_cast<uintptr_t>(dev_blob) & 255) == 0, "conv_chain: device blob must be 256-byte aligned");
  int rc = conv_ensure_debug_word();
  if (rc != ME_OK) return rc;
  unsigned char* dbase = static_cast<unsigned char*>(dev_blob);
  ChainParams p{};
  p.layers = reinterpret_cast<const ChainLayer*>(dbase + H->layers_off);
  p.work = reinterpret_cast<const int*>(dbase + H->work_off);
  p.work_stride = H->work_stride;
  p.stages = H->stages;
  p.counters = reinterpret_cast<int*>(dbase + H->counters_off);
  p.debug = conv_debug_word();
  p.trace = conv_trace_buffer();
  int dev = 0;
  ME_CUDA(cudaGetDevice(&dev));
  static bool attr_set[64] = {false};
  if (dev < 64 && !attr_set[dev]) {
    ME_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  ME_CUDA(cudaMemsetAsync(p.counters, 0, static_cast<size_t>(H->n_counters) * sizeof(int), stream));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * H->npairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = H->smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 0;   // the memset node in front of the kernel is a full dependency anyway
  ME_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel, p));
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
