// CTA-pair (cta_group::2) variant of the fused conv implicit GEMM for the large layers.
//
// Why: with one CTA per 128 x 128 tile every tcgen05.mma (K=16) reads 8 KB of operands from shared
// memory in ~64 clocks while TMA refills the same 8 KB - 256 B/clk against a 128 B/clk shared-memory
// port, which is what held the 3x3 layers at ~65 % of the tensor peak (profiles/round1).  A CTA pair
// computes a 256 x BN tile with ONE instruction stream: each CTA stages its own 128 rows of A and only
// BN/2 rows of the weights, the pair's tensor cores read both halves of B, so operand bytes per flop
// halve.  BN = 256 gives 64 B/clk of reads + 64 B/clk of fills per CTA.
//
// Structure (per CTA, roles as in conv_gemm.cu):
//   warp 0   TMA producer for this CTA's A rows (im2col or tiled) and its half of the B rows; the byte
//            count lands on the LEADER CTA's full barrier (cp.async.bulk.tensor ... .cta_group::2).
//   warp 1   leader only: waits the full barrier, issues tcgen05.mma.cta_group::2 (M=256, N=BN) and
//            multicast-commits to the empty / tmem_full barriers of both CTAs.
//   warps 2-9 epilogue over this CTA's 128 accumulator rows (own TMEM): warps 2-5 take the tile's first BN/2
//            columns, warps 6-9 the rest (the last tile's epilogue is exposed at the end of the kernel, so its
//            latency matters); same fused epilogue + TMA store; both CTAs release the accumulator on the
//            leader's tmem_empty barrier (16 arrivals).
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace me {

// defined in conv_gemm.cu
unsigned long long* conv_debug_word();
int conv_ensure_debug_word();
bool conv_pdl_enabled();
unsigned long long* conv_trace_buffer();

namespace {

constexpr int kBM = 128;       // rows per CTA (256 per pair)
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;   // 8 epilogue warps: two per TMEM lane quarter, each takes half of the tile's columns
constexpr uint32_t kEpiBarrierId = 1;
constexpr int kMaxStages = 8;
constexpr int kBK = 64;

struct PairParams {
  int M;
  int Ho, Wo;
  int stride, pad;
  int kb_per_tap;
  int num_kb;
  int tiles_m, tiles_n;   // in pair tiles: 256 x BN
  int act;
  int has_res;
  int im2col;
  int stages;
  int prefetch_kb;  // weight K blocks prefetched into L2 before griddepcontrol.wait
  int dbg;  // ME_CONV_DBG bit mask for attribution runs: 1 skip epilogue, 2 skip operand loads, 4 skip MMAs, 32 / 64 skip B / A loads
  const float* bias;
  unsigned long long* debug;
  unsigned long long* trace;  // see conv_gemm.cu
  // Tail split (see PairWork): the tiles of the last, partial wave are cut into `slices` K ranges
  int whole_limit;   // tiles [0, whole_limit) are computed whole, tile = pair + i * npairs
  int tail_tiles;    // tiles [whole_limit, whole_limit + tail_tiles) are split
  int slices;        // 1 = no split (then whole_limit = all tiles)
  float* ws;         // fp32 partial tiles [tail_tile][slice][cta rank][128][BN]
  int* counters;     // arrivals per [tail_tile][cta rank]; the last arriver reduces and resets
};

// Work items of one CTA pair.  A layer whose tile count is not a multiple of the 74 pairs used to end with a wave in
// which most pairs idle (13^2 x 32 frames, 512->1024: 88 tiles = one full wave + 14 tiles, i.e. 2 tile times for 1.19
// tiles of work per pair).  The tail tiles are now cut along K into `slices` ranges, one per otherwise idle pair; every
// slice writes its fp32 partial tile to a workspace, and the CTA that arrives last at the tile's counter adds the
// partials in slice order (deterministic) and runs the normal epilogue.  Nobody waits for anybody.
struct PairWork {
  int pair, npairs, whole_limit, tail_tiles, slices, num_kb;
  __device__ bool whole(int i, int* tile) const {
    *tile = pair + i * npairs;
    return *tile < whole_limit;
  }
  __device__ bool tail(int* tile, int* tail_tile, int* slice, int* kb0, int* kb1) const {
    if (slices <= 1 || pair >= tail_tiles * slices) return false;
    *tail_tile = pair % tail_tiles;
    *slice = pair / tail_tiles;
    *tile = whole_limit + *tail_tile;
    *kb0 = static_cast<int>(static_cast<long long>(num_kb) * *slice / slices);
    *kb1 = static_cast<int>(static_cast<long long>(num_kb) * (*slice + 1) / slices);
    return true;
  }
};

#define ME_TRACED_WAIT(acc, ...)            \
  do {                                      \
    if (p.trace) {                          \
      const long long t_ = clock64();       \
      mbar_wait(__VA_ARGS__);               \
      (acc) += clock64() - t_;              \
    } else {                                \
      mbar_wait(__VA_ARGS__);               \
    }                                       \
  } while (0)

template <int BN>
struct PCfg {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = (BN / 2) * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SUB_COLS = 64;                 // fp16, 128-byte staging rows
  static constexpr int NUM_SUB = BN / SUB_COLS;
  static constexpr int SUB_BYTES = kBM * 128;
  static constexpr int STAGING_BYTES = NUM_SUB * SUB_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int TAIL_BYTES = BN * 4 + 64 * 8 + 16;
  static constexpr int smem_bytes(int stages) { return 1024 + stages * STAGE_BYTES + STAGING_BYTES + TAIL_BYTES; }
  static_assert(BN == 128 || BN == 256, "pair tile N");
};

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long* dbg, uint32_t tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
      if (dbg) {
        *reinterpret_cast<volatile unsigned long long*>(dbg) =
            (static_cast<unsigned long long>(tag | 0x8000u) << 32) | (static_cast<unsigned long long>(blockIdx.x) << 8) | parity | 0x80u;
        __threadfence_system();
      }
      __trap();
    }
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ME_ACT_LEAKY) return v > 0.f ? v : 0.1f * v;
  if (act == ME_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                      const PairParams p) {
  using C = PCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + p.stages * C::STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(staging + C::STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + BN);
  uint64_t* full_bar = bars;                    // used in the leader CTA only
  uint64_t* empty_bar = bars + kMaxStages;      // per CTA, multicast-arrived by the leader's commits
  uint64_t* tmem_full = bars + 2 * kMaxStages;  // per CTA
  uint64_t* tmem_empty = tmem_full + 2;         // leader CTA only, 8 arrivals
  uint64_t* res_full = tmem_empty + 2;          // per CTA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = static_cast<int>(ptx::cluster_id_x());
  const int npairs = static_cast<int>(ptx::num_clusters_x());
  unsigned long long* tr = p.trace ? p.trace + 16ull * blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  ptx::pdl_launch_dependents();  // see conv_gemm.cu
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    if (p.has_res) ptx::prefetch_tmap(&tmR);
  }
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
      ptx::mbar_init(res_full, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  }
  ptx::cluster_sync();  // both CTAs' barriers exist before any remote arrive / multicast commit
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_ptr, C::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tr && threadIdx.x == 0) tr[1] = clock64();
  // weights of the first stages -> L2 while the previous layer drains (see conv_gemm.cu)
  if (warp == 0 && p.prefetch_kb > 0 && pair < p.tiles_m * p.tiles_n && ptx::elect_one()) {
    const int tn0 = pair % p.tiles_n;
    for (int kb = 0; kb < p.prefetch_kb; ++kb)
      ptx::tma_prefetch_2d(&tmB, kb * kBK, tn0 * BN + static_cast<int>(rank) * (BN / 2));
  }
  ptx::pdl_wait();
  if (tr && threadIdx.x == 0) tr[2] = clock64();

  const PairWork work{pair, npairs, p.whole_limit, p.tail_tiles, p.slices, p.num_kb};
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      auto produce = [&](int tile, int kb0, int kb1) {
        const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
        int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM;         // this CTA's 128 rows
        if (m0 >= p.M) m0 = 0;  // ragged last pair: rows are discarded by the epilogue, load something valid
        const int nb = tn * BN + static_cast<int>(rank) * (BN / 2);   // this CTA's half of the weight rows
        int cw = 0, ch = 0, cn = 0;
        if (p.im2col) {
          const int q0 = m0 % p.Wo;
          const int t = m0 / p.Wo;
          cw = q0 * p.stride - p.pad;
          ch = (t % p.Ho) * p.stride - p.pad;
          cn = t / p.Ho;
        }
        int tap = kb0 / p.kb_per_tap, cb = kb0 - tap * p.kb_per_tap;
        for (int kb = kb0; kb < kb1; ++kb) {
          ME_TRACED_WAIT(w_empty, &empty_bar[stage], phase ^ 1, p.debug, 0x100u + stage);
          uint8_t* sa = stage_base + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
          if (p.dbg & 2) {  // attribution run: no operand traffic, the barrier protocol stays intact
            if (leader) ptx::mbar_arrive(&full_bar[stage]);
          } else {
            // attribution bits 32 / 64: no weight (B) / no activation (A) traffic; the stage keeps stale bytes
            const bool load_a = !(p.dbg & 64), load_b = !(p.dbg & 32);
            if (leader)
              ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * ((load_a ? C::A_BYTES : 0) + (load_b ? C::B_BYTES : 0)));
            if (load_a) {
              if (p.im2col) {
                const int r = tap / 3, s = tap - r * 3;
                ptx::tma_load_im2col_4d_pair(&tmA, full_leader, sa, cb * kBK, cw, ch, cn, (uint16_t)s, (uint16_t)r);
              } else {
                ptx::tma_load_2d_pair(&tmA, full_leader, sa, cb * kBK, m0);
              }
            }
            if (load_b) ptx::tma_load_2d_pair(&tmB, full_leader, sb, kb * kBK, nb);
          }
          if (++cb == p.kb_per_tap) { cb = 0; ++tap; }
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      };
      int tile, tt, sl, kb0, kb1;
      for (int i = 0; work.whole(i, &tile); ++i) produce(tile, 0, p.num_kb);
      if (work.tail(&tile, &tt, &sl, &kb0, &kb1)) produce(tile, kb0, kb1);
      if (tr) { tr[3] = w_empty; tr[4] = clock64(); }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * kBM, BN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      long long w_full = 0, w_acc = 0, t_first = 0;
      auto issue = [&](int kb0, int kb1) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ME_TRACED_WAIT(w_acc, &tmem_empty[acc], acc_phase ^ 1, p.debug, 0x200u + acc);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ME_TRACED_WAIT(w_full, &full_bar[stage], phase, p.debug, 0x300u + stage);
          if (tr && t_first == 0) { t_first = clock64(); w_full = 0; }
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stage_base + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            if (p.dbg & 4) break;  // attribution run: no tensor work
            const uint64_t adesc = ptx::make_kmajor_desc(a_addr + k * 32, kBK * 2);
            const uint64_t bdesc = ptx::make_kmajor_desc(b_addr + k * 32, kBK * 2);
            ptx::umma_f16_ss_pair(d_tmem, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          ptx::umma_commit_pair(&empty_bar[stage], 0b11);   // frees the stage in both CTAs
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pair(&tmem_full[acc], 0b11);       // accumulator ready in both CTAs
        ++it;
      };
      int tile, tt, sl, kb0, kb1;
      for (int i = 0; work.whole(i, &tile); ++i) issue(0, p.num_kb);
      if (work.tail(&tile, &tt, &sl, &kb0, &kb1)) issue(kb0, kb1);
      if (tr) { tr[5] = w_full; tr[6] = w_acc; tr[7] = t_first; tr[8] = clock64(); tr[14] = it; }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = (warp - 2) >> 2;   // which BN/2 columns this warp converts
    const int etid = threadIdx.x - 64;
    const bool eleader = (threadIdx.x == 64);
    constexpr int NCH = BN / 64;        // 32-column chunks per warp
    const int c_base = half * (BN / 2);
    long long w_tfull = 0, w_other = 0, t_work = 0, t_first_full = 0;
    auto load_residual = [&](int tile) {
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      const int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM;
      ptx::mbar_arrive_expect_tx(res_full, C::STAGING_BYTES);
#pragma unroll
      for (int sub = 0; sub < C::NUM_SUB; ++sub)
        ptx::tma_load_2d(&tmR, res_full, staging + sub * C::SUB_BYTES, tn * BN + sub * C::SUB_COLS, m0 < p.M ? m0 : 0);
    };
    // bias + activation (+ residual from the staging tile) -> fp16 -> swizzled staging tile, 32 columns of this row
    auto convert_chunk = [&](int c, float (&v)[32]) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c + 4 * j4);
        v[4 * j4 + 0] += b4.x;
        v[4 * j4 + 1] += b4.y;
        v[4 * j4 + 2] += b4.z;
        v[4 * j4 + 3] += b4.w;
      }
      if (p.act == ME_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
      } else if (p.act == ME_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ME_ACT_SIGMOID);
      }
      uint8_t* sub = staging + (c / C::SUB_COLS) * C::SUB_BYTES;
      const uint32_t rbase = row * 128 + (c % C::SUB_COLS) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t off = rbase + j * 16;
        off ^= ((off >> 7) & 7u) << 4;
        uint4* dst = reinterpret_cast<uint4*>(sub + off);
        float* vv = v + 8 * j;
        if (p.has_res) {
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
    auto store_tile = [&](int m0, int n0) {
      if (m0 < p.M && !(p.dbg & 1)) {
#pragma unroll
        for (int sub = 0; sub < C::NUM_SUB; ++sub)
          ptx::tma_store_2d(&tmC, staging + sub * C::SUB_BYTES, n0 + sub * C::SUB_COLS, m0);
        ptx::tma_store_commit();
      }
    };
    // residual tiles are prefetched into the staging tile as soon as the previous store has drained it
    int tile, next_tile;
    if (p.has_res && eleader && work.whole(0, &tile)) load_residual(tile);
    int it = 0;
    for (int i = 0; work.whole(i, &tile); ++i, ++it) {
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      const int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM, n0 = tn * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long te0 = (tr && eleader) ? clock64() : 0;
      if (eleader && !p.has_res) ptx::tma_store_wait_read0();
      for (int k = etid; k < BN; k += kEpiThreads) s_bias[k] = p.bias[n0 + k];
      ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);
      const long long te1 = (tr && eleader) ? clock64() : 0;
      mbar_wait(&tmem_full[acc], acc_phase, p.debug, 0x400u + acc);
      ptx::tc_fence_after();
      const long long te2 = (tr && eleader) ? clock64() : 0;
      if (p.has_res) mbar_wait(res_full, it & 1, p.debug, 0x500u);
      const long long te3 = (tr && eleader) ? clock64() : 0;
      if (tr && eleader) {
        w_other += (te1 - te0) + (te3 - te2);
        w_tfull += te2 - te1;
        if (it == 0) t_first_full = te2;
      }

      if (!(p.dbg & 1)) {  // attribution run: accumulators are released unread
        const uint32_t t_row = tmem_base + acc * BN + c_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t r[2][32];
        ptx::tmem_ld_32x32b_x32(t_row, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          ptx::tmem_ld_wait_regs(r[ci & 1]);
          if (ci + 1 < NCH) ptx::tmem_ld_32x32b_x32(t_row + (ci + 1) * 32, r[(ci + 1) & 1]);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[ci & 1][j]);
          convert_chunk(c_base + ci * 32, v);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);
      if (eleader) {
        store_tile(m0, n0);
        if (p.has_res && work.whole(i + 1, &next_tile)) {
          ptx::tma_store_wait_read0();
          load_residual(next_tile);
        }
      }
      if (tr && eleader) t_work += clock64() - te3;
    }
    int tail_tile, slice, kb0, kb1;
    if (work.tail(&tile, &tail_tile, &slice, &kb0, &kb1)) {
      // ---- one K slice of a tail tile: partial sums to the workspace; the last arriver reduces and finishes the tile
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      const int m0 = tm * 2 * kBM + static_cast<int>(rank) * kBM, n0 = tn * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const size_t part = static_cast<size_t>(kBM) * BN;   // floats per partial tile of one CTA
      float* my_part = p.ws + ((static_cast<size_t>(tail_tile) * p.slices + slice) * 2 + rank) * part;
      for (int k = etid; k < BN; k += kEpiThreads) s_bias[k] = p.bias[n0 + k];
      mbar_wait(&tmem_full[acc], acc_phase, p.debug, 0x400u + acc);
      ptx::tc_fence_after();
      // Partial-tile layout [32-column chunk][float4 of the chunk][row]: a warp's 32 rows are 512 contiguous bytes for
      // every store / load instruction (row-major rows of 1 KB made each lane touch its own sector: the reduction alone
      // cost more than the split saved).
      const int chunk0 = c_base / 32;
      {
        const uint32_t t_row = tmem_base + acc * BN + c_base + (static_cast<uint32_t>(q * 32) << 16);
        float4* dst = reinterpret_cast<float4*>(my_part) + static_cast<size_t>(chunk0) * 8 * kBM + row;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_row + ci * 32, r);
          ptx::tmem_ld_wait_regs(r);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[(ci * 8 + j) * kBM] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                  __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));
      __threadfence();                                        // partial tile visible before the arrival is counted
      if (eleader) ptx::tma_store_wait_read0();               // the last whole tile's store has drained the staging tile
      ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);
      int* s_last = reinterpret_cast<int*>(tmem_ptr + 1);
      if (eleader) {
        int* ctr = p.counters + tail_tile * 2 + rank;
        const int old = atomicAdd(ctr, 1);
        const int last = old == p.slices - 1;
        if (last) *ctr = 0;                                   // every slice has arrived: ready for the next launch
        *s_last = last;
        if (last && p.has_res) load_residual(tile);
      }
      ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);
      if (*s_last) {
        __threadfence();
        if (p.has_res) mbar_wait(res_full, it & 1, p.debug, 0x500u);
        const float4* parts = reinterpret_cast<const float4*>(p.ws + (static_cast<size_t>(tail_tile) * p.slices * 2 + rank) * part) +
                              static_cast<size_t>(chunk0) * 8 * kBM + row;
#pragma unroll 1
        for (int ci = 0; ci < NCH; ++ci) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          for (int sidx = 0; sidx < p.slices; sidx += 2) {     // fixed order: the sum does not depend on who came last
            const float4* src = parts + static_cast<size_t>(sidx) * 2 * (part / 4) + ci * 8 * kBM;
            const bool two = sidx + 1 < p.slices;               // two slices' loads in flight per round trip
            float4 t0[8], t1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t0[j] = __ldcg(src + j * kBM);
            if (two) {
#pragma unroll
              for (int j = 0; j < 8; ++j) t1[j] = __ldcg(src + 2 * (part / 4) + j * kBM);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += t0[j].x;
              v[4 * j + 1] += t0[j].y;
              v[4 * j + 2] += t0[j].z;
              v[4 * j + 3] += t0[j].w;
            }
            if (two) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += t1[j].x;
                v[4 * j + 1] += t1[j].y;
                v[4 * j + 2] += t1[j].z;
                v[4 * j + 3] += t1[j].w;
              }
            }
          }
          convert_chunk(c_base + ci * 32, v);
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(kEpiBarrierId, kEpiThreads);
        if (eleader) store_tile(m0, n0);
      }
      ++it;
    }
    if (eleader) ptx::tma_store_wait_all0();
    if (tr && eleader) { tr[9] = w_tfull; tr[10] = w_other; tr[11] = t_work; tr[12] = clock64(); tr[15] = t_first_full; }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();  // the peer may still be reading our barriers / issuing MMAs on our TMEM
  ptx::tc_fence_after();
  if (tr && threadIdx.x == 0) tr[13] = clock64();
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
}

// Workspace of the tail split: at most one work item per pair, so 96 x 2 partial tiles of 128 x 256 fp32 (25 MB) and
// 96 x 2 counters bound it for every layer.  It is the caller's (me_conv_gemm_ws): the library allocates nothing,
// and kernels that may overlap on different streams must be given different workspaces.
constexpr int kMaxPairs = 96;
constexpr size_t kWsPartialBytes = static_cast<size_t>(kMaxPairs) * 2 * kBM * 256 * sizeof(float);
constexpr size_t kWsBytes = kWsPartialBytes + kMaxPairs * 2 * sizeof(int);

// ME_PAIR_SPLIT=0 disables the tail split (A/B measurements).
bool tail_split_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ME_PAIR_SPLIT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int BN>
int launch_pair(const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                void* tail_ws, cudaStream_t stream) {
  using C = PCfg<BN>;
  const int pad = (d->ksize - 1) / 2;
  const int Ho = (d->h + 2 * pad - d->ksize) / d->stride + 1;
  const int Wo = (d->w + 2 * pad - d->ksize) / d->stride + 1;
  const int M = d->n * Ho * Wo;
  const int taps = d->ksize * d->ksize;
  const int cin_pad = round_up(d->cin, kBK);
  const int ktot = taps * cin_pad;

  PairParams p{};
  p.M = M;
  p.Ho = Ho;
  p.Wo = Wo;
  p.stride = d->stride;
  p.pad = pad;
  p.kb_per_tap = cin_pad / kBK;
  p.num_kb = taps * p.kb_per_tap;
  p.tiles_m = ceil_div(M, 2 * kBM);
  p.tiles_n = ceil_div(d->cout, BN);
  p.act = d->act;
  p.has_res = (d->res_pitch > 0 && residual != nullptr) ? 1 : 0;
  p.im2col = (d->ksize == 3) ? 1 : 0;
  p.bias = bias;
  int rc = conv_ensure_debug_word();
  if (rc != ME_OK) return rc;
  p.debug = conv_debug_word();
  p.trace = conv_trace_buffer();
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("ME_CONV_DBG");
      dbg = e ? atoi(e) : 0;
    }
    p.dbg = dbg;
  }

  int stages = (227 * 1024 - 1024 - C::STAGING_BYTES - C::TAIL_BYTES) / C::STAGE_BYTES;
  if (stages > kMaxStages) stages = kMaxStages;
  {
    static int cap = -1;  // ME_PAIR_STAGES: pipeline-depth sensitivity runs
    if (cap < 0) {
      const char* e = getenv("ME_PAIR_STAGES");
      cap = e ? atoi(e) : 0;
    }
    if (cap >= 2 && cap < stages) stages = cap;
  }
  ME_REQUIRE(stages >= 2, "conv(pair): not enough shared memory");
  p.stages = stages;
  {
    static int pf = -1;
    if (pf < 0) {
      const char* e = getenv("ME_CONV_PREFETCH");
      pf = (e && e[0] == '0') ? 0 : 1;
    }
    p.prefetch_kb = pf ? (p.num_kb < stages ? p.num_kb : stages) : 0;
  }
  const int smem = C::smem_bytes(stages);

  CUtensorMap tmA, tmB, tmC, tmR;
  if (p.im2col) {
    rc = encode_im2col_nhwc(&tmA, x, d->n, d->h, d->w, d->cin, d->in_pitch, d->ksize, pad, d->stride, kBK, kBM,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    rc = encode_tiled_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x, d->cin, M, d->in_pitch, kBK, kBM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc != ME_OK) return rc;
  rc = encode_tiled_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, ktot, d->cout, ktot, kBK, BN / 2,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != ME_OK) return rc;
  rc = encode_tiled_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y, d->cout, M, d->out_pitch, C::SUB_COLS, kBM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != ME_OK) return rc;
  if (p.has_res) {
    rc = encode_tiled_2d(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, residual, d->cout, M, d->res_pitch, C::SUB_COLS, kBM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != ME_OK) return rc;
  } else {
    tmR = tmC;
  }

  auto kern = conv_gemm_pair_kernel<BN>;
  static bool attr_seen[64] = {false};   // per instantiation and per device
  if (first_use_on_device(attr_seen))
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int total = p.tiles_m * p.tiles_n;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int pairs = sms / 2;
  if (pairs > total) pairs = total;
  // tail split: worth it when the last wave leaves at least half of the pairs idle and a slice keeps >= 4 K blocks
  p.whole_limit = total;
  p.tail_tiles = 0;
  p.slices = 1;
  const int tail = total % pairs;
  // (measured, profiles/round1/attr_r1i.log: 13^2 512->1024, 14 tail tiles x 4 slices: 48.5 vs 52.8 us; 26^2 256->512,
  // 22 tail tiles x 3 slices: 46.9 vs 44.2 us - the reduction by the last arriver eats the gain unless the tail is small)
  if (tail_split_enabled() && tail > 0 && 4 * tail <= pairs && pairs <= kMaxPairs) {
    int slices = pairs / tail;
    if (slices > p.num_kb / 4) slices = p.num_kb / 4;
    if (slices > 4) slices = 4;   // the finisher reads every slice's partial tile: beyond 4 that costs what the split saves
    if (slices >= 4 && tail_ws != nullptr) {
      p.whole_limit = total - tail;
      p.tail_tiles = tail;
      p.slices = slices;
      p.ws = static_cast<float*>(tail_ws);
      p.counters = reinterpret_cast<int*>(static_cast<unsigned char*>(tail_ws) + kWsPartialBytes);
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = conv_pdl_enabled() ? 1 : 0;
  ME_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, p));
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // namespace

// Entry used by me_conv_gemm's dispatcher. bn is 128 or 256.
int conv_gemm_pair(int bn, const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                   void* y, void* tail_ws, cudaStream_t stream) {
  if (bn == 256) return launch_pair<256>(d, x, w, bias, residual, y, tail_ws, stream);
  return launch_pair<128>(d, x, w, bias, residual, y, tail_ws, stream);
}

size_t conv_pair_workspace_bytes() { return kWsBytes; }

}  // namespace me

extern "C" {

size_t me_conv_workspace_bytes(void) { return me::conv_pair_workspace_bytes(); }

}  // extern "C"
