// Fused conv + (folded BN) bias + activation (+ residual) as an implicit GEMM on tcgen05.
//
//   D[M = N*Ho*Wo pixels][Cout] = A[M][K = taps*Cin] * W[Cout][K]^T
//
// One persistent CTA per SM walks 128 x BN output tiles.  Warp roles:
//   warp 0   : TMA producer  - A tile: im2col-mode TMA (3x3, any stride, zero fill at the image
//              border) or tiled TMA (1x1); B tile: tiled TMA over the packed weights.
//   warp 1   : allocates TMEM, one elected lane issues tcgen05.mma (M=128, N=BN, K=16) and
//              commits to the mbarriers that free smem stages / publish the accumulator.
//   warps 2-5: epilogue - tcgen05.ld (lane quarter = warp % 4) -> +bias -> activation ->
//              (+ residual, TMA-loaded into the staging tile) -> fp16/fp32 -> swizzled smem
//              staging -> TMA store.  Accumulators are double-buffered in TMEM so the epilogue
//              of tile i overlaps the main loop of tile i+1.
//
// Replaces the conv/BN/LeakyReLU Sequential of yolov3/models.py:22-41 executed at :252-253,
// the shortcut add at :258-260, cnn_layers_1 (my_models.py:47-77), cnn_layers_3 (:130-157)
// and refinement_head.net0 (:238-241).
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace me {

int conv_gemm_pair(int bn, const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                   void* y, void* tail_ws, cudaStream_t stream);  // conv_gemm_pair.cu
size_t conv_pair_workspace_bytes();
bool conv_thin_enabled();                          // conv_thin.cu
bool conv_thin_supported(const me_conv_desc* d);
bool conv_thin_pool_supported(const me_conv_desc* d);
int conv_thin_pool(const me_conv_desc* d, const void* x, const void* w, const float* bias, void* y, cudaStream_t stream);
int conv_thin(const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
              cudaStream_t stream);

namespace {

constexpr int kBM = 128;
constexpr int kEpiThreads = 128;        // one epilogue group = 4 warps = the four TMEM lane quarters
constexpr uint32_t kEpiBarrierId = 1;   // named barrier of group g is kEpiBarrierId + g
constexpr int kMaxStages = 8;

// YOLO decode fused into the epilogue of a head conv (me_conv_gemm_yolo): out == nullptr means a plain conv.
struct DecodeParams {
  float* out;        // [n][rows_total][attrs] fp32
  int na, attrs;     // anchors of this head, 5 + classes
  int g, gg;         // grid size, g * g
  int rows_total, row_offset;
  float stride;
  float aw[8], ah[8];  // anchors / stride, rounded to fp32 like models.py:127
};

struct ConvParams {
  int M;
  int Ho, Wo;
  int stride, pad;
  int kb_per_tap;  // K blocks per filter tap (= cin_pad / BK)
  int num_kb;      // taps * kb_per_tap
  int tiles_m, tiles_n;
  int act;
  int has_res;
  int im2col;
  int stages;
  int bres_bytes;  // weight-stationary mode: bytes of the resident weight slab (num_kb * B_BYTES)
  int kps;         // K blocks per pipeline stage (one mbarrier round trip)
  int prefetch_kb; // weight K blocks prefetched into L2 before griddepcontrol.wait
  int dbg;         // ME_CONV_DBG attribution mask: 1 skip epilogue, 2 skip operand loads, 4 skip MMAs, 16 no stage commits
  const float* bias;
  unsigned long long* debug;  // host-mapped word, written before a watchdog trap
  unsigned long long* trace;  // me_conv_set_trace: 16 clock64 words per CTA (tools/conv_trace.py), else nullptr
  DecodeParams dec;
};

// Timed wait for the trace mode: adds the cycles spent in the wait to *acc.
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

// Tile order of one persistent CTA.  Default: tile = blockIdx.x + i * gridDim.x over the (m, n) tile grid,
// n fastest, so the CTAs that share an A tile run together.  Weight-stationary: the CTA keeps ONE n tile
// (its weight slab stays in shared memory) and strides over the m tiles.
template <bool WS>
struct TileWalk {
  int tm, tn, step, tiles_m, tiles_n, tile;
  __device__ TileWalk(int tiles_m_, int tiles_n_) : tiles_m(tiles_m_), tiles_n(tiles_n_) {
    if (WS) {
      tn = blockIdx.x % tiles_n;
      tm = blockIdx.x / tiles_n;
      step = gridDim.x / tiles_n;
      tile = 0;
    } else {
      tile = blockIdx.x;
      step = gridDim.x;
      tm = tile / tiles_n;
      tn = tile - tm * tiles_n;
    }
  }
  __device__ bool valid() const { return WS ? tm < tiles_m : tile < tiles_m * tiles_n; }
  __device__ void next() {
    if (WS) {
      tm += step;
    } else {
      tile += step;
      tm = tile / tiles_n;
      tn = tile - tm * tiles_n;
    }
  }
};

template <int BN, int BK, bool OUT_F32, int EG_>
struct Cfg {
  static constexpr int A_BYTES = kBM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ESIZE = OUT_F32 ? 4 : 2;
  static constexpr int SUB_ROW_BYTES = (BN * ESIZE >= 128) ? 128 : BN * ESIZE;
  static constexpr int SUB_COLS = SUB_ROW_BYTES / ESIZE;
  static constexpr int NUM_SUB = BN / SUB_COLS;
  static constexpr int SUB_BYTES = kBM * SUB_ROW_BYTES;
  static constexpr int STAGING_BYTES = NUM_SUB * SUB_BYTES;  // per epilogue group
  // Two epilogue groups alternate tiles (group g owns TMEM accumulator g and its own staging tile): on the
  // thin layers one group's tcgen05.ld -> math -> st.shared -> TMA store chain (plus the residual load) took
  // longer than the tile's MMAs, so the tensor pipe idled on tmem_empty (tools/conv_trace.py, profiles/round1).
  // A second staging tile costs pipeline stages, so the dispatcher (me_conv_gemm) picks EG per layer.
  static constexpr int EG = EG_;
  static_assert(EG == 1 || EG == 2, "epilogue groups");
  // Eight epilogue warps either way: with one group both warps of a TMEM lane quarter work on the same tile and
  // split its columns (halves the exposed epilogue of the last tile and the accumulator hold time).
  static constexpr int THREADS = 64 + 2 * kEpiThreads;
  static constexpr int GROUP_THREADS = (EG == 2) ? kEpiThreads : 2 * kEpiThreads;
  static constexpr int WARP_COLS = (EG == 2) ? BN : BN / 2;   // columns one epilogue warp converts per tile
  static_assert(WARP_COLS % 32 == 0, "an epilogue warp converts whole 32-column chunks");
  static constexpr int SWZ_BITS = (SUB_ROW_BYTES == 128) ? 3 : (SUB_ROW_BYTES == 64) ? 2 : 1;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int TAIL_BYTES = EG * BN * 4 + 64 * 8 + 16;  // bias (per group) + barriers + tmem ptr
  static constexpr int EPI_BYTES = EG * STAGING_BYTES + TAIL_BYTES;
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
  static_assert(BK == 16 || BK == 32 || BK == 64, "BK");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "tiles must keep 1024B alignment");
};

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long* dbg, uint32_t tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
      if (dbg) {
        *reinterpret_cast<volatile unsigned long long*>(dbg) =
            (static_cast<unsigned long long>(tag) << 32) | (static_cast<unsigned long long>(blockIdx.x) << 8) | parity | 0x80u;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// WS (weight stationary): the layer's whole weight slab for this CTA's n tile (num_kb x BN x BK) is loaded
// into shared memory once and the pipeline stages carry A only.  For the thin-channel layers (32..128
// channels, 208^2..52^2 pixels) re-fetching the weights with every tile doubled the L2 -> SM traffic.
template <int BN, int BK, bool OUT_F32, bool WS, int EG>
__global__ void __launch_bounds__((Cfg<BN, BK, OUT_F32, EG>::THREADS), 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const ConvParams p) {
  using C = Cfg<BN, BK, OUT_F32, EG>;
  constexpr int STAGE = WS ? C::A_BYTES : C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bres = smem;  // WS only
  uint8_t* stage_base = smem + (WS ? p.bres_bytes : 0);
  uint8_t* staging = stage_base + p.stages * p.kps * STAGE;
  float* s_bias = reinterpret_cast<float*>(staging + C::EG * C::STAGING_BYTES);  // [EG][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + C::EG * BN);
  uint64_t* full_bar = bars;                     // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;       // [kMaxStages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* res_full = tmem_empty + 2;           // [2] one per epilogue group
  uint64_t* b_full = res_full + 2;               // [1] WS: resident weights landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* tr = p.trace ? p.trace + 16ull * blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  // PDL: let the next layer's CTAs get scheduled as ours retire; everything up to pdl_wait() below touches
  // only our own smem / TMEM / kernel parameters, so it overlaps the previous layer's tail.
  ptx::pdl_launch_dependents();
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
        ptx::mbar_init(&tmem_empty[a], C::GROUP_THREADS / 32);  // one arrive per epilogue warp of the group
      }
      ptx::mbar_init(&res_full[0], 1);
      ptx::mbar_init(&res_full[1], 1);
      ptx::mbar_init(b_full, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tr && threadIdx.x == 0) tr[1] = clock64();
  // The weights do not depend on the previous layer: pull this CTA's first weight tiles into L2 while the previous
  // kernel drains (ME_CONV_PREFETCH=0 disables; p.prefetch_kb = 0).
  if (warp == 0 && p.prefetch_kb > 0 && ptx::elect_one()) {
    TileWalk<WS> tw0(p.tiles_m, p.tiles_n);
    if (tw0.valid())
      for (int kb = 0; kb < p.prefetch_kb; ++kb) ptx::tma_prefetch_2d(&tmB, kb * BK, tw0.tn * BN);
  }
  ptx::pdl_wait();  // inputs written by the previous kernel are complete and visible from here on
  if (tr && threadIdx.x == 0) tr[2] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      TileWalk<WS> tw(p.tiles_m, p.tiles_n);
      if (WS && tw.valid()) {
        ptx::mbar_arrive_expect_tx(b_full, p.bres_bytes);
        for (int kb = 0; kb < p.num_kb; ++kb) ptx::tma_load_2d(&tmB, b_full, bres + kb * C::B_BYTES, kb * BK, tw.tn * BN);
      }
      for (; tw.valid(); tw.next()) {
        const int m0 = tw.tm * kBM, n0 = tw.tn * BN;
        int cw = 0, ch = 0, cn = 0;
        if (p.im2col) {
          const int q0 = m0 % p.Wo;
          const int t = m0 / p.Wo;
          cw = q0 * p.stride - p.pad;
          ch = (t % p.Ho) * p.stride - p.pad;
          cn = t / p.Ho;
        }
        int tap = 0, cb = 0;
        // One barrier round trip (wait empty -> expect_tx -> loads) covers p.kps consecutive K blocks: a thin
        // layer's K block is only a few dozen MMA clocks, far less than the ~300 clocks a round trip costs.
        for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.kps) {
          if (!(p.dbg & 16)) ME_TRACED_WAIT(w_empty, &empty_bar[stage], phase ^ 1, p.debug, 0x100u + stage);
          if (p.dbg & 16) {  // attribution run: no stage barriers at all (with bit 2)
          } else if (p.dbg & 2) {  // attribution run: no operand traffic, barrier protocol intact
            ptx::mbar_arrive(&full_bar[stage]);
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], p.kps * STAGE);
          }
          for (int j = 0; j < p.kps; ++j) {
            const int kb = kb0 + j;
            uint8_t* sa = stage_base + (stage * p.kps + j) * STAGE;
            if (!(p.dbg & 2)) {
              if (p.im2col) {
                const int r = tap / 3, s = tap - r * 3;
                ptx::tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * BK, cw, ch, cn, (uint16_t)s, (uint16_t)r);
              } else {
                ptx::tma_load_2d(&tmA, &full_bar[stage], sa, cb * BK, m0);
              }
              if (!WS) ptx::tma_load_2d(&tmB, &full_bar[stage], sa + C::A_BYTES, kb * BK, n0);
            }
            if (++cb == p.kb_per_tap) { cb = 0; ++tap; }
          }
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (tr) { tr[3] = w_empty; tr[4] = clock64(); }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, BN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      long long w_full = 0, w_acc = 0, t_first = 0;
      TileWalk<WS> tw(p.tiles_m, p.tiles_n);
      if (WS && tw.valid()) mbar_wait(b_full, 0, p.debug, 0x600u);
      for (; tw.valid(); tw.next(), ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ME_TRACED_WAIT(w_acc, &tmem_empty[acc], acc_phase ^ 1, p.debug, 0x200u + acc);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.kps) {
          if (!(p.dbg & 16)) ME_TRACED_WAIT(w_full, &full_bar[stage], phase, p.debug, 0x300u + stage);
          if (tr && t_first == 0) { t_first = clock64(); w_full = 0; }
          ptx::tc_fence_after();
          for (int j = 0; j < p.kps; ++j) {
            const int kb = kb0 + j;
            const uint32_t a_addr = ptx::smem_u32(stage_base + (stage * p.kps + j) * STAGE);
            const uint32_t b_addr = WS ? ptx::smem_u32(bres + kb * C::B_BYTES) : a_addr + C::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (p.dbg & 4) break;  // attribution run: no tensor work
              const uint64_t adesc = ptx::make_kmajor_desc(a_addr + k * 32, BK * 2);
              const uint64_t bdesc = ptx::make_kmajor_desc(b_addr + k * 32, BK * 2);
              ptx::umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          // (attribution bit 16, with bit 2: no per-stage commit and the producer does not wait for one)
          if (!(p.dbg & 16)) ptx::umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
      if (tr) { tr[5] = w_full; tr[6] = w_acc; tr[7] = t_first; tr[8] = clock64(); tr[14] = it; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 [, 6..9])
    const int g = (C::EG == 2) ? ((warp - 2) >> 2) : 0;  // epilogue group
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                       // tile row == TMEM lane
    const int c_base = (C::EG == 2) ? 0 : ((warp - 2) >> 2) * C::WARP_COLS;  // first column of this warp's share
    const int etid = (threadIdx.x - 64) & (C::GROUP_THREADS - 1);
    const bool leader = (etid == 0);
    const bool tleader = leader && g == 0;               // trace words come from group 0
    uint8_t* stg = staging + g * C::STAGING_BYTES;
    float* bias_s = s_bias + g * BN;
    uint64_t* res_bar = &res_full[g];
    const uint32_t bar_id = kEpiBarrierId + g;
    long long w_tfull = 0, w_other = 0, t_work = 0, t_first_full = 0;
    TileWalk<WS> tw(p.tiles_m, p.tiles_n);
    if (C::EG == 2 && g == 1) tw.next();
    // The residual tile is TMA-loaded into the staging tile the result will overwrite.  The load for a group's
    // NEXT tile is issued right after the store of the current one has drained the staging tile, a whole tile
    // period before it is needed.
    auto load_residual = [&](const TileWalk<WS>& t) {
      ptx::mbar_arrive_expect_tx(res_bar, C::STAGING_BYTES);
#pragma unroll
      for (int sub = 0; sub < C::NUM_SUB; ++sub)
        ptx::tma_load_2d(&tmR, res_bar, stg + sub * C::SUB_BYTES, t.tn * BN + sub * C::SUB_COLS, t.tm * kBM);
    };
    if (p.has_res && leader && tw.valid()) load_residual(tw);
    for (int lit = 0; tw.valid(); ++lit) {
      const int m0 = tw.tm * kBM, n0 = tw.tn * BN;
      const int acc = (C::EG == 2) ? g : (lit & 1);
      const uint32_t acc_phase = (C::EG == 2) ? (lit & 1) : ((lit >> 1) & 1);
      const long long te0 = (tr && tleader) ? clock64() : 0;
      if (leader && !p.has_res) ptx::tma_store_wait_read0();  // previous store has drained the staging tile
      for (int i = etid; i < BN; i += C::GROUP_THREADS) bias_s[i] = p.bias[n0 + i];
      ptx::named_bar_sync(bar_id, C::GROUP_THREADS);
      const long long te1 = (tr && tleader) ? clock64() : 0;
      mbar_wait(&tmem_full[acc], acc_phase, p.debug, 0x400u + acc);
      ptx::tc_fence_after();
      const long long te2 = (tr && tleader) ? clock64() : 0;
      if (p.has_res) mbar_wait(res_bar, lit & 1, p.debug, 0x500u + g);
      const long long te3 = (tr && tleader) ? clock64() : 0;
      if (tr && tleader) {
        w_other += (te1 - te0) + (te3 - te2);
        w_tfull += te2 - te1;
        if (lit == 0) t_first_full = te2;
      }

      const uint32_t t_row = tmem_base + acc * BN + c_base + (static_cast<uint32_t>(q * 32) << 16);
      if (!(p.dbg & 1)) {  // attribution run: accumulators released unread
        constexpr int NCH = C::WARP_COLS / 32;
        uint32_t r[2][32];
        ptx::tmem_ld_32x32b_x32(t_row, r[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = c_base + ci * 32;
          ptx::tmem_ld_wait_regs(r[ci & 1]);
          if (ci + 1 < NCH) ptx::tmem_ld_32x32b_x32(t_row + (ci + 1) * 32, r[(ci + 1) & 1]);  // in flight during the math
          float v[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + 4 * j4);
            v[4 * j4 + 0] = __uint_as_float(r[ci & 1][4 * j4 + 0]) + b4.x;
            v[4 * j4 + 1] = __uint_as_float(r[ci & 1][4 * j4 + 1]) + b4.y;
            v[4 * j4 + 2] = __uint_as_float(r[ci & 1][4 * j4 + 2]) + b4.z;
            v[4 * j4 + 3] = __uint_as_float(r[ci & 1][4 * j4 + 3]) + b4.w;
          }
          if (p.act == ME_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);  // == v > 0 ? v : 0.1 v
          } else if (p.act == ME_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
          }
          uint8_t* sub = stg + (c / C::SUB_COLS) * C::SUB_BYTES;
          if constexpr (OUT_F32) {
            // 32 fp32 columns = one 128B staging row of sub-tile c/32
            const uint32_t rbase = row * C::SUB_ROW_BYTES + (c % C::SUB_COLS) * 4;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t off = rbase + j * 16;
              off ^= ((off >> 7) & ((1u << C::SWZ_BITS) - 1)) << 4;
              *reinterpret_cast<float4*>(sub + off) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          } else {
            const uint32_t rbase = row * C::SUB_ROW_BYTES + (c % C::SUB_COLS) * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t off = rbase + j * 16;
              off ^= ((off >> 7) & ((1u << C::SWZ_BITS) - 1)) << 4;
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
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);  // accumulator may be overwritten
      ptx::fence_proxy_async_smem();                       // staging writes -> visible to TMA
      ptx::named_bar_sync(bar_id, C::GROUP_THREADS);
      tw.next();
      if (C::EG == 2) tw.next();
      if constexpr (OUT_F32) {
        if (p.dec.out != nullptr) {
          // YOLOLayer.forward (models.py:142-177) on the staged logits: the group's warps walk the tile row by row,
          // a lane per head channel, so the writes of one row are contiguous over the 5+C attributes of an output
          // row (anchor-major, then gy, gx).  Same arithmetic as yolo_decode_kernel (simt_ops.cu).
          const DecodeParams& dp = p.dec;
          const int head_ch = dp.na * dp.attrs;
          for (int r = etid >> 5; r < kBM; r += C::GROUP_THREADS / 32) {
            const int m = m0 + r;
            if (m >= p.M) break;
            const int img = m / dp.gg, rem = m - img * dp.gg;
            const int gy = rem / dp.g, gx = rem - gy * dp.g;
            float* orow = dp.out + (static_cast<size_t>(img) * dp.rows_total + dp.row_offset + rem) * dp.attrs;
            for (int c = lane; c < BN; c += 32) {
              const int col = n0 + c;
              if (col >= head_ch) break;
              const int a = col / dp.attrs, k = col - a * dp.attrs;
              uint32_t off = r * C::SUB_ROW_BYTES + (c % C::SUB_COLS) * 4;
              off ^= ((off >> 7) & ((1u << C::SWZ_BITS) - 1)) << 4;
              const float v = *reinterpret_cast<const float*>(stg + (c / C::SUB_COLS) * C::SUB_BYTES + off);
              float o;
              if (k < 2) o = (1.f / (1.f + expf(-v)) + static_cast<float>(k == 0 ? gx : gy)) * dp.stride;
              else if (k < 4) o = (expf(v) * (k == 2 ? dp.aw[a] : dp.ah[a])) * dp.stride;
              else o = 1.f / (1.f + expf(-v));
              orow[static_cast<size_t>(a) * dp.gg * dp.attrs + k] = o;
            }
          }
          ptx::named_bar_sync(bar_id, C::GROUP_THREADS);  // the staging tile may be overwritten
        }
      }
      if (leader) {
        if (!(p.dbg & 1) && !(OUT_F32 && p.dec.out != nullptr)) {
#pragma unroll
          for (int sub = 0; sub < C::NUM_SUB; ++sub)
            ptx::tma_store_2d(&tmC, stg + sub * C::SUB_BYTES, n0 + sub * C::SUB_COLS, m0);
          ptx::tma_store_commit();
        }
        if (p.has_res && tw.valid()) {
          ptx::tma_store_wait_read0();
          load_residual(tw);
        }
      }
      if (tr && tleader) t_work += clock64() - te3;
    }
    if (leader) ptx::tma_store_wait_all0();
    if (tr && tleader) { tr[9] = w_tfull; tr[10] = w_other; tr[11] = t_work; tr[12] = clock64(); tr[15] = t_first_full; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tr && threadIdx.x == 0) tr[13] = clock64();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ME_PDL=0 disables programmatic dependent launch (A/B measurements).
bool pdl_enabled_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ME_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
inline bool pdl_enabled() { return pdl_enabled_impl(); }

bool conv_kps_enabled();

// K blocks handled per pipeline stage (= per mbarrier round trip of producer and MMA issuer).  Attribution runs
// (profiles/round1/attribution_early.log) showed a round trip costs ~300 clocks on each side while a 128 x BN x 64
// K block is 64..256 tensor clocks on one CTA, so thin layers were bound by barrier latency, not by loads or MMAs.
// Prefer a whole filter row (3 taps) for 3x3 layers whose tap is one K block, else the largest of 4 / 2 that
// divides the K loop, as long as >= 3 stages (>= 2 for the filter-row case) still fit.  ME_CONV_KPS=1: off.
int choose_kps(int ksize, int kb_per_tap, int num_kb, int stage_bytes, int budget) {
  if (!conv_kps_enabled()) return 1;
  if (ksize == 3 && kb_per_tap == 1 && budget / (3 * stage_bytes) >= 2) return 3;
  for (int k : {4, 2})
    if (num_kb % k == 0 && budget / (k * stage_bytes) >= 3) return k;
  return 1;
}

bool conv_kps_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ME_CONV_KPS");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

unsigned long long* g_debug_host = nullptr;
unsigned long long* g_debug_dev = nullptr;
unsigned long long* g_trace_dev = nullptr;  // me_conv_set_trace

int ensure_debug_word() {
  if (g_debug_host) return ME_OK;
  ME_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_debug_host), sizeof(unsigned long long), cudaHostAllocMapped));
  *g_debug_host = 0;
  ME_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_debug_dev), g_debug_host, 0));
  return ME_OK;
}

template <int BN, int BK, bool OUT_F32, bool WS, int EG>
int launch(const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
           cudaStream_t stream, const DecodeParams* dec = nullptr) {
  using C = Cfg<BN, BK, OUT_F32, EG>;
  const int pad = (d->ksize - 1) / 2;
  const int Ho = (d->h + 2 * pad - d->ksize) / d->stride + 1;
  const int Wo = (d->w + 2 * pad - d->ksize) / d->stride + 1;
  const long long M64 = 1LL * d->n * Ho * Wo;
  ME_REQUIRE(M64 > 0 && M64 < (1LL << 31), "conv: pixel count out of range");
  const int M = static_cast<int>(M64);
  const int taps = d->ksize * d->ksize;
  const int cin_pad = round_up(d->cin, BK);
  const int ktot = taps * cin_pad;

  ConvParams p{};
  p.M = M;
  p.Ho = Ho;
  p.Wo = Wo;
  p.stride = d->stride;
  p.pad = pad;
  p.kb_per_tap = cin_pad / BK;
  p.num_kb = taps * p.kb_per_tap;
  p.tiles_m = ceil_div(M, kBM);
  p.tiles_n = ceil_div(d->cout, BN);
  p.act = d->act;
  p.has_res = (d->res_pitch > 0 && residual != nullptr) ? 1 : 0;
  p.im2col = (d->ksize == 3) ? 1 : 0;
  p.bias = bias;
  int rc = ensure_debug_word();
  if (rc != ME_OK) return rc;
  p.debug = g_debug_dev;
  p.trace = g_trace_dev;
  if (dec) p.dec = *dec;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("ME_CONV_DBG");
      dbg = e ? atoi(e) : 0;
    }
    p.dbg = dbg;
  }

  constexpr int STAGE = WS ? C::A_BYTES : C::STAGE_BYTES;
  p.bres_bytes = WS ? p.num_kb * C::B_BYTES : 0;
  const int budget = 227 * 1024 - 1024 - C::EPI_BYTES - p.bres_bytes;
  p.kps = choose_kps(d->ksize, p.kb_per_tap, p.num_kb, STAGE, budget);
  int stages = budget / (p.kps * STAGE);
  if (stages > kMaxStages) stages = kMaxStages;
  ME_REQUIRE(stages >= 2, "conv: not enough shared memory for a 2-stage pipeline");
  p.stages = stages;
  {
    static int pf = -1;
    if (pf < 0) {
      const char* e = getenv("ME_CONV_PREFETCH");
      pf = (e && e[0] == '0') ? 0 : 1;
    }
    p.prefetch_kb = pf ? (p.num_kb < stages * p.kps ? p.num_kb : stages * p.kps) : 0;
  }
  const int smem = 1024 + p.bres_bytes + stages * p.kps * STAGE + C::EPI_BYTES;

  CUtensorMap tmA, tmB, tmC, tmR;
  const CUtensorMapSwizzle swz_k = swizzle_for_row_bytes(BK * 2);
  if (p.im2col) {
    rc = encode_im2col_nhwc(&tmA, x, d->n, d->h, d->w, d->cin, d->in_pitch, d->ksize, pad, d->stride, BK, kBM, swz_k);
  } else {
    rc = encode_tiled_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x, d->cin, M, d->in_pitch, BK, kBM, swz_k);
  }
  if (rc != ME_OK) return rc;
  // weights: [cout rows][ktot cols]; rows past cout are zero-filled by TMA.
  rc = encode_tiled_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, ktot, d->cout, ktot, BK, BN, swz_k);
  if (rc != ME_OK) return rc;
  const CUtensorMapSwizzle swz_o = swizzle_for_row_bytes(C::SUB_ROW_BYTES);
  rc = encode_tiled_2d(&tmC, OUT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C::ESIZE, y,
                       d->cout, M, d->out_pitch, C::SUB_COLS, kBM, swz_o);
  if (rc != ME_OK) return rc;
  if (p.has_res) {
    ME_REQUIRE(!OUT_F32, "conv: residual add only with fp16 output");
    rc = encode_tiled_2d(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, residual, d->cout, M, d->res_pitch, C::SUB_COLS, kBM,
                         swz_o);
    if (rc != ME_OK) return rc;
  } else {
    tmR = tmC;
  }

  auto kern = conv_gemm_kernel<BN, BK, OUT_F32, WS, EG>;
  static bool attr_seen[64] = {false};   // per instantiation and per device
  if (first_use_on_device(attr_seen))
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int total = p.tiles_m * p.tiles_n;
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (WS) {
    int per_n = grid / p.tiles_n;  // CTAs that share one n tile
    if (per_n > p.tiles_m) per_n = p.tiles_m;
    ME_REQUIRE(per_n >= 1, "conv(ws): more n tiles than SMs");
    grid = per_n * p.tiles_n;
  } else if (grid > total) {
    grid = total;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  ME_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, p));
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // namespace

unsigned long long* conv_debug_word() { return g_debug_dev; }
unsigned long long* conv_trace_buffer() { return g_trace_dev; }
bool conv_pdl_enabled() { return pdl_enabled_impl(); }
// ME_CONV_WS: 0 = never, 1 (default) = when profitable, 2 = whenever the slab fits (tests force it on small shapes)
static int conv_ws_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ME_CONV_WS");
    // per-layer probes (profiles/round1/kps_ws_probe.log): 208^2 3x3 32->64 goes 222 -> 165 us with resident weights
    v = e ? (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)) : 1;
  }
  return v;
}
static bool conv_ws_enabled() { return conv_ws_mode() != 0; }

int conv_ensure_debug_word() { return ensure_debug_word(); }

// 0 = automatic, 1 = single-CTA tiles only, 2 = CTA pairs with N=128, 3 = CTA pairs with N=256 (where legal)
static int conv_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("ME_CONV_MODE");
    mode = 0;
    if (e) {
      if (!strcmp(e, "single")) mode = 1;
      else if (!strcmp(e, "pair128")) mode = 2;
      else if (!strcmp(e, "pair256")) mode = 3;
    }
  }
  return mode;
}

}  // namespace me

extern "C" {

int me_conv_k_block(int cin) { return cin > 32 ? 64 : (cin > 16 ? 32 : 16); }
int me_conv_cin_pad(int cin) { return me::round_up(cin, me_conv_k_block(cin)); }

int me_conv_set_trace(unsigned long long* dev_words) {
  me::g_trace_dev = dev_words;
  return ME_OK;
}

int me_debug_status(unsigned long long* host_word) {
  if (host_word) *host_word = me::g_debug_host ? *me::g_debug_host : 0ull;
  return ME_OK;
}

int me_conv_pool_supported(const me_conv_desc* d) { return (d && me::conv_thin_pool_supported(d)) ? 1 : 0; }

// 3x3 / stride-1 conv + bias + activation followed by MaxPool2d(2, 2) in one kernel (thin layers: 16 / 32 input channels,
// 32 / 64 filters); y is the pooled (n, h/2, w/2, out_pitch) tensor.
int me_conv_pool(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y, me_stream_t stream_) {
  using namespace me;
  ME_REQUIRE(d && x && w_packed && bias && y, "conv_pool: null argument");
  ME_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "conv_pool: empty input");
  if (!conv_thin_pool_supported(d))
    return fail(ME_ERR_UNSUPPORTED, "conv_pool: unsupported layer (3x3 stride 1, cin 16/32, cout 32/64, no residual, "
                "w %% 8 == 0, h even); run me_conv_gemm + me_maxpool2");
  ME_REQUIRE(d->out_pitch >= d->cout && d->out_pitch % 8 == 0, "conv_pool: bad out_pitch %d", d->out_pitch);
  return conv_thin_pool(d, x, w_packed, bias, y, static_cast<cudaStream_t>(stream_));
}

static int conv_dispatch(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                         const void* residual, void* y, me_stream_t stream_, const me::DecodeParams* dec,
                         void* tail_ws = nullptr) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(d && x && w_packed && bias && y, "conv: null argument");
  ME_REQUIRE(d->ksize == 1 || d->ksize == 3, "conv: ksize %d unsupported (1 or 3)", d->ksize);
  ME_REQUIRE(d->stride == 1 || (d->stride == 2 && d->ksize == 3), "conv: stride %d with ksize %d unsupported", d->stride,
             d->ksize);
  ME_REQUIRE(d->cout > 0 && d->cout % 32 == 0, "conv: cout %d must be a positive multiple of 32", d->cout);
  ME_REQUIRE(d->cin > 0 && d->in_pitch >= d->cin && d->in_pitch % 8 == 0, "conv: bad cin/in_pitch %d/%d", d->cin,
             d->in_pitch);
  ME_REQUIRE(d->out_pitch >= d->cout && (d->out_pitch * (d->out_f32 ? 4 : 2)) % 16 == 0, "conv: bad out_pitch %d",
             d->out_pitch);
  ME_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "conv: empty input");
  // 3x3 layers over 16 / 32 channels: SIMT im2col producers (TMA delivers their 32- / 64-byte pixel rows too slowly)
  if (dec == nullptr && conv_thin_enabled() && conv_thin_supported(d)) return conv_thin(d, x, w_packed, bias, residual, y, stream);
  const int bk = me_conv_k_block(d->cin);
  const bool f32 = d->out_f32 != 0;
  const int cout = d->cout;
  const int pad_ = (d->ksize - 1) / 2;
  const long long m = 1LL * d->n * ((d->h + 2 * pad_ - d->ksize) / d->stride + 1) * ((d->w + 2 * pad_ - d->ksize) / d->stride + 1);
  const int num_kb = d->ksize * d->ksize * (round_up(d->cin, bk) / bk);
  // Weight-stationary tiles: the n tile's whole weight slab stays in shared memory.  Used when it fits next to
  // >= 3 A stages and every CTA gets >= 4 m tiles to amortise the one-off load (ME_CONV_WS=0 disables).
  auto ws_ok = [&](int bn, int eg) {
    if (f32 || !conv_ws_enabled()) return false;
    const int tiles_n = ceil_div(cout, bn), tiles_m = static_cast<int>((m + kBM - 1) / kBM);
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int a_bytes = kBM * bk * 2, b_bytes = bn * bk * 2;
    const int budget = 227 * 1024 - 1024 - eg * kBM * bn * 2 - (eg * bn * 4 + 64 * 8 + 16);  // Cfg::EPI_BYTES, fp16
    const int budget_ws = budget - num_kb * b_bytes;
    if (budget_ws < 3 * a_bytes || tiles_n > sms) return false;
    if (conv_ws_mode() == 2) return true;
    // keep the weights resident only if that does not cost K-block grouping or pipeline depth
    const int kb_per_tap = round_up(d->cin, bk) / bk;
    const int kps_ws = choose_kps(d->ksize, kb_per_tap, num_kb, a_bytes, budget_ws);
    const int kps_plain = choose_kps(d->ksize, kb_per_tap, num_kb, a_bytes + b_bytes, budget);
    if (kps_ws < kps_plain || budget_ws / (kps_ws * a_bytes) < 3) return false;
    return tiles_n <= 4 && tiles_m >= 4 * (sms / tiles_n);
  };
#define ME_GO(BN, BK, F32, EG)                                                                            \
  do {                                                                                                    \
    if (!F32 && ws_ok(BN, EG)) return launch<BN, BK, false, true, EG>(d, x, w_packed, bias, residual, y, stream); \
    return launch<BN, BK, F32, false, EG>(d, x, w_packed, bias, residual, y, stream, dec);                \
  } while (0)
  if (bk == 64 && !f32 && cout >= 128) {
    // CTA pairs (256 x BN tiles) halve the shared-memory bytes per flop; worth it once the layer has
    // enough pair tiles to fill the 74 SM pairs.
    const int mode = conv_mode();
    int pair_bn = 0;
    if (mode == 2) pair_bn = 128;
    else if (mode == 3) pair_bn = cout >= 256 ? 256 : 128;
    // measured on B200 (profiles/round1): pairs win on 3x3 layers with >= 256 output channels; 1x1 layers and
    // 128-channel layers are faster on single-CTA tiles.  (13^2 x 32 frames = 5408 rows: pair256 52.7 us vs
    // 57.2 us on single-CTA tiles, profiles/round1/attr_r1e.log)
    else if (mode == 0 && d->ksize == 3 && cout >= 256 && m >= 4096) pair_bn = 256;
    if (pair_bn) return conv_gemm_pair(pair_bn, d, x, w_packed, bias, residual, y, tail_ws, stream);
  }
  // Epilogue groups (see Cfg): two for the thin tiles (N <= 64), whose epilogue outlasts their MMAs, and for short-K
  // 1x1 layers with N = 128 (52^2 256->128: -23 %); the 3x3 and long-K N = 128 layers keep one group because the
  // second 32 KB staging tile would cost them the K-block grouping (26^2 512->256: +24 % with two groups).
  // An SS-mode tcgen05.mma costs ~100-130 clocks whatever N <= 128 is (the A-operand fetch; tools/conv_trace.py,
  // profiles/round1/attr_r1f.log), which is why the 3x3 layers with >= 256 output channels run as N = 256 pair tiles.
  // Single-CTA 128 x 256 tiles for the 1x1 layers were measured and dropped (26^2 512->256: 23.5k vs 18k clocks,
  // the 128 x 256 epilogue tail costs more than the faster main loop saves; profiles/round1/trace_r1g.txt).
  const bool eg2_128 = d->ksize == 1 && num_kb <= 6;
  if (bk == 64) {
    if (f32) {
      if (cout >= 128) ME_GO(128, 64, true, 1);
      if (cout >= 64) ME_GO(64, 64, true, 2);
      ME_GO(32, 64, true, 2);
    }
    if (cout >= 128) {
      if (eg2_128) ME_GO(128, 64, false, 2);
      ME_GO(128, 64, false, 1);
    }
    if (cout >= 64) ME_GO(64, 64, false, 2);
    ME_GO(32, 64, false, 2);
  } else if (bk == 32) {
    ME_REQUIRE(!f32, "conv: fp32 output needs cin > 32");
    if (cout >= 64) ME_GO(64, 32, false, 2);
    ME_GO(32, 32, false, 2);
  } else {
    ME_REQUIRE(!f32, "conv: fp32 output needs cin > 32");
    if (cout >= 64) ME_GO(64, 16, false, 2);
    ME_GO(32, 16, false, 2);
  }
#undef ME_GO
  return ME_OK;
}

int me_conv_gemm(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, const void* residual,
                 void* y, me_stream_t stream) {
  return conv_dispatch(d, x, w_packed, bias, residual, y, stream, nullptr);
}

int me_conv_gemm_ws(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, const void* residual,
                    void* y, void* workspace, size_t workspace_bytes, me_stream_t stream) {
  using namespace me;
  if (workspace != nullptr) {
    ME_REQUIRE(workspace_bytes >= conv_pair_workspace_bytes(), "conv workspace: %zu bytes given, %zu needed", workspace_bytes,
               conv_pair_workspace_bytes());
    ME_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "conv workspace must be 256-byte aligned");
  }
  return conv_dispatch(d, x, w_packed, bias, residual, y, stream, nullptr, workspace);
}

int me_conv_gemm_yolo(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, int g,
                      int num_anchors, int num_classes, const float* host_anchors_wh, float stride, int rows_total,
                      int row_offset, float* pred, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(d && pred && host_anchors_wh, "conv_yolo: null argument");
  ME_REQUIRE(d->out_f32 == 1 && d->res_pitch == 0 && d->act == ME_ACT_LINEAR, "conv_yolo: the head conv is linear, fp32, no residual");
  ME_REQUIRE(num_anchors >= 1 && num_anchors <= 8, "conv_yolo: 1..8 anchors");
  const int attrs = 5 + num_classes;
  ME_REQUIRE(d->cout >= num_anchors * attrs, "conv_yolo: cout %d < %d head channels", d->cout, num_anchors * attrs);
  const int pad = (d->ksize - 1) / 2;
  ME_REQUIRE((d->h + 2 * pad - d->ksize) / d->stride + 1 == g && (d->w + 2 * pad - d->ksize) / d->stride + 1 == g,
             "conv_yolo: the conv output is not a %d x %d grid", g, g);
  ME_REQUIRE(rows_total >= row_offset + num_anchors * g * g, "conv_yolo: rows_total too small");
  DecodeParams dec{};
  dec.out = pred;
  dec.na = num_anchors;
  dec.attrs = attrs;
  dec.g = g;
  dec.gg = g * g;
  dec.rows_total = rows_total;
  dec.row_offset = row_offset;
  dec.stride = stride;
  for (int a = 0; a < num_anchors; ++a) {
    // models.py:127 keeps anchors / stride as float32 (FloatTensor of python doubles), :171 multiplies back
    dec.aw[a] = static_cast<float>(static_cast<double>(host_anchors_wh[2 * a]) / static_cast<double>(stride));
    dec.ah[a] = static_cast<float>(static_cast<double>(host_anchors_wh[2 * a + 1]) / static_cast<double>(stride));
  }
  // the output tensor map is never used for a store in this mode; pred only gives it a valid base address
  me_conv_desc dd = *d;
  dd.out_pitch = d->cout;
  return conv_dispatch(&dd, x, w_packed, bias, nullptr, pred, stream, &dec);
}

}  // extern "C"
