// 3x3 conv + (folded BN) bias + activation (+ residual) for THIN inputs (16 or 32 channels, 32 or 64 filters) on
// tcgen05, with SIMT im2col producers instead of im2col-mode TMA.
//
// Why: a pixel of these layers is a 32- or 64-byte row, and TMA delivers im2col rows at ~0.4 rows per clock per SM
// whatever their length (profiles/round1: 208^2 3x3 32->64 spent 2765 of its ~3100 clocks per 128-pixel tile waiting
// for 9 x 128 rows; the MMAs need ~2000, the HBM floor is 1740).  Eight producer warps gather the same bytes with
// 16-byte cp.async copies (L1/L2 resident: every input pixel is read 9 times but fetched from HBM once; synchronous
// loads through registers were tried first and were latency-bound at 2x the TMA time) and build the operand
// tile themselves, one filter row (3 taps) per pipeline stage: per tap a [128 rows][CIN] K-major tile in the same
// 64-byte / 32-byte swizzled layout TMA would have produced (chunk ^= (row / 2) & 3 resp. (row / 4) & 1; the first
// version used the no-swizzle canonical layout, whose operand fetch costs ~250 clocks per MMA instead of ~110).
// The layer's whole weight slab (<= 36 KB) stays in shared memory, one swizzled [COUT][CIN] tile per tap.  Warp roles:
//   warps 0-7   producers (256 threads: lanes run over a pixel's 16-byte chunks first, then over tile rows)
//   warp 8      TMEM owner + MMA issuer: 3 x CIN/16 tcgen05.mma (M=128, N=COUT, K=16) per stage, commit per stage
//   warps 9-16  two epilogue groups alternating tiles over FOUR TMEM accumulators: tcgen05.ld -> +bias -> activation ->
//               (+ residual, loaded straight from global memory before the accumulator is awaited) -> fp16 -> 32-byte
//               global stores, a lane per output pixel (a warp's rows are one contiguous span).  No staging tile, no
//               block barrier, no TMA store: with 3 K stages per tile the per-tile hand-shakes of the staged epilogue
//               were half of the kernel's time (ME_THIN_DBG attribution, tools/thin_probe2.py).
//   POOL        (stride 1, the 2x2 / stride-2 max-pool that follows the layer in the tiny cfgs): a tile is a block of
//               bw x 128/bw pixels (bw = 16 or 8), so a warp's 32 rows hold whole 2x2 windows (shfl_xor 1 and bw); the
//               pooled pixels are written instead of the conv output.  max commutes with the fp16 rounding.
// Replaces the same reference lines as conv_gemm.cu (yolov3/models.py:22-41,252,258-260) for modules 1 and 3 of
// Darknet-53 and the first two 3x3 layers after the pools of the tiny cfgs.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace me {

unsigned long long* conv_debug_word();  // conv_gemm.cu
int conv_ensure_debug_word();
bool conv_pdl_enabled();

namespace {

constexpr int kBM = 128;
constexpr int kProdThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kThreads = kProdThreads + 32 + 2 * kEpiThreads;  // 544
constexpr int kMmaWarp = kProdThreads / 32;
constexpr uint32_t kEpiBarrierId = 1;

struct ThinParams {
  const __half* x;
  const __half* w;      // packed [cout][9 * CIN] fp16 (me_pack_conv_weights)
  const float* bias;
  int M, H, W, Ho, Wo, in_pitch, stride;
  int tiles, act, has_res;
  const __half* res;    // residual rows (res_pitch halves apart) or nullptr
  __half* y;            // output rows (out_pitch halves apart); POOL: the pooled tensor
  int out_pitch, res_pitch;
  int pool_bw;          // POOL: tile block width (16 or 8); 0 otherwise
  int dbg;              // ME_THIN_DBG attribution bits: 1 no output stores, 2 no input copies, 4 no MMAs
  unsigned long long* debug;
};

template <int CIN, int COUT>
struct TCfg {
  static constexpr int RPS = CIN == 16 ? 3 : 1;        // filter rows per pipeline stage: the whole 3x3 window when it is
                                                       // 36 KB (16 channels) - one barrier round trip per tile, not three
  static constexpr int KROW = RPS * 3 * CIN;           // K of one stage
  static constexpr int CHUNKS = KROW / 8;              // 16-byte chunks per tile row and stage
  static constexpr int STAGE_BYTES = CHUNKS * kBM * 16;
  static constexpr int W_BYTES = COUT * 9 * CIN * 2;
  static constexpr int ACCS = 4;                        // TMEM accumulators in flight
  static constexpr int TAIL_BYTES = COUT * 4 + 32 * 8 + 16;
  static constexpr int FIXED = 1024 + W_BYTES + TAIL_BYTES;
  static constexpr int STAGES = (227 * 1024 - FIXED) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - FIXED) / STAGE_BYTES;
  static constexpr int SMEM = FIXED + STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = ACCS * COUT <= 128 ? 128 : 256;
  static_assert(CIN == 16 || CIN == 32, "thin conv: 16 or 32 input channels");
  static_assert(COUT == 32 || COUT == 64, "thin conv: 32 or 64 filters");
  static_assert(STAGES >= 3, "thin conv: pipeline too shallow");
};

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long* dbg, uint32_t tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
      if (dbg) {
        *reinterpret_cast<volatile unsigned long long*>(dbg) =
            (static_cast<unsigned long long>(tag | 0x4000u) << 32) | (static_cast<unsigned long long>(blockIdx.x) << 8) | parity | 0x80u;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// Byte offset of 16-byte chunk c8 of row `row` in a K-major tile with ROWB-byte rows, TMA swizzle of that width.
template <int ROWB>
__device__ __forceinline__ uint32_t swz_off(int row, int c8) {
  uint32_t off = row * ROWB + c8 * 16;
  off ^= ((off >> 7) & (ROWB / 16 - 1)) << 4;
  return off;
}

// Output pixel of one tile row, advanced from tile to tile without divisions (17 warps doing four runtime div/mod per tile
// were a third of the kernel's instruction issue).  Flat tiles: pixel index tile * 128 + i decomposed over (Wo, Ho).
// POOL tiles: the BLOCK index decomposed over (blocks per row, block rows per image); the row's offset inside the block
// is added by the caller.
struct PixIter {
  int a, b, c;        // flat: ox, oy, img;  pool: bx, by, img
  int sa, sb, sc, A, B;
  __device__ __forceinline__ void init(int first, int step, int A_, int B_) {
    A = A_;
    B = B_;
    a = first % A;
    int t = first / A;
    b = t % B;
    c = t / B;
    sa = step % A;
    t = step / A;
    sb = t % B;
    sc = t / B;
  }
  __device__ __forceinline__ void next() {
    a += sa;
    if (a >= A) { a -= A; ++b; }
    b += sb;
    if (b >= B) { b -= B; ++c; }
    c += sc;
  }
};

template <int CIN, int COUT, bool POOL>
__global__ void __launch_bounds__(kThreads, 1)
conv_thin_kernel(const ThinParams p) {
  using C = TCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ROWB = CIN * 2;                          // bytes of one tile row (one pixel's channels / one filter's tap)
  uint8_t* s_w = smem;                                   // 9 taps x [COUT][ROWB] swizzled
  uint8_t* s_a = s_w + C::W_BYTES;                        // STAGES x 3 taps x [128][ROWB] swizzled
  float* s_bias = reinterpret_cast<float*>(s_a + C::STAGES * C::STAGE_BYTES);  // [COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + COUT);
  uint64_t* full_bar = bars;            // [8]
  uint64_t* empty_bar = bars + 8;       // [8]
  uint64_t* tmem_full = bars + 16;      // [4]
  uint64_t* tmem_empty = bars + 20;     // [4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();
  if (warp == kMmaWarp) {
    if (ptx::elect_one()) {
      for (int s = 0; s < C::STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], kProdThreads / 32);   // one arrival per producer warp
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < C::ACCS; ++a) {
        ptx::mbar_init(&tmem_full[a], 1);
        ptx::mbar_init(&tmem_empty[a], 4);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  // weights -> canonical no-swizzle layout (they are parameters, not written by the previous layer: before pdl_wait)
  for (int i = threadIdx.x; i < COUT * (9 * CIN / 8); i += kThreads) {
    const int o = i / (9 * CIN / 8), k8 = i - o * (9 * CIN / 8);
    const int tap = k8 / (CIN / 8), c8 = k8 - tap * (CIN / 8);
    *reinterpret_cast<uint4*>(s_w + tap * (COUT * ROWB) + swz_off<ROWB>(o, c8)) =
        __ldg(reinterpret_cast<const uint4*>(p.w + static_cast<size_t>(o) * 9 * CIN + k8 * 8));
  }
  for (int i = threadIdx.x; i < COUT; i += kThreads) s_bias[i] = p.bias[i];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::pdl_wait();

  // Tile row i of a tile -> output pixel.  Flat tiles: pixel tile * 128 + i.  POOL tiles: block (img, by, bx) of
  // bw x 128/bw pixels, row i = (i / bw, i % bw) inside it (W % bw == 0; the last block row of an image may be partial).
  const int n_img = p.M / (p.Ho * p.Wo);
  const int bw_log = p.pool_bw == 16 ? 4 : 3;
  auto iter_init = [&](PixIter& it, int first_tile, int tile_step, int i) {
    if constexpr (POOL) {
      it.init(first_tile, tile_step, p.Wo >> bw_log, (p.Ho + (kBM >> bw_log) - 1) / (kBM >> bw_log));
    } else {
      it.init(first_tile * kBM + i, tile_step * kBM, p.Wo, p.Ho);
    }
  };
  auto iter_pixel = [&](const PixIter& it, int i, int& ox, int& oy, int& img) -> bool {
    img = it.c;
    if constexpr (POOL) {
      ox = (it.a << bw_log) + (i & (p.pool_bw - 1));
      oy = it.b * (kBM >> bw_log) + (i >> bw_log);
      return oy < p.Ho;
    } else {
      ox = it.a;
      oy = it.b;
      return img < n_img;
    }
  };

  if (warp < kMmaWarp) {
    // ------------------------------------------------------------------ im2col producers
    // Lanes run over the 16-byte chunks of a pixel first, then over pixels: a warp instruction reads 512 contiguous
    // bytes (stride 1) instead of 32 half-used sectors (a lane per pixel was bound by L1 sector throughput).
    constexpr int CPP = CIN / 8;                      // chunks per pixel
    constexpr int PIX_PER_PASS = kProdThreads / CPP;  // tile rows covered by one pass of the 256 threads
    constexpr int PASSES = kBM / PIX_PER_PASS;        // 2 (CIN 32) or 1 (CIN 16)
    const int c8 = threadIdx.x % CPP;
    const int p0 = threadIdx.x / CPP;
    uint32_t stage = 0, phase = 0, ann = 0;
    int issued = 0;
    constexpr int LAG = C::STAGES >= 6 ? 3 : (C::STAGES >= 4 ? 2 : 1);
    PixIter pit[PASSES];
#pragma unroll
    for (int q = 0; q < PASSES; ++q) iter_init(pit[q], blockIdx.x, gridDim.x, p0 + q * PIX_PER_PASS);
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const __half* base[PASSES];
      int ix0[PASSES], iy0[PASSES];
      bool live[PASSES];
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        int ox, oy, img;
        live[q] = iter_pixel(pit[q], p0 + q * PIX_PER_PASS, ox, oy, img);
        pit[q].next();
        ix0[q] = ox * p.stride - 1;
        iy0[q] = oy * p.stride - 1;
        base[q] = p.x + (static_cast<size_t>(img) * p.H * p.W) * p.in_pitch + c8 * 8;
      }
#pragma unroll 1
      for (int r0 = 0; r0 < 3; r0 += C::RPS) {
        mbar_wait(&empty_bar[stage], phase ^ 1, p.debug, 0x100u + stage);
        const uint32_t dst0 = ptx::smem_u32(s_a + stage * C::STAGE_BYTES);
#pragma unroll
        for (int rr = 0; rr < C::RPS; ++rr) {
#pragma unroll
          for (int q = 0; q < PASSES; ++q) {
            const int iy = iy0[q] + r0 + rr;
            const bool yok = live[q] && iy >= 0 && iy < p.H;
            const __half* line = base[q] + static_cast<size_t>(iy) * p.W * p.in_pitch;
            const uint32_t dst = dst0 + rr * 3 * (kBM * ROWB) + swz_off<ROWB>(p0 + q * PIX_PER_PASS, c8);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const int ix = ix0[q] + s;
              const bool ok = yok && ix >= 0 && ix < p.W;
              // cp.async: the copies of several stages are in flight per thread, no registers held; a pixel outside the
              // image is zero-filled (the source address is then only a placeholder)
              if (!(p.dbg & 2))
                ptx::cp_async_16(dst + s * (kBM * ROWB), ok ? line + static_cast<size_t>(ix) * p.in_pitch : p.x, ok);
            }
          }
        }
        // A stage is announced LAG stages after its copies were issued: every lane waits until its own copies of that
        // stage have landed (cp.async.wait_group), makes them visible to the tensor core's proxy (writer-side fence), and
        // one lane per warp arrives: 8 arrivals per stage instead of 256 cp.async.mbarrier.arrive (measured neutral in
        // time; kept for the writer-side fence).
        ptx::cp_async_commit_group();
        ++issued;
        if (issued > LAG) {
          ptx::cp_async_wait_group<LAG>();
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&full_bar[ann]);
          if (++ann == (uint32_t)C::STAGES) ann = 0;
        }
        if (++stage == (uint32_t)C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    // drain: the last LAG stages
    ptx::cp_async_wait_group<0>();
    ptx::fence_proxy_async_smem();
    __syncwarp();
    for (int k = issued > LAG ? LAG : issued; k > 0; --k) {
      if (lane == 0) ptx::mbar_arrive(&full_bar[ann]);
      if (++ann == (uint32_t)C::STAGES) ann = 0;
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, COUT);
      const uint32_t w_addr = ptx::smem_u32(s_w);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int acc = it & (C::ACCS - 1);
        mbar_wait(&tmem_empty[acc], ((it / C::ACCS) & 1) ^ 1, p.debug, 0x200u + acc);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * COUT;
#pragma unroll 1
        for (int r0 = 0; r0 < 3; r0 += C::RPS) {
          mbar_wait(&full_bar[stage], phase, p.debug, 0x300u + stage);
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(s_a + stage * C::STAGE_BYTES);
#pragma unroll
          for (int rs = 0; rs < C::RPS * 3; ++rs) {
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k) {
              const uint64_t adesc = ptx::make_kmajor_desc(a_addr + rs * (kBM * ROWB) + k * 32, ROWB);
              const uint64_t bdesc = ptx::make_kmajor_desc(w_addr + (r0 * 3 + rs) * (COUT * ROWB) + k * 32, ROWB);
              if (!(p.dbg & 4)) ptx::umma_f16_ss(d_tmem, adesc, bdesc, idesc, (r0 | rs | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == (uint32_t)C::STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: two groups alternate tiles
    const int g = (warp - (kMmaWarp + 1)) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter (warp index % 4)
    const int row = q * 32 + lane;
    constexpr int NV = COUT / 8;            // 16-byte vectors per output row
    const bool wide = (p.out_pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
    int it = g;
    PixIter eit;
    iter_init(eit, blockIdx.x + g * gridDim.x, 2 * gridDim.x, row);
    for (int tile = blockIdx.x + g * gridDim.x; tile < p.tiles; tile += 2 * gridDim.x, it += 2) {
      const int acc = it & (C::ACCS - 1);
      int ox, oy, img;
      const bool live = iter_pixel(eit, row, ox, oy, img);
      eit.next();
      const long long pix = (static_cast<long long>(img) * p.Ho + oy) * p.Wo + ox;
      // the residual row travels while the MMAs of this tile are still running
      uint4 res[NV];
      if (p.has_res && live) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_pitch);
#pragma unroll
        for (int j = 0; j < NV; ++j) res[j] = __ldg(rp + j);
      }
      mbar_wait(&tmem_full[acc], (it / C::ACCS) & 1, p.debug, 0x400u + acc);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + acc * COUT + (static_cast<uint32_t>(q * 32) << 16);
      constexpr int NCH = COUT / 32;
      uint32_t r[2][32];
      uint4 o[NV];
      ptx::tmem_ld_32x32b_x32(t_row, r[0]);
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = ci * 32;
        ptx::tmem_ld_wait_regs(r[ci & 1]);
        if (ci + 1 < NCH) ptx::tmem_ld_32x32b_x32(t_row + c + 32, r[(ci + 1) & 1]);
        float v[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c + 4 * j4);
          v[4 * j4 + 0] = __uint_as_float(r[ci & 1][4 * j4 + 0]) + b4.x;
          v[4 * j4 + 1] = __uint_as_float(r[ci & 1][4 * j4 + 1]) + b4.y;
          v[4 * j4 + 2] = __uint_as_float(r[ci & 1][4 * j4 + 2]) + b4.z;
          v[4 * j4 + 3] = __uint_as_float(r[ci & 1][4 * j4 + 3]) + b4.w;
        }
        if (p.act == ME_ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
        } else if (p.act == ME_ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* vv = v + 8 * j;
          if (p.has_res && live) {
            const __half2* rh = reinterpret_cast<const __half2*>(&res[ci * 4 + j]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(rh[e]);
              vv[2 * e] += f.x;
              vv[2 * e + 1] += f.y;
            }
          }
          __half2* oh = reinterpret_cast<__half2*>(&o[ci * 4 + j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(vv[2 * e], vv[2 * e + 1]);
        }
      }
      // the accumulator is in registers: hand it back before the stores
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      long long opix = pix;
      bool writer = live;
      if constexpr (POOL) {
        // a warp's 32 rows are 32/bw image rows x bw columns of the block: the 2x2 window of an even (row, column) lane
        // is lanes +1, +bw, +bw+1
        const int bw = p.pool_bw;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          uint32_t* w4 = reinterpret_cast<uint32_t*>(&o[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t u = w4[e];
            uint32_t n1 = __shfl_xor_sync(0xffffffffu, u, 1);
            __half2 m2 = __hmax2(*reinterpret_cast<__half2*>(&u), *reinterpret_cast<__half2*>(&n1));
            u = *reinterpret_cast<uint32_t*>(&m2);
            uint32_t n2 = __shfl_xor_sync(0xffffffffu, u, bw);
            m2 = __hmax2(m2, *reinterpret_cast<__half2*>(&n2));
            w4[e] = *reinterpret_cast<uint32_t*>(&m2);
          }
        }
        writer = live && ((ox | oy) & 1) == 0;
        opix = (static_cast<long long>(img) * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1);
      }
      if (writer && !(p.dbg & 1)) {
        __half* dst = p.y + opix * p.out_pitch;
        if (wide) {
#pragma unroll
          for (int j = 0; j < NV; j += 2) ptx::st_global_256(dst + j * 8, o[j], o[j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < NV; ++j) *reinterpret_cast<uint4*>(dst + j * 8) = o[j];
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// pool_bw: 0 = plain conv; 16 / 8 = fused 2x2 max-pool with bw-wide tile blocks (y is the pooled tensor)
template <int CIN, int COUT, bool POOL>
int launch_thin(const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                int pool_bw, cudaStream_t stream) {
  using C = TCfg<CIN, COUT>;
  const int Ho = (d->h + 2 - 3) / d->stride + 1, Wo = (d->w + 2 - 3) / d->stride + 1;
  const long long M64 = 1LL * d->n * Ho * Wo;
  ME_REQUIRE(M64 > 0 && M64 < (1LL << 31) - kBM, "conv(thin): pixel count out of range");
  ThinParams p{};
  p.x = static_cast<const __half*>(x);
  p.w = static_cast<const __half*>(w);
  p.bias = bias;
  p.M = static_cast<int>(M64);
  p.H = d->h;
  p.W = d->w;
  p.Ho = Ho;
  p.Wo = Wo;
  p.in_pitch = d->in_pitch;
  p.stride = d->stride;
  p.tiles = POOL ? d->n * ceil_div(Ho, kBM / pool_bw) * (Wo / pool_bw) : ceil_div(p.M, kBM);
  p.act = d->act;
  p.has_res = (d->res_pitch > 0 && residual != nullptr) ? 1 : 0;
  p.res = static_cast<const __half*>(residual);
  p.res_pitch = d->res_pitch;
  p.y = static_cast<__half*>(y);
  p.out_pitch = d->out_pitch;
  p.pool_bw = pool_bw;
  ME_REQUIRE(d->out_pitch % 8 == 0 && (!p.has_res || d->res_pitch % 8 == 0), "conv(thin): pitches must be multiples of 8");
  ME_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
             "conv(thin): output / residual must be 16-byte aligned");
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("ME_THIN_DBG");
    dbg = e ? atoi(e) : 0;
  }
  p.dbg = dbg;
  int rc = conv_ensure_debug_word();
  if (rc != ME_OK) return rc;
  p.debug = conv_debug_word();
  auto kern = conv_thin_kernel<CIN, COUT, POOL>;
  static bool attr_seen[64] = {false};   // per instantiation and per device
  if (first_use_on_device(attr_seen))
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (grid > p.tiles) grid = p.tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = conv_pdl_enabled() ? 1 : 0;
  ME_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  ME_LAUNCH_CHECK();
  return ME_OK;
}

// Block width of the fused-pool tiles for an output of Ho x Wo pixels (0: the shape has none).
int pool_block_width(int Ho, int Wo) {
  if (Ho % 2 != 0) return 0;
  if (Wo % 16 == 0) return 16;   // blocks of 8 rows x 16 columns
  if (Wo % 8 == 0) return 8;     // 16 rows x 8 columns
  return 0;
}

}  // namespace

// ME_CONV_THIN: 0 = never, 1 (default) = 16-channel inputs, 2 = 16- and 32-channel inputs.
// Measured on B200 at batch 32 (profiles/round1/thin_probe.log): 208^2 16->32: 88 us vs 124 us through im2col TMA;
// with 32 input channels both paths take the same time (208^2 32->64: 132 vs 131 us, 104^2: 36.9 vs 37.8 us) because
// there the ~110-130 clocks per tcgen05.mma, not the operand delivery, set the pace - so those stay on the TMA kernel.
static int conv_thin_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ME_CONV_THIN");
    v = e ? (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)) : 1;
  }
  return v;
}
bool conv_thin_enabled() { return conv_thin_mode() != 0; }

// True if the layer goes to the thin kernel (me_conv_gemm's dispatcher asks before falling back).
bool conv_thin_supported(const me_conv_desc* d) {
  const bool cin_ok = d->cin == 16 || (d->cin == 32 && conv_thin_mode() == 2);
  return d->ksize == 3 && (d->stride == 1 || d->stride == 2) && !d->out_f32 && cin_ok &&
         (d->cout == 32 || d->cout == 64) && d->in_pitch % 8 == 0;
}

int conv_thin(const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
              cudaStream_t stream) {
  if (d->cin == 16 && d->cout == 32) return launch_thin<16, 32, false>(d, x, w, bias, residual, y, 0, stream);
  if (d->cin == 16 && d->cout == 64) return launch_thin<16, 64, false>(d, x, w, bias, residual, y, 0, stream);
  if (d->cin == 32 && d->cout == 32) return launch_thin<32, 32, false>(d, x, w, bias, residual, y, 0, stream);
  return launch_thin<32, 64, false>(d, x, w, bias, residual, y, 0, stream);
}

// conv (3x3, stride 1, 16 / 32 input channels, 32 / 64 filters) + MaxPool2d(2, 2): y is the pooled tensor.
bool conv_thin_pool_supported(const me_conv_desc* d) {
  return d->ksize == 3 && d->stride == 1 && !d->out_f32 && d->res_pitch == 0 && (d->cin == 16 || d->cin == 32) &&
         (d->cout == 32 || d->cout == 64) && d->in_pitch % 8 == 0 && pool_block_width(d->h, d->w) != 0;
}

int conv_thin_pool(const me_conv_desc* d, const void* x, const void* w, const float* bias, void* y, cudaStream_t stream) {
  const int bw = pool_block_width(d->h, d->w);
  if (d->cin == 16 && d->cout == 32) return launch_thin<16, 32, true>(d, x, w, bias, nullptr, y, bw, stream);
  if (d->cin == 16 && d->cout == 64) return launch_thin<16, 64, true>(d, x, w, bias, nullptr, y, bw, stream);
  if (d->cin == 32 && d->cout == 32) return launch_thin<32, 32, true>(d, x, w, bias, nullptr, y, bw, stream);
  return launch_thin<32, 64, true>(d, x, w, bias, nullptr, y, bw, stream);
}

}  // namespace me
