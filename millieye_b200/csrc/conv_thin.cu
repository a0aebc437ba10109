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
//   warps 9-16  two epilogue groups alternating tiles, as in conv_gemm.cu: tcgen05.ld -> +bias -> activation ->
//               (+ residual, TMA-prefetched into the staging tile) -> fp16 -> swizzled staging -> TMA store
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
  int dbg;              // ME_THIN_DBG attribution bits: 1 no output stores, 2 no input copies, 4 no MMAs
  unsigned long long* debug;
};

template <int CIN, int COUT>
struct TCfg {
  static constexpr int KROW = 3 * CIN;                 // K of one stage (one filter row)
  static constexpr int CHUNKS = KROW / 8;              // 16-byte chunks per tile row and stage
  static constexpr int STAGE_BYTES = CHUNKS * kBM * 16;
  static constexpr int W_BYTES = COUT * 9 * CIN * 2;
  static constexpr int STAGING_BYTES = kBM * COUT * 2;  // per epilogue group; rows of 64 (COUT=32) or 128 bytes
  static constexpr int ROW_BYTES = COUT * 2;
  static constexpr int SWZ_BITS = ROW_BYTES == 128 ? 3 : 2;
  static constexpr int TAIL_BYTES = 2 * COUT * 4 + 32 * 8 + 16;
  static constexpr int FIXED = 1024 + W_BYTES + 2 * STAGING_BYTES + TAIL_BYTES;
  static constexpr int STAGES = (227 * 1024 - FIXED) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - FIXED) / STAGE_BYTES;
  static constexpr int SMEM = FIXED + STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * COUT <= 64 ? 64 : 128;
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

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
conv_thin_kernel(const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const ThinParams p) {
  using C = TCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ROWB = CIN * 2;                          // bytes of one tile row (one pixel's channels / one filter's tap)
  uint8_t* s_w = smem;                                   // 9 taps x [COUT][ROWB] swizzled
  uint8_t* s_a = s_w + C::W_BYTES;                        // STAGES x 3 taps x [128][ROWB] swizzled
  uint8_t* staging = s_a + C::STAGES * C::STAGE_BYTES;    // 2 x [128][ROW_BYTES] swizzled
  float* s_bias = reinterpret_cast<float*>(staging + 2 * C::STAGING_BYTES);  // [2][COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 2 * COUT);
  uint64_t* full_bar = bars;            // [8]
  uint64_t* empty_bar = bars + 8;       // [8]
  uint64_t* tmem_full = bars + 16;      // [2]
  uint64_t* tmem_empty = bars + 18;     // [2]
  uint64_t* res_full = bars + 20;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();
  if (warp == kMmaWarp) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmC);
      if (p.has_res) ptx::prefetch_tmap(&tmR);
      for (int s = 0; s < C::STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], kProdThreads);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(&tmem_full[a], 1);
        ptx::mbar_init(&tmem_empty[a], 4);
        ptx::mbar_init(&res_full[a], 1);
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
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::pdl_wait();

  if (warp < kMmaWarp) {
    // ------------------------------------------------------------------ im2col producers
    // Lanes run over the 16-byte chunks of a pixel first, then over pixels: a warp instruction reads 512 contiguous
    // bytes (stride 1) instead of 32 half-used sectors (a lane per pixel was bound by L1 sector throughput).
    constexpr int CPP = CIN / 8;                      // chunks per pixel
    constexpr int PIX_PER_PASS = kProdThreads / CPP;  // tile rows covered by one pass of the 256 threads
    constexpr int PASSES = kBM / PIX_PER_PASS;        // 2 (CIN 32) or 1 (CIN 16)
    const int c8 = threadIdx.x % CPP;
    const int p0 = threadIdx.x / CPP;
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const __half* base[PASSES];
      int ix0[PASSES], iy0[PASSES];
      bool live[PASSES];
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        const int m = tile * kBM + p0 + q * PIX_PER_PASS;
        live[q] = m < p.M;
        const int ox = m % p.Wo;
        const int t = m / p.Wo;
        const int oy = t % p.Ho, img = t / p.Ho;
        ix0[q] = ox * p.stride - 1;
        iy0[q] = oy * p.stride - 1;
        base[q] = p.x + (static_cast<size_t>(img) * p.H * p.W) * p.in_pitch + c8 * 8;
      }
#pragma unroll 1
      for (int r = 0; r < 3; ++r) {
        mbar_wait(&empty_bar[stage], phase ^ 1, p.debug, 0x100u + stage);
        const uint32_t dst0 = ptx::smem_u32(s_a + stage * C::STAGE_BYTES);
#pragma unroll
        for (int q = 0; q < PASSES; ++q) {
          const int iy = iy0[q] + r;
          const bool yok = live[q] && iy >= 0 && iy < p.H;
          const __half* line = base[q] + static_cast<size_t>(iy) * p.W * p.in_pitch;
          const uint32_t dst = dst0 + swz_off<ROWB>(p0 + q * PIX_PER_PASS, c8);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int ix = ix0[q] + s;
            const bool ok = yok && ix >= 0 && ix < p.W;
            // cp.async: the copies of up to STAGES filter rows are in flight per thread, no registers held; a pixel
            // outside the image is zero-filled (the source address is then only a placeholder)
            if (!(p.dbg & 2))
              ptx::cp_async_16(dst + s * (kBM * ROWB), ok ? line + static_cast<size_t>(ix) * p.in_pitch : p.x, ok);
          }
        }
        ptx::cp_async_mbar_arrive_noinc(&full_bar[stage]);
        if (++stage == (uint32_t)C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, COUT);
      const uint32_t w_addr = ptx::smem_u32(s_w);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, p.debug, 0x200u + acc);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * COUT;
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
          mbar_wait(&full_bar[stage], phase, p.debug, 0x300u + stage);
          ptx::fence_proxy_async_smem();   // the producers' cp.async writes (generic proxy) -> visible to the tensor core
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(s_a + stage * C::STAGE_BYTES);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k) {
              const uint64_t adesc = ptx::make_kmajor_desc(a_addr + s * (kBM * ROWB) + k * 32, ROWB);
              const uint64_t bdesc = ptx::make_kmajor_desc(w_addr + (r * 3 + s) * (COUT * ROWB) + k * 32, ROWB);
              if (!(p.dbg & 4)) ptx::umma_f16_ss(d_tmem, adesc, bdesc, idesc, (r | s | k) != 0 ? 1u : 0u);
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
    const int etid = (threadIdx.x - (kProdThreads + 32)) & (kEpiThreads - 1);
    const bool leader = etid == 0;
    uint8_t* stg = staging + g * C::STAGING_BYTES;
    float* bias_s = s_bias + g * COUT;
    uint64_t* res_bar = &res_full[g];
    const uint32_t bar_id = kEpiBarrierId + g;
    auto load_residual = [&](int tile) {
      ptx::mbar_arrive_expect_tx(res_bar, C::STAGING_BYTES);
      ptx::tma_load_2d(&tmR, res_bar, stg, 0, tile * kBM);
    };
    for (int i = etid; i < COUT; i += kEpiThreads) bias_s[i] = p.bias[i];
    int tile = blockIdx.x + g * gridDim.x;
    if (p.has_res && leader && tile < p.tiles) load_residual(tile);
    for (int lit = 0; tile < p.tiles; ++lit, tile += 2 * gridDim.x) {
      if (leader && !p.has_res) ptx::tma_store_wait_read0();  // previous store has drained the staging tile
      ptx::named_bar_sync(bar_id, kEpiThreads);
      mbar_wait(&tmem_full[g], lit & 1, p.debug, 0x400u + g);
      ptx::tc_fence_after();
      if (p.has_res) mbar_wait(res_bar, lit & 1, p.debug, 0x500u + g);
      const uint32_t t_row = tmem_base + g * COUT + (static_cast<uint32_t>(q * 32) << 16);
      constexpr int NCH = COUT / 32;
      uint32_t r[2][32];
      ptx::tmem_ld_32x32b_x32(t_row, r[0]);
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = ci * 32;
        ptx::tmem_ld_wait_regs(r[ci & 1]);
        if (ci + 1 < NCH) ptx::tmem_ld_32x32b_x32(t_row + c + 32, r[(ci + 1) & 1]);
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
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
        } else if (p.act == ME_ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        }
        const uint32_t rbase = row * C::ROW_BYTES + c * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t off = rbase + j * 16;
          off ^= ((off >> 7) & ((1u << C::SWZ_BITS) - 1)) << 4;
          uint4* dst = reinterpret_cast<uint4*>(stg + off);
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
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[g]);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(bar_id, kEpiThreads);
      if (leader) {
        if (!(p.dbg & 1)) ptx::tma_store_2d(&tmC, stg, 0, tile * kBM);
        ptx::tma_store_commit();
        const int next = tile + 2 * gridDim.x;
        if (p.has_res && next < p.tiles) {
          ptx::tma_store_wait_read0();
          load_residual(next);
        }
      }
    }
    if (leader) ptx::tma_store_wait_all0();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int CIN, int COUT>
int launch_thin(const me_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                cudaStream_t stream) {
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
  p.tiles = ceil_div(p.M, kBM);
  p.act = d->act;
  p.has_res = (d->res_pitch > 0 && residual != nullptr) ? 1 : 0;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("ME_THIN_DBG");
    dbg = e ? atoi(e) : 0;
  }
  p.dbg = dbg;
  int rc = conv_ensure_debug_word();
  if (rc != ME_OK) return rc;
  p.debug = conv_debug_word();
  CUtensorMap tmC, tmR;
  const CUtensorMapSwizzle swz = swizzle_for_row_bytes(C::ROW_BYTES);
  rc = encode_tiled_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y, COUT, p.M, d->out_pitch, COUT, kBM, swz);
  if (rc != ME_OK) return rc;
  if (p.has_res) {
    rc = encode_tiled_2d(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, residual, COUT, p.M, d->res_pitch, COUT, kBM, swz);
    if (rc != ME_OK) return rc;
  } else {
    tmR = tmC;
  }
  auto kern = conv_thin_kernel<CIN, COUT>;
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
  ME_CUDA(cudaLaunchKernelEx(&cfg, kern, tmC, tmR, p));
  ME_LAUNCH_CHECK();
  return ME_OK;
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
  if (d->cin == 16 && d->cout == 32) return launch_thin<16, 32>(d, x, w, bias, residual, y, stream);
  if (d->cin == 16 && d->cout == 64) return launch_thin<16, 64>(d, x, w, bias, residual, y, stream);
  if (d->cin == 32 && d->cout == 32) return launch_thin<32, 32>(d, x, w, bias, residual, y, stream);
  return launch_thin<32, 64>(d, x, w, bias, residual, y, stream);
}

}  // namespace me
