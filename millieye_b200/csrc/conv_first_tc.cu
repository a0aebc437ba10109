// First conv (3x3 / stride 1 / pad 1, cin <= 3) on tcgen05 with an in-kernel im2col producer.
//
// The layer reads the caller's NCHW fp32 image directly: K = 9*cin = 27 (padded to 32) is too thin
// for a TMA-fed pipeline, so producer warps build the 128 x 32 fp16 A tile in shared memory themselves,
// in the no-swizzle K-major canonical layout  addr(row, k8) = k8 * 2048 + row * 16  (8-row core
// matrices are 128 contiguous bytes, LBO = 2048, SBO = 128), fence it to the async proxy and hand it to
// the MMA warp through an mbarrier.  Two tcgen05.mma (K = 16 each) per tile accumulate into TMEM; 4
// epilogue warps add the folded-BN bias, apply the activation and write NHWC fp16 rows with 16-byte
// stores (a warp's 32 rows are one contiguous span).  Replaces the Conv2d/BN/LeakyReLU of module 0
// (yolov3/models.py:22-41, :252) - 0.3 GFLOP/frame of SIMT work that used to cost as much as ten
// tensor-core layers.
//
// Two producer flavours (the layer is HBM-bound: 12 B in, 2*COUT B out per pixel, so what matters is
// bytes in flight and instructions per pixel):
//   VEC  (image width % 4 == 0): one warp builds a whole tile, each lane 4 consecutive pixels from one
//        128-bit load + two edge scalars per (channel, row) - 6.75 loads per pixel; 8 warps / 8 stages.
//   !VEC (any width): one pixel per thread, 27 scalar loads; 4 groups of 4 warps / 4 stages.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace me {
namespace {

constexpr int kTileM = 128;
constexpr int kK = 32;                 // 27 real + 5 zero
constexpr int kABytes = kTileM * kK * 2;

template <bool VEC, int EPI = 2>
struct FCfg {
  static constexpr int STAGES = VEC ? 8 : 4;
  static constexpr int PROD_WARPS = VEC ? 8 : 16;
  static constexpr int ARRIVALS = VEC ? 32 : 128;  // producer threads per tile
  static constexpr int EPI_GROUPS = EPI;           // 4-warp epilogue groups taking tiles in turn
  // TMEM accumulators in flight.  They form ACCS / TPS sets (one per stage in flight); the number of sets must be a
  // multiple of the number of epilogue groups so that a group meets each set's barrier phases in order (a group that
  // skips a phase would pass a parity wait on a phase that has not started)
  static constexpr int ACCS = VEC ? 2 * EPI : 4;   // (TPS = 1: 4 or 6 sets for 2 or 3 groups)
  static constexpr int THREADS = 32 * (PROD_WARPS + 1 + 4 * EPI_GROUPS);
  // Tiles per pipeline stage / barrier round trip.  With everything but the barriers stripped the kernel takes 330 clocks
  // per 128-pixel tile (ME_FIRST_DBG=31), and TPS = 2 halves that floor (49 -> 29 us at 416^2 x 32) - but the whole
  // kernel does not gain (16 channels 100 -> 94 us, 32 channels 134 -> 143 us, pooled 95 -> 95 us): the epilogue then
  // owns two accumulators per wake-up and the MMAs run less far ahead of it.  Kept at 1; the loops are written for any TPS.
  static constexpr int TPS = 1;
  static constexpr int SMEM = STAGES * TPS * kABytes + 1024;
};

__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!ptx::mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1u) << 46;  // descriptor version; layout_type 0 = no swizzle
  return d;
}

// 27 taps of one pixel (k = c*9 + r*3 + s) -> one 64-byte K-major row, written as four 16-byte chunks
__device__ __forceinline__ void store_row(uint8_t* tile, int row, const float (&v)[27]) {
  uint8_t* dst = tile + row * 16;
#pragma unroll
  for (int k8 = 0; k8 < 4; ++k8) {
    uint4 pk;
    __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k8 * 8 + 2 * e;
      ph[e] = __floats2half2_rn(k < 27 ? v[k] : 0.f, k + 1 < 27 ? v[k + 1] : 0.f);
    }
    *reinterpret_cast<uint4*>(dst + k8 * (kTileM * 16)) = pk;
  }
}

// POOL (VEC only; w % 32 == 0, h % 4 == 0): the 2x2 / stride-2 max-pool that follows the layer in the tiny cfgs
// (yolov3-tiny-12.cfg blocks 0-1) runs in the epilogue.  A tile is a 4-row x 32-column block of one image; its 128 MMA
// rows are ordered so that each epilogue warp holds 2 rows x 16 columns (lane = 16 * dy + dx): the 2x2 window is a
// shfl_xor 1 / shfl_xor 16 away, and the 8 pooled pixels of a warp are one contiguous span of the pooled NHWC output.
// max commutes with the fp16 rounding, so the result equals conv -> fp16 -> maxpool bit for bit.
template <int COUT, bool VEC, bool POOL, int EPI>
__global__ void __launch_bounds__(FCfg<VEC, EPI>::THREADS, 1)
conv_first_tc_kernel(const float* __restrict__ x, const __half* __restrict__ wk, const float* __restrict__ bias,
                     __half* __restrict__ y, int n, int h, int w, int cin, int out_pitch, int act, int tiles, int dbg) {
  using F = FCfg<VEC, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* s_a = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(128) uint8_t s_b[COUT * kK * 2];
  __shared__ __align__(16) float s_bias[COUT];
  __shared__ __align__(8) uint64_t full_bar[F::STAGES], empty_bar[F::STAGES], tmem_full[F::ACCS], tmem_empty[F::ACCS];
  __shared__ uint32_t tmem_ptr;
  constexpr int TMEM_COLS = (F::ACCS * COUT <= 32) ? 32 : (F::ACCS * COUT <= 64) ? 64 : (F::ACCS * COUT <= 128) ? 128
                            : (F::ACCS * COUT <= 256) ? 256 : 512;
  static_assert(F::ACCS * COUT <= 512, "first conv: accumulators exceed TMEM");
  static_assert((F::ACCS / F::TPS) % F::EPI_GROUPS == 0, "first conv: accumulator sets must be a multiple of the epilogue groups");
  constexpr int MMA_WARP = F::PROD_WARPS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_px = n * h * w;  // < 2^31, checked by the launcher

  ptx::pdl_launch_dependents();
  // weights: wk is [COUT][32] fp16 (k = c*9 + r*3 + s, zero padded) -> canonical layout, chunk-major
  for (int i = threadIdx.x; i < COUT * 4; i += blockDim.x) {
    const int o = i >> 2, k8 = i & 3;
    *reinterpret_cast<uint4*>(s_b + k8 * (COUT * 16) + o * 16) = *reinterpret_cast<const uint4*>(wk + o * kK + k8 * 8);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_bias[i] = bias[i];
  if (warp == MMA_WARP) {
    if (ptx::elect_one()) {
      for (int s = 0; s < F::STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], F::ARRIVALS);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < F::ACCS; ++a) {
        ptx::mbar_init(&tmem_full[a], 1);
        ptx::mbar_init(&tmem_empty[a], 4);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&tmem_ptr, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();  // s_b was written through the generic proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  ptx::pdl_wait();

  if (warp < F::PROD_WARPS) {
    // ------------------------------------------------------------ im2col producers
    // producer unit u (a warp when VEC, a 4-warp group otherwise) owns stage u and every STAGES-th tile
    const uint32_t stage = VEC ? warp : (warp >> 2);
    uint32_t phase = 0;
    const int supers = (tiles + F::TPS - 1) / F::TPS;
    for (int st = blockIdx.x + stage * gridDim.x; st < supers; st += F::STAGES * gridDim.x) {
     for (int h2 = 0; h2 < F::TPS; ++h2) {
      const int tile = st * F::TPS + h2;
      uint8_t* a_tile = s_a + (stage * F::TPS + h2) * kABytes;
      if constexpr (VEC) {
        // 4 consecutive pixels of one image row per lane (w % 4 == 0).  Their MMA rows are row0 + j * rstep: for a fixed
        // j the 32 lanes write 32 rows whose 16-byte slots cover every bank group 4 times - the minimum (lane-major
        // rows 4 * lane + j would hit two bank groups 16 times each).
        int pix, row0, rstep;
        if constexpr (POOL) {
          const int cbs = w >> 5, rgs = h >> 2;
          const int cb = tile % cbs, t2 = tile / cbs;
          const int ry = lane >> 3, c4 = lane & 7;
          pix = ((t2 / rgs) * h + (t2 % rgs) * 4 + ry) * w + cb * 32 + c4 * 4;
          // row(ry, j) = 32 * (2 * (j >> 1) + (ry >> 1)) + 16 * (ry & 1) + 8 * (j & 1) + c4   (see the epilogue)
          row0 = 32 * (ry >> 1) + 16 * (ry & 1) + c4;
          rstep = 0;   // not affine in j: handled below
        } else {
          pix = tile * kTileM + lane * 4;
          row0 = lane;
          rstep = 32;
        }
        float v[4][27];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 27; ++i) v[j][i] = 0.f;
        if (pix < total_px && !(dbg & 2)) {  // ME_FIRST_DBG bit 2: attribution run without the image loads
          const int px = pix % w;
          const int t = pix / w;
          const int py = t % h;
          const int img = t / h;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c >= cin) break;
            const float* plane = x + (1LL * img * cin + c) * h * w;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int yy = py + r - 1;
              if (yy < 0 || yy >= h) continue;
              const float* rowp = plane + 1LL * yy * w + px;
              const float4 m = __ldg(reinterpret_cast<const float4*>(rowp));
              const float left = px > 0 ? __ldg(rowp - 1) : 0.f;
              const float right = px + 4 < w ? __ldg(rowp + 4) : 0.f;
              const float win[6] = {left, m.x, m.y, m.z, m.w, right};
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int s = 0; s < 3; ++s) v[j][c * 9 + r * 3 + s] = win[j + s];
            }
          }
        }
        mbar_wait_spin(&empty_bar[stage], phase ^ 1);
        if (!(dbg & 4)) {  // bit 4: no tile build either
#pragma unroll
          for (int j = 0; j < 4; ++j)
            store_row(a_tile, POOL ? row0 + 64 * (j >> 1) + 8 * (j & 1) : row0 + j * rstep, v[j]);
        }
      } else {
        const int row = threadIdx.x & 127;
        const int pix = tile * kTileM + row;
        float v[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) v[i] = 0.f;
        if (pix < total_px) {
          const int px = pix % w;
          const int t = pix / w;
          const int py = t % h;
          const int img = t / h;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c >= cin) break;
            const float* plane = x + (1LL * img * cin + c) * h * w;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int yy = py + r - 1;
              const bool yok = yy >= 0 && yy < h;
#pragma unroll
              for (int s = 0; s < 3; ++s) {
                const int xx = px + s - 1;
                if (yok && xx >= 0 && xx < w) v[c * 9 + r * 3 + s] = __ldg(plane + 1LL * yy * w + xx);
              }
            }
          }
        }
        mbar_wait_spin(&empty_bar[stage], phase ^ 1);
        store_row(a_tile, row, v);
      }
     }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full_bar[stage]);
      phase ^= 1;
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kTileM, COUT);
      const uint32_t b_addr = ptx::smem_u32(s_b);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      constexpr int SETS = F::ACCS / F::TPS;      // accumulator sets: one per super tile in flight
      const int supers = (tiles + F::TPS - 1) / F::TPS;
      for (int st = blockIdx.x; st < supers; st += gridDim.x, ++it) {
        const int set = it % SETS;
        mbar_wait_spin(&tmem_empty[set], ((it / SETS) & 1) ^ 1);
        mbar_wait_spin(&full_bar[stage], phase);
        ptx::tc_fence_after();
#pragma unroll
        for (int h2 = 0; h2 < F::TPS; ++h2) {
          const uint32_t a_addr = ptx::smem_u32(s_a + (stage * F::TPS + h2) * kABytes);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t adesc = make_nosw_desc(a_addr + k * 2 * (kTileM * 16), kTileM * 16, 128);
            const uint64_t bdesc = make_nosw_desc(b_addr + k * 2 * (COUT * 16), COUT * 16, 128);
            if (!(dbg & 16))   // bit 16: no MMAs
              ptx::umma_f16_ss(tmem_base + (set * F::TPS + h2) * COUT, adesc, bdesc, idesc, k);
          }
        }
        ptx::umma_commit(&empty_bar[stage]);
        ptx::umma_commit(&tmem_full[set]);
        if (++stage == F::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (last 4 warps)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int group = (warp - (MMA_WARP + 1)) >> 2;  // this group takes tiles it = group, group + 2, ...
    constexpr int SETS = F::ACCS / F::TPS;
    const int supers = (tiles + F::TPS - 1) / F::TPS;
    for (int it = group; blockIdx.x + it * gridDim.x < supers; it += F::EPI_GROUPS) {
     const int set = it % SETS;
     mbar_wait_spin(&tmem_full[set], (it / SETS) & 1);
     ptx::tc_fence_after();
     for (int h2 = 0; h2 < F::TPS; ++h2) {
      const int tile = (blockIdx.x + it * gridDim.x) * F::TPS + h2;
      if (tile >= tiles || (dbg & 8)) continue;   // bit 8: the epilogue only hands the accumulators back
      const int acc = set * F::TPS + h2;
      const uint32_t taddr = tmem_base + acc * COUT + (static_cast<uint32_t>(q * 32) << 16);
      // TMEM lane q * 32 + lane holds: VEC producers -> pixel 4 * lane + q of the tile; !VEC -> pixel q * 32 + lane
      int pix = tile * kTileM + (VEC ? 4 * lane + q : row);
      bool writer = pix < total_px;
      if constexpr (POOL) {
        // warp q: image rows 2 * (q & 1) + {0, 1} (lane bit 4), columns 4 * c4 + 2 * (q >> 1) + {0, 1} (lane bit 3),
        // c4 = lane & 7: the 2x2 window of pooled pixel (q & 1, 2 * c4 + (q >> 1)) sits in lanes c4 + {0, 8, 16, 24}
        const int cbs = w >> 5, rgs = h >> 2;
        const int cb = tile % cbs, t2 = tile / cbs;
        const int prow = (t2 % rgs) * 2 + (q & 1), pcol = cb * 16 + 2 * (lane & 7) + (q >> 1);
        pix = ((t2 / rgs) * (h >> 1) + prow) * (w >> 1) + pcol;
        writer = (lane & 24) == 0;
      }
      __half* dst = y + 1LL * pix * out_pitch;
      // 32-byte sectors are written whole: one 256-bit store per 16 channels when the row pitch allows it
      const bool wide = (out_pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 31) == 0;
      uint32_t r[2][16];
      ptx::tmem_ld_32x32b_x16(taddr, r[0]);
#pragma unroll
      for (int ci = 0; ci < COUT / 16; ++ci) {
        const int c = ci * 16;
        ptx::tmem_ld_wait_regs16(r[ci & 1]);
        if (ci + 1 < COUT / 16) ptx::tmem_ld_32x32b_x16(taddr + c + 16, r[(ci + 1) & 1]);
        uint4 o[2];
        __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c + 4 * e4);
          float v0 = __uint_as_float(r[ci & 1][4 * e4 + 0]) + b4.x;
          float v1 = __uint_as_float(r[ci & 1][4 * e4 + 1]) + b4.y;
          float v2 = __uint_as_float(r[ci & 1][4 * e4 + 2]) + b4.z;
          float v3 = __uint_as_float(r[ci & 1][4 * e4 + 3]) + b4.w;
          if (act == ME_ACT_LEAKY) {
            v0 = fmaxf(v0, 0.1f * v0);
            v1 = fmaxf(v1, 0.1f * v1);
            v2 = fmaxf(v2, 0.1f * v2);
            v3 = fmaxf(v3, 0.1f * v3);
          } else if (act == ME_ACT_SIGMOID) {
            v0 = 1.f / (1.f + __expf(-v0));
            v1 = 1.f / (1.f + __expf(-v1));
            v2 = 1.f / (1.f + __expf(-v2));
            v3 = 1.f / (1.f + __expf(-v3));
          }
          oh[2 * e4] = __floats2half2_rn(v0, v1);
          oh[2 * e4 + 1] = __floats2half2_rn(v2, v3);
        }
        if constexpr (POOL) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            __half2 m = oh[e];
            uint32_t u = *reinterpret_cast<uint32_t*>(&m);
            uint32_t o1 = __shfl_xor_sync(0xffffffffu, u, 8);
            m = __hmax2(m, *reinterpret_cast<__half2*>(&o1));
            u = *reinterpret_cast<uint32_t*>(&m);
            uint32_t o2 = __shfl_xor_sync(0xffffffffu, u, 16);
            oh[e] = __hmax2(m, *reinterpret_cast<__half2*>(&o2));
          }
        }
        if (writer && !(dbg & 1)) {  // bit 1: attribution run without the output stores
          if (wide) {
            ptx::st_global_256(dst + c, o[0], o[1]);
          } else {
            *reinterpret_cast<uint4*>(dst + c) = o[0];
            *reinterpret_cast<uint4*>(dst + c + 8) = o[1];
          }
        }
      }
     }
     ptx::tc_fence_before();
     __syncwarp();
     if (lane == 0) ptx::mbar_arrive(&tmem_empty[set]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == MMA_WARP) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

__global__ void pack_first_tc_kernel(const float* __restrict__ w_folded, int cout, int per_out, __half* __restrict__ wk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * kK) return;
  const int o = i / kK, k = i - o * kK;
  wk[i] = __float2half_rn(k < per_out ? w_folded[o * per_out + k] : 0.f);
}

template <int COUT, bool VEC, bool POOL = false, int EPI = 2>
int launch_first(const float* x, const __half* wk, const float* bias, __half* y, int n, int h, int w, int cin,
                 int out_pitch, int act, int tiles, int grid, cudaStream_t stream) {
  using F = FCfg<VEC, EPI>;
  auto kern = conv_first_tc_kernel<COUT, VEC, POOL, EPI>;
  static bool attr_seen[64] = {false};   // per instantiation and per device
  if (first_use_on_device(attr_seen))
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM));
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("ME_FIRST_DBG");
    dbg = e ? atoi(e) : 0;
  }
  kern<<<grid, F::THREADS, F::SMEM, stream>>>(x, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, dbg);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // namespace
}  // namespace me

namespace me {
namespace {

int conv_first_tc_impl(const float* x_nchw, const float* w_folded, const float* bias, void* wk_scratch, void* y_nhwc,
                       int n, int h, int w, int cin, int cout, int out_pitch, int act, bool pool, cudaStream_t stream) {
  ME_REQUIRE(x_nchw && w_folded && bias && wk_scratch && y_nhwc, "conv_first_tc: null argument");
  ME_REQUIRE(cin >= 1 && cin <= 3, "conv_first_tc: cin %d must be 1..3", cin);
  ME_REQUIRE(out_pitch >= cout && out_pitch % 8 == 0, "conv_first_tc: bad out_pitch %d", out_pitch);
  __half* wk = static_cast<__half*>(wk_scratch);
  pack_first_tc_kernel<<<ceil_div(cout * kK, 128), 128, 0, stream>>>(w_folded, cout, cin * 9, wk);
  ME_LAUNCH_CHECK();
  const long long total = 1LL * n * h * w;
  ME_REQUIRE(total > 0 && total < (1LL << 31) - kTileM, "conv_first_tc: pixel count out of range");
  const int tiles = static_cast<int>((total + kTileM - 1) / kTileM);
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (grid > tiles) grid = tiles;
  __half* y = static_cast<__half*>(y_nhwc);
  const bool vec = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0);
  if (pool) {
    ME_REQUIRE(vec && w % 32 == 0 && h % 4 == 0, "conv_first_tc_pool: needs a 16-byte aligned image with w %% 32 == 0 and "
               "h %% 4 == 0 (got %d x %d); run me_conv_first_tc + me_maxpool instead", h, w);
    static int epi3 = -1;
    if (epi3 < 0) {
      const char* e = getenv("ME_FIRST_EPI");
      epi3 = e ? atoi(e) : 3;   // three epilogue groups: the pooled epilogue's latency is what bounds the kernel with two
    }
    switch (cout) {   // tiles are exact 4 x 32 blocks: n * (h / 4) * (w / 32) == total / 128
      case 16:
        if (epi3 == 3) return launch_first<16, true, true, 3>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream);
        if (epi3 == 4) return launch_first<16, true, true, 4>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream);
        return launch_first<16, true, true>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream);
      case 32: return launch_first<32, true, true>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream);
      default: return fail(ME_ERR_UNSUPPORTED, "conv_first_tc_pool: cout %d unsupported (16 or 32)", cout);
    }
  }
  static int epi_plain = -1;
  if (epi_plain < 0) {
    const char* e = getenv("ME_FIRST_EPI_PLAIN");
    epi_plain = e ? atoi(e) : 2;
  }
  if (vec && epi_plain == 3 && cout == 32)
    return launch_first<32, true, false, 3>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream);
  if (vec && epi_plain == 3 && cout == 16)
    return launch_first<16, true, false, 3>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream);
#define ME_FIRST(C)                                                                                              \
  return vec ? launch_first<C, true>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream)     \
             : launch_first<C, false>(x_nchw, wk, bias, y, n, h, w, cin, out_pitch, act, tiles, grid, stream)
  switch (cout) {
    case 16: ME_FIRST(16);
    case 32: ME_FIRST(32);
    case 64: ME_FIRST(64);
    default: return fail(ME_ERR_UNSUPPORTED, "conv_first_tc: cout %d unsupported (16, 32 or 64)", cout);
  }
#undef ME_FIRST
}

}  // namespace
}  // namespace me

extern "C" {

// wk: fp16 [cout][32] scratch owned by the caller (filled here from the folded fp32 weights).
int me_conv_first_tc(const float* x_nchw, const float* w_folded, const float* bias, void* wk_scratch, void* y_nhwc,
                     int n, int h, int w, int cin, int cout, int out_pitch, int act, me_stream_t stream_) {
  return me::conv_first_tc_impl(x_nchw, w_folded, bias, wk_scratch, y_nhwc, n, h, w, cin, cout, out_pitch, act, false,
                                static_cast<cudaStream_t>(stream_));
}

// The same layer with the 2x2 / stride-2 max-pool that follows it applied in the epilogue: y is the POOLED
// (n, h/2, w/2, out_pitch) NHWC tensor.
int me_conv_first_tc_pool(const float* x_nchw, const float* w_folded, const float* bias, void* wk_scratch, void* y_nhwc,
                          int n, int h, int w, int cin, int cout, int out_pitch, int act, me_stream_t stream_) {
  return me::conv_first_tc_impl(x_nchw, w_folded, bias, wk_scratch, y_nhwc, n, h, w, cin, cout, out_pitch, act, true,
                                static_cast<cudaStream_t>(stream_));
}

}  // extern "C"
