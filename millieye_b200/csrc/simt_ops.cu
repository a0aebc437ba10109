// HBM-bound glue of the Darknet forward: weight folding/packing, the 3-channel first conv,
// max-pool, nearest upsample into a concat slice, layout converts and the YOLO anchor decode.
// All NHWC fp16 kernels move 16 bytes (8 channels) per thread so warps issue full 128-bit
// coalesced requests.
#include "common.cuh"

namespace me {
namespace {

// ------------------------------------------------------------------ weight folding / packing
__device__ __forceinline__ float bn_scale(const float* gamma, const float* var, float eps, int o) {
  return gamma ? gamma[o] / sqrtf(var[o] + eps) : 1.f;
}

__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ var, float eps, int cout, int cin, int ksize,
                                    int cin_pad, int cout_pad, __half* __restrict__ out) {
  const int taps = ksize * ksize;
  const long long ktot = 1LL * taps * cin_pad;
  const long long total = ktot * cout_pad;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int o = static_cast<int>(i / ktot);
    const int k = static_cast<int>(i - o * ktot);
    const int tap = k / cin_pad;
    const int c = k - tap * cin_pad;
    float v = 0.f;
    if (o < cout && c < cin) v = w[(1LL * o * cin + c) * taps + tap] * bn_scale(gamma, var, eps, o);
    out[i] = __float2half_rn(v);
  }
}

__global__ void fold_bias_kernel(const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                 const float* __restrict__ var, float eps, int cout, int n_out,
                                 float* __restrict__ bias_out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  float b = 0.f;
  if (o < cout) {
    const float cb = conv_bias ? conv_bias[o] : 0.f;
    if (gamma) {
      const float s = bn_scale(gamma, var, eps, o);
      b = (cb - mean[o]) * s + beta[o];
    } else {
      b = cb;
    }
  }
  bias_out[o] = b;
}

__global__ void fold_first_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                  const float* __restrict__ var, float eps, int cout, int per_out,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * per_out) return;
  out[i] = w[i] * bn_scale(gamma, var, eps, i / per_out);
}

// ------------------------------------------------------------------ first conv (cin <= 4)
// One thread = one output pixel, all COUT channels in registers; weights broadcast from smem.
template <int COUT>
__global__ void __launch_bounds__(128) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ wf,
                                                         const float* __restrict__ bias, __half* __restrict__ y,
                                                         int n, int h, int w, int cin, int out_pitch, int act) {
  __shared__ float s_w[4 * 9 * COUT];
  __shared__ float s_b[COUT];
  const int kk = cin * 9;
  for (int i = threadIdx.x; i < kk * COUT; i += blockDim.x) {
    const int o = i / kk, k = i - o * kk;  // wf is [o][c][r][s]
    s_w[k * COUT + o] = wf[i];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_b[i] = bias[i];
  __syncthreads();
  const long long total = 1LL * n * h * w;
  const long long pix = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int px = static_cast<int>(pix % w);
  const int py = static_cast<int>((pix / w) % h);
  const int img = static_cast<int>(pix / (1LL * w * h));
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = s_b[o];
  for (int c = 0; c < cin; ++c) {
    const float* plane = x + (1LL * img * cin + c) * h * w;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = py + r - 1;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int xx = px + s - 1;
        const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(plane + 1LL * yy * w + xx) : 0.f;
        const float* wk = s_w + ((c * 3 + r) * 3 + s) * COUT;
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(v, wk[o], acc[o]);
      }
    }
  }
  __half* dst = y + pix * out_pitch;
#pragma unroll
  for (int o = 0; o < COUT; o += 8) {
    uint4 pk;
    __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = acc[o + 2 * e], b = acc[o + 2 * e + 1];
      if (act == ME_ACT_LEAKY) {
        a = a > 0.f ? a : 0.1f * a;
        b = b > 0.f ? b : 0.1f * b;
      } else if (act == ME_ACT_SIGMOID) {
        a = 1.f / (1.f + expf(-a));
        b = 1.f / (1.f + expf(-b));
      }
      ph[e] = __floats2half2_rn(a, b);
    }
    *reinterpret_cast<uint4*>(dst + o) = pk;
  }
}

// ------------------------------------------------------------------ pooling / upsample / copies
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__global__ void maxpool2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int c8,
                                int in_pitch, int out_pitch, int stride, int ho, int wo) {
  const long long total = 1LL * n * ho * wo * c8;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    const int iy = oy * stride, ix = ox * stride;
    const __half* base = x + ((1LL * img * h + iy) * w + ix) * in_pitch + cg * 8;
    const uint4 zero = make_uint4(0, 0, 0, 0);  // ZeroPad2d((0,1,0,1)) ahead of the stride-1 pool
    const bool has_r = ix + 1 < w, has_b = iy + 1 < h;
    uint4 m = *reinterpret_cast<const uint4*>(base);
    m = hmax8(m, has_r ? *reinterpret_cast<const uint4*>(base + in_pitch) : zero);
    m = hmax8(m, has_b ? *reinterpret_cast<const uint4*>(base + 1LL * w * in_pitch) : zero);
    m = hmax8(m, (has_r && has_b) ? *reinterpret_cast<const uint4*>(base + 1LL * (w + 1) * in_pitch) : zero);
    *reinterpret_cast<uint4*>(y + ((1LL * img * ho + oy) * wo + ox) * out_pitch + cg * 8) = m;
  }
}

__global__ void upsample2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int c8,
                                 int in_pitch, int out_pitch) {
  const long long total = 1LL * n * h * w * c8;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ix = static_cast<int>(t % w);
    t /= w;
    const int iy = static_cast<int>(t % h);
    const int img = static_cast<int>(t / h);
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((1LL * img * h + iy) * w + ix) * in_pitch + cg * 8);
    __half* o = y + ((1LL * img * 2 * h + 2 * iy) * (2 * w) + 2 * ix) * out_pitch + cg * 8;
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + out_pitch) = v;
    *reinterpret_cast<uint4*>(o + 2LL * w * out_pitch) = v;
    *reinterpret_cast<uint4*>(o + (2LL * w + 1) * out_pitch) = v;
  }
}

// ToTensor on the device: uint8 0..255 -> fp32 0..1 (x / 255, the division torchvision's ToTensor performs), 16 values per thread
__global__ void u8_to_unit_f32_kernel(const uint4* __restrict__ x, float4* __restrict__ y, long long n16, const unsigned char* tail_x,
                                      float* tail_y, int tail) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n16; i += 1LL * gridDim.x * blockDim.x) {
    const uint4 v = x[i];
    const unsigned int wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int u = wds[k];
      y[i * 4 + k] = make_float4(static_cast<float>(u & 0xffu) / 255.f, static_cast<float>((u >> 8) & 0xffu) / 255.f,
                                 static_cast<float>((u >> 16) & 0xffu) / 255.f, static_cast<float>(u >> 24) / 255.f);
    }
  }
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < tail) tail_y[threadIdx.x] = static_cast<float>(tail_x[threadIdx.x]) / 255.f;
}

__global__ void copy_channels_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long pixels, int c8,
                                     int in_pitch, int out_pitch) {
  const long long total = pixels * c8;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    const long long p = i / c8;
    *reinterpret_cast<uint4*>(y + p * out_pitch + cg * 8) = *reinterpret_cast<const uint4*>(x + p * in_pitch + cg * 8);
  }
}

// [pixels][pitch] fp16 -> [n][c][hw] fp32 through a 32x33 smem tile (both sides coalesced).
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ x, float* __restrict__ y, int hw, int c, int in_pitch) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, ch = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < hw && ch < c) ? __half2float(x[(1LL * img * hw + p) * in_pitch + ch]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ch = c0 + r, p = p0 + threadIdx.x;
    if (ch < c && p < hw) y[(1LL * img * c + ch) * hw + p] = tile[threadIdx.x][r];
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __half* __restrict__ y, int hw, int c, int out_pitch) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ch = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (ch < c && p < hw) ? x[(1LL * img * c + ch) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, ch = c0 + threadIdx.x;
    if (p < hw && ch < out_pitch) y[(1LL * img * hw + p) * out_pitch + ch] = __float2half_rn(tile[threadIdx.x][r]);
  }
}

// ------------------------------------------------------------------ YOLO decode
struct Anchors {
  float w[8];
  float h[8];
};

// One block = kDecodeCells consecutive cells of one grid row, one head channel per thread: reads are contiguous over
// the head's channel axis, writes contiguous over the 5+C attributes of an output row, and every thread has
// kDecodeCells independent loads in flight.  The cell -> (image, gy, gx) split is done once per block; with the
// div/mod chain per element the 52x52 head was instruction bound (~150 instructions per element, 1.1 TB/s).
constexpr int kDecodeCells = 8;

__device__ __forceinline__ float decode_one(float v, int k, float g_xy, float anc_wh, float stride) {
  if (k < 2) return (1.f / (1.f + expf(-v)) + g_xy) * stride;
  if (k < 4) return (expf(v) * anc_wh) * stride;
  return 1.f / (1.f + expf(-v));
}

__global__ void __launch_bounds__(256)
yolo_decode_kernel(const float* __restrict__ logits, int pitch, float* __restrict__ out, int g, int chunks, int na,
                   int attrs, Anchors anc, float stride, int rows_total, int row_offset) {
  const int per_cell = na * attrs;
  const int chunk = blockIdx.x % chunks;
  const int line = blockIdx.x / chunks;  // img * g + gy
  const int gy = line % g, img = line / g;
  const int gx0 = chunk * kDecodeCells;
  const int ncell = min(kDecodeCells, g - gx0);
  const float* src = logits + (1LL * line * g + gx0) * pitch;
  for (int ch = threadIdx.x; ch < per_cell; ch += blockDim.x) {
    const int a = ch / attrs, k = ch - a * attrs;
    float v[kDecodeCells];
#pragma unroll
    for (int j = 0; j < kDecodeCells; ++j) v[j] = j < ncell ? __ldg(src + 1LL * j * pitch + ch) : 0.f;
    const float anc_wh = k == 2 ? anc.w[a] : anc.h[a];
    float* dst = out + (1LL * img * rows_total + row_offset + (a * g + gy) * g + gx0) * attrs + k;
#pragma unroll
    for (int j = 0; j < kDecodeCells; ++j) {
      if (j >= ncell) break;
      const float g_xy = static_cast<float>(k == 0 ? gx0 + j : gy);
      dst[1LL * j * attrs] = decode_one(v[j], k, g_xy, anc_wh, stride);
    }
  }
}

inline int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace
}  // namespace me

extern "C" {

int me_pack_conv_weights(const float* w_oihw, const float* conv_bias, const float* bn_gamma, const float* bn_beta,
                         const float* bn_mean, const float* bn_var, float bn_eps, int cout, int cin, int ksize,
                         int cout_pad, void* w_packed, float* bias_out, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(w_oihw && w_packed && bias_out, "pack: null argument");
  ME_REQUIRE(cout > 0 && cin > 0 && cout_pad >= cout, "pack: bad sizes");
  ME_REQUIRE((bn_gamma == nullptr) == (bn_var == nullptr) && (bn_gamma == nullptr) == (bn_mean == nullptr) &&
                 (bn_gamma == nullptr) == (bn_beta == nullptr),
             "pack: batch-norm pointers must be all set or all null");
  const int cin_pad = me_conv_cin_pad(cin);
  const long long total = 1LL * ksize * ksize * cin_pad * cout_pad;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(w_oihw, bn_gamma, bn_var, bn_eps, cout, cin, ksize,
                                                                cin_pad, cout_pad, static_cast<__half*>(w_packed));
  ME_LAUNCH_CHECK();
  const int n_bias = round_up(cout_pad, 256);  // the GEMM epilogue reads whole BN-wide slices
  fold_bias_kernel<<<ceil_div(n_bias, 256), 256, 0, stream>>>(conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, bn_eps,
                                                              cout, n_bias, bias_out);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_fold_first_weights(const float* w_oihw, const float* conv_bias, const float* bn_gamma, const float* bn_beta,
                          const float* bn_mean, const float* bn_var, float bn_eps, int cout, int cin, float* w_folded,
                          float* bias_out, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(w_oihw && w_folded && bias_out, "fold_first: null argument");
  ME_REQUIRE(cin >= 1 && cin <= 4 && cout >= 1, "fold_first: cin must be 1..4");
  const int per_out = cin * 9;
  fold_first_kernel<<<ceil_div(cout * per_out, 256), 256, 0, stream>>>(w_oihw, bn_gamma, bn_var, bn_eps, cout, per_out,
                                                                       w_folded);
  ME_LAUNCH_CHECK();
  fold_bias_kernel<<<ceil_div(cout, 256), 256, 0, stream>>>(conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, bn_eps, cout,
                                                            cout, bias_out);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_conv_first(const float* x_nchw, const float* w_folded, const float* bias, void* y_nhwc, int n, int h, int w,
                  int cin, int cout, int out_pitch, int act, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(x_nchw && w_folded && bias && y_nhwc, "conv_first: null argument");
  ME_REQUIRE(cin >= 1 && cin <= 4, "conv_first: cin %d must be 1..4", cin);
  ME_REQUIRE(out_pitch >= cout && out_pitch % 8 == 0, "conv_first: bad out_pitch %d", out_pitch);
  const long long total = 1LL * n * h * w;
  const int blocks = static_cast<int>((total + 127) / 128);
  __half* y = static_cast<__half*>(y_nhwc);
  switch (cout) {
    case 16: conv_first_kernel<16><<<blocks, 128, 0, stream>>>(x_nchw, w_folded, bias, y, n, h, w, cin, out_pitch, act); break;
    case 32: conv_first_kernel<32><<<blocks, 128, 0, stream>>>(x_nchw, w_folded, bias, y, n, h, w, cin, out_pitch, act); break;
    case 64: conv_first_kernel<64><<<blocks, 128, 0, stream>>>(x_nchw, w_folded, bias, y, n, h, w, cin, out_pitch, act); break;
    default: return fail(ME_ERR_UNSUPPORTED, "conv_first: cout %d unsupported (16, 32 or 64)", cout);
  }
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_maxpool2(const void* x, void* y, int n, int h, int w, int c, int in_pitch, int out_pitch, int stride,
                me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(x && y, "maxpool: null argument");
  ME_REQUIRE(c % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "maxpool: channels/pitches must be multiples of 8");
  ME_REQUIRE(stride == 1 || stride == 2, "maxpool: stride %d unsupported", stride);
  const int ho = stride == 2 ? h / 2 : h, wo = stride == 2 ? w / 2 : w;
  const long long total = 1LL * n * ho * wo * (c / 8);
  maxpool2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(y), n, h,
                                                            w, c / 8, in_pitch, out_pitch, stride, ho, wo);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_upsample2(const void* x, void* y, int n, int h, int w, int c, int in_pitch, int out_pitch, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(x && y, "upsample: null argument");
  ME_REQUIRE(c % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "upsample: channels/pitches must be multiples of 8");
  const long long total = 1LL * n * h * w * (c / 8);
  upsample2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(y), n,
                                                             h, w, c / 8, in_pitch, out_pitch);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_copy_channels(const void* x, void* y, long long pixels, int c, int in_pitch, int out_pitch, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(x && y, "copy_channels: null argument");
  ME_REQUIRE(c % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "copy_channels: multiples of 8 required");
  copy_channels_kernel<<<grid_for(pixels * (c / 8), 256), 256, 0, stream>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), pixels, c / 8, in_pitch, out_pitch);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_nhwc_to_nchw_f32(const void* x, float* y, int n, int h, int w, int c, int in_pitch, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(x && y, "nhwc_to_nchw: null argument");
  const int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, stream>>>(static_cast<const __half*>(x), y, hw, c, in_pitch);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_u8_to_unit_f32(const void* x, float* y, long long count, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (count <= 0) return ME_OK;
  ME_REQUIRE(x && y, "u8_to_unit_f32: null argument");
  ME_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "u8_to_unit_f32: buffers must be 16-byte aligned");
  const long long n16 = count / 16;
  const int tail = static_cast<int>(count - n16 * 16);
  const unsigned char* xb = static_cast<const unsigned char*>(x);
  u8_to_unit_f32_kernel<<<grid_for(n16 > 0 ? n16 : 1, 256), 256, 0, stream>>>(static_cast<const uint4*>(x), reinterpret_cast<float4*>(y),
                                                                             n16, xb + n16 * 16, y + n16 * 16, tail);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_nchw_f32_to_nhwc(const float* x, void* y, int n, int h, int w, int c, int out_pitch, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(x && y, "nchw_to_nhwc: null argument");
  const int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(out_pitch, 32), n), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, stream>>>(x, static_cast<__half*>(y), hw, c, out_pitch);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_yolo_decode(const float* logits, int pitch, float* out, int n, int g, int num_anchors, int num_classes,
                   const float* host_anchors_wh, float stride, int rows_total, int row_offset, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(logits && out && host_anchors_wh, "yolo_decode: null argument");
  ME_REQUIRE(num_anchors >= 1 && num_anchors <= 8, "yolo_decode: 1..8 anchors");
  const int attrs = 5 + num_classes;
  ME_REQUIRE(pitch >= num_anchors * attrs, "yolo_decode: pitch %d < %d head channels", pitch, num_anchors * attrs);
  Anchors anc{};
  for (int a = 0; a < num_anchors; ++a) {
    // models.py:127 keeps anchors / stride as float32 (FloatTensor of python doubles), :171 multiplies back
    anc.w[a] = static_cast<float>(static_cast<double>(host_anchors_wh[2 * a]) / static_cast<double>(stride));
    anc.h[a] = static_cast<float>(static_cast<double>(host_anchors_wh[2 * a + 1]) / static_cast<double>(stride));
  }
  const long long cells = 1LL * n * g * g;
  ME_REQUIRE(cells < (1LL << 31), "yolo_decode: too many cells");
  const int block = num_anchors * attrs >= 192 ? 256 : (num_anchors * attrs >= 96 ? 128 : 64);
  const int chunks = (g + kDecodeCells - 1) / kDecodeCells;
  const long long blocks = 1LL * n * g * chunks;
  yolo_decode_kernel<<<static_cast<int>(blocks), block, 0, stream>>>(logits, pitch, out, g, chunks, num_anchors, attrs,
                                                                     anc, stride, rows_total, row_offset);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
