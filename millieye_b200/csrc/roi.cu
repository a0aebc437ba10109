// RoI gathers of Network.forward (my_models.py:495-496), restating the torchvision CPU kernels
// the reference calls (third-party, torchvision 0.26 ps_roi_align_kernel.cpp / roi_align_kernel.cpp;
// semantics listed in SURVEY.md §8a A10/A11):
//   ps_roi_align: start = coord*scale - 0.5, size = end - start (no clamp), channel of output
//                 (c, ph, pw) is (c*P + ph)*P + pw, grid = ceil(size / P), mean of bilinear samples.
//   roi_align   : aligned=False -> start = coord*scale, size = max(end - start, 1).
//   bilinear    : 0 outside [-1, H] x [-1, W]; clamp to >= 0; high index clamped to H-1 / W-1.
// One thread per output element; the score maps (a few hundred KB per frame) stay L2 resident,
// so the gathers are latency- not bandwidth-bound.  Output rows are fp16 and feed the
// tensor-core GEMM of refinement_head.net0 directly (flatten order of my_models.py:262).
#include "common.cuh"

namespace me {
namespace {

__device__ __forceinline__ float bilinear(const __half* __restrict__ plane, int h, int w, int pitch, float y, float x) {
  if (y < -1.0f || y > static_cast<float>(h) || x < -1.0f || x > static_cast<float>(w)) return 0.f;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = static_cast<int>(y), x_low = static_cast<int>(x);
  int y_high, x_high;
  if (y_low >= h - 1) {
    y_high = y_low = h - 1;
    y = static_cast<float>(y_low);
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= w - 1) {
    x_high = x_low = w - 1;
    x = static_cast<float>(x_low);
  } else {
    x_high = x_low + 1;
  }
  const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
  const float v1 = __half2float(plane[(1LL * y_low * w + x_low) * pitch]);
  const float v2 = __half2float(plane[(1LL * y_low * w + x_high) * pitch]);
  const float v3 = __half2float(plane[(1LL * y_high * w + x_low) * pitch]);
  const float v4 = __half2float(plane[(1LL * y_high * w + x_high) * pitch]);
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

// BIN_MAJOR: the output row is ordered [bin = ph * P + pw][c] instead of the reference's [c][ph][pw], and the
// position-sensitive score map holds channel (ph * P + pw) * C + c: consecutive threads then read consecutive channels
// of ONE pixel neighbourhood (a single 32-byte sector serves a bin's 10 channels) instead of 2-byte loads from 10
// different sectors.  The caller permutes the producing conv's output channels and the consuming layer's input
// columns once at pack time, so the result is the reference's up to that fixed permutation.
template <bool POSITION_SENSITIVE, bool BIN_MAJOR>
__global__ void roi_gather_kernel(const __half* __restrict__ feat, int n, int h, int w, int pitch, int channels,
                                  int pooled, float scale, const float* __restrict__ rois,
                                  const int* __restrict__ roi_count, int count_index, int cap,
                                  __half* __restrict__ out, int out_pitch) {
  const int live = min(roi_count[count_index], cap);
  const int per_roi = channels * pooled * pooled;
  const long long total = 1LL * cap * out_pitch;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int r = static_cast<int>(i / out_pitch);
    const int e = static_cast<int>(i - 1LL * r * out_pitch);
    float val = 0.f;
    if (r < live && e < per_roi) {
      const int c = BIN_MAJOR ? e % channels : e / (pooled * pooled);
      const int bin = BIN_MAJOR ? e / channels : e % (pooled * pooled);
      const int ph = bin / pooled;
      const int pw = bin % pooled;
      const float* roi = rois + r * 5;
      int b = static_cast<int>(roi[0]);
      b = b < 0 ? 0 : (b >= n ? n - 1 : b);
      const float off = POSITION_SENSITIVE ? 0.5f : 0.f;
      const float sw = roi[1] * scale - off, sh = roi[2] * scale - off;
      const float ew = roi[3] * scale - off, eh = roi[4] * scale - off;
      float rw = ew - sw, rh = eh - sh;
      if (!POSITION_SENSITIVE) {
        rw = fmaxf(rw, 1.f);
        rh = fmaxf(rh, 1.f);
      }
      const float bin_h = rh / pooled, bin_w = rw / pooled;
      const int gh = static_cast<int>(ceilf(rh / pooled)), gw = static_cast<int>(ceilf(rw / pooled));
      const int ch = POSITION_SENSITIVE ? (BIN_MAJOR ? bin * channels + c : (c * pooled + ph) * pooled + pw) : c;
      const __half* plane = feat + 1LL * b * h * w * pitch + ch;
      const float hstart = ph * bin_h + sh, wstart = pw * bin_w + sw;
      float sum = 0.f;
      for (int iy = 0; iy < gh; ++iy) {
        const float y = hstart + (iy + 0.5f) * bin_h / static_cast<float>(gh);
        for (int ix = 0; ix < gw; ++ix) {
          const float x = wstart + (ix + 0.5f) * bin_w / static_cast<float>(gw);
          sum += bilinear(plane, h, w, pitch, y, x);
        }
      }
      // ps_roi_align divides by gh*gw as is (0/0 -> NaN for an inverted box, as torchvision
      // does); roi_align uses max(gh*gw, 1).
      const float cnt = POSITION_SENSITIVE ? static_cast<float>(gh * gw) : static_cast<float>(max(gh * gw, 1));
      val = sum / cnt;
    }
    out[i] = __float2half_rn(val);
  }
}

template <bool PS, bool BM = false>
int launch(const void* feat, int n, int h, int w, int pitch, int channels, int pooled, float scale, const float* rois,
           const int* roi_count, int cap, void* out, int out_pitch, cudaStream_t stream) {
  ME_REQUIRE(feat && rois && roi_count && out, "roi: null argument");
  ME_REQUIRE(cap > 0 && pooled > 0 && channels > 0, "roi: empty problem");
  ME_REQUIRE(out_pitch >= channels * pooled * pooled, "roi: out_pitch %d < %d", out_pitch, channels * pooled * pooled);
  ME_REQUIRE(pitch >= (PS ? channels * pooled * pooled : channels), "roi: feature pitch %d too small", pitch);
  const long long total = 1LL * cap * out_pitch;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  roi_gather_kernel<PS, BM><<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const __half*>(feat), n, h, w, pitch,
                                                                      channels, pooled, scale, rois, roi_count, 0, cap,
                                                                      static_cast<__half*>(out), out_pitch);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // namespace
}  // namespace me

extern "C" {

int me_psroi_align(const void* feat, int n, int h, int w, int pitch, int out_channels, int pooled, float spatial_scale,
                   const float* rois, const int* roi_count, int cap, void* out, int out_pitch, me_stream_t stream) {
  return me::launch<true>(feat, n, h, w, pitch, out_channels, pooled, spatial_scale, rois, roi_count, cap, out, out_pitch,
                          static_cast<cudaStream_t>(stream));
}

int me_roi_align(const void* feat, int n, int h, int w, int pitch, int channels, int pooled, float spatial_scale,
                 const float* rois, const int* roi_count, int cap, void* out, int out_pitch, me_stream_t stream) {
  return me::launch<false>(feat, n, h, w, pitch, channels, pooled, spatial_scale, rois, roi_count, cap, out, out_pitch,
                           static_cast<cudaStream_t>(stream));
}

int me_roi_gather_bin_major(const void* feat, int n, int h, int w, int pitch, int channels, int pooled, float spatial_scale,
                            const float* rois, const int* roi_count, int cap, void* out, int out_pitch,
                            int position_sensitive, me_stream_t stream) {
  if (position_sensitive)
    return me::launch<true, true>(feat, n, h, w, pitch, channels, pooled, spatial_scale, rois, roi_count, cap, out, out_pitch,
                                  static_cast<cudaStream_t>(stream));
  return me::launch<false, true>(feat, n, h, w, pitch, channels, pooled, spatial_scale, rois, roi_count, cap, out, out_pitch,
                                 static_cast<cudaStream_t>(stream));
}

}  // extern "C"
