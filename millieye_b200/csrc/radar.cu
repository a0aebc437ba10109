// Radar point cloud -> network input map, on the device (SURVEY.md §8f rows f1 + f2):
//   from_3d_to_2d / projection_xyr_to_uv   data_collection/utils/utils.py:81-120  (pinhole + plumb-bob, float64,
//                                           astype(int64) truncation)
//   FOV / depth / velocity filter           data_collection/prepare_data.py:108, run_mp.py:83
//   plot_radar_heatmap                      utils/datasets.py:56-106 (three np.histogram2d maps: count, mean depth
//                                           with the <1 -> 100 sentinel, |mean velocity|; clip ranges (0,5),(12,0),(0,4))
//   pad_to_square                           utils/datasets.py:16-26
//   bilinear resize, align_corners=True     utils/datasets.py:320-322 (collate_fn)
// One block per frame.  All reference arithmetic here is float64 numpy; it is replayed in fp64 with the _rn
// intrinsics (no FMA contraction) and per-bin sums run in point order like np.bincount, so the integer pixel
// coordinates, the bin assignment and the histogram values are reproduced exactly; the final resize is fp32.
#include "common.cuh"

namespace me {
namespace {

constexpr int kRadarThreads = 256;
constexpr int kMaxPoints = 1024;
constexpr int kMaxMap = 32;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// np.searchsorted(edges, x, side="right") - 1 with the last edge folded into the last bin (np.histogramdd)
__device__ __forceinline__ int bin_of(const double* edges, int bins, double x) {
  if (!(x >= edges[0]) || !(x <= edges[bins])) return -1;
  int k = 0;
  for (int e = 1; e <= bins; ++e) k += (edges[e] <= x) ? 1 : 0;
  return k >= bins ? bins - 1 : k;
}

__global__ void __launch_bounds__(kRadarThreads)
radar_maps_kernel(const float* __restrict__ points, const int* __restrict__ counts, int cap, me_radar_cfg cfg,
                  float* __restrict__ maps_out, float* __restrict__ uvzv_out, int* __restrict__ kept_out) {
  __shared__ int s_bin[kMaxPoints];          // flattened bin of every kept point, -1 otherwise
  __shared__ double s_depth[kMaxPoints], s_vel[kMaxPoints];
  __shared__ int s_u[kMaxPoints], s_v[kMaxPoints];
  __shared__ float s_map[3][kMaxMap][kMaxMap];  // after clip + pad_to_square
  const int frame = blockIdx.x;
  const int tid = threadIdx.x;
  const int count = min(counts[frame], min(cap, kMaxPoints));
  const double fx = cfg.calib[0], cx = cfg.calib[1], fy = cfg.calib[2], cy = cfg.calib[3];
  const double k1 = cfg.calib[4], k2 = cfg.calib[5], t1 = cfg.calib[6], t2 = cfg.calib[7], k3 = cfg.calib[8];
  const double tx = cfg.calib[9], ty = cfg.calib[10], tz = cfg.calib[11];

  for (int i = tid; i < count; i += kRadarThreads) {
    const float* pt = points + (1LL * frame * cap + i) * 4;
    // radar (x, y, z) -> camera (x, -z, y)   utils.py:113
    const double px = pt[0], py = -static_cast<double>(pt[2]), pz = pt[1], vel = pt[3];
    const double zt = dadd(pz, tz);
    const double x = dadd(px, tx) / zt, y = dadd(py, ty) / zt;
    const double x2 = dmul(x, x), y2 = dmul(y, y);
    const double r2 = dadd(x2, y2);
    const double r4 = dmul(r2, r2);
    const double r6 = pow(r2, 3.0);
    const double tmp = dadd(dadd(dadd(1.0, dmul(k1, r2)), dmul(k2, r4)), dmul(k3, r6));
    const double xu = dadd(dadd(dmul(x, tmp), dmul(dmul(dmul(2.0, t1), x), y)), dmul(t2, dadd(r2, dmul(2.0, x2))));
    const double yu = dadd(dadd(dmul(y, tmp), dmul(dmul(dmul(2.0, t2), x), y)), dmul(t1, dadd(r2, dmul(2.0, y2))));
    const double uf = dadd(dmul(xu, fx), cx), vf = dadd(dmul(yu, fy), cy);
    const long long u = static_cast<long long>(uf), v = static_cast<long long>(vf);  // astype(np.int64)
    // numpy turns a NaN coordinate (0/0 for a point at the sensor origin) into INT64_MIN, which the FOV test drops;
    // CUDA's conversion would give 0, so NaNs are rejected explicitly
    const bool keep = (uf == uf) && (vf == vf) && u >= 0 && u < cfg.img_w && v >= 0 && v < cfg.img_h &&
                      zt < cfg.max_depth && fabs(vel) >= cfg.min_velocity;
    int b = -1;
    if (keep) {
      const int bw = bin_of(cfg.edges_w, cfg.bin_w, static_cast<double>(u));
      const int bh = bin_of(cfg.edges_h, cfg.bin_h, static_cast<double>(v));
      if (bw >= 0 && bh >= 0) b = bh * cfg.bin_w + bw;
    }
    s_bin[i] = keep ? (b >= 0 ? b : -2) : -1;
    s_u[i] = static_cast<int>(u);
    s_v[i] = static_cast<int>(v);
    s_depth[i] = zt;
    s_vel[i] = vel;
  }
  for (int i = tid; i < 3 * kMaxMap * kMaxMap; i += kRadarThreads) (&s_map[0][0][0])[i] = 0.f;
  __syncthreads();

  // filtered point list (u, v, depth, velocity), in point order: prepare_data.py:109-110
  if (tid == 0 && uvzv_out != nullptr) {
    int k = 0;
    for (int i = 0; i < count; ++i) {
      if (s_bin[i] == -1) continue;
      float* o = uvzv_out + (1LL * frame * cap + k) * 4;
      o[0] = static_cast<float>(s_u[i]);
      o[1] = static_cast<float>(s_v[i]);
      o[2] = static_cast<float>(s_depth[i]);
      o[3] = static_cast<float>(s_vel[i]);
      ++k;
    }
    if (kept_out) kept_out[frame] = k;
  }

  // histograms: one thread per bin, points visited in order (np.bincount summation order)
  const int pad_h = cfg.bin_h <= cfg.bin_w ? (cfg.bin_w - cfg.bin_h) / 2 : 0;  // pad_to_square: (upper/left) = diff // 2
  const int pad_w = cfg.bin_h <= cfg.bin_w ? 0 : (cfg.bin_h - cfg.bin_w) / 2;
  for (int b = tid; b < cfg.bin_h * cfg.bin_w; b += kRadarThreads) {
    double h0 = 0.0, sd = 0.0, sv = 0.0;
    for (int i = 0; i < count; ++i) {
      if (s_bin[i] == b) {
        h0 = dadd(h0, 1.0);
        sd = dadd(sd, s_depth[i]);
        sv = dadd(sv, s_vel[i]);
      }
    }
    const double den = dadd(h0, 1e-6);
    double h1 = sd / den;
    if (h1 < 1.0) h1 = 100.0;
    const double h2 = fabs(sv / den);
    const double c0 = fmin(fmax(dadd(h0, -0.0) / 5.0, 0.0), 1.0);
    const double c1 = fmin(fmax(dadd(h1, -12.0) / -12.0, 0.0), 1.0);
    const double c2 = fmin(fmax(dadd(h2, -0.0) / 4.0, 0.0), 1.0);
    const int bh = b / cfg.bin_w, bw = b - bh * cfg.bin_w;
    s_map[0][bh + pad_h][bw + pad_w] = static_cast<float>(c0);
    s_map[1][bh + pad_h][bw + pad_w] = static_cast<float>(c1);
    s_map[2][bh + pad_h][bw + pad_w] = static_cast<float>(c2);
  }
  __syncthreads();

  // bilinear resize of the padded square map, align_corners=True
  const int in = max(cfg.bin_h, cfg.bin_w), out = cfg.out_size;
  const float scale = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  float* dst = maps_out + 1LL * frame * 3 * out * out;
  for (int i = tid; i < 3 * out * out; i += kRadarThreads) {
    const int c = i / (out * out), oy = (i / out) % out, ox = i % out;
    if (in == out) {
      dst[i] = s_map[c][oy][ox];
      continue;
    }
    const float sy = scale * oy, sx = scale * ox;
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = y0 + (y0 < in - 1 ? 1 : 0), x1 = x0 + (x0 < in - 1 ? 1 : 0);
    const float ly = fminf(fmaxf(sy - y0, 0.f), 1.f), lx = fminf(fmaxf(sx - x0, 0.f), 1.f);
    const float hy = 1.f - ly, hx = 1.f - lx;
    dst[i] = hy * (hx * s_map[c][y0][x0] + lx * s_map[c][y0][x1]) + ly * (hx * s_map[c][y1][x0] + lx * s_map[c][y1][x1]);
  }
}

}  // namespace
}  // namespace me

extern "C" {

int me_radar_maps(const float* points, const int* counts, int n, int cap, const me_radar_cfg* cfg, float* maps_out,
                  float* uvzv_out, int* kept_out, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(points && counts && cfg && maps_out, "radar_maps: null argument");
  ME_REQUIRE(n > 0 && cap > 0 && cap <= kMaxPoints, "radar_maps: 1..%d points per frame", kMaxPoints);
  ME_REQUIRE(cfg->bin_w >= 1 && cfg->bin_h >= 1 && cfg->bin_w <= kMaxMap && cfg->bin_h <= kMaxMap,
             "radar_maps: histogram of %d x %d bins unsupported (max %d)", cfg->bin_w, cfg->bin_h, kMaxMap);
  ME_REQUIRE(cfg->out_size >= 1 && cfg->out_size <= 64, "radar_maps: out_size %d out of range", cfg->out_size);
  radar_maps_kernel<<<n, kRadarThreads, 0, stream>>>(points, counts, cap, *cfg, maps_out, uvzv_out, kept_out);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
