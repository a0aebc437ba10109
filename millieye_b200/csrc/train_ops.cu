// fp32 kernels of the stage-3 TRAINING step (reference module3_our_dataset/train.py:169-191: model.train(),
// base_detector.eval(), forward with targets, loss.backward(), Adam) for the heads that train:
// img_cnn_layers (1x1 conv + BatchNorm + LeakyReLU, my_models.py:47-77), radar_cnn_layers (:130-157),
// refinement_head (:213-284) and ensemble_head (:176-210), with the loss of :610-635.
//
// The detector is frozen and runs on the tensor-core engine; everything that receives a gradient is small
// (<= 0.7 GMAC per op at 8 frames per GPU) and has to match the reference's fp32 autograd to ~1e-5, so this file
// is plain fp32 SIMT: a tiled GEMM with generic strides (forward, dX and dW products of every linear / conv layer;
// 3x3 convs go through an explicit im2col), batch-statistics BatchNorm forward / backward with double accumulation
// (torch CPU accumulates float statistics in double), RoIAlign / PS-RoIAlign on fp32 maps with their adjoints
// (atomic scatter), the per-proposal tail of the two heads with its hand-derived backward, and Adam.
// The op-by-op derivation these kernels transcribe is oracle/stage3_backward.py (checked against torch.autograd and
// the reference's own gradients).  All matrices are row-major [rows][channels] ("NHWC flattened").
#include "common.cuh"
#include <cooperative_groups.h>

namespace me {
namespace {

constexpr float kSlope = 0.1f;

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

// ------------------------------------------------------------------------------------------------ GEMM
// C[i][j] (ldc) = sum_k A(i,k) * B(k,j) (+ bias[j]) (then act), A(i,k) = A[i*sai + k*sak], B(k,j) = B[k*sbk + j*sbj].
// accumulate != 0: C += result (no bias / act).  BM x BN x 16 tiles, 256 threads, (BM/16) x (BN/16) outputs per thread.
template <int BM, int BN>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, long long sai, long long sak,
                const float* __restrict__ B, long long sbk, long long sbj, float* __restrict__ C, long long ldc,
                const float* __restrict__ bias, int act, int accumulate, int k_chunk) {
  // blockIdx.z selects a K range of k_chunk (split-K for the weight-gradient products: a few output tiles, tens of
  // thousands of rows to reduce over); the slices add their partial tiles with atomicAdd into a zero-filled C
  constexpr int BK = 16, TM = BM / 16, TN = BN / 16;
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
  // the faster-varying index of a tile load follows the operand's unit stride, so either layout loads coalesced
  const bool a_k_fast = sak == 1, b_j_fast = sbj == 1;
  const bool split = gridDim.z > 1;
  const int k_begin = blockIdx.z * k_chunk;
  if (split) K = min(K, k_begin + k_chunk);
  for (int k0 = k_begin; k0 < K; k0 += BK) {
    for (int e = threadIdx.x; e < BM * BK; e += 256) {
      const int i = a_k_fast ? e / BK : e % BM, k = a_k_fast ? e % BK : e / BM;
      const int gi = i0 + i, gk = k0 + k;
      As[k][i] = (gi < M && gk < K) ? A[gi * sai + gk * sak] : 0.f;
    }
    for (int e = threadIdx.x; e < BN * BK; e += 256) {
      const int j = b_j_fast ? e % BN : e / BK, k = b_j_fast ? e / BN : e % BK;
      const int gj = j0 + j, gk = k0 + k;
      Bs[k][j] = (gj < N && gk < K) ? B[gk * sbk + gj * sbj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int a = 0; a < TM; ++a) av[a] = As[k][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < TN; ++b) bv[b] = Bs[k][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < TM; ++a) {
    const int gi = i0 + ty + 16 * a;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      const int gj = j0 + tx + 16 * b;
      if (gj >= N) continue;
      float v = acc[a][b];
      float* c = C + gi * ldc + gj;
      if (split) {
        atomicAdd(c, v);
      } else if (accumulate) {
        *c += v;
      } else {
        if (bias) v += bias[gj];
        if (act == ME_ACT_LEAKY) v = v > 0.f ? v : kSlope * v;
        else if (act == ME_ACT_SIGMOID) v = sigmoid_f(v);
        *c = v;
      }
    }
  }
}

// The same product for the large shapes (pixels x channels forward / input-gradient products, split-K weight gradients):
// 128 x 128 x 8 tiles, 256 threads, an 8 x 8 micro-tile per thread read from shared memory as float4 (16 FMAs per shared
// load instead of 2), the next K slab prefetched into registers while the current one is multiplied.
__global__ void __launch_bounds__(256, 2)
gemm_f32_big_kernel(int M, int N, int K, const float* __restrict__ A, long long sai, long long sak,
                    const float* __restrict__ B, long long sbk, long long sbj, float* __restrict__ C, long long ldc,
                    const float* __restrict__ bias, int act, int accumulate, int k_chunk) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const bool a_k_fast = sak == 1, b_j_fast = sbj == 1;
  const bool split = gridDim.z > 1;
  const int k_begin = blockIdx.z * k_chunk;
  if (split) K = min(K, k_begin + k_chunk);
  // element e (0..1023) of a slab handled by this thread: 4 per operand; the faster index follows the unit stride
  int ai[4], ak[4], bj[4], bk[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = threadIdx.x + 256 * u;
    ai[u] = a_k_fast ? e / BK : e % BM;
    ak[u] = a_k_fast ? e % BK : e / BM;
    bj[u] = b_j_fast ? e % BN : e / BK;
    bk[u] = b_j_fast ? e / BN : e % BK;
  }
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gi = i0 + ai[u], gk = k0 + ak[u];
      ra[u] = (gi < M && gk < K) ? A[gi * sai + gk * sak] : 0.f;
      const int gj = j0 + bj[u], gk2 = k0 + bk[u];
      rb[u] = (gj < N && gk2 < K) ? B[gk2 * sbk + gj * sbj] : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      As[buf][ak[u]][ai[u]] = ra[u];
      Bs[buf][bk[u]][bj[u]] = rb[u];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  int buf = 0;
  if (k_begin < K) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int gi = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + a - 4);
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int gj = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + b - 4);
      if (gj >= N) continue;
      float v = acc[a][b];
      float* c = C + gi * ldc + gj;
      if (split) {
        atomicAdd(c, v);
      } else if (accumulate) {
        *c += v;
      } else {
        if (bias) v += bias[gj];
        if (act == ME_ACT_LEAKY) v = v > 0.f ? v : kSlope * v;
        else if (act == ME_ACT_SIGMOID) v = sigmoid_f(v);
        *c = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[c] = sum_r X[r][c] * (Y ? Y[r][c] : 1), double accumulation.  One block of 32 x 32 threads per 32 columns: a warp
// reads 32 consecutive columns of one row (128 bytes), the 32 warps take rows r, r+32, ... four at a time (memory-level
// parallelism: with 8 row lanes and one load in flight the 43 264-row sums of a batch-64 step took ~0.3 ms each).
// Fixed reduction order: deterministic.
constexpr int kColsumLanes = 32;
constexpr int kRowSplit = 8;     // CTAs of one thread-block cluster share a 32-column group, each takes an eighth of the rows
// Column sums over [rows][cols] matrices have cols / 32 column groups - 1 to 16 blocks for the head layers, i.e. 1 to 16 of
// 148 SMs reading tens of MB.  A cluster of 8 CTAs per column group splits the rows; the partial sums meet in CTA 0 through
// distributed shared memory and are added in rank order: still a fixed reduction order (deterministic), no workspace.
__global__ void __cluster_dims__(kRowSplit, 1, 1) __launch_bounds__(32 * kColsumLanes)
colsum_kernel(const float* __restrict__ X, const float* __restrict__ Y, long long rows, int cols, long long ldx,
              long long ldy, float* __restrict__ out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s[kColsumLanes][33];
  __shared__ double s_tot[32];
  const int rank = static_cast<int>(cluster.block_rank());
  const int c = (blockIdx.x / kRowSplit) * 32 + (threadIdx.x & 31);
  const int lane_r = threadIdx.x >> 5;
  const long long per = (rows + kRowSplit - 1) / kRowSplit;
  const long long r_end = min(rows, (rank + 1) * per);
  double acc = 0.0;
  if (c < cols) {
    long long r = rank * per + lane_r;
    for (; r + 3 * kColsumLanes < r_end; r += 4 * kColsumLanes) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = X[(r + u * kColsumLanes) * ldx + c];
      if (Y) {
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = Y[(r + u * kColsumLanes) * ldy + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += Y ? static_cast<double>(x[u]) * static_cast<double>(y[u]) : static_cast<double>(x[u]);
    }
    for (; r < r_end; r += kColsumLanes) {
      const float x = X[r * ldx + c];
      acc += Y ? static_cast<double>(x) * static_cast<double>(Y[r * ldy + c]) : static_cast<double>(x);
    }
  }
  s[lane_r][threadIdx.x & 31] = acc;
  __syncthreads();
  if (lane_r == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kColsumLanes; ++k) t += s[k][threadIdx.x & 31];
    s_tot[threadIdx.x & 31] = t;
  }
  cluster.sync();
  if (rank == 0 && lane_r == 0 && c < cols) {
    double t = 0.0;
    for (int k = 0; k < kRowSplit; ++k) t += cluster.map_shared_rank(s_tot, k)[threadIdx.x & 31];
    out[c] = static_cast<float>(t);
  }
  cluster.sync();   // the peers' shared memory stays alive until CTA 0 has read it
}

// ------------------------------------------------------------------------------------------------ layout / im2col
__global__ void half_rows_to_float_kernel(const __half* __restrict__ x, long long rows, int cols, int pitch,
                                          float* __restrict__ y) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    y[i] = __half2float(x[r * pitch + c]);
  }
}

__global__ void nchw_to_rows_kernel(const float* __restrict__ x, int n, int c, int hw, float* __restrict__ y) {
  const long long total = 1LL * n * c * hw;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const long long p = i / c;
    const int img = static_cast<int>(p / hw), pix = static_cast<int>(p % hw);
    y[i] = x[(1LL * img * c + ch) * hw + pix];
  }
}

// cols[p][ci*9 + tap] = x[n, y+dy, x+dx, ci] (zero outside): the k index matches an OIHW weight row flattened.
// One thread per (pixel, channel): the index is split once and the nine taps are written as one 36-byte run (a thread per
// output element spent ~7 64-bit divisions on every float it copied).
__global__ void im2col3_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ cols) {
  const long long total = 1LL * n * h * w * c;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int p = static_cast<int>(i / c);            // pixel index < 2^31 (checked by the launcher)
    const int ci = static_cast<int>(i - 1LL * p * c);
    const int px = p % w, t = p / w;
    const int py = t % h, img = t / h;
    float v[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      v[tap] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? x[((1LL * img * h + yy) * w + xx) * c + ci] : 0.f;
    }
    float* dst = cols + i * 9;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) dst[tap] = v[tap];
  }
}

// dx[n, y, x, ci] = sum_tap dcols[(n, y-dy, x-dx)][ci*9 + tap]  (gather form of col2im: deterministic)
__global__ void col2im3_kernel(const float* __restrict__ dcols, int n, int h, int w, int c, float* __restrict__ dx) {
  const long long total = 1LL * n * h * w * c;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int ci = static_cast<int>(i % c);
    const long long p = i / c;
    const int px = static_cast<int>(p % w), py = static_cast<int>((p / w) % h);
    const long long img = p / (1LL * w * h);
    float s = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py - (tap / 3 - 1), xx = px - (tap % 3 - 1);
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) s += dcols[(((img * h + yy) * w + xx) * c + ci) * 9 + tap];
    }
    dx[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm (batch statistics)
// Batch statistics in two steps so that the ranks of a sharded batch can add their partial sums in between
// (synchronised BatchNorm: the step then equals the reference's single-process step on the whole batch):
//   partial : sums[c] = sum_r z[r][c], sums[cols + c] = sum_r z[r][c]^2 (double), sums[2 cols] = rows
//   finalize: mean, biased variance -> inv_std; running statistics updated like nn.BatchNorm2d in training mode
//             (momentum m, unbiased variance into running_var).  The row count is read from device memory.
__global__ void __cluster_dims__(kRowSplit, 1, 1) __launch_bounds__(32 * kColsumLanes)
bn_partial_kernel(const float* __restrict__ z, long long rows, int cols, double* __restrict__ sums) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s1[kColsumLanes][33], s2[kColsumLanes][33];
  __shared__ double t1s[32], t2s[32];
  const int rank = static_cast<int>(cluster.block_rank());
  const int c = (blockIdx.x / kRowSplit) * 32 + (threadIdx.x & 31);
  const int lr = threadIdx.x >> 5;
  const long long per = (rows + kRowSplit - 1) / kRowSplit;
  const long long r_end = min(rows, (rank + 1) * per);
  double a = 0.0, b = 0.0;
  if (c < cols) {
    long long r = rank * per + lr;
    for (; r + 3 * kColsumLanes < r_end; r += 4 * kColsumLanes) {
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = z[(r + u * kColsumLanes) * cols + c];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double v = static_cast<double>(x[u]);
        a += v;
        b += v * v;
      }
    }
    for (; r < r_end; r += kColsumLanes) {
      const double v = static_cast<double>(z[r * cols + c]);
      a += v;
      b += v * v;
    }
  }
  s1[lr][threadIdx.x & 31] = a;
  s2[lr][threadIdx.x & 31] = b;
  __syncthreads();
  if (lr == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < kColsumLanes; ++k) {
      t1 += s1[k][threadIdx.x & 31];
      t2 += s2[k][threadIdx.x & 31];
    }
    t1s[threadIdx.x & 31] = t1;
    t2s[threadIdx.x & 31] = t2;
  }
  cluster.sync();
  if (rank == 0 && lr == 0 && c < cols) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < kRowSplit; ++k) {     // rank order: deterministic
      t1 += cluster.map_shared_rank(t1s, k)[threadIdx.x & 31];
      t2 += cluster.map_shared_rank(t2s, k)[threadIdx.x & 31];
    }
    sums[c] = t1;
    sums[cols + c] = t2;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * cols] = static_cast<double>(rows);
  cluster.sync();
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int cols, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_out, float* __restrict__ inv_std) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const double rows = sums[2 * cols];
  if (rows < 1.0) {   // no rows anywhere (no proposals in the whole batch): nothing to normalise, statistics untouched
    mean_out[c] = 0.f;
    inv_std[c] = rsqrtf(eps);
    return;
  }
  const double mean = sums[c] / rows;
  double biased = sums[cols + c] / rows - mean * mean;
  if (biased < 0.0) biased = 0.0;
  const double unbiased = rows > 1.0 ? biased * rows / (rows - 1.0) : biased;
  mean_out[c] = static_cast<float>(mean);
  inv_std[c] = static_cast<float>(1.0 / sqrt(biased + static_cast<double>(eps)));
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// x_hat = (z - mean) * inv_std;  a = leaky(gamma * x_hat + beta)
__global__ void bn_apply_kernel(const float* __restrict__ z, long long rows, int cols, const float* __restrict__ mean,
                                const float* __restrict__ inv_std, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ xhat, float* __restrict__ a) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % cols);
    const float xh = (z[i] - mean[c]) * inv_std[c];
    const float y = gamma[c] * xh + beta[c];
    xhat[i] = xh;
    a[i] = y > 0.f ? y : kSlope * y;
  }
}

// dy = da * leaky'(a) written over da; then dgamma / dbeta are column sums (colsum_kernel), then
// dz = inv_std * gamma * (dy - dbeta / rows - x_hat * dgamma / rows)
__global__ void leaky_bwd_kernel(float* __restrict__ da, const float* __restrict__ a, long long total) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x)
    da[i] = a[i] > 0.f ? da[i] : kSlope * da[i];
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ xhat, long long rows, int cols,
                                    const float* __restrict__ gamma, const float* __restrict__ inv_std,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                    const double* __restrict__ total_rows, float* __restrict__ dz) {
  const long long total = rows * cols;
  // dgamma / dbeta are sums over ALL rows of the batch (all ranks); total_rows is that row count (device memory)
  const float inv_rows = static_cast<float>(1.0 / fmax(*total_rows, 1.0));
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % cols);
    const float g = gamma[c];
    dz[i] = inv_std[c] * g * (dy[i] - dbeta[c] * inv_rows - xhat[i] * dgamma[c] * inv_rows);
  }
}
// ds * s * (1 - s) in place
__global__ void sigmoid_bwd_kernel(float* __restrict__ ds, const float* __restrict__ s, long long total) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x)
    ds[i] = ds[i] * s[i] * (1.f - s[i]);
}

// ------------------------------------------------------------------------------------------------ RoIAlign on fp32 maps
struct RoiGeom {
  int b, gh, gw;
  float sh, sw, bin_h, bin_w, cnt;
};
template <bool PS>
__device__ __forceinline__ RoiGeom roi_geom(const float* roi, int n, int pooled, float scale) {
  RoiGeom g;
  int b = static_cast<int>(roi[0]);
  g.b = b < 0 ? 0 : (b >= n ? n - 1 : b);
  const float off = PS ? 0.5f : 0.f;
  g.sw = roi[1] * scale - off;
  g.sh = roi[2] * scale - off;
  const float ew = roi[3] * scale - off, eh = roi[4] * scale - off;
  float rw = ew - g.sw, rh = eh - g.sh;
  if (!PS) {
    rw = fmaxf(rw, 1.f);
    rh = fmaxf(rh, 1.f);
  }
  g.bin_h = rh / pooled;
  g.bin_w = rw / pooled;
  g.gh = static_cast<int>(ceilf(rh / pooled));
  g.gw = static_cast<int>(ceilf(rw / pooled));
  g.cnt = PS ? static_cast<float>(g.gh * g.gw) : static_cast<float>(max(g.gh * g.gw, 1));
  return g;
}
struct Bilin {
  int yl, yh, xl, xh;
  float hy, ly, hx, lx;
  bool valid;
};
__device__ __forceinline__ Bilin bilin_setup(int h, int w, float y, float x) {
  Bilin q;
  q.valid = !(y < -1.0f || y > static_cast<float>(h) || x < -1.0f || x > static_cast<float>(w));
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  q.yl = static_cast<int>(y);
  q.xl = static_cast<int>(x);
  if (q.yl >= h - 1) {
    q.yh = q.yl = h - 1;
    y = static_cast<float>(q.yl);
  } else {
    q.yh = q.yl + 1;
  }
  if (q.xl >= w - 1) {
    q.xh = q.xl = w - 1;
    x = static_cast<float>(q.xl);
  } else {
    q.xh = q.xl + 1;
  }
  q.ly = y - q.yl;
  q.lx = x - q.xl;
  q.hy = 1.f - q.ly;
  q.hx = 1.f - q.lx;
  return q;
}

// One thread per output element (r, c, ph, pw); feat [n][h][w][chan_total] fp32; out [rows][channels*P*P].
// BWD: grad_out has the same layout; the sample weights are scattered into dfeat with atomicAdd.
template <bool PS, bool BWD>
__global__ void roi_f32_kernel(const float* __restrict__ feat, float* __restrict__ dfeat, int n, int h, int w, int chan_total,
                               int channels, int pooled, float scale, const float* __restrict__ rois, int num_rois,
                               float* __restrict__ out, const float* __restrict__ grad_out) {
  const int per_roi = channels * pooled * pooled;
  const long long total = 1LL * num_rois * per_roi;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int r = static_cast<int>(i / per_roi);
    const int e = static_cast<int>(i - 1LL * r * per_roi);
    const int c = e / (pooled * pooled), ph = (e / pooled) % pooled, pw = e % pooled;
    const RoiGeom g = roi_geom<PS>(rois + r * 5, n, pooled, scale);
    const int ch = PS ? (c * pooled + ph) * pooled + pw : c;
    const long long plane = 1LL * g.b * h * w * chan_total + ch;
    const float hstart = ph * g.bin_h + g.sh, wstart = pw * g.bin_w + g.sw;
    float sum = 0.f;
    const float go = BWD ? grad_out[i] / g.cnt : 0.f;
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = hstart + (iy + 0.5f) * g.bin_h / static_cast<float>(g.gh);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = wstart + (ix + 0.5f) * g.bin_w / static_cast<float>(g.gw);
        const Bilin q = bilin_setup(h, w, y, x);
        if (!q.valid) continue;
        const long long o1 = plane + (1LL * q.yl * w + q.xl) * chan_total, o2 = plane + (1LL * q.yl * w + q.xh) * chan_total;
        const long long o3 = plane + (1LL * q.yh * w + q.xl) * chan_total, o4 = plane + (1LL * q.yh * w + q.xh) * chan_total;
        if (BWD) {
          atomicAdd(dfeat + o1, go * q.hy * q.hx);
          atomicAdd(dfeat + o2, go * q.hy * q.lx);
          atomicAdd(dfeat + o3, go * q.ly * q.hx);
          atomicAdd(dfeat + o4, go * q.ly * q.lx);
        } else {
          sum += q.hy * q.hx * feat[o1] + q.hy * q.lx * feat[o2] + q.ly * q.hx * feat[o3] + q.ly * q.lx * feat[o4];
        }
      }
    }
    if (!BWD) out[i] = sum / g.cnt;
  }
}

// ------------------------------------------------------------------------------------------------ per-proposal tail
// Forward (my_models.py:268-284 tail, :202-210, :513-514): one thread per proposal.
//   rc = sigmoid(r2); conf = sigmoid(rc + cls[:,0]); refine = [conf, cls[:,1]]
//   image proposals: u = [[conf, yolo_conf], [cls1, yolo_cls]] -> fc1 (2->32, leaky) per row -> flatten 64 -> fc2 -> softmax p
//   mask = p[0] (image) | conf (radar)
__global__ void stage3_tail_fwd_kernel(const float* __restrict__ r2, const float* __restrict__ cls, int cls_pitch,
                                       const float* __restrict__ img_boxes, int n_img, int n_all,
                                       const float* __restrict__ fc1_w, const float* __restrict__ fc1_b,
                                       const float* __restrict__ fc2_w, const float* __restrict__ fc2_b,
                                       float* __restrict__ rc_out, float* __restrict__ refine, float* __restrict__ mask,
                                       float* __restrict__ p_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_all) return;
  const float rc = sigmoid_f(r2[r]);
  const float conf = sigmoid_f(rc + cls[1LL * r * cls_pitch]);
  const float c1 = cls[1LL * r * cls_pitch + 1];
  rc_out[r] = rc;
  refine[2 * r] = conf;
  refine[2 * r + 1] = c1;
  if (r >= n_img) {
    mask[r] = conf;
    return;
  }
  const float u[2][2] = {{conf, img_boxes[r * 9 + 5]}, {c1, img_boxes[r * 9 + 8]}};
  float o0 = fc2_b[0], o1 = fc2_b[1];
  for (int c = 0; c < 2; ++c)
    for (int j = 0; j < 32; ++j) {
      float hp = fc1_w[j * 2] * u[c][0] + fc1_w[j * 2 + 1] * u[c][1] + fc1_b[j];
      hp = hp > 0.f ? hp : kSlope * hp;
      o0 = fmaf(fc2_w[c * 32 + j], hp, o0);
      o1 = fmaf(fc2_w[64 + c * 32 + j], hp, o1);
    }
  const float mx = fmaxf(o0, o1);
  const float e0 = expf(o0 - mx), e1 = expf(o1 - mx);
  const float p0 = e0 / (e0 + e1);
  p_out[2 * r] = p0;
  p_out[2 * r + 1] = e1 / (e0 + e1);
  mask[r] = p0;
}

// Backward of the loss  focal(masks of the sampled image proposals) + BCE(conf of the sample) / lambda  down to the
// pre-activations (oracle/stage3_backward.py backward(), the part above the GEMMs).  One thread per proposal; writes
//   d_o [n_img][2], hl [n_img][64], dhp [2*n_img][32], u [2*n_img][2]  (operands of the ensemble head's dW GEMMs)
//   dr2 [n_all]  (gradient at radar_net's last pre-activation),  dz2 [n_all][13] (at net2's pre-activation)
__global__ void stage3_tail_bwd_kernel(const float* __restrict__ rc, const float* __restrict__ refine,
                                       const float* __restrict__ cls, int cls_pitch, const float* __restrict__ p,
                                       const float* __restrict__ img_boxes, int n_img, int n_all,
                                       const unsigned char* __restrict__ pos, const unsigned char* __restrict__ sel,
                                       float alpha, float lambda_conf, const float* __restrict__ fc1_w,
                                       const float* __restrict__ fc1_b, const float* __restrict__ fc2_w,
                                       float* __restrict__ d_o, float* __restrict__ hl_out, float* __restrict__ dhp_out,
                                       float* __restrict__ u_out, float* __restrict__ dr2, float* __restrict__ dz2) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_all) return;
  const float conf = refine[2 * r], c1 = refine[2 * r + 1];
  const bool is_pos = pos[r] != 0, in_sample = sel[r] != 0;
  const float y = is_pos ? 1.f : 0.f;
  float dconf = in_sample ? -(y / conf - (1.f - y) / (1.f - conf)) / lambda_conf : 0.f;
  float dcls1 = 0.f;
  if (r < n_img) {
    const float m = p[2 * r];
    const float prob = is_pos ? m : 1.f - m;
    const float a_f = is_pos ? alpha : 1.f - alpha;
    const float dprob = -a_f * (-2.f * (1.f - prob) * logf(prob) + (1.f - prob) * (1.f - prob) / prob);
    const float dm = in_sample ? (is_pos ? dprob : -dprob) : 0.f;
    // softmax: do = p * (dp - sum(dp * p)), dp = [dm, 0]
    const float p0 = p[2 * r], p1 = p[2 * r + 1];
    const float dot = dm * p0;
    const float do0 = p0 * (dm - dot), do1 = p1 * (0.f - dot);
    d_o[2 * r] = do0;
    d_o[2 * r + 1] = do1;
    const float u[2][2] = {{conf, img_boxes[r * 9 + 5]}, {c1, img_boxes[r * 9 + 8]}};
    float du[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int c = 0; c < 2; ++c) {
      u_out[(2 * r + c) * 2] = u[c][0];
      u_out[(2 * r + c) * 2 + 1] = u[c][1];
      for (int j = 0; j < 32; ++j) {
        const float hp = fc1_w[j * 2] * u[c][0] + fc1_w[j * 2 + 1] * u[c][1] + fc1_b[j];
        hl_out[r * 64 + c * 32 + j] = hp > 0.f ? hp : kSlope * hp;
        const float dhl = do0 * fc2_w[c * 32 + j] + do1 * fc2_w[64 + c * 32 + j];
        const float dhp = hp > 0.f ? dhl : kSlope * dhl;
        dhp_out[(2 * r + c) * 32 + j] = dhp;
        du[c][0] = fmaf(dhp, fc1_w[j * 2], du[c][0]);
        du[c][1] = fmaf(dhp, fc1_w[j * 2 + 1], du[c][1]);
      }
    }
    dconf += du[0][0];   // u[0][0] is the refined confidence
    dcls1 = du[1][0];    // u[1][0] is cls[:, 1]
  }
  const float drc = dconf * conf * (1.f - conf);   // also the gradient at cls[:, 0] (same pre-activation sum)
  const float rcv = rc[r];
  dr2[r] = drc * rcv * (1.f - rcv);
  const float* cr = cls + 1LL * r * cls_pitch;
  float* dz = dz2 + 1LL * r * 13;
#pragma unroll
  for (int k = 0; k < 13; ++k) dz[k] = 0.f;
  dz[0] = drc * cr[0] * (1.f - cr[0]);
  dz[1] = dcls1 * cr[1] * (1.f - cr[1]);
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam (train.py:158: lr 5e-4, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) over a flat buffer.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float beta1, float beta2, float eps, float bc1, float bc2) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

inline int grid_for(long long total, int block = 256) {
  long long b = (total + block - 1) / block;
  if (b > 148LL * 32) b = 148LL * 32;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace
}  // namespace me

extern "C" {

int me_gemm_f32(int M, int N, int K, const float* A, long long sai, long long sak, const float* B, long long sbk,
                long long sbj, float* C, long long ldc, const float* bias, int act, int accumulate, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0) return ME_OK;
  ME_REQUIRE(A && B && C && K >= 0, "gemm_f32: null argument");
  ME_REQUIRE(!accumulate || (!bias && act == ME_ACT_LINEAR), "gemm_f32: accumulate excludes bias / activation");
  const long long tiles64 = 1LL * ((M + 63) / 64) * ((N + 63) / 64);
  const long long tiles128 = 1LL * ((M + 127) / 128) * ((N + 127) / 128);
  // Few output tiles and a long reduction (the weight gradients dW = dZ^T x over all pixels / proposals of the batch):
  // split K over blockIdx.z so that ~4 waves of CTAs work, partial tiles added with atomicAdd into a zero-filled C.
  int splits = 1, k_chunk = K;
  if (tiles64 < 148 && K >= 1024 && !bias && act == ME_ACT_LINEAR) {
    // chunks of >= 128 (K < 8192: the proposals' products, a few thousand rows) or >= 512 reduction steps
    const int min_chunk = K >= 8192 ? 512 : 128;
    splits = static_cast<int>((4 * 148 + tiles64 - 1) / tiles64);
    if (splits > K / min_chunk) splits = K / min_chunk;
    if (splits < 1) splits = 1;
    k_chunk = (((K + splits - 1) / splits) + 15) / 16 * 16;
    splits = (K + k_chunk - 1) / k_chunk;
  }
  if (splits > 1) {
    if (!accumulate) ME_CUDA(cudaMemset2DAsync(C, static_cast<size_t>(ldc) * sizeof(float), 0, static_cast<size_t>(N) * sizeof(float), M, stream));
    if (M >= 96 && N >= 96) {   // 128 x 128 tiles, K split so that ~4 waves of them work
      int sp = static_cast<int>((4 * 148 + tiles128 - 1) / tiles128);
      if (sp > K / 256) sp = K / 256;
      if (sp < 1) sp = 1;
      const int kc = (((K + sp - 1) / sp) + 7) / 8 * 8;
      sp = (K + kc - 1) / kc;
      dim3 grid((N + 127) / 128, (M + 127) / 128, sp);
      gemm_f32_big_kernel<<<grid, 256, 0, stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc, nullptr, ME_ACT_LINEAR, 0, kc);
    } else {
      dim3 grid((N + 63) / 64, (M + 63) / 64, splits);
      gemm_f32_kernel<64, 64><<<grid, 256, 0, stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc, nullptr, ME_ACT_LINEAR, 0, k_chunk);
    }
  } else if (tiles128 >= 120 && N >= 64) {
    dim3 grid((N + 127) / 128, (M + 127) / 128);
    gemm_f32_big_kernel<<<grid, 256, 0, stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc, bias, act, accumulate, K);
  } else if (tiles64 >= 148) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    gemm_f32_kernel<64, 64><<<grid, 256, 0, stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc, bias, act, accumulate, K);
  } else {
    // few output tiles, short K: 32 x 32 tiles put more CTAs on the machine
    dim3 grid((N + 31) / 32, (M + 31) / 32);
    gemm_f32_kernel<32, 32><<<grid, 256, 0, stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc, bias, act, accumulate, K);
  }
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_colsum_f32(const float* X, const float* Y, long long rows, int cols, long long ldx, long long ldy, float* out,
                  me_stream_t stream) {
  using namespace me;
  if (cols <= 0) return ME_OK;
  ME_REQUIRE(X && out, "colsum_f32: null argument");
  colsum_kernel<<<kRowSplit * ((cols + 31) / 32), 32 * kColsumLanes, 0, static_cast<cudaStream_t>(stream)>>>(X, Y, rows, cols, ldx, ldy, out);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_half_rows_to_float(const void* x, long long rows, int cols, int pitch, float* y, me_stream_t stream) {
  using namespace me;
  if (rows <= 0) return ME_OK;
  ME_REQUIRE(x && y && pitch >= cols, "half_rows_to_float: bad argument");
  half_rows_to_float_kernel<<<grid_for(rows * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), rows, cols, pitch, y);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_nchw_to_rows_f32(const float* x, int n, int c, int hw, float* y, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(x && y && n > 0 && c > 0 && hw > 0, "nchw_to_rows: bad argument");
  nchw_to_rows_kernel<<<grid_for(1LL * n * c * hw), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, c, hw, y);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_im2col3_f32(const float* x, int n, int h, int w, int c, float* cols, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(x && cols && n > 0 && h > 0 && w > 0 && c > 0, "im2col3: bad argument");
  ME_REQUIRE(1LL * n * h * w < (1LL << 31), "im2col3: pixel count out of range");
  im2col3_kernel<<<grid_for(1LL * n * h * w * c), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, h, w, c, cols);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_col2im3_f32(const float* dcols, int n, int h, int w, int c, float* dx, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(dcols && dx && n > 0 && h > 0 && w > 0 && c > 0, "col2im3: bad argument");
  col2im3_kernel<<<grid_for(1LL * n * h * w * c), 256, 0, static_cast<cudaStream_t>(stream)>>>(dcols, n, h, w, c, dx);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_bn_partial_stats(const float* z, long long rows, int cols, double* sums, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(sums && cols > 0 && rows >= 0 && (rows == 0 || z), "bn_partial_stats: bad argument");
  bn_partial_kernel<<<kRowSplit * ((cols + 31) / 32), 32 * kColsumLanes, 0, static_cast<cudaStream_t>(stream)>>>(z, rows, cols, sums);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_bn_finalize(const double* sums, int cols, float eps, float momentum, float* running_mean, float* running_var,
                   float* mean_out, float* inv_std, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(sums && mean_out && inv_std && cols > 0, "bn_finalize: bad argument");
  bn_finalize_kernel<<<(cols + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, cols, eps, momentum, running_mean,
                                                                                        running_var, mean_out, inv_std);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_bn_apply(const float* z, long long rows, int cols, const float* mean, const float* inv_std, const float* gamma,
                const float* beta, float* xhat, float* a, me_stream_t stream) {
  using namespace me;
  if (rows <= 0) return ME_OK;
  ME_REQUIRE(z && mean && inv_std && gamma && beta && xhat && a && cols > 0, "bn_apply: bad argument");
  bn_apply_kernel<<<grid_for(rows * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, rows, cols, mean, inv_std, gamma, beta,
                                                                                        xhat, a);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_bn_train_fwd(const float* z, long long rows, int cols, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, double* sums_ws, float* mean_ws, float* inv_std,
                    float* xhat, float* a, me_stream_t stream) {
  int rc = me_bn_partial_stats(z, rows, cols, sums_ws, stream);
  if (rc != ME_OK) return rc;
  rc = me_bn_finalize(sums_ws, cols, eps, momentum, running_mean, running_var, mean_ws, inv_std, stream);
  if (rc != ME_OK) return rc;
  return me_bn_apply(z, rows, cols, mean_ws, inv_std, gamma, beta, xhat, a, stream);
}

int me_bn_bwd_sums(float* da_inout, const float* a, const float* xhat, long long rows, int cols, float* dgamma, float* dbeta,
                   me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(dgamma && dbeta && cols > 0 && rows >= 0 && (rows == 0 || (da_inout && a && xhat)), "bn_bwd_sums: bad argument");
  if (rows > 0) leaky_bwd_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(da_inout, a, rows * cols);
  colsum_kernel<<<kRowSplit * ((cols + 31) / 32), 32 * kColsumLanes, 0, stream>>>(da_inout, xhat, rows, cols, cols, cols, dgamma);
  colsum_kernel<<<kRowSplit * ((cols + 31) / 32), 32 * kColsumLanes, 0, stream>>>(da_inout, nullptr, rows, cols, cols, cols, dbeta);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_bn_bwd_apply(const float* dy, const float* xhat, long long rows, int cols, const float* gamma, const float* inv_std,
                    const float* dgamma_total, const float* dbeta_total, const double* total_rows, float* dz,
                    me_stream_t stream) {
  using namespace me;
  if (rows <= 0) return ME_OK;
  ME_REQUIRE(dy && xhat && gamma && inv_std && dgamma_total && dbeta_total && total_rows && dz && cols > 0,
             "bn_bwd_apply: bad argument");
  bn_bwd_apply_kernel<<<grid_for(rows * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, xhat, rows, cols, gamma, inv_std, dgamma_total, dbeta_total, total_rows, dz);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_leaky_bwd_f32(float* d_inout, const float* a, long long total, me_stream_t stream) {
  using namespace me;
  if (total <= 0) return ME_OK;
  ME_REQUIRE(d_inout && a, "leaky_bwd: null argument");
  leaky_bwd_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_inout, a, total);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_sigmoid_bwd_f32(float* d_inout, const float* s, long long total, me_stream_t stream) {
  using namespace me;
  if (total <= 0) return ME_OK;
  ME_REQUIRE(d_inout && s, "sigmoid_bwd: null argument");
  sigmoid_bwd_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_inout, s, total);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_roi_align_f32(int position_sensitive, int backward, const float* feat, float* dfeat, int n, int h, int w,
                     int chan_total, int channels, int pooled, float scale, const float* rois, int num_rois, float* out,
                     const float* grad_out, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_rois <= 0) return ME_OK;
  ME_REQUIRE(rois && n > 0 && h > 0 && w > 0 && channels > 0 && pooled > 0, "roi_align_f32: bad argument");
  ME_REQUIRE(chan_total >= (position_sensitive ? channels * pooled * pooled : channels), "roi_align_f32: too few channels");
  ME_REQUIRE(backward ? (dfeat && grad_out) : (feat && out), "roi_align_f32: null argument");
  const int grid = grid_for(1LL * num_rois * channels * pooled * pooled);
#define ME_ROI(PS, BW) \
  roi_f32_kernel<PS, BW><<<grid, 256, 0, stream>>>(feat, dfeat, n, h, w, chan_total, channels, pooled, scale, rois, num_rois, out, grad_out)
  if (position_sensitive) {
    if (backward) ME_ROI(true, true); else ME_ROI(true, false);
  } else {
    if (backward) ME_ROI(false, true); else ME_ROI(false, false);
  }
#undef ME_ROI
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_stage3_tail_fwd(const float* r2, const float* cls, int cls_pitch, const float* img_boxes, int n_img, int n_all,
                       const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b, float* rc,
                       float* refine, float* mask, float* p, me_stream_t stream) {
  using namespace me;
  if (n_all <= 0) return ME_OK;
  ME_REQUIRE(r2 && cls && fc1_w && fc1_b && fc2_w && fc2_b && rc && refine && mask && p && (n_img == 0 || img_boxes),
             "stage3_tail_fwd: null argument");
  stage3_tail_fwd_kernel<<<(n_all + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      r2, cls, cls_pitch, img_boxes, n_img, n_all, fc1_w, fc1_b, fc2_w, fc2_b, rc, refine, mask, p);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_stage3_tail_bwd(const float* rc, const float* refine, const float* cls, int cls_pitch, const float* p,
                       const float* img_boxes, int n_img, int n_all, const unsigned char* pos, const unsigned char* sel,
                       float alpha, float lambda_conf, const float* fc1_w, const float* fc1_b, const float* fc2_w, float* d_o,
                       float* hl, float* dhp, float* u, float* dr2, float* dz2, me_stream_t stream) {
  using namespace me;
  if (n_all <= 0) return ME_OK;
  ME_REQUIRE(rc && refine && cls && p && pos && sel && fc1_w && fc1_b && fc2_w && d_o && hl && dhp && u && dr2 && dz2,
             "stage3_tail_bwd: null argument");
  stage3_tail_bwd_kernel<<<(n_all + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      rc, refine, cls, cls_pitch, p, img_boxes, n_img, n_all, pos, sel, alpha, lambda_conf, fc1_w, fc1_b, fc2_w, d_o, hl, dhp,
      u, dr2, dz2);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                 float beta2, float eps, int step, me_stream_t stream) {
  using namespace me;
  if (n <= 0) return ME_OK;
  ME_REQUIRE(params && grads && exp_avg && exp_avg_sq && step >= 1, "adam_step: bad argument");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step)), bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adam_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                          eps, bc1, bc2);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
