// Confidence filter + class arg-max + batched NMS on the device, bit-compatible with
// non_max_suppression_cpp (utils/utils.py:337-378) and the torchvision CPU ops it calls
// (torchvision 0.26 ops/boxes.py:51-120 batched_nms; csrc/ops/cpu/nms_kernel.cpp greedy loop):
//   * rows with conf >= conf_thresh, in row order, are the candidates;
//   * class_conf / class_pred = first arg-max over the class scores (torch.max tie rule);
//   * candidates are ordered by objectness, descending, stable (ties keep row order);
//   * <= 1000 candidates: "coordinate trick" - boxes are offset by class * (max_coord + 1) in
//     fp32 and one class-agnostic NMS runs on the offset boxes (the offsets perturb IoUs, so
//     the same fp32 operations are replayed here);  > 1000: per-class NMS on the raw boxes;
//   * IoU = inter / (area_i + area_j - inter), strict '>' against the threshold (as double);
//   * the first max_det survivors in score order are returned.
// All fp32 arithmetic uses the _rn intrinsics so no FMA contraction changes a rounding.
#include "common.cuh"

namespace me {
namespace {

constexpr int kSelThreads = 1024;
constexpr int kMaxSmemKeys = 16384;
constexpr int kSmemBoxes = 2048;
constexpr int kMatrixK = 512;     // up to this many candidates the suppression matrix is built in shared memory
constexpr int kCompactIters = 16;  // rows / kSelThreads handled by the single-scan compaction  // sorted candidate boxes are staged in shared memory up to this many

struct Layout {
  // per-image slices of the workspace (all sized by `rows`, keys by pow2(rows))
  float* box;        // [rows][4] xyxy
  float* conf;       // [rows]
  float* cls_conf;   // [rows]
  int* cls_idx;      // [rows]
  float* sbox;       // [rows][4] boxes in sorted order (offset boxes on the coordinate-trick path)
  float* sarea;      // [rows]
  int* scls;         // [rows]
  unsigned long long* keys;  // [pow2]
};

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

__host__ __device__ inline size_t per_image_bytes(int rows) {
  const size_t r = (static_cast<size_t>(rows) + 3) & ~size_t(3);  // keeps every array 16B aligned
  size_t b = r * 4 * 4 + r * 4 + r * 4 + r * 4 + r * 4 * 4 + r * 4 + r * 4;
  b = (b + 15) & ~size_t(15);
  b += static_cast<size_t>(next_pow2(rows)) * 8;
  return (b + 255) & ~size_t(255);
}

__device__ inline Layout layout_for(void* ws, int rows, int img) {
  uint8_t* p = static_cast<uint8_t*>(ws) + per_image_bytes(rows) * img;
  Layout L;
  const size_t r = (static_cast<size_t>(rows) + 3) & ~size_t(3);
  L.box = reinterpret_cast<float*>(p);       p += r * 16;
  L.conf = reinterpret_cast<float*>(p);      p += r * 4;
  L.cls_conf = reinterpret_cast<float*>(p);  p += r * 4;
  L.cls_idx = reinterpret_cast<int*>(p);     p += r * 4;
  L.sbox = reinterpret_cast<float*>(p);      p += r * 16;
  L.sarea = reinterpret_cast<float*>(p);     p += r * 4;
  L.scls = reinterpret_cast<int*>(p);        p += r * 4;
  p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  L.keys = reinterpret_cast<unsigned long long*>(p);
  return L;
}

// One thread per prediction row: cxcywh -> xyxy (in place if asked), objectness filter, and - for the few rows that
// pass - class max / first arg-max.  (A warp per row left 31 lanes idle on the ~99 % of rows that fail the filter
// and serialised the row's five loads behind one lane: 71 us for 32 x 10647 rows.)
__global__ void __launch_bounds__(256)
nms_prepare_kernel(float* __restrict__ pred, int n, int rows, int nc, float conf_thresh, int inplace, void* ws) {
  const long long total = 1LL * n * rows;
  const int attrs = 5 + nc;
  for (long long rr = blockIdx.x * 1LL * blockDim.x + threadIdx.x; rr < total; rr += 1LL * gridDim.x * blockDim.x) {
    const int img = static_cast<int>(rr / rows);
    const int row = static_cast<int>(rr - 1LL * img * rows);
    float* src = pred + rr * attrs;
    Layout L = layout_for(ws, rows, img);
    const float cx = src[0], cy = src[1], w = src[2], h = src[3], conf = src[4];
    // xywh2xyxy, utils.py:68-74: x - w / 2, x + w / 2 (the halving is exact)
    const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
    const float x1 = __fsub_rn(cx, hw), y1 = __fsub_rn(cy, hh);
    const float x2 = __fadd_rn(cx, hw), y2 = __fadd_rn(cy, hh);
    *reinterpret_cast<float4*>(L.box + row * 4) = make_float4(x1, y1, x2, y2);
    L.conf[row] = conf;
    if (inplace) {
      src[0] = x1;
      src[1] = y1;
      src[2] = x2;
      src[3] = y2;
    }
    if (!(conf >= conf_thresh)) continue;
    float best = src[5];
    int best_i = 0;
    for (int c = 1; c < nc; ++c) {
      const float v = src[5 + c];
      if (v > best) {  // strict: the first maximum wins (torch.max tie rule)
        best = v;
        best_i = c;
      }
    }
    L.cls_conf[row] = best;
    L.cls_idx[row] = best_i;
  }
}

__device__ __forceinline__ unsigned int score_desc_bits(float s) {
  // monotone map float -> uint (ascending), then invert so that ascending keys = descending scores
  unsigned int u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}

// One block per image: compaction (row order) -> bitonic sort of (score desc, row asc) keys ->
// greedy suppression, stopping once max_det boxes are kept.
__global__ void __launch_bounds__(kSelThreads, 1)
nms_select_kernel(const float* __restrict__ pred, int rows, int nc, float conf_thresh, float nms_thresh_f, int max_det,
                  float* __restrict__ det, int* __restrict__ det_count, int* __restrict__ det_index, void* ws,
                  int keys_in_smem) {
  extern __shared__ unsigned long long s_dyn[];
  __shared__ int s_scan[kSelThreads / 32];
  __shared__ int s_k;
  __shared__ float s_red[kSelThreads / 32];
  __shared__ float s_maxc;
  __shared__ int s_keep[256];
  __shared__ int s_cnt[kCompactIters * (kSelThreads / 32)];

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  Layout L = layout_for(ws, rows, img);
  const int attrs = 5 + nc;
  const int det_cols = 7 + nc;
  const int pow2 = next_pow2(rows);
  // the host decides from the whole shared-memory budget whether the sort keys fit next to the fixed staging areas
  unsigned long long* keys = keys_in_smem ? s_dyn : L.keys;
  unsigned char* supp = reinterpret_cast<unsigned char*>(s_dyn + (keys_in_smem ? pow2 : 0));
  // staging area for the sorted boxes: the greedy loop reads box i once per survivor, and a global/L2 round trip
  // per survivor (~0.7 us x up to 200) was most of this kernel's 84 us
  float* sm_box = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(supp) + rows + 15) & ~uintptr_t(15));
  float* sm_area = sm_box + kSmemBoxes * 4;
  int* sm_cls = reinterpret_cast<int*>(sm_area + kSmemBoxes);
  uint32_t* sm_mask = reinterpret_cast<uint32_t*>(sm_cls + kSmemBoxes);  // [kMatrixK][kMatrixK / 32]

  // ---- 1. order-preserving compaction of candidate rows into keys[0..k)
  const int iters = (rows + kSelThreads - 1) / kSelThreads;
  if (iters <= kCompactIters) {
    // all loads first, one scan over the (iteration, warp) pass counts: 3 block barriers instead of 4 per 1024 rows
    float cv[kCompactIters];
    unsigned int bal[kCompactIters];
#pragma unroll
    for (int it = 0; it < kCompactIters; ++it) {
      const int row = it * kSelThreads + tid;
      cv[it] = (it < iters && row < rows) ? L.conf[row] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < kCompactIters; ++it) {
      const int row = it * kSelThreads + tid;
      const bool pass = it < iters && row < rows && (cv[it] >= conf_thresh);
      bal[it] = __ballot_sync(0xffffffffu, pass);
      if (lane == 0) s_cnt[it * (kSelThreads / 32) + wid] = __popc(bal[it]);
    }
    __syncthreads();
    if (wid == 0) {
      // entries are ordered (iteration, warp) = row order; lane l owns `iters` consecutive entries
      int sum = 0;
      for (int e = 0; e < iters; ++e) sum += s_cnt[lane * iters + e];
      int incl = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      int run = incl - sum;
      for (int e = 0; e < iters; ++e) {
        const int c = s_cnt[lane * iters + e];
        s_cnt[lane * iters + e] = run;
        run += c;
      }
      if (lane == 31) s_k = incl;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kCompactIters; ++it) {
      if ((bal[it] >> lane) & 1u) {
        const int row = it * kSelThreads + tid;
        const int pos = s_cnt[it * (kSelThreads / 32) + wid] + __popc(bal[it] & ((1u << lane) - 1));
        keys[pos] = (static_cast<unsigned long long>(score_desc_bits(cv[it])) << 32) | static_cast<unsigned int>(row);
      }
    }
    __syncthreads();
  } else {
    if (tid == 0) s_k = 0;
    __syncthreads();
    for (int base = 0; base < rows; base += kSelThreads) {
      const int row = base + tid;
      const bool pass = row < rows && (L.conf[row] >= conf_thresh);
      const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
      if (lane == 0) s_scan[wid] = __popc(ballot);
      __syncthreads();
      if (wid == 0) {
        int v = s_scan[lane];
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += t;
        }
        s_scan[lane] = incl - v;  // exclusive prefix of the warp totals
        if (lane == 31) s_red[0] = __int_as_float(incl);
      }
      __syncthreads();
      const int k0 = s_k;
      if (pass) {
        const int pos = k0 + s_scan[wid] + __popc(ballot & ((1u << lane) - 1));
        keys[pos] = (static_cast<unsigned long long>(score_desc_bits(L.conf[row])) << 32) | static_cast<unsigned int>(row);
      }
      __syncthreads();
      if (tid == 0) s_k = k0 + __float_as_int(s_red[0]);
      __syncthreads();
    }
  }
  const int k = s_k;
  if (k == 0) {
    if (tid == 0) det_count[img] = 0;
    return;
  }

  // ---- 2. bitonic sort of the k keys (padded with +inf keys to a power of two)
  const int kp = next_pow2(k);
  for (int i = k + tid; i < kp; i += kSelThreads) keys[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (kp >> 1); t += kSelThreads) {
        const int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` clear
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == asc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }

  // ---- 3. boxes in sorted order; coordinate trick when boxes.numel() <= 4000
  const bool trick = (k * 4) <= 4000;
  float maxc = -INFINITY;
  if (trick) {
    for (int i = tid; i < k; i += kSelThreads) {
      const int row = static_cast<int>(keys[i] & 0xffffffffu);
      const float4 b = *reinterpret_cast<const float4*>(L.box + row * 4);
      maxc = fmaxf(fmaxf(maxc, fmaxf(b.x, b.y)), fmaxf(b.z, b.w));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) maxc = fmaxf(maxc, __shfl_xor_sync(0xffffffffu, maxc, off));
    if (lane == 0) s_red[wid] = maxc;
    __syncthreads();
    if (wid == 0) {
      float v = s_red[lane];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
      if (lane == 0) s_maxc = v;
    }
    __syncthreads();
  }
  const float span = trick ? __fadd_rn(s_maxc, 1.f) : 0.f;  // max_coordinate + 1
  const bool in_smem = k <= kSmemBoxes;
  float* sbox = in_smem ? sm_box : L.sbox;
  float* sarea = in_smem ? sm_area : L.sarea;
  int* scls = in_smem ? sm_cls : L.scls;
  for (int i = tid; i < k; i += kSelThreads) {
    const int row = static_cast<int>(keys[i] & 0xffffffffu);
    float4 b = *reinterpret_cast<const float4*>(L.box + row * 4);
    const int cls = L.cls_idx[row];
    if (trick) {
      const float off = __fmul_rn(static_cast<float>(cls), span);  // idxs.to(boxes) * (max + 1)
      b.x = __fadd_rn(b.x, off);
      b.y = __fadd_rn(b.y, off);
      b.z = __fadd_rn(b.z, off);
      b.w = __fadd_rn(b.w, off);
    }
    *reinterpret_cast<float4*>(sbox + i * 4) = b;
    sarea[i] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    scls[i] = cls;
    supp[i] = 0;
  }
  __syncthreads();

  // ---- 4. greedy suppression in score order
  int kept = 0;
  if (k <= kMatrixK) {
    // 4a. suppression matrix: one warp per (box i, 32-candidate word), lanes = candidates, ballot = the word.
    // Bit j of row i says "i suppresses j" (j > i).  All pairs are independent, so 32 warps stay busy; the greedy
    // loop with one block barrier per survivor cost ~930 clocks per survivor (105 k clocks for k = 200).
    const int W = (k + 31) >> 5;
    for (int t = wid; t < k * W; t += kSelThreads / 32) {
      const int i = t / W, w = t - i * W;
      const int j = w * 32 + lane;
      bool hit = false;
      if (w >= (i >> 5) && j > i && j < k && (trick || scls[j] == scls[i])) {
        const float4 bi = *reinterpret_cast<const float4*>(sbox + i * 4);
        const float4 bj = *reinterpret_cast<const float4*>(sbox + j * 4);
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float bw = fmaxf(0.f, __fsub_rn(xx2, xx1)), bh = fmaxf(0.f, __fsub_rn(yy2, yy1));
        const float inter = __fmul_rn(bw, bh);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sarea[i], sarea[j]), inter));
        hit = ovr > nms_thresh_f;
      }
      const unsigned int word = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) sm_mask[t] = word;
    }
    __syncthreads();
    // 4b. one warp walks the candidates in score order; lane w carries word w of the "removed" set and the
    // next row of the matrix is fetched while the current candidate is examined
    if (wid == 0) {
      unsigned int removed = 0;
      unsigned int next_row = lane < W ? sm_mask[lane] : 0u;
      for (int i = 0; i < k; ++i) {
        const unsigned int row_i = next_row;
        if (i + 1 < k) next_row = lane < W ? sm_mask[(i + 1) * W + lane] : 0u;
        const unsigned int word = __shfl_sync(0xffffffffu, removed, i >> 5);
        if ((word >> (i & 31)) & 1u) continue;
        if (lane == 0) s_keep[kept] = i;
        ++kept;
        if (kept == max_det) break;
        removed |= row_i;
      }
      if (lane == 0) s_k = kept;
    }
    __syncthreads();
    kept = s_k;
  } else {
    for (int i = 0; i < k; ++i) {
      if (supp[i]) continue;  // block-uniform
      if (tid == 0) s_keep[kept] = i;
      ++kept;
      if (kept == max_det) break;
      const float4 bi = *reinterpret_cast<const float4*>(sbox + i * 4);
      const float ai = sarea[i];
      const int ci = scls[i];
      for (int j = i + 1 + tid; j < k; j += kSelThreads) {
        if (supp[j]) continue;
        if (!trick && scls[j] != ci) continue;
        const float4 bj = *reinterpret_cast<const float4*>(sbox + j * 4);
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
        const float inter = __fmul_rn(w, h);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, sarea[j]), inter));
        if (ovr > nms_thresh_f) supp[j] = 1;
      }
      __syncthreads();
    }
    __syncthreads();
  }

  // ---- 5. gather survivors: [x1,y1,x2,y2,conf,class_conf,class_pred,cls...]
  if (tid == 0) det_count[img] = kept;
  float* out = det + 1LL * img * max_det * det_cols;
  for (int e = tid; e < kept * det_cols; e += kSelThreads) {
    const int d = e / det_cols, c = e - d * det_cols;
    const int row = static_cast<int>(keys[s_keep[d]] & 0xffffffffu);
    float v;
    if (c < 4) v = L.box[row * 4 + c];
    else if (c == 4) v = L.conf[row];
    else if (c == 5) v = L.cls_conf[row];
    else if (c == 6) v = static_cast<float>(L.cls_idx[row]);
    else v = pred[(1LL * img * rows + row) * attrs + 5 + (c - 7)];
    out[e] = v;
  }
  for (int d = tid; d < kept; d += kSelThreads)
    det_index[1LL * img * max_det + d] = static_cast<int>(keys[s_keep[d]] & 0xffffffffu);
}

}  // namespace
}  // namespace me

extern "C" {

size_t me_filter_nms_workspace(int n, int rows, int num_classes) {
  (void)num_classes;
  if (n <= 0 || rows <= 0) return 0;
  return me::per_image_bytes(rows) * static_cast<size_t>(n);
}

int me_filter_nms(float* pred, int n, int rows, int num_classes, float conf_thresh, double nms_thresh, int max_det,
                  int xyxy_inplace, float* det, int* det_count, int* det_index, void* workspace,
                  size_t workspace_bytes, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(pred && det && det_count && det_index && workspace, "filter_nms: null argument");
  ME_REQUIRE(n > 0 && rows > 0 && num_classes >= 1, "filter_nms: empty input");
  ME_REQUIRE(max_det >= 1 && max_det <= 256, "filter_nms: max_det %d out of range (1..256)", max_det);
  ME_REQUIRE(rows <= (1 << 20), "filter_nms: too many rows");
  ME_REQUIRE(workspace_bytes >= me_filter_nms_workspace(n, rows, num_classes), "filter_nms: workspace too small");
  long long blocks = (1LL * n * rows + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  nms_prepare_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(pred, n, rows, num_classes, conf_thresh,
                                                                   xyxy_inplace, workspace);
  ME_LAUNCH_CHECK();
  const int pow2 = next_pow2(rows);
  // Sort keys live in shared memory when they fit NEXT TO the box stage and the suppression matrix (the whole budget,
  // not just pow2 <= kMaxSmemKeys: 12 257..16 384 rows - YOLOv3 at 448 / 480 / 512 - need 128 KB of keys plus 80 KB of
  // fixed areas and used to fail the limit below); otherwise the kernel sorts in the global workspace.
  constexpr size_t kSmemLimit = 220 * 1024;
  const size_t fixed = static_cast<size_t>(rows) + 32 + static_cast<size_t>(kSmemBoxes) * 24 +
                       static_cast<size_t>(kMatrixK) * (kMatrixK / 32) * 4;
  const int keys_in_smem = (pow2 <= kMaxSmemKeys && static_cast<size_t>(pow2) * 8 + fixed <= kSmemLimit) ? 1 : 0;
  const size_t smem = (keys_in_smem ? static_cast<size_t>(pow2) * 8 : 0) + fixed;
  // (double)ovr > nms_thresh for a float ovr  <=>  ovr > the largest float that is <= nms_thresh
  float thresh_f = static_cast<float>(nms_thresh);
  if (static_cast<double>(thresh_f) > nms_thresh) thresh_f = nextafterf(thresh_f, -INFINITY);
  int dev = 0;
  ME_CUDA(cudaGetDevice(&dev));
  static bool attr_set[64] = {false};   // per device: the attribute belongs to the device's copy of the function
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    ME_CUDA(cudaFuncSetAttribute(nms_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  ME_REQUIRE(smem <= kSmemLimit, "filter_nms: %d rows need %zu B of shared memory", rows, smem);
  nms_select_kernel<<<n, kSelThreads, smem, stream>>>(pred, rows, num_classes, conf_thresh, thresh_f, max_det, det,
                                                      det_count, det_index, workspace, keys_in_smem);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
