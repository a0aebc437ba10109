// Proposal assembly, the small per-RoI heads and the output stage of Network.forward
// (my_models.py:459-539).  Everything is capacity-bounded with device-side counts, so the
// whole fusion forward runs without a host synchronisation until the caller reads the result.
#include "common.cuh"

namespace me {
namespace {

constexpr int kBlock = 1024;

// Block-wide, order-preserving compaction step: returns this thread's output slot (or -1) and
// advances *s_total.  s_scan must hold kBlock/32 ints.
__device__ __forceinline__ int block_compact(bool pass, int* s_scan, int* s_total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
  if (lane == 0) s_scan[wid] = __popc(ballot);
  __syncthreads();
  const int base = *s_total;
  __syncthreads();
  if (wid == 0) {
    const int v = s_scan[lane];
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    s_scan[lane] = incl - v;
    if (lane == 31) *s_total = base + incl;
  }
  __syncthreads();
  return pass ? base + s_scan[wid] + __popc(ballot & ((1u << lane) - 1)) : -1;
}

// my_models.py:459-473 + :490-492.  One block.
__global__ void __launch_bounds__(kBlock)
build_proposals_kernel(const float* __restrict__ det, const int* __restrict__ det_count, int n, int max_det, int det_cols,
                       int class_idx, const float* __restrict__ radar, int num_radar, const int* __restrict__ num_radar_dev,
                       float img_size, float* __restrict__ img_boxes, float* __restrict__ rois, int* __restrict__ counts,
                       int cap) {
  if (num_radar_dev) num_radar = max(0, min(*num_radar_dev, num_radar));   // device-side count, host value = capacity
  __shared__ int s_scan[kBlock / 32];
  __shared__ int s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  const int total = n * max_det;
  for (int base = 0; base < total; base += kBlock) {
    const int i = base + threadIdx.x;
    bool pass = false;
    const float* d = nullptr;
    int img = 0;
    if (i < total) {
      img = i / max_det;
      const int k = i - img * max_det;
      if (k < det_count[img]) {
        d = det + (1LL * img * max_det + k) * det_cols;
        // m3: detection_i[:, 6] == self.class_idx (my_models.py:463); stage 2 keeps every class (m2 :325-330)
        pass = class_idx < 0 || (d[6] == static_cast<float>(class_idx));
      }
    }
    const int slot = block_compact(pass, s_scan, &s_total);
    if (slot >= 0 && slot < cap) {
      if (class_idx < 0) {
        float* b = img_boxes + 1LL * slot * (1 + det_cols);
        b[0] = static_cast<float>(img);
        for (int c = 0; c < det_cols; ++c) b[1 + c] = d[c];
      } else {
        float* b = img_boxes + slot * 9;
        b[0] = static_cast<float>(img);
#pragma unroll
        for (int c = 0; c < 7; ++c) b[1 + c] = d[c];
        b[8] = d[7 + class_idx];
      }
      float* r = rois + slot * 5;
      r[0] = static_cast<float>(img);
      r[1] = d[0];
      r[2] = d[1];
      r[3] = d[2];
      r[4] = d[3];
    }
  }
  __syncthreads();
  const int n_img = min(s_total, cap);
  for (int j = threadIdx.x; j < num_radar; j += kBlock) {
    const int slot = n_img + j;
    if (slot < cap) {
      float* r = rois + slot * 5;
      r[0] = radar[j * 5];
#pragma unroll
      for (int c = 1; c < 5; ++c) r[c] = radar[j * 5 + c] * img_size;
    }
  }
  if (threadIdx.x == 0) {
    counts[0] = n_img;
    counts[1] = min(n_img + num_radar, cap);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float leakyf_(float v) { return v > 0.f ? v : 0.1f * v; }

// One warp per RoI.  refinement_head.forward tail (my_models.py:264-284) and ensemble_head.forward
// (:202-210) on image proposals.
__global__ void __launch_bounds__(256)
fusion_heads_kernel(const __half* __restrict__ hidden, int hidden_pitch, const __half* __restrict__ crop, int crop_pitch,
                    me_head_weights hw, const float* __restrict__ img_boxes, const int* __restrict__ counts, int cap,
                    float* __restrict__ regress, float* __restrict__ refine, float* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_img = counts[0], n_all = min(counts[1], cap);
  for (int r = warp; r < n_all; r += nwarps) {
    // hidden vector: 256 values, 8 per lane
    float hv[8];
    {
      const uint4 raw = *reinterpret_cast<const uint4*>(hidden + 1LL * r * hidden_pitch + lane * 8);
      const __half2* p = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(p[e]);
        hv[2 * e] = f.x;
        hv[2 * e + 1] = f.y;
      }
    }
    float reg[4], cls[13];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float* wrow = hw.net1_w + o * 256 + lane * 8;
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(hv[e], wrow[e], s);
      reg[o] = warp_sum(s) + hw.net1_b[o];
    }
#pragma unroll
    for (int o = 0; o < 13; ++o) {
      const float* wrow = hw.net2_w + o * 256 + lane * 8;
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(hv[e], wrow[e], s);
      cls[o] = sigmoidf_(warp_sum(s) + hw.net2_b[o]);
    }
    // radar branch: 7x7 conv over the 10x7x7 crop == 490-long dot product per output channel
    float t[10];
#pragma unroll
    for (int o = 0; o < 10; ++o) t[o] = 0.f;
    for (int e = lane; e < 490; e += 32) {
      const float x = __half2float(crop[1LL * r * crop_pitch + e]);
#pragma unroll
      for (int o = 0; o < 10; ++o) t[o] = fmaf(x, hw.radar_w[o * 490 + e], t[o]);
    }
    float rc = hw.radar2_b[0];
#pragma unroll
    for (int o = 0; o < 10; ++o) rc = fmaf(leakyf_(warp_sum(t[o]) + hw.radar_b[o]), hw.radar2_w[o], rc);
    rc = sigmoidf_(rc);
    const float conf = sigmoidf_(rc + cls[0]);
    const float cscore = cls[1];
    float m = conf;
    if (r < n_img) {
      // ensemble: x[j] = (refinement[j], yolo[j]); fc1 2->32 + leaky per j; flatten 64; fc2 64->2; softmax
      const float yolo[2] = {img_boxes[r * 9 + 5], img_boxes[r * 9 + 8]};
      const float refv[2] = {conf, cscore};
      float o0 = 0.f, o1 = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float hcell = leakyf_(fmaf(hw.fc1_w[lane * 2], refv[j], fmaf(hw.fc1_w[lane * 2 + 1], yolo[j], hw.fc1_b[lane])));
        o0 = fmaf(hcell, hw.fc2_w[j * 32 + lane], o0);
        o1 = fmaf(hcell, hw.fc2_w[64 + j * 32 + lane], o1);
      }
      o0 = warp_sum(o0) + hw.fc2_b[0];
      o1 = warp_sum(o1) + hw.fc2_b[1];
      const float mx = fmaxf(o0, o1);
      const float e0 = expf(o0 - mx), e1 = expf(o1 - mx);
      m = e0 / (e0 + e1);  // masks_img_proposals[:, :1] (column 0), my_models.py:513
    }
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 4; ++o) regress[r * 4 + o] = reg[o];
      refine[r * 2] = conf;
      refine[r * 2 + 1] = cscore;
      mask[r] = m;
    }
  }
}

// Stage-2 variant (module2_mixed/my_models.py): refinement_head.forward :121-126 (FC 256->4, sigmoid FC 256->NV)
// and ensemble_head.forward :153-161 (stack -> FC 2->32 + leaky -> flatten 32*NV -> FC -> LeakyReLU -> softmax);
// the new confidence is column 1 of the softmax (:352).  One warp per RoI, lane = hidden unit of fc1.
template <int NV>
__global__ void __launch_bounds__(256)
stage2_heads_kernel(const __half* __restrict__ hidden, int hidden_pitch, me_stage2_weights hw,
                    const float* __restrict__ boxes, int box_pitch, const int* __restrict__ counts, int cap,
                    float* __restrict__ regress, float* __restrict__ mask, float* __restrict__ refine) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_all = min(counts[1], cap);
  for (int r = warp; r < n_all; r += nwarps) {
    float hv[8];
    {
      const uint4 raw = *reinterpret_cast<const uint4*>(hidden + 1LL * r * hidden_pitch + lane * 8);
      const __half2* p = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(p[e]);
        hv[2 * e] = f.x;
        hv[2 * e + 1] = f.y;
      }
    }
    float reg[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float* wrow = hw.net1_w + o * 256 + lane * 8;
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(hv[e], wrow[e], s);
      reg[o] = warp_sum(s) + hw.net1_b[o];
    }
    const float* box = boxes + 1LL * r * box_pitch;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float* wrow = hw.net2_w + j * 256 + lane * 8;
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(hv[e], wrow[e], s);
      const float refv = sigmoidf_(warp_sum(s) + hw.net2_b[j]);
      if (refine != nullptr && lane == 0) refine[static_cast<size_t>(r) * NV + j] = refv;   // kept for the loss branch
      const float yolo = j == 0 ? box[5] : box[8 + (j - 1)];  // (obj_conf, class scores) m2 :347
      const float hcell = leakyf_(fmaf(hw.fc1_w[lane * 2], refv, fmaf(hw.fc1_w[lane * 2 + 1], yolo, hw.fc1_b[lane])));
      o0 = fmaf(hcell, hw.fc2_w[j * 32 + lane], o0);
      o1 = fmaf(hcell, hw.fc2_w[NV * 32 + j * 32 + lane], o1);
    }
    o0 = leakyf_(warp_sum(o0) + hw.fc2_b[0]);
    o1 = leakyf_(warp_sum(o1) + hw.fc2_b[1]);
    const float mx = fmaxf(o0, o1);
    const float e0 = expf(o0 - mx), e1 = expf(o1 - mx);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 4; ++o) regress[r * 4 + o] = reg[o];
      mask[r] = e1 / (e0 + e1);
    }
  }
}

__device__ __forceinline__ unsigned int desc_bits(float s) {
  unsigned int u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}

// my_models.py:516-539.  One block: threshold -> compaction -> bitonic sort (priority desc) -> rows.
__global__ void __launch_bounds__(kBlock)
finalize_kernel(const float* __restrict__ img_boxes, const float* __restrict__ rois, const float* __restrict__ refine,
                const float* __restrict__ regress, const float* __restrict__ mask, const int* __restrict__ counts,
                int cap, int box_pitch, float thr_img, float thr_radar, int do_regress, float* __restrict__ out,
                int* __restrict__ out_count, unsigned long long* __restrict__ keys) {
  __shared__ int s_scan[kBlock / 32];
  __shared__ int s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  const int n_img = counts[0], n_all = min(counts[1], cap);
  for (int base = 0; base < n_all; base += kBlock) {
    const int r = base + threadIdx.x;
    bool pass = false;
    float pri = 0.f;
    if (r < n_all) {
      const float m = mask[r];
      pass = r < n_img ? (m > thr_img) : (m > thr_radar);
      pri = r < n_img ? m : m / 5.f;  // masks_tmp[num_img_boxes:, 1] /= 5
    }
    const int slot = block_compact(pass, s_scan, &s_total);
    if (slot >= 0) keys[slot] = (static_cast<unsigned long long>(desc_bits(pri)) << 32) | static_cast<unsigned int>(r);
  }
  __syncthreads();
  const int k = s_total;
  if (threadIdx.x == 0) out_count[0] = k;
  if (k == 0) return;
  int kp = 1;
  while (kp < k) kp <<= 1;
  for (int i = k + threadIdx.x; i < kp; i += kBlock) keys[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (kp >> 1); t += kBlock) {
        const int lo = 2 * t - (t & (stride - 1));
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
  for (int i = threadIdx.x; i < k; i += kBlock) {
    const int r = static_cast<int>(keys[i] & 0xffffffffu);
    const float* roi = rois + r * 5;
    float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
    if (do_regress) {
      // box_regress, my_models.py:378-391 via xyxy2xywh / xywh2xyxy (utils.py:58-74)
      const float cx = (x1 + x2) / 2.f, cy = (y1 + y2) / 2.f, w = x2 - x1, h = y2 - y1;
      const float* g = regress + r * 4;
      const float nx = g[0] * w + cx, ny = g[1] * h + cy;
      const float nw = expf(g[2]) * w, nh = expf(g[3]) * h;
      x1 = nx - nw / 2.f;
      y1 = ny - nh / 2.f;
      x2 = nx + nw / 2.f;
      y2 = ny + nh / 2.f;
    }
    float* o = out + i * 8;
    o[0] = roi[0];
    o[1] = x1;
    o[2] = y1;
    o[3] = x2;
    o[4] = y2;
    o[5] = mask[r];
    if (r < n_img) {
      o[6] = img_boxes[1LL * r * box_pitch + 6];  // class score
      o[7] = img_boxes[1LL * r * box_pitch + 7];  // class pred
    } else {
      o[6] = refine[r * 2 + 1];
      o[7] = 0.f;
    }
  }
}

}  // namespace
}  // namespace me

extern "C" {

static int build_proposals_impl(const float* det, const int* det_count, int n, int max_det, int det_cols, int class_idx,
                                const float* radar_boxes, int num_radar, const int* num_radar_dev, float img_size,
                                float* img_boxes, float* rois, int* counts, int cap, cudaStream_t stream) {
  using namespace me;
  ME_REQUIRE(det && det_count && img_boxes && rois && counts, "build_proposals: null argument");
  ME_REQUIRE(num_radar == 0 || radar_boxes, "build_proposals: radar boxes missing");
  ME_REQUIRE(det_cols >= 8 && 7 + class_idx < det_cols, "build_proposals: bad det_cols/class_idx");
  ME_REQUIRE(cap > 0 && n > 0 && max_det > 0, "build_proposals: empty problem");
  build_proposals_kernel<<<1, kBlock, 0, stream>>>(det, det_count, n, max_det, det_cols, class_idx, radar_boxes, num_radar,
                                                   num_radar_dev, img_size, img_boxes, rois, counts, cap);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_build_proposals(const float* det, const int* det_count, int n, int max_det, int det_cols, int class_idx,
                       const float* radar_boxes, int num_radar, float img_size, float* img_boxes, float* rois,
                       int* counts, int cap, me_stream_t stream_) {
  return build_proposals_impl(det, det_count, n, max_det, det_cols, class_idx, radar_boxes, num_radar, nullptr, img_size,
                              img_boxes, rois, counts, cap, static_cast<cudaStream_t>(stream_));
}

int me_build_proposals_dev(const float* det, const int* det_count, int n, int max_det, int det_cols, int class_idx,
                           const float* radar_boxes, int radar_cap, const int* num_radar_dev, float img_size,
                           float* img_boxes, float* rois, int* counts, int cap, me_stream_t stream_) {
  using namespace me;
  ME_REQUIRE(num_radar_dev, "build_proposals_dev: null radar count");
  return build_proposals_impl(det, det_count, n, max_det, det_cols, class_idx, radar_boxes, radar_cap, num_radar_dev, img_size,
                              img_boxes, rois, counts, cap, static_cast<cudaStream_t>(stream_));
}

int me_fusion_heads(const void* hidden, int hidden_pitch, const void* radar_crop, int radar_pitch,
                    const me_head_weights* hw, const float* img_boxes, const int* counts, int cap, float* regress,
                    float* refine, float* mask, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(hidden && radar_crop && hw && img_boxes && counts && regress && refine && mask, "fusion_heads: null argument");
  ME_REQUIRE(hidden_pitch >= 256 && hidden_pitch % 8 == 0 && radar_pitch >= 490, "fusion_heads: bad pitches");
  int blocks = ceil_div(cap, 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  fusion_heads_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __half*>(hidden), hidden_pitch,
                                                  static_cast<const __half*>(radar_crop), radar_pitch, *hw, img_boxes,
                                                  counts, cap, regress, refine, mask);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_stage2_heads(const void* hidden, int hidden_pitch, const me_stage2_weights* hw, const float* boxes, int box_pitch,
                    int num_vec, const int* counts, int cap, float* regress, float* mask, float* refine_out,
                    me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(hidden && hw && boxes && counts && regress && mask, "stage2_heads: null argument");
  ME_REQUIRE(hidden_pitch >= 256 && hidden_pitch % 8 == 0, "stage2_heads: bad hidden pitch");
  ME_REQUIRE(num_vec == 13, "stage2_heads: %d-entry class vectors unsupported (13 = 12 classes + objectness)", num_vec);
  ME_REQUIRE(box_pitch >= 8 + (num_vec - 1), "stage2_heads: box rows too narrow");
  int blocks = ceil_div(cap, 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  stage2_heads_kernel<13><<<blocks, 256, 0, stream>>>(static_cast<const __half*>(hidden), hidden_pitch, *hw, boxes,
                                                      box_pitch, counts, cap, regress, mask, refine_out);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

size_t me_finalize_workspace(int cap) {
  if (cap <= 0) return 0;
  int kp = 1;
  while (kp < cap) kp <<= 1;
  return static_cast<size_t>(kp) * 8;
}

int me_finalize_output(const float* img_boxes, const float* rois, const float* refine, const float* regress,
                       const float* mask, const int* counts, int cap, int box_pitch, float thr_img, float thr_radar,
                       int regress_boxes,
                       float* out, int* out_count, void* workspace, size_t workspace_bytes, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(img_boxes && rois && refine && regress && mask && counts && out && out_count && workspace,
             "finalize: null argument");
  ME_REQUIRE(workspace_bytes >= me_finalize_workspace(cap), "finalize: workspace too small");
  ME_REQUIRE(box_pitch >= 8, "finalize: box_pitch %d < 8", box_pitch);
  finalize_kernel<<<1, kBlock, 0, stream>>>(img_boxes, rois, refine, regress, mask, counts, cap, box_pitch, thr_img, thr_radar,
                                            regress_boxes, out, out_count, static_cast<unsigned long long*>(workspace));
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
