// Stage-3 labelling and loss of Network.forward with targets (my_models.py:545-640) on the device buffers of
// the fusion forward: obtain_iou_labels (:317-375), FocalLoss (:287-314), the confidence / category BCE
// (:614-633) and regression_loss (:394-408).  The reference runs the labelling as an O(R*T) Python loop on the
// CPU after copying every proposal to the host; here it is one thread per proposal and one block for the sums.
#include "common.cuh"

namespace me {
namespace {

constexpr int kBlock = 1024;

// bbox_iou(x1y1x2y2=True), utils/utils.py:255-281.  Every operation is rounded to fp32 on its own (no FMA
// contraction) so the label equals the reference's torch-CPU value bit for bit.
__device__ __forceinline__ float iou_plus_one(float ax1, float ay1, float ax2, float ay2, float bx1, float by1,
                                              float bx2, float by2) {
  const float ix1 = fmaxf(ax1, bx1), iy1 = fmaxf(ay1, by1);
  const float ix2 = fminf(ax2, bx2), iy2 = fminf(ay2, by2);
  const float iw = fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.f), 0.f);
  const float ih = fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.f), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float a1 = __fmul_rn(__fadd_rn(__fsub_rn(ax2, ax1), 1.f), __fadd_rn(__fsub_rn(ay2, ay1), 1.f));
  const float a2 = __fmul_rn(__fadd_rn(__fsub_rn(bx2, bx1), 1.f), __fadd_rn(__fsub_rn(by2, by1), 1.f));
  const float uni = __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-16f);
  return __fdiv_rn(inter, uni);
}

// One thread per proposal row: the best-overlapping target of the same image and class (first maximum).
__global__ void stage3_labels_kernel(const float* __restrict__ img_boxes, int box_pitch, const float* __restrict__ rois,
                                     const int* __restrict__ counts, int cap, const float* __restrict__ targets,
                                     int num_targets, float* __restrict__ iou_labels,
                                     float* __restrict__ target_location) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cap) return;
  const int n_img = counts[0], n_all = min(counts[1], cap);
  float best = 0.f, loc[4] = {0.f, 0.f, 0.f, 0.f};
  if (r < n_all) {
    const float img = rois[r * 5];
    const float cls = r < n_img ? img_boxes[static_cast<size_t>(r) * box_pitch + 7] : 0.f;  // radar rows: class 0 (:503)
    const float x1 = rois[r * 5 + 1], y1 = rois[r * 5 + 2], x2 = rois[r * 5 + 3], y2 = rois[r * 5 + 4];
    bool any = false;
    for (int t = 0; t < num_targets; ++t) {
      const float* g = targets + t * 6;
      if (g[0] != img || g[1] != cls) continue;
      const float v = iou_plus_one(x1, y1, x2, y2, g[2], g[3], g[4], g[5]);
      if (!any || v > best) {  // torch.max: first maximum
        any = true;
        best = v;
        loc[0] = g[2];
        loc[1] = g[3];
        loc[2] = g[4];
        loc[3] = g[5];
      }
    }
    if (!any) best = 0.f;
  }
  iou_labels[r] = best;
#pragma unroll
  for (int k = 0; k < 4; ++k) target_location[r * 4 + k] = loc[k];
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (wid == 0) {
    t = s_red[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
    if (lane == 0) s_red[0] = t;
  }
  __syncthreads();
  return s_red[0];
}

__device__ __forceinline__ float smooth_l1(float a, float b) {
  const float d = fabsf(a - b);
  return d < 1.f ? 0.5f * d * d : d - 0.5f;
}

// One block.  Per-term sums in double (the reference sums fp32 values in torch's reduction order; the terms
// themselves are computed in fp32 like the reference's).
__global__ void __launch_bounds__(kBlock)
stage3_loss_kernel(const float* __restrict__ rois, const float* __restrict__ refine, const float* __restrict__ regress,
                   const float* __restrict__ mask, const int* __restrict__ counts, int cap,
                   const float* __restrict__ iou_labels, const float* __restrict__ target_location,
                   const unsigned char* __restrict__ sample_filter, me_stage3_loss_cfg cfg, float* __restrict__ out) {
  __shared__ double s_red[kBlock / 32];
  const int n_img = counts[0], n_all = min(counts[1], cap);
  double n_pos = 0.0;
  for (int r = threadIdx.x; r < n_all; r += kBlock) n_pos += iou_labels[r] > cfg.iou_hi ? 1.0 : 0.0;
  const int P = static_cast<int>(block_sum(n_pos, s_red));

  double focal = 0.0, conf = 0.0, lxy = 0.0, lwh = 0.0, cat = 0.0, positive = 0.0, tp = 0.0;
  for (int r = threadIdx.x; r < n_all; r += kBlock) {
    const bool pos = iou_labels[r] > cfg.iou_hi;
    const bool sel = sample_filter[r] != 0;
    const float m = mask[r];
    const bool is_pos_mask = r < n_img ? (m > cfg.thr_img) : (m > cfg.thr_radar);
    positive += is_pos_mask ? 1.0 : 0.0;
    tp += (is_pos_mask && pos) ? 1.0 : 0.0;
    if (sel && r < n_img) {
      // FocalLoss on masks = [1 - m, m] against the one-hot label (:296-307)
      const float p = pos ? m : 1.f - m;
      const float a = pos ? cfg.alpha : 1.f - cfg.alpha;
      const float q = 1.f - p;
      focal += static_cast<double>(-a * (q * q) * logf(p));
    }
    if (sel) {
      const float x = refine[r * 2];
      const float y = pos ? 1.f : 0.f;
      conf += static_cast<double>(-(y * fmaxf(logf(x), -100.f) + (1.f - y) * fmaxf(logf(1.f - x), -100.f)));
    }
    if (pos) {
      const float x1 = rois[r * 5 + 1], y1 = rois[r * 5 + 2], x2 = rois[r * 5 + 3], y2 = rois[r * 5 + 4];
      const float* g = target_location + r * 4;
      const float x = (x1 + x2) / 2.f, y = (y1 + y2) / 2.f, w = x2 - x1, h = y2 - y1;
      const float xt = (g[0] + g[2]) / 2.f, yt = (g[1] + g[3]) / 2.f, wt = g[2] - g[0], ht = g[3] - g[1];
      const float* q = regress + r * 4;
      lxy += static_cast<double>(smooth_l1((xt - x) / (w + 1e-16f), q[0]) + smooth_l1((yt - y) / (h + 1e-16f), q[1]));
      lwh += static_cast<double>(smooth_l1(logf(wt / w + 1e-16f), q[2]) + smooth_l1(logf(ht / h + 1e-16f), q[3]));
      // class_label row i (i-th positive, not row idx) is set, then rows are picked by pos_filter (:628-633)
      const float yc = r < P ? 1.f : 0.f;
      const float xc = refine[r * 2 + 1];
      cat += static_cast<double>(-(yc * fmaxf(logf(xc), -100.f) + (1.f - yc) * fmaxf(logf(1.f - xc), -100.f)));
    }
  }
  focal = block_sum(focal, s_red);
  conf = block_sum(conf, s_red);
  lxy = block_sum(lxy, s_red);
  lwh = block_sum(lwh, s_red);
  cat = block_sum(cat, s_red);
  positive = block_sum(positive, s_red);
  tp = block_sum(tp, s_red);
  if (threadIdx.x == 0) {
    const float masks_loss = static_cast<float>(focal), conf_loss = static_cast<float>(conf);
    out[0] = masks_loss;
    out[1] = conf_loss;
    out[2] = static_cast<float>(lxy);
    out[3] = static_cast<float>(lwh);
    out[4] = static_cast<float>(cat);
    out[5] = masks_loss + conf_loss / cfg.lambda_conf;  // loss (:635)
    out[6] = static_cast<float>(P);
    out[7] = static_cast<float>(positive);
    out[8] = static_cast<float>(tp);
    out[9] = static_cast<float>(n_all);
  }
}


// Stage-2 training branch (reference module2_mixed/my_models.py:363-461) on the stage-2 forward's buffers: every
// proposal is an image proposal with a 13-entry refinement vector [conf, 12 class scores].  One block.
//   out[0..4] = masks (focal), conf, xy, wh, category;  out[5] = masks + (conf + category) / l0 + (xy + wh) / l1 (:445)
//   out[6..9] = true (labels > iou_hi), refined (mask > thr), tp, rows
__global__ void __launch_bounds__(kBlock)
stage2_loss_kernel(const float* __restrict__ boxes, int box_pitch, const float* __restrict__ rois,
                   const float* __restrict__ refine, int refine_pitch, const float* __restrict__ regress,
                   const float* __restrict__ mask, const int* __restrict__ counts, int cap,
                   const float* __restrict__ iou_labels, const float* __restrict__ target_location,
                   const unsigned char* __restrict__ sample_filter, me_stage2_loss_cfg cfg, int* __restrict__ pos_list,
                   float* __restrict__ out) {
  __shared__ double s_red[kBlock / 32];
  __shared__ int s_npos;
  const int n_all = min(counts[1], cap);
  const int nc = refine_pitch - 1;
  // ordered list of the positive rows: class_label row i takes the class of the i-th positive (:440-441)
  if (threadIdx.x < 32) {
    int base = 0;
    for (int r0 = 0; r0 < n_all; r0 += 32) {
      const int r = r0 + threadIdx.x;
      const bool pos = r < n_all && iou_labels[r] > cfg.iou_hi;
      const unsigned int b = __ballot_sync(0xffffffffu, pos);
      if (pos) pos_list[base + __popc(b & ((1u << threadIdx.x) - 1))] = r;
      base += __popc(b);
    }
    if (threadIdx.x == 0) s_npos = base;
  }
  __syncthreads();
  const int P = s_npos;
  double focal = 0.0, conf = 0.0, lxy = 0.0, lwh = 0.0, cat = 0.0, positive = 0.0, tp = 0.0;
  for (int r = threadIdx.x; r < n_all; r += kBlock) {
    const bool pos = iou_labels[r] > cfg.iou_hi;
    const bool sel = sample_filter[r] != 0;
    const float m = mask[r];
    const bool refined = m > cfg.thr;
    positive += refined ? 1.0 : 0.0;
    tp += (refined && pos) ? 1.0 : 0.0;
    if (sel) {
      const float p = pos ? m : 1.f - m;   // masks = softmax output [1 - m, m]
      const float a = pos ? cfg.alpha : 1.f - cfg.alpha;
      const float q = 1.f - p;
      focal += static_cast<double>(-a * (q * q) * logf(p));
      const float x = refine[static_cast<size_t>(r) * refine_pitch];
      const float y = pos ? 1.f : 0.f;
      conf += static_cast<double>(-(y * fmaxf(logf(x), -100.f) + (1.f - y) * fmaxf(logf(1.f - x), -100.f)));
    }
    if (pos) {
      const float x1 = rois[r * 5 + 1], y1 = rois[r * 5 + 2], x2 = rois[r * 5 + 3], y2 = rois[r * 5 + 4];
      const float* g = target_location + r * 4;
      const float x = (x1 + x2) / 2.f, y = (y1 + y2) / 2.f, w = x2 - x1, h = y2 - y1;
      const float xt = (g[0] + g[2]) / 2.f, yt = (g[1] + g[3]) / 2.f, wt = g[2] - g[0], ht = g[3] - g[1];
      const float* q = regress + r * 4;
      lxy += static_cast<double>(smooth_l1((xt - x) / (w + 1e-16f), q[0]) + smooth_l1((yt - y) / (h + 1e-16f), q[1]));
      lwh += static_cast<double>(smooth_l1(logf(wt / w + 1e-16f), q[2]) + smooth_l1(logf(ht / h + 1e-16f), q[3]));
      // class_label[r] is non-zero only for r < P: a one at the class of the r-th positive row
      const int hot = r < P ? static_cast<int>(boxes[static_cast<size_t>(pos_list[r]) * box_pitch + 7]) : -1;
      for (int c = 0; c < nc; ++c) {
        const float xc = refine[static_cast<size_t>(r) * refine_pitch + 1 + c];
        const float yc = c == hot ? 1.f : 0.f;
        cat += static_cast<double>(-(yc * fmaxf(logf(xc), -100.f) + (1.f - yc) * fmaxf(logf(1.f - xc), -100.f)));
      }
    }
  }
  focal = block_sum(focal, s_red);
  conf = block_sum(conf, s_red);
  lxy = block_sum(lxy, s_red);
  lwh = block_sum(lwh, s_red);
  cat = block_sum(cat, s_red);
  positive = block_sum(positive, s_red);
  tp = block_sum(tp, s_red);
  if (threadIdx.x == 0) {
    out[0] = static_cast<float>(focal);
    out[1] = static_cast<float>(conf);
    out[2] = static_cast<float>(lxy);
    out[3] = static_cast<float>(lwh);
    out[4] = static_cast<float>(cat);
    out[5] = out[0] + (out[1] + out[4]) / cfg.lambda0 + (out[2] + out[3]) / cfg.lambda1;
    out[6] = static_cast<float>(P);
    out[7] = static_cast<float>(positive);
    out[8] = static_cast<float>(tp);
    out[9] = static_cast<float>(n_all);
  }
}

}  // namespace
}  // namespace me

extern "C" {

int me_stage3_labels(const float* img_boxes, int box_pitch, const float* rois, const int* counts, int cap,
                     const float* targets_xyxy, int num_targets, float* iou_labels, float* target_location,
                     me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(img_boxes && rois && counts && iou_labels && target_location, "stage3_labels: null argument");
  ME_REQUIRE(num_targets == 0 || targets_xyxy, "stage3_labels: %d targets but no target array", num_targets);
  ME_REQUIRE(cap > 0 && box_pitch >= 8 && num_targets >= 0, "stage3_labels: bad cap/box_pitch/num_targets");
  stage3_labels_kernel<<<ceil_div(cap, 256), 256, 0, stream>>>(img_boxes, box_pitch, rois, counts, cap, targets_xyxy,
                                                              num_targets, iou_labels, target_location);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_stage3_loss(const float* rois, const float* refine, const float* regress, const float* mask, const int* counts,
                   int cap, const float* iou_labels, const float* target_location, const unsigned char* sample_filter,
                   const me_stage3_loss_cfg* cfg, float* out10, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(rois && refine && regress && mask && counts && iou_labels && target_location && sample_filter && cfg && out10,
             "stage3_loss: null argument");
  ME_REQUIRE(cap > 0, "stage3_loss: cap %d", cap);
  stage3_loss_kernel<<<1, kBlock, 0, stream>>>(rois, refine, regress, mask, counts, cap, iou_labels, target_location,
                                               sample_filter, *cfg, out10);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

int me_stage2_loss(const float* boxes, int box_pitch, const float* rois, const float* refine, int refine_pitch,
                   const float* regress, const float* mask, const int* counts, int cap, const float* iou_labels,
                   const float* target_location, const unsigned char* sample_filter, const me_stage2_loss_cfg* cfg,
                   int* pos_ws, float* out10, me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(boxes && rois && refine && regress && mask && counts && iou_labels && target_location && sample_filter && cfg &&
                 pos_ws && out10, "stage2_loss: null argument");
  ME_REQUIRE(cap > 0 && box_pitch >= 8 && refine_pitch >= 2, "stage2_loss: bad cap / pitches");
  stage2_loss_kernel<<<1, kBlock, 0, stream>>>(boxes, box_pitch, rois, refine, refine_pitch, regress, mask, counts, cap,
                                               iou_labels, target_location, sample_filter, *cfg, pos_ws, out10);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
