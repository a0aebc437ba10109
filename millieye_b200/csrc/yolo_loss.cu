// YOLO training loss of one detection layer: YOLOLayer.forward with targets (reference yolov3/models.py:180-232) and
// build_targets (utils/utils.py:381-440), on the fp32 head logits the detector forward leaves on the device.
//
//   targets  : one thread per target - the anchor whose SHAPE fits the box best owns it at the cell of its centre
//              (utils.py:403-418).  Several targets on one (image, anchor, cell): PyTorch's indexed assignment keeps the
//              LAST one for tx / ty / tw / th / class_mask / iou_scores (atomicMax on the target index) but sets the
//              one-hot class target of ALL of them (atomicOr on a class bit mask); every anchor whose shape IoU exceeds
//              ignore_thres is taken out of the no-object set at that cell (:421-422).
//   cells    : one thread per (image, anchor, gy, gx) - decode, squared errors / BCE terms (log clamped at -100 like
//              torch's binary_cross_entropy), metric counters; block reduction, then double atomics into 16 sums.
//   finalize : means, loss = x + y + w + h + obj_scale * conf_obj + noobj_scale * conf_noobj + cls, the 13 metrics.
// Forward value only: the reference's autograd through the detector is not part of the accelerated path (no reference
// script trains stage 1).
#include "common.cuh"

namespace me {
namespace {

constexpr int kMaxAnchors = 8;
constexpr int kClassWords = 4;   // up to 128 classes

struct YoloLossParams {
  int n, g, na, nc, pitch;
  float aw[kMaxAnchors], ah[kMaxAnchors];   // anchors / stride, rounded to fp32 like models.py:127
  float ignore_thres;
};

enum { S_X, S_Y, S_W, S_H, S_NOBJ, S_BCE_OBJ, S_BCE_NOOBJ, S_NNOOBJ, S_CLS, S_CMASK, S_CONF50, S_DET50, S_DET75, S_CONF_OBJ,
       S_CONF_NOOBJ, S_COUNT };

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float bce_term(float p, float t) {
  const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
  return -(t * lp + (1.f - t) * lq);
}

__global__ void yolo_targets_kernel(const float* __restrict__ targets, int m, YoloLossParams P, int* __restrict__ owner,
                                    unsigned char* __restrict__ ignore, unsigned int* __restrict__ cls_bits) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const float* tg = targets + t * 6;
  const int b = static_cast<int>(tg[0]), label = static_cast<int>(tg[1]);
  const float gx = tg[2] * P.g, gy = tg[3] * P.g, gw = tg[4] * P.g, gh = tg[5] * P.g;
  if (b < 0 || b >= P.n) return;
  const int gi = static_cast<int>(gx), gj = static_cast<int>(gy);   // .long(): truncation
  if (gi < 0 || gi >= P.g || gj < 0 || gj >= P.g) return;
  int best = 0;
  float best_iou = -1.f;
  for (int a = 0; a < P.na; ++a) {
    const float inter = fminf(P.aw[a], gw) * fminf(P.ah[a], gh);
    const float iou = inter / ((P.aw[a] * P.ah[a] + 1e-16f) + gw * gh - inter);
    if (iou > best_iou) {   // torch.max: first maximum
      best_iou = iou;
      best = a;
    }
    if (iou > P.ignore_thres) ignore[((b * P.na + a) * P.g + gj) * P.g + gi] = 1;
  }
  const int cell = ((b * P.na + best) * P.g + gj) * P.g + gi;
  atomicMax(owner + cell, t);
  if (label >= 0 && label < 32 * kClassWords) atomicOr(cls_bits + cell * kClassWords + (label >> 5), 1u << (label & 31));
}

__global__ void __launch_bounds__(256)
yolo_cells_kernel(const float* __restrict__ logits, const float* __restrict__ targets, YoloLossParams P,
                  const int* __restrict__ owner, const unsigned char* __restrict__ ignore,
                  const unsigned int* __restrict__ cls_bits, double* __restrict__ sums) {
  __shared__ double s_part[S_COUNT][8];
  double acc[S_COUNT];
#pragma unroll
  for (int k = 0; k < S_COUNT; ++k) acc[k] = 0.0;
  const long long total = 1LL * P.n * P.na * P.g * P.g;
  const int attrs = 5 + P.nc;
  for (long long cell = blockIdx.x * 1LL * blockDim.x + threadIdx.x; cell < total; cell += 1LL * gridDim.x * blockDim.x) {
    const int gx = static_cast<int>(cell % P.g), gy = static_cast<int>((cell / P.g) % P.g);
    const int a = static_cast<int>((cell / (P.g * P.g)) % P.na), b = static_cast<int>(cell / (1LL * P.g * P.g * P.na));
    const float* lg = logits + ((1LL * b * P.g + gy) * P.g + gx) * P.pitch + a * attrs;
    const float conf = sigmoid_f(lg[4]);
    if (conf > 0.5f) acc[S_CONF50] += 1.0;
    const int t = owner[cell];
    if (t >= 0) {
      const float* tg = targets + t * 6;
      const float tgx = tg[2] * P.g, tgy = tg[3] * P.g, tgw = tg[4] * P.g, tgh = tg[5] * P.g;
      const float x = sigmoid_f(lg[0]), y = sigmoid_f(lg[1]), w = lg[2], h = lg[3];
      const float tx = tgx - floorf(tgx), ty = tgy - floorf(tgy);
      const float tw = logf(tgw / P.aw[a] + 1e-16f), th = logf(tgh / P.ah[a] + 1e-16f);
      acc[S_X] += static_cast<double>((x - tx) * (x - tx));
      acc[S_Y] += static_cast<double>((y - ty) * (y - ty));
      acc[S_W] += static_cast<double>((w - tw) * (w - tw));
      acc[S_H] += static_cast<double>((h - th) * (h - th));
      acc[S_NOBJ] += 1.0;
      acc[S_BCE_OBJ] += static_cast<double>(bce_term(conf, 1.f));
      acc[S_CONF_OBJ] += static_cast<double>(conf);
      // class terms: BCE against the OR of the labels of every target of this cell; arg-max against the last one's
      const int label = static_cast<int>(tg[1]);
      float best = -1.f;
      int best_c = 0;
      for (int c = 0; c < P.nc; ++c) {
        const float pc = sigmoid_f(lg[5 + c]);
        const bool on = c < 32 * kClassWords && ((cls_bits[cell * kClassWords + (c >> 5)] >> (c & 31)) & 1u);
        acc[S_CLS] += static_cast<double>(bce_term(pc, on ? 1.f : 0.f));
        if (pc > best) {
          best = pc;
          best_c = c;
        }
      }
      const float cmask = best_c == label ? 1.f : 0.f;
      acc[S_CMASK] += cmask;
      // iou_scores: predicted box (grid units) against the target box, cxcywh with the +1 convention (utils.py:249-281)
      const float bx = x + gx, by = y + gy, bw = expf(w) * P.aw[a], bh = expf(h) * P.ah[a];
      const float b1x1 = bx - bw / 2, b1x2 = bx + bw / 2, b1y1 = by - bh / 2, b1y2 = by + bh / 2;
      const float b2x1 = tgx - tgw / 2, b2x2 = tgx + tgw / 2, b2y1 = tgy - tgh / 2, b2y2 = tgy + tgh / 2;
      const float iw = fmaxf(fminf(b1x2, b2x2) - fmaxf(b1x1, b2x1) + 1.f, 0.f);
      const float ih = fmaxf(fminf(b1y2, b2y2) - fmaxf(b1y1, b2y1) + 1.f, 0.f);
      const float inter = iw * ih;
      const float a1 = (b1x2 - b1x1 + 1.f) * (b1y2 - b1y1 + 1.f), a2 = (b2x2 - b2x1 + 1.f) * (b2y2 - b2y1 + 1.f);
      const float iou = inter / (a1 + a2 - inter + 1e-16f);
      const float detected = (conf > 0.5f ? 1.f : 0.f) * cmask;
      if (iou > 0.5f) acc[S_DET50] += detected;
      if (iou > 0.75f) acc[S_DET75] += detected;
    } else if (!ignore[cell]) {
      acc[S_NNOOBJ] += 1.0;
      acc[S_BCE_NOOBJ] += static_cast<double>(bce_term(conf, 0.f));
      acc[S_CONF_NOOBJ] += static_cast<double>(conf);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < S_COUNT; ++k) {
    double v = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_part[k][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < S_COUNT) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_part[threadIdx.x][w];
    if (v != 0.0) atomicAdd(sums + threadIdx.x, v);
  }
}

// out[14] = loss, x, y, w, h, conf, cls, cls_acc, recall50, recall75, precision, conf_obj, conf_noobj, grid_size
__global__ void yolo_finalize_kernel(const double* __restrict__ s, int nc, int g, float obj_scale, float noobj_scale,
                                     float* __restrict__ out) {
  const double nobj = s[S_NOBJ], nnoobj = s[S_NNOOBJ];
  // the mean of an empty selection is NaN in torch (0 / 0): keep that behaviour
  const float lx = static_cast<float>(s[S_X] / nobj), ly = static_cast<float>(s[S_Y] / nobj);
  const float lw = static_cast<float>(s[S_W] / nobj), lh = static_cast<float>(s[S_H] / nobj);
  const float conf_obj = static_cast<float>(s[S_BCE_OBJ] / nobj), conf_noobj = static_cast<float>(s[S_BCE_NOOBJ] / nnoobj);
  const float lconf = obj_scale * conf_obj + noobj_scale * conf_noobj;
  const float lcls = static_cast<float>(s[S_CLS] / (nobj * nc));
  out[0] = lx + ly + lw + lh + lconf + lcls;
  out[1] = lx;
  out[2] = ly;
  out[3] = lw;
  out[4] = lh;
  out[5] = lconf;
  out[6] = lcls;
  out[7] = static_cast<float>(100.0 * s[S_CMASK] / nobj);
  out[8] = static_cast<float>(s[S_DET50] / (nobj + 1e-16));
  out[9] = static_cast<float>(s[S_DET75] / (nobj + 1e-16));
  out[10] = static_cast<float>(s[S_DET50] / (s[S_CONF50] + 1e-16));
  out[11] = static_cast<float>(s[S_CONF_OBJ] / nobj);
  out[12] = static_cast<float>(s[S_CONF_NOOBJ] / nnoobj);
  out[13] = static_cast<float>(g);
}

}  // namespace
}  // namespace me

extern "C" {

size_t me_yolo_loss_workspace(int n, int g, int num_anchors) {
  const size_t cells = static_cast<size_t>(n) * num_anchors * g * g;
  // owner (int) + class bits (4 words) + ignore flags, then 16 double sums (256-byte aligned)
  size_t bytes = cells * 4 + cells * 4 * me::kClassWords + cells;
  bytes = (bytes + 255) & ~size_t(255);
  return bytes + 256;
}

int me_yolo_loss(const float* logits, int pitch, int n, int g, int num_anchors, int num_classes,
                 const float* host_anchors_wh, float stride, const float* targets, int num_targets, float ignore_thres,
                 float obj_scale, float noobj_scale, void* workspace, size_t workspace_bytes, float* out14,
                 me_stream_t stream_) {
  using namespace me;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ME_REQUIRE(logits && host_anchors_wh && workspace && out14, "yolo_loss: null argument");
  ME_REQUIRE(n > 0 && g > 0 && num_anchors >= 1 && num_anchors <= kMaxAnchors, "yolo_loss: bad shape");
  ME_REQUIRE(num_classes >= 1 && num_classes <= 32 * kClassWords, "yolo_loss: 1..%d classes", 32 * kClassWords);
  ME_REQUIRE(pitch >= num_anchors * (5 + num_classes), "yolo_loss: pitch %d < head channels", pitch);
  ME_REQUIRE(num_targets == 0 || targets, "yolo_loss: null targets");
  ME_REQUIRE(workspace_bytes >= me_yolo_loss_workspace(n, g, num_anchors), "yolo_loss: workspace too small");
  YoloLossParams P{};
  P.n = n;
  P.g = g;
  P.na = num_anchors;
  P.nc = num_classes;
  P.pitch = pitch;
  P.ignore_thres = ignore_thres;
  for (int a = 0; a < num_anchors; ++a) {
    P.aw[a] = static_cast<float>(static_cast<double>(host_anchors_wh[2 * a]) / static_cast<double>(stride));
    P.ah[a] = static_cast<float>(static_cast<double>(host_anchors_wh[2 * a + 1]) / static_cast<double>(stride));
  }
  const size_t cells = static_cast<size_t>(n) * num_anchors * g * g;
  unsigned char* base = static_cast<unsigned char*>(workspace);
  int* owner = reinterpret_cast<int*>(base);
  unsigned int* cls_bits = reinterpret_cast<unsigned int*>(base + cells * 4);
  unsigned char* ignore = base + cells * 4 + cells * 4 * kClassWords;
  double* sums = reinterpret_cast<double*>(base + ((cells * 4 + cells * 4 * kClassWords + cells + 255) & ~size_t(255)));
  ME_CUDA(cudaMemsetAsync(owner, 0xff, cells * 4, stream));                               // -1: no target
  ME_CUDA(cudaMemsetAsync(cls_bits, 0, cells * 4 * kClassWords + cells, stream));          // class bits + ignore flags
  ME_CUDA(cudaMemsetAsync(sums, 0, 16 * sizeof(double), stream));
  if (num_targets > 0)
    yolo_targets_kernel<<<(num_targets + 127) / 128, 128, 0, stream>>>(targets, num_targets, P, owner, ignore, cls_bits);
  long long blocks = (static_cast<long long>(cells) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  yolo_cells_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(logits, targets, P, owner, ignore, cls_bits, sums);
  yolo_finalize_kernel<<<1, 1, 0, stream>>>(sums, num_classes, g, obj_scale, noobj_scale, out14);
  ME_LAUNCH_CHECK();
  return ME_OK;
}

}  // extern "C"
