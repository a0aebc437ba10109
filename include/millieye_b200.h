/*
 * millieye_b200 — C-ABI of the B200 (sm_100a) detection-and-fusion hot path.
 *
 * Every entry point replaces one stretch of the reference's PyTorch forward
 * (sxontheway/milliEye, paths relative to /root/reference/module3_our_dataset):
 * the reference has no FFI of its own (it is pure Python calling torch / torchvision),
 * so each function cites the Python lines whose arithmetic it takes over.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless
 *    a parameter is named host_*; the library never allocates or frees on the hot path
 *    (workspaces are passed in) and never synchronises: work is enqueued on `stream`
 *    (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *  - activations are NHWC fp16 with a row pitch (elements between consecutive pixels),
 *    so route/concat is a pitch + channel offset, not a copy.
 *  - return value: ME_OK or an ME_ERR_* code; me_last_error() gives the text.
 */
#ifndef MILLIEYE_B200_H_
#define MILLIEYE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ME_OK 0
#define ME_ERR_ARG 1
#define ME_ERR_CUDA 2
#define ME_ERR_UNSUPPORTED 3

#define ME_ACT_LINEAR 0
#define ME_ACT_LEAKY 1   /* LeakyReLU(0.1): yolov3/models.py:40-41 */
#define ME_ACT_SIGMOID 2 /* my_models.py:153-157 (radar score map) */

typedef void* me_stream_t; /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------- */
int me_version(void);
/* Copies the calling thread's last error text into buf (NUL-terminated); returns its length. */
int me_last_error(char* buf, size_t n);
/* SM count / compute capability of the current device (fails loudly off sm_100). */
int me_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Debug word written by a kernel whose pipeline wait timed out (0 = never). */
int me_debug_status(unsigned long long* host_word);
/* Profiling aid (tools/conv_trace.py): while dev_words != NULL every me_conv_gemm launch writes 16
 * clock64 words per CTA (phase time stamps and barrier-wait totals) to dev_words[16 * blockIdx.x ...];
 * NULL switches it off.  Not part of the reference's interface. */
int me_conv_set_trace(unsigned long long* dev_words);

/* ---- conv + BN + activation (A2/A8/A9 of SURVEY §8a) ------------------------------ */
/* K-block (channels per pipeline stage) the GEMM uses for `cin` input channels; the packed
 * weight's per-tap channel count is cin rounded up to it. */
int me_conv_k_block(int cin);
int me_conv_cin_pad(int cin);

/* Folds eval-mode BatchNorm into the conv and packs OIHW fp32 weights into the kernel's
 * K-major fp16 layout  w_packed[cout_pad][ksize*ksize*cin_pad]  (tap-major, then channel);
 * bias_out[cout_pad] fp32 = beta - mean*gamma/sqrt(var+eps) (+ conv bias).
 * Replaces nn.Conv2d + nn.BatchNorm2d parameter use at yolov3/models.py:27-39 and
 * my_models.py:62-73,133-150.  Any of conv_bias / bn_* may be NULL. */
int me_pack_conv_weights(const float* w_oihw, const float* conv_bias, const float* bn_gamma,
                         const float* bn_beta, const float* bn_mean, const float* bn_var, float bn_eps,
                         int cout, int cin, int ksize, int cout_pad, void* w_packed, float* bias_out,
                         me_stream_t stream);

typedef struct me_conv_desc {
  int n, h, w;       /* input batch / height / width                                      */
  int cin;           /* input channels read (view width)                                  */
  int in_pitch;      /* elements between consecutive input pixels (>= cin, multiple of 8) */
  int cout;          /* output channels written (multiple of 16; padded channels get 0+bias) */
  int out_pitch;     /* elements between consecutive output pixels                        */
  int ksize;         /* 1 or 3; padding is (ksize-1)/2 as in models.py:25                 */
  int stride;        /* 1 or 2                                                            */
  int act;           /* ME_ACT_*                                                          */
  int out_f32;       /* 0: fp16 output, 1: fp32 output (YOLO head logits)                 */
  int res_pitch;     /* >0: add residual[pixel][cout] (fp16, this pitch) AFTER the activation
                        (shortcut layer, models.py:258-260); 0: none                      */
} me_conv_desc;

/* y = act(conv(x, w) + bias) (+ residual).  Implicit GEMM on tcgen05 tensor cores: TMA
 * (im2col mode for 3x3, tiled for 1x1) -> swizzled smem -> tcgen05.mma -> TMEM -> epilogue
 * -> TMA store.  Replaces the convolutional branch of Darknet.forward (models.py:252-253),
 * cnn_layers_1 (my_models.py:76-77), cnn_layers_3 (my_models.py:152-157) and, with
 * ksize=1,h=w=1, the Linear 490->256 of refinement_head.net0 (my_models.py:263). */
int me_conv_gemm(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                 const void* residual, void* y, me_stream_t stream);

/* me_conv_gemm for a thin 3x3 / stride-1 layer FOLLOWED BY MaxPool2d(2, 2) (blocks 2-3, 4-5 of yolov3-tiny*.cfg,
 * models.py:22-51) in one kernel: the pool runs on the epilogue's registers, y is the pooled (n, h/2, w/2, out_pitch)
 * tensor and the full-resolution activation is never written.  Bit-identical to me_conv_gemm + me_maxpool2.
 * me_conv_pool_supported: 1 if the layer qualifies (cin 16 / 32, cout 32 / 64, no residual, fp16 output, w % 8 == 0,
 * h even), else 0 and me_conv_pool returns ME_ERR_UNSUPPORTED. */
int me_conv_pool_supported(const me_conv_desc* d);
int me_conv_pool(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y_pooled,
                 me_stream_t stream);

/* me_conv_gemm with a workspace for the split-K tail: layers whose last wave of 256 x 256 tiles would leave most SM pairs
 * idle cut those tiles along K; partial sums and arrival counters live in `workspace` (zero-filled by the caller once,
 * me_conv_workspace_bytes() long, 256-byte aligned; NULL: no split).  The workspace belongs to the call, not to the
 * library: launches that may overlap (different streams, different host threads) simply pass different buffers - there is
 * no process-global state.  Results are deterministic and equal to me_conv_gemm's up to fp32 summation order. */
size_t me_conv_workspace_bytes(void);
int me_conv_gemm_ws(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, const void* residual,
                    void* y, void* workspace, size_t workspace_bytes, me_stream_t stream);

/* ---- a run of conv layers as one persistent kernel --------------------------------------------------------------
 * Darknet.forward walks its module list one layer at a time (yolov3/models.py:247-262); launched that way a third of
 * the conv time of Darknet-53 was launch gaps, pipeline fill / drain and partial last waves.  me_conv_chain_* runs a
 * list of layers - each described exactly like a me_conv_gemm call - as ONE persistent kernel of CTA pairs in which a
 * 256 x BN tile of layer l starts as soon as the tiles of the layers it reads (dep_layer: producer of x, res_layer:
 * producer of residual; -1 = complete before the launch) have been stored.  Results are those of the per-layer calls.
 *   eligible:  1x1 / 3x3 stride 1 / 3x3 stride 2, cin % 64 == 0, cout % 128 == 0, fp16 output.
 *   build:     fills a caller-owned HOST blob (me_conv_chain_blob_bytes long, 128-byte aligned) with the layer table
 *              (tensor maps), the per-pair work lists and the counter area; the caller copies it to a 256-byte aligned
 *              DEVICE buffer of the same size once.
 *   run:       enqueues the counter reset + the kernel on `stream` (graph capturable, no allocation, re-entrant per
 *              device blob).  Layers must appear in execution order. */
typedef struct me_chain_layer {
  me_conv_desc d;
  const void* x;
  const void* w_packed;
  const float* bias;
  const void* residual;
  void* y;
  int dep_layer;
  int res_layer;
} me_chain_layer;
int me_conv_chain_eligible(const me_conv_desc* d);
size_t me_conv_chain_blob_bytes(const me_chain_layer* layers, int n_layers);
int me_conv_chain_build(const me_chain_layer* layers, int n_layers, void* host_blob, size_t blob_bytes);
int me_conv_chain_run(const void* host_blob, void* dev_blob, me_stream_t stream);
/* me_conv_chain_verify: replays a blob's work lists under the rules the kernel waits by (every pair takes its items in
 * order; an item runs once the m tiles it reads and its residual tile are complete) and returns ME_ERR_ARG with the
 * blocked item if the schedule cannot complete, lists a tile twice or not at all; me_conv_chain_build runs it on every
 * blob it writes.  me_conv_chain_plan: me_conv_chain_build without tensor maps - needs no CUDA driver (tests,
 * inspection of the schedule); me_conv_chain_run refuses such a blob. */
int me_conv_chain_plan(const me_chain_layer* layers, int n_layers, void* host_blob, size_t blob_bytes);
int me_conv_chain_verify(const void* host_blob);

/* YOLO head: the linear 1x1 head conv (models.py:252, blocks followed by a [yolo] block) with YOLOLayer.forward's
 * decode (models.py:142-177, see me_yolo_decode) fused into its epilogue: the fp32 logits never go to memory, the
 * decoded rows are written straight into pred [n][rows_total][5+C] at row_offset.  d->out_f32 must be 1, no
 * residual, linear activation; d->out_pitch is ignored.  Same results as me_conv_gemm + me_yolo_decode, bit for bit. */
int me_conv_gemm_yolo(const me_conv_desc* d, const void* x, const void* w_packed, const float* bias, int g,
                      int num_anchors, int num_classes, const float* host_anchors_wh, float stride, int rows_total,
                      int row_offset, float* pred, me_stream_t stream);

/* First layer: 3x3 / stride 1 / pad 1 conv straight from the caller's NCHW fp32 image
 * (cin <= 4) to NHWC fp16, BN folded, activation applied.  w_first is fp32 [cout][cin][3][3]
 * already BN-folded (me_fold_first_weights).  models.py:252 for module 0, my_models.py:133. */
int me_fold_first_weights(const float* w_oihw, const float* conv_bias, const float* bn_gamma,
                          const float* bn_beta, const float* bn_mean, const float* bn_var, float bn_eps,
                          int cout, int cin, float* w_folded, float* bias_out, me_stream_t stream);
int me_conv_first(const float* x_nchw, const float* w_folded, const float* bias, void* y_nhwc, int n, int h,
                  int w, int cin, int cout, int out_pitch, int act, me_stream_t stream);
/* Same layer on tcgen05: producer warps im2col the fp32 image into a 128x32 fp16 tile in shared memory
 * (K = 9*cin padded to 32), two tcgen05.mma per tile, fused bias/activation epilogue.  cin <= 3,
 * cout in {16, 32, 64}; wk_scratch: fp16 [cout][32] device buffer owned by the caller. */
int me_conv_first_tc(const float* x_nchw, const float* w_folded, const float* bias, void* wk_scratch,
                     void* y_nhwc, int n, int h, int w, int cin, int cout, int out_pitch, int act,
                     me_stream_t stream);
/* me_conv_first_tc followed by MaxPool2d(2, 2) (blocks 0-1 of yolov3-tiny*.cfg, models.py:22-51) in one kernel: the
 * pool runs in the epilogue, y_nhwc is the pooled (n, h/2, w/2, out_pitch) tensor and the full-resolution activation is
 * never written.  Bit-identical to me_conv_first_tc + me_maxpool2.  Needs w % 32 == 0, h % 4 == 0, a 16-byte aligned
 * image and cout in {16, 32}; anything else is ME_ERR_UNSUPPORTED / ME_ERR_ARG (run the two calls instead). */
int me_conv_first_tc_pool(const float* x_nchw, const float* w_folded, const float* bias, void* wk_scratch,
                          void* y_nhwc, int n, int h, int w, int cin, int cout, int out_pitch, int act,
                          me_stream_t stream);

/* ---- glue layers (A3) -------------------------------------------------------------- */
/* MaxPool2d(2, stride) on NHWC fp16; stride 1 uses the right/bottom zero pad of
 * models.py:46-49 (ZeroPad2d((0,1,0,1)) then pool). */
int me_maxpool2(const void* x, void* y, int n, int h, int w, int c, int in_pitch, int out_pitch, int stride,
                me_stream_t stream);
/* Nearest x2 upsample (models.py:82-92) written into a channel slice of a concat buffer. */
int me_upsample2(const void* x, void* y, int n, int h, int w, int c, int in_pitch, int out_pitch,
                 me_stream_t stream);
/* Strided channel-slice copy (route of a tensor that could not be produced in place). */
int me_copy_channels(const void* x, void* y, long long pixels, int c, int in_pitch, int out_pitch,
                     me_stream_t stream);
/* NHWC fp16 (pitch) -> NCHW fp32 contiguous: the featuremap handed back to PyTorch callers
 * (models.py:254-255). */
int me_nhwc_to_nchw_f32(const void* x, float* y, int n, int h, int w, int c, int in_pitch, me_stream_t stream);
/* uint8 image bytes -> fp32 in 0..1 (x / 255): torchvision ToTensor, which the reference applies on the HOST before the
 * upload (utils/datasets.py:209; run_sp.py:205-211 for camera frames).  Doing it here lets callers upload bytes: a quarter
 * of the host -> device traffic.  Same layout in and out; both buffers 16-byte aligned. */
int me_u8_to_unit_f32(const void* x, float* y, long long count, me_stream_t stream);
/* NCHW fp32 -> NHWC fp16 (pitch, zero padded channels). */
int me_nchw_f32_to_nhwc(const float* x, void* y, int n, int h, int w, int c, int out_pitch, me_stream_t stream);

/* ---- YOLO decode (A4) -------------------------------------------------------------- */
/* logits: NHWC fp32 [n][g][g][pitch] holding A*(5+C) head channels; out: fp32
 * [n][rows_total][5+C], this head's rows start at row_offset, ordered anchor-major, then gy,
 * then gx (YOLOLayer.forward, models.py:142-177).  anchors_wh: host array of 2*A floats in
 * pixels.  stride = img_dim / g. */
int me_yolo_decode(const float* logits, int pitch, float* out, int n, int g, int num_anchors, int num_classes,
                   const float* host_anchors_wh, float stride, int rows_total, int row_offset,
                   me_stream_t stream);

/* ---- confidence filter + class arg-max + per-class NMS (A6) ------------------------ */
/* Bytes of workspace me_filter_nms needs for a (n, rows, 5+C) prediction tensor. */
size_t me_filter_nms_workspace(int n, int rows, int num_classes);
/* pred: fp32 [n][rows][5+C] rows [cx,cy,w,h,conf,cls...].  Reproduces
 * non_max_suppression_cpp (utils/utils.py:337-378) incl. torchvision.batched_nms's two code
 * paths (coordinate trick for <= 1000 candidates, per-class otherwise).  Output:
 * det [n][max_det][7+C] = [x1,y1,x2,y2,conf,class_conf,class_pred,cls...], det_count[n],
 * det_index [n][max_det] = source row of every survivor (score-descending).
 * If xyxy_inplace != 0 the first four columns of pred are rewritten to x1y1x2y2 as the
 * reference does (utils.py:354).  nms_thresh is a double because torchvision compares the fp32
 * IoU against the Python float as a double. */
int me_filter_nms(float* pred, int n, int rows, int num_classes, float conf_thresh, double nms_thresh,
                  int max_det, int xyxy_inplace, float* det, int* det_count, int* det_index, void* workspace,
                  size_t workspace_bytes, me_stream_t stream);

/* ---- RoI gathers (A10/A11) ---------------------------------------------------------- */
/* rois: fp32 [cap][5] = [batch_idx,x1,y1,x2,y2]; only the first *roi_count rows are live.
 * feat: NHWC fp16 [n][h][w][pitch].  Output fp16 [cap][out_pitch] with element
 * c*pooled*pooled + ph*pooled + pw (the flatten order of my_models.py:262), zero padded. */
int me_psroi_align(const void* feat, int n, int h, int w, int pitch, int out_channels, int pooled,
                   float spatial_scale, const float* rois, const int* roi_count, int cap, void* out,
                   int out_pitch, me_stream_t stream);
int me_roi_align(const void* feat, int n, int h, int w, int pitch, int channels, int pooled, float spatial_scale,
                 const float* rois, const int* roi_count, int cap, void* out, int out_pitch,
                 me_stream_t stream);
/* The same two gathers in BIN-MAJOR order: output element (ph*pooled + pw)*channels + c, and for the position-sensitive
 * variant score-map channel (ph*pooled + pw)*channels + c.  Same values as me_psroi_align / me_roi_align under that
 * fixed permutation (the caller permutes the producing conv's output channels and the consuming layer's input columns
 * once when packing weights); consecutive threads read consecutive channels of one pixel neighbourhood, so one
 * 32-byte sector serves a bin's channels instead of one sector per 2-byte element. */
int me_roi_gather_bin_major(const void* feat, int n, int h, int w, int pitch, int channels, int pooled,
                            float spatial_scale, const float* rois, const int* roi_count, int cap, void* out,
                            int out_pitch, int position_sensitive, me_stream_t stream);

/* ---- proposal assembly, heads, output (A7, A12-A14) -------------------------------- */
typedef struct me_head_weights {
  /* refinement_head (my_models.py:213-258), fp32 device pointers, row-major [out][in] */
  const float* net1_w; const float* net1_b;   /* 4 x 256  box regression          */
  const float* net2_w; const float* net2_b;   /* 13 x 256 class vector (sigmoid)  */
  const float* radar_w; const float* radar_b; /* 10 x 490 = 7x7 conv, BN folded   */
  const float* radar2_w; const float* radar2_b; /* 1 x 10 1x1 conv                */
  /* ensemble_head (my_models.py:176-210) */
  const float* fc1_w; const float* fc1_b;     /* 32 x 2  */
  const float* fc2_w; const float* fc2_b;     /* 2 x 64  */
} me_head_weights;

/* Builds the RoI list of Network.forward (my_models.py:459-492): per-image NMS survivors
 * with class_pred == class_idx (image proposals, image-major order) followed by the radar
 * boxes scaled by img_size (class_idx < 0: keep every class, rows of 1 + det_cols floats, see
 * me_stage2_heads).  img_boxes: fp32 [cap][9] = [i,x1,y1,x2,y2,conf,class_conf,
 * class_pred,cls[class_idx]] for the image proposals; rois [cap][5]; counts[0] = image
 * proposals, counts[1] = image + radar proposals. */
int me_build_proposals(const float* det, const int* det_count, int n, int max_det, int det_cols, int class_idx,
                       const float* radar_boxes, int num_radar, float img_size, float* img_boxes, float* rois,
                       int* counts, int cap, me_stream_t stream);
/* The same with the number of radar boxes read from DEVICE memory at run time (*num_radar_dev, clipped to
 * [0, radar_cap]): the launch has no per-call host scalar, so a captured CUDA graph serves every batch
 * (millieye_b200.my_models.FusionPipeline). */
int me_build_proposals_dev(const float* det, const int* det_count, int n, int max_det, int det_cols, int class_idx,
                           const float* radar_boxes, int radar_cap, const int* num_radar_dev, float img_size,
                           float* img_boxes, float* rois, int* counts, int cap, me_stream_t stream);

/* refinement_head tail + ensemble_head + masks (my_models.py:264-284, 498-514): from
 * hidden = leaky(net0(psroi)) [cap][hidden_pitch] fp16 and the radar crop [cap][radar_pitch]
 * fp16, computes regress[cap][4], refine[cap][2] (confidence, class score) and mask[cap]
 * (foreground probability: ensemble column 0 for image proposals, refinement confidence for
 * radar proposals). */
int me_fusion_heads(const void* hidden, int hidden_pitch, const void* radar_crop, int radar_pitch,
                    const me_head_weights* hw, const float* img_boxes, const int* counts, int cap, float* regress,
                    float* refine, float* mask, me_stream_t stream);

/* Stage-2 model (module2_mixed/my_models.py:96-163, 299-361): every class is kept (me_build_proposals with
 * class_idx = -1 writes rows [i, x1,y1,x2,y2, conf, class_conf, class_pred, cls...] of 1 + det_cols floats),
 * no radar branch; from hidden = leaky(net0(psroi)) computes regress[cap][4] and mask[cap] = column 1 of
 * softmax(LeakyReLU(fc2(flatten(leaky(fc1(stack(sigmoid(net2(hidden)), yolo_vector))))))). */
typedef struct me_stage2_weights {
  const float* net1_w; const float* net1_b;   /* 4 x 256                         */
  const float* net2_w; const float* net2_b;   /* num_vec x 256                   */
  const float* fc1_w;  const float* fc1_b;    /* 32 x 2                          */
  const float* fc2_w;  const float* fc2_b;    /* 2 x (32 * num_vec)              */
} me_stage2_weights;
/* refine_out (may be NULL): [cap][num_vec] refinement vectors (confidence, class scores), kept for the loss branch. */
int me_stage2_heads(const void* hidden, int hidden_pitch, const me_stage2_weights* hw, const float* boxes,
                    int box_pitch, int num_vec, const int* counts, int cap, float* regress, float* mask,
                    float* refine_out, me_stream_t stream);

/* Threshold, box regression, priority sort (my_models.py:516-539, box_regress :378-391).
 * out [cap][8] = [i,x1,y1,x2,y2,new_conf,class_score,class_pred] sorted by mask descending with
 * radar masks divided by 5; out_count[0] = rows; box_pitch = floats per img_boxes row (9, or 1 + det_cols).  regress_boxes = 0 skips the regression
 * (model_mode 2).  workspace >= me_finalize_workspace(cap) bytes. */
size_t me_finalize_workspace(int cap);
int me_finalize_output(const float* img_boxes, const float* rois, const float* refine, const float* regress,
                       const float* mask, const int* counts, int cap, int box_pitch, float thr_img, float thr_radar,
                       int regress_boxes, float* out, int* out_count, void* workspace, size_t workspace_bytes,
                       me_stream_t stream);

/* ---- stage-3 labelling and loss (f3; my_models.py:545-640) --------------------- */
/* obtain_iou_labels (my_models.py:317-375, multi_boxes true as Network.forward calls it, :555) over the proposal
 * buffers of the forward: for each of the counts[1] proposals the best IoU (bbox_iou with the +1 pixel convention,
 * utils/utils.py:255-281, first maximum) among the targets of the same image and predicted class, and that target's
 * box; 0 / zeros when no target matches.  targets_xyxy: fp32 [num_targets][6] = [image, class, x1, y1, x2, y2] in
 * pixels (the rewrite of :548-549 is the caller's).  iou_labels [cap], target_location [cap][4]; bit-exact. */
int me_stage3_labels(const float* img_boxes, int box_pitch, const float* rois, const int* counts, int cap,
                     const float* targets_xyxy, int num_targets, float* iou_labels, float* target_location,
                     me_stream_t stream);

typedef struct me_stage3_loss_cfg {
  float iou_hi;       /* positives: iou_label > iou_hi (iou_thresh[1] = 0.7, my_models.py:556)            */
  float alpha;        /* FocalLoss alpha (0.75, :420); gamma is 2                                          */
  float lambda_conf;  /* loss = masks_loss + conf_loss / lambda_conf (loss_lambda[0] = 6, :635)            */
  float thr_img;      /* refine_threshold_img / _radar: positive_masks of the metric (:516-519)            */
  float thr_radar;
} me_stage3_loss_cfg;

/* Losses and counters of my_models.py:586-635 from the forward's buffers, the labels above and the caller's
 * sample_filter (uint8 [cap]: every positive plus the negatives drawn at :600 - the draw stays on the host so that
 * python's `random` state governs it, as in the reference).  out10 (device, fp32):
 *   [0] masks_loss (FocalLoss, sum, image proposals in the sample)   [1] conf_loss (BCE sum over the sample)
 *   [2] loss_xy  [3] loss_wh (SmoothL1 sums over positives)          [4] category_loss (BCE sum over positives)
 *   [5] loss = [0] + [1] / lambda_conf                               [6] positives  [7] positive_masks  [8] true positives
 *   [9] proposals. */
int me_stage3_loss(const float* rois, const float* refine, const float* regress, const float* mask, const int* counts,
                   int cap, const float* iou_labels, const float* target_location, const unsigned char* sample_filter,
                   const me_stage3_loss_cfg* cfg, float* out10, me_stream_t stream);

/* Stage-2 training branch (module2_mixed/my_models.py:363-461) on the stage-2 forward's buffers (labels from
 * me_stage3_labels - obtain_iou_labels is the same routine in both modules; sample_filter drawn on the host as above):
 * FocalLoss over the sampled proposals' masks (:420), confidence BCE (:424-429), regression_loss (:432-435), category BCE
 * over the num_vec - 1 class scores of the positives with the reference's row-i label indexing (:438-443).  out10:
 *   [0] masks  [1] conf  [2] xy  [3] wh  [4] category  [5] loss = [0] + ([1] + [4]) / lambda0 + ([2] + [3]) / lambda1 (:445)
 *   [6] true  [7] refined (mask > thr)  [8] true positives  [9] proposals.   pos_ws: cap ints of scratch. */
typedef struct me_stage2_loss_cfg {
  float iou_hi, alpha, lambda0, lambda1, thr;
} me_stage2_loss_cfg;
int me_stage2_loss(const float* boxes, int box_pitch, const float* rois, const float* refine, int refine_pitch,
                   const float* regress, const float* mask, const int* counts, int cap, const float* iou_labels,
                   const float* target_location, const unsigned char* sample_filter, const me_stage2_loss_cfg* cfg,
                   int* pos_ws, float* out10, me_stream_t stream);

/* ---- radar point cloud -> network input map (f1 + f2) ----------------------- */
typedef struct me_radar_cfg {
  double calib[12];        /* fx, cx, fy, cy, k1, k2, t1, t2, k3, trans_x, trans_y, trans_z
                              (load_calib, data_collection/utils/utils.py:63-76)                        */
  int img_w, img_h;        /* camera frame: FOV filter and histogram range (640 x 480)                  */
  double max_depth;        /* keep depth < max_depth      (prepare_data.py:41,108; run_mp.py:248)       */
  double min_velocity;     /* keep |v| >= min_velocity    (prepare_data.py:39,108)                      */
  int bin_w, bin_h;        /* round(img/scale), scale = max(img_w, img_h)/32 (utils/datasets.py:70-71)  */
  double edges_w[33];      /* np.linspace(0, img_w, bin_w + 1) - histogram2d's bin edges               */
  double edges_h[33];      /* np.linspace(0, img_h, bin_h + 1)                                          */
  int out_size;            /* S/16 side of the map fed to the network (datasets.py:320-322)             */
} me_radar_cfg;

/* points: fp32 [n][cap][4] = radar (x, y, z, velocity), counts[n] live points per frame.
 * Per frame: from_3d_to_2d (float64, int64 truncation) -> FOV/depth/velocity filter -> plot_radar_heatmap
 * (count / mean depth / |mean velocity| histograms, clipped to 0..1) -> pad_to_square -> bilinear resize
 * (align_corners=True).  maps_out: fp32 [n][3][out_size][out_size].  Optional uvzv_out [n][cap][4] +
 * kept_out[n]: the filtered (u, v, depth, velocity) list the reference hands to DBSCAN (prepare_data.py:110). */
int me_radar_maps(const float* points, const int* counts, int n, int cap, const me_radar_cfg* cfg,
                  float* maps_out, float* uvzv_out, int* kept_out, me_stream_t stream);

/* ---- YOLO training loss --------------------------------------------------------------------------
 * YOLOLayer.forward with targets (yolov3/models.py:180-232) + build_targets (utils/utils.py:381-440) for ONE detection
 * layer, on the fp32 head logits [n][g][g][pitch] the forward leaves on the device; targets: device (m,6)
 * [image, class, cx, cy, w, h] in 0..1.  out14 = loss, x, y, w, h, conf, cls, cls_acc, recall50, recall75, precision,
 * conf_obj, conf_noobj, grid_size (the layer's `metrics` dictionary, in the reference's order).  Forward value only. */
size_t me_yolo_loss_workspace(int n, int g, int num_anchors);
int me_yolo_loss(const float* logits, int pitch, int n, int g, int num_anchors, int num_classes,
                 const float* host_anchors_wh, float stride, const float* targets, int num_targets, float ignore_thres,
                 float obj_scale, float noobj_scale, void* workspace, size_t workspace_bytes, float* out14,
                 me_stream_t stream);

/* ---- stage-3 training step (train.py:169-191, my_models.py:545-640 + autograd) -----------
 * fp32 building blocks of the train-mode forward and the hand-derived backward of the heads that train
 * (img_cnn_layers, radar_cnn_layers, refinement_head, ensemble_head); the frozen detector stays on the tensor-core
 * engine.  Matrices are row-major [rows][channels]; every call is stream-ordered and allocation-free.
 * Host side: millieye_b200/stage3_train.py; derivation: oracle/stage3_backward.py. */
/* C[i][j] = sum_k A(i,k) * B(k,j) (+ bias[j], then ME_ACT_*), A(i,k) = A[i*sai + k*sak], B(k,j) = B[k*sbk + j*sbj];
 * accumulate != 0: C += product.  Covers x W^T (nn.Linear / 1x1 conv / im2col conv forward), dZ W (input gradient)
 * and dZ^T x (weight gradient) of my_models.py:62,133-150,238-251 by the choice of strides. */
int me_gemm_f32(int M, int N, int K, const float* A, long long sai, long long sak, const float* B, long long sbk,
                long long sbj, float* C, long long ldc, const float* bias, int act, int accumulate, me_stream_t stream);
/* out[c] = sum_r X[r][c] * (Y ? Y[r][c] : 1), double accumulation (bias gradients, BatchNorm dgamma / dbeta). */
int me_colsum_f32(const float* X, const float* Y, long long rows, int cols, long long ldx, long long ldy, float* out,
                  me_stream_t stream);
/* fp16 rows with a pitch (the detector's NHWC feature map) -> dense fp32 rows; NCHW fp32 -> rows [n*hw][c]. */
int me_half_rows_to_float(const void* x, long long rows, int cols, int pitch, float* y, me_stream_t stream);
int me_nchw_to_rows_f32(const float* x, int n, int c, int hw, float* y, me_stream_t stream);
/* 3x3 / stride 1 / pad 1 im2col on rows: cols[p][ci*9 + tap] (the k order of an OIHW weight row), and its adjoint in
 * gather form (deterministic). */
int me_im2col3_f32(const float* x, int n, int h, int w, int c, float* cols, me_stream_t stream);
int me_col2im3_f32(const float* dcols, int n, int h, int w, int c, float* dx, me_stream_t stream);
/* nn.BatchNorm2d / BatchNorm1d in training mode + LeakyReLU(0.1) (my_models.py:63-75,136-150,248-249), in steps so that
 * the ranks of a sharded batch can add their partial sums in between (synchronised BatchNorm: the sharded step then
 * equals the reference's single-process step on the whole batch):
 *   partial_stats: sums[0..cols) = sum z, sums[cols..2cols) = sum z^2 (double), sums[2 cols] = rows   (2 cols + 1 doubles)
 *   finalize:      batch mean / biased variance -> mean_out, inv_std; running statistics updated in place with
 *                  `momentum` (unbiased variance; NULL: no update); the row count is read from sums[2 cols]
 *   apply:         x_hat = (z - mean) * inv_std, a = leaky(gamma * x_hat + beta)
 *   train_fwd:     the three in a row (single process).
 *   bwd_sums:      da_inout holds dL/da on entry and dL/dy on return; dgamma = sum dy * x_hat, dbeta = sum dy (this rank)
 *   bwd_apply:     dz = inv_std * gamma * (dy - dbeta_total / N - x_hat * dgamma_total / N), N = *total_rows (device). */
int me_bn_partial_stats(const float* z, long long rows, int cols, double* sums, me_stream_t stream);
int me_bn_finalize(const double* sums, int cols, float eps, float momentum, float* running_mean, float* running_var,
                   float* mean_out, float* inv_std, me_stream_t stream);
int me_bn_apply(const float* z, long long rows, int cols, const float* mean, const float* inv_std, const float* gamma,
                const float* beta, float* xhat, float* a, me_stream_t stream);
int me_bn_train_fwd(const float* z, long long rows, int cols, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, double* sums_ws, float* mean_ws, float* inv_std,
                    float* xhat, float* a, me_stream_t stream);
int me_bn_bwd_sums(float* da_inout, const float* a, const float* xhat, long long rows, int cols, float* dgamma, float* dbeta,
                   me_stream_t stream);
int me_bn_bwd_apply(const float* dy, const float* xhat, long long rows, int cols, const float* gamma, const float* inv_std,
                    const float* dgamma_total, const float* dbeta_total, const double* total_rows, float* dz,
                    me_stream_t stream);
int me_leaky_bwd_f32(float* d_inout, const float* a, long long total, me_stream_t stream);
int me_sigmoid_bwd_f32(float* d_inout, const float* s, long long total, me_stream_t stream);
/* torchvision roi_align (aligned=False) / ps_roi_align (sampling_ratio=-1, see me_roi_align / me_psroi_align) on an fp32
 * map [n][h][w][chan_total]; out / grad_out [num_rois][channels*pooled*pooled] in (c, ph, pw) order.  backward != 0:
 * the adjoint - grad_out is scattered into dfeat (zero-filled by the caller) with atomicAdd. */
int me_roi_align_f32(int position_sensitive, int backward, const float* feat, float* dfeat, int n, int h, int w,
                     int chan_total, int channels, int pooled, float scale, const float* rois, int num_rois, float* out,
                     const float* grad_out, me_stream_t stream);
/* Per-proposal tail of refinement_head + ensemble_head in fp32 (my_models.py:268-284, 202-210, 513-514) and the
 * backward of  FocalLoss(masks of the sampled image proposals) + BCE(conf of the sample) / lambda  (:610-635) down to
 * the pre-activations: d_o [n_img][2], hl [n_img][64], dhp [2 n_img][32], u [2 n_img][2] (operands of the ensemble
 * head's weight-gradient products), dr2 [n_all] (radar_net's last pre-activation), dz2 [n_all][13] (net2's). */
int me_stage3_tail_fwd(const float* r2, const float* cls, int cls_pitch, const float* img_boxes, int n_img, int n_all,
                       const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b, float* rc,
                       float* refine, float* mask, float* p, me_stream_t stream);
int me_stage3_tail_bwd(const float* rc, const float* refine, const float* cls, int cls_pitch, const float* p,
                       const float* img_boxes, int n_img, int n_all, const unsigned char* pos, const unsigned char* sel,
                       float alpha, float lambda_conf, const float* fc1_w, const float* fc1_b, const float* fc2_w, float* d_o,
                       float* hl, float* dhp, float* u, float* dr2, float* dz2, me_stream_t stream);
/* torch.optim.Adam step (train.py:158; no weight decay / amsgrad) over flat fp32 buffers; step counts from 1. */
int me_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                 float beta2, float eps, int step, me_stream_t stream);

/* ---- copy-engine collective plumbing (millieye_b200/dist.py::PeerGather) ------------------------------------------ */
/* The detection path has no exchange step (frames are independent); when a caller wants every rank's
 * detections on every GPU, the shards are PUSHED into the peers' buffers (opened through CUDA IPC) with copy engines and
 * completion is signalled / awaited with stream memory operations - no kernel, so nothing competes with the persistent
 * convolution kernels for SMs (an NCCL all-gather kernel has to wait until they leave).
 * me_stream_wait_value32: blocks `stream` until *(uint32*)addr >= value.  me_stream_write_value32: stream-ordered store.
 * me_peer_copy: cudaMemcpyAsync between device allocations of (possibly) different GPUs, ordered on `stream`. */
int me_stream_wait_value32(const void* addr, unsigned int value, me_stream_t stream);
int me_stream_write_value32(void* addr, unsigned int value, me_stream_t stream);
int me_peer_copy(void* dst, const void* src, size_t bytes, me_stream_t stream);
/* cudaDeviceEnablePeerAccess(current device -> peer_device); ME_OK when already enabled or the same device. */
int me_peer_enable(int peer_device);

#ifdef __cplusplus
}
#endif
#endif /* MILLIEYE_B200_H_ */
