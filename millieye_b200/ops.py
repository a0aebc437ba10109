"""torch-tensor front ends of the C-ABI entry points (include/millieye_b200.h).

These are plumbing only: they check devices/dtypes, pass raw device pointers and the current
CUDA stream to the library and raise on any error code.  No arithmetic happens here.
"""
import ctypes
from ctypes import byref, c_float

import torch

from . import _lib
from ._lib import ME_ACT_LEAKY, ME_ACT_LINEAR, ME_ACT_SIGMOID, ConvDesc, HeadWeights, check, ptr, stream_ptr

__all__ = [
    "ME_ACT_LINEAR", "ME_ACT_LEAKY", "ME_ACT_SIGMOID", "round_up", "PackedConv", "pack_conv", "conv_gemm", "conv_gemm_yolo",
    "conv_desc", "conv_chain_eligible", "ConvChain", "FirstConv", "pack_first_conv", "conv_first", "maxpool2", "upsample2", "copy_channels", "nhwc_to_nchw_f32",
    "nchw_f32_to_nhwc", "yolo_decode", "filter_nms", "psroi_align", "roi_align", "build_proposals", "fusion_heads",
    "finalize_output",
]


def round_up(a, b):
    return (a + b - 1) // b * b


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.MeError("millieye_b200 runs on CUDA tensors only (no CPU fallback exists)")


def _f32(t):
    return None if t is None else t.detach().to(dtype=torch.float32).contiguous()


class PackedConv:
    """Device-resident, BN-folded fp16 weights of one conv in the GEMM kernel's layout."""

    def __init__(self, w, bias, cin, cout, cout_pad, ksize):
        self.w, self.bias, self.cin, self.cout, self.cout_pad, self.ksize = w, bias, cin, cout, cout_pad, ksize


def pack_conv(weight, conv_bias=None, bn=None, cout_pad=None):
    """weight: OIHW fp32 (cuda). bn: None or (gamma, beta, running_mean, running_var, eps)."""
    L = _lib.lib()
    weight = _f32(weight)
    _need_cuda(weight)
    cout, cin, kh, kw = weight.shape
    assert kh == kw
    cout_pad = cout_pad or round_up(cout, 32)
    cin_pad = L.me_conv_cin_pad(cin)
    wp = torch.empty((cout_pad, kh * kw * cin_pad), dtype=torch.float16, device=weight.device)
    bias = torch.empty((round_up(cout_pad, 256),), dtype=torch.float32, device=weight.device)
    cb = _f32(conv_bias)
    g = b = m = v = None
    eps = 0.0
    if bn is not None:
        g, b, m, v = (_f32(t) for t in bn[:4])
        eps = float(bn[4])
    check(L.me_pack_conv_weights(ptr(weight), ptr(cb), ptr(g), ptr(b), ptr(m), ptr(v), eps, cout, cin, kh, cout_pad,
                                 ptr(wp), ptr(bias), stream_ptr()), "me_pack_conv_weights")
    return PackedConv(wp, bias, cin, cout, cout_pad, kh)


def conv_gemm(x, packed, n, h, w, in_pitch, out, out_pitch, stride=1, act=ME_ACT_LEAKY, residual=None, res_pitch=0,
              cin=None, cout=None, out_f32=False, workspace=None):
    """x / out / residual are NHWC fp16 buffers (any tensor whose data_ptr is the first pixel of the view).
    workspace: conv_workspace() buffer that lets the pair kernel split its tail tiles along K (per call: launches that may
    overlap pass different buffers)."""
    _need_cuda(x, out, residual, workspace)
    d = ConvDesc(n=n, h=h, w=w, cin=cin or packed.cin, in_pitch=in_pitch, cout=cout or packed.cout_pad,
                 out_pitch=out_pitch, ksize=packed.ksize, stride=stride, act=act, out_f32=1 if out_f32 else 0,
                 res_pitch=res_pitch if residual is not None else 0)
    if workspace is None:
        check(_lib.lib().me_conv_gemm(byref(d), ptr(x), ptr(packed.w), ptr(packed.bias), ptr(residual), ptr(out),
                                      stream_ptr()), "me_conv_gemm")
    else:
        check(_lib.lib().me_conv_gemm_ws(byref(d), ptr(x), ptr(packed.w), ptr(packed.bias), ptr(residual), ptr(out),
                                         ptr(workspace), workspace.numel() * workspace.element_size(), stream_ptr()),
              "me_conv_gemm_ws")
    return out


def conv_pool_supported(packed, n, h, w, in_pitch, out_pitch, act=ME_ACT_LEAKY, cin=None, cout=None):
    """True if conv_pool() can run this 3x3 / stride-1 layer with the 2x2 / stride-2 max-pool behind it in one kernel."""
    d = conv_desc(packed, n, h, w, in_pitch, out_pitch, 1, act, 0, cin, cout)
    return bool(_lib.lib().me_conv_pool_supported(byref(d)))


def conv_pool(x, packed, n, h, w, in_pitch, out, out_pitch, act=ME_ACT_LEAKY, cin=None, cout=None):
    """conv_gemm (3x3, stride 1, thin layer) + maxpool2(stride 2) in one kernel; out is the pooled (n, h/2, w/2, out_pitch)
    tensor.  Bit-identical to the two calls (me_conv_pool)."""
    _need_cuda(x, out)
    d = conv_desc(packed, n, h, w, in_pitch, out_pitch, 1, act, 0, cin, cout)
    check(_lib.lib().me_conv_pool(byref(d), ptr(x), ptr(packed.w), ptr(packed.bias), ptr(out), stream_ptr()), "me_conv_pool")
    return out


def conv_gemm_yolo(x, packed, n, h, w, in_pitch, pred, g, anchors, num_classes, yolo_stride, rows_total, row_offset,
                   cin=None, cout=None):
    """Head conv (1x1, linear, bias) with the YOLO decode fused into its epilogue: writes decoded rows into
    pred (n, rows_total, 5+C) fp32 at row_offset (reference models.py:252 + :142-177)."""
    _need_cuda(x, pred)
    assert pred.dtype == torch.float32
    d = ConvDesc(n=n, h=h, w=w, cin=cin or packed.cin, in_pitch=in_pitch, cout=cout or packed.cout_pad,
                 out_pitch=cout or packed.cout_pad, ksize=packed.ksize, stride=1, act=ME_ACT_LINEAR, out_f32=1, res_pitch=0)
    flat = [float(v) for wh in anchors for v in wh]
    arr = (c_float * len(flat))(*flat)
    check(_lib.lib().me_conv_gemm_yolo(byref(d), ptr(x), ptr(packed.w), ptr(packed.bias), g, len(anchors), num_classes,
                                       arr, float(yolo_stride), rows_total, row_offset, ptr(pred), stream_ptr()),
          "me_conv_gemm_yolo")
    return pred


def conv_desc(packed, n, h, w, in_pitch, out_pitch, stride=1, act=ME_ACT_LEAKY, res_pitch=0, cin=None, cout=None,
              out_f32=False):
    return ConvDesc(n=n, h=h, w=w, cin=cin or packed.cin, in_pitch=in_pitch, cout=cout or packed.cout_pad,
                    out_pitch=out_pitch, ksize=packed.ksize, stride=stride, act=act, out_f32=1 if out_f32 else 0,
                    res_pitch=res_pitch)


def conv_chain_eligible(desc):
    return bool(_lib.lib().me_conv_chain_eligible(byref(desc)))


class ConvChain:
    """A run of conv layers executed as one persistent kernel (include/millieye_b200.h: me_conv_chain_*).

    layers: list of dicts with desc (ConvDesc), x, packed (PackedConv), y, residual (tensor or None), dep, res (index of
    the layer in this list that produces x / residual, or -1 when it is complete before the launch)."""

    def __init__(self, layers, device):
        L = _lib.lib()
        n = len(layers)
        arr = (_lib.ChainLayer * n)()
        self._keep = layers
        for a, l in zip(arr, layers):
            _need_cuda(l["x"], l["y"], l.get("residual"))
            a.d = l["desc"]
            a.x = l["x"].data_ptr()
            a.w_packed = l["packed"].w.data_ptr()
            a.bias = l["packed"].bias.data_ptr()
            a.residual = l["residual"].data_ptr() if l.get("residual") is not None else None
            a.y = l["y"].data_ptr()
            a.dep_layer = l.get("dep", -1)
            a.res_layer = l.get("res", -1)
        with torch.cuda.device(device):
            nbytes = L.me_conv_chain_blob_bytes(arr, n)
            if nbytes == 0:
                raise _lib.MeError("me_conv_chain_blob_bytes failed")
            raw = torch.zeros((nbytes + 128,), dtype=torch.uint8)
            shift = (-raw.data_ptr()) % 128
            self.host = raw[shift:shift + nbytes]
            check(L.me_conv_chain_build(arr, n, ctypes.c_void_p(self.host.data_ptr()), nbytes), "me_conv_chain_build")
            self.dev = torch.empty((nbytes,), dtype=torch.uint8, device=device)
            self.dev.copy_(self.host)
        self.n_layers = n

    def run(self):
        check(_lib.lib().me_conv_chain_run(ctypes.c_void_p(self.host.data_ptr()), ptr(self.dev), stream_ptr()),
              "me_conv_chain_run")


def conv_workspace(device):
    """Zero-filled workspace for the conv kernels' split-K tail (include/millieye_b200.h: me_conv_gemm_ws)."""
    return torch.zeros((_lib.lib().me_conv_workspace_bytes(),), dtype=torch.uint8, device=device)


class FirstConv:
    def __init__(self, w, bias, cin, cout):
        self.w, self.bias, self.cin, self.cout = w, bias, cin, cout
        self.wk = torch.zeros((cout, 32), dtype=torch.float16, device=w.device)  # tensor-core weight tile


def pack_first_conv(weight, conv_bias=None, bn=None):
    L = _lib.lib()
    weight = _f32(weight)
    _need_cuda(weight)
    cout, cin, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    wf = torch.empty_like(weight)
    bias = torch.empty((cout,), dtype=torch.float32, device=weight.device)
    cb = _f32(conv_bias)
    g = b = m = v = None
    eps = 0.0
    if bn is not None:
        g, b, m, v = (_f32(t) for t in bn[:4])
        eps = float(bn[4])
    check(L.me_fold_first_weights(ptr(weight), ptr(cb), ptr(g), ptr(b), ptr(m), ptr(v), eps, cout, cin, ptr(wf),
                                  ptr(bias), stream_ptr()), "me_fold_first_weights")
    return FirstConv(wf, bias, cin, cout)


def conv_first(x_nchw, first, out, out_pitch, act=ME_ACT_LEAKY, tensor_cores=True, pool=False):
    """First 3x3 conv from the NCHW fp32 image; pool=True also applies the 2x2 / stride-2 max-pool that follows it in the
    tiny cfgs (out is then the pooled (n, h/2, w/2, out_pitch) tensor)."""
    _need_cuda(x_nchw, out)
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    n, c, h, w = x_nchw.shape
    assert c == first.cin
    if pool:
        check(_lib.lib().me_conv_first_tc_pool(ptr(x_nchw), ptr(first.w), ptr(first.bias), ptr(first.wk), ptr(out), n, h, w,
                                               c, first.cout, out_pitch, act, stream_ptr()), "me_conv_first_tc_pool")
        return out
    if tensor_cores and c <= 3 and first.cout in (16, 32, 64):
        check(_lib.lib().me_conv_first_tc(ptr(x_nchw), ptr(first.w), ptr(first.bias), ptr(first.wk), ptr(out), n, h, w,
                                          c, first.cout, out_pitch, act, stream_ptr()), "me_conv_first_tc")
        return out
    check(_lib.lib().me_conv_first(ptr(x_nchw), ptr(first.w), ptr(first.bias), ptr(out), n, h, w, c, first.cout,
                                   out_pitch, act, stream_ptr()), "me_conv_first")
    return out


def maxpool2(x, out, n, h, w, c, in_pitch, out_pitch, stride):
    _need_cuda(x, out)
    check(_lib.lib().me_maxpool2(ptr(x), ptr(out), n, h, w, c, in_pitch, out_pitch, stride, stream_ptr()),
          "me_maxpool2")
    return out


def upsample2(x, out, n, h, w, c, in_pitch, out_pitch):
    _need_cuda(x, out)
    check(_lib.lib().me_upsample2(ptr(x), ptr(out), n, h, w, c, in_pitch, out_pitch, stream_ptr()), "me_upsample2")
    return out


def copy_channels(x, out, pixels, c, in_pitch, out_pitch):
    _need_cuda(x, out)
    check(_lib.lib().me_copy_channels(ptr(x), ptr(out), pixels, c, in_pitch, out_pitch, stream_ptr()),
          "me_copy_channels")
    return out


def nhwc_to_nchw_f32(x, n, h, w, c, in_pitch, out=None):
    _need_cuda(x)
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().me_nhwc_to_nchw_f32(ptr(x), ptr(out), n, h, w, c, in_pitch, stream_ptr()), "me_nhwc_to_nchw_f32")
    return out


def u8_to_unit_f32(x, out):
    """uint8 image bytes -> fp32 0..1 (x / 255, ToTensor) on the device; same shape in and out."""
    _need_cuda(x, out)
    assert x.dtype == torch.uint8 and out.dtype == torch.float32 and x.numel() == out.numel() and x.is_contiguous()
    check(_lib.lib().me_u8_to_unit_f32(ptr(x), ptr(out), x.numel(), stream_ptr()), "me_u8_to_unit_f32")
    return out


def nchw_f32_to_nhwc(x, out_pitch=None, out=None):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c, h, w = x.shape
    out_pitch = out_pitch or round_up(c, 8)
    if out is None:
        out = torch.empty((n, h, w, out_pitch), dtype=torch.float16, device=x.device)
    check(_lib.lib().me_nchw_f32_to_nhwc(ptr(x), ptr(out), n, h, w, c, out_pitch, stream_ptr()), "me_nchw_f32_to_nhwc")
    return out


def yolo_decode(logits, pitch, out, n, g, anchors, num_classes, stride, rows_total, row_offset):
    _need_cuda(logits, out)
    assert logits.dtype == torch.float32 and out.dtype == torch.float32
    flat = [float(v) for wh in anchors for v in wh]
    arr = (c_float * len(flat))(*flat)
    check(_lib.lib().me_yolo_decode(ptr(logits), pitch, ptr(out), n, g, len(anchors), num_classes, arr, float(stride),
                                    rows_total, row_offset, stream_ptr()), "me_yolo_decode")
    return out


METRIC_KEYS = ("loss", "x", "y", "w", "h", "conf", "cls", "cls_acc", "recall50", "recall75", "precision", "conf_obj",
               "conf_noobj", "grid_size")


def yolo_loss(logits, pitch, n, g, anchors, num_classes, stride, targets, out14, workspace, ignore_thres=0.5, obj_scale=1.0,
              noobj_scale=100.0):
    """Loss + metrics of one YOLO layer (reference models.py:180-232, utils.py:381-440) -> out14 (fp32 cuda, 14 values)."""
    _need_cuda(logits, out14, workspace)
    m = 0 if targets is None else int(targets.shape[0])
    if m:
        _need_cuda(targets)
        assert targets.dtype == torch.float32 and targets.is_contiguous()
    flat = [float(v) for wh in anchors for v in wh]
    arr = (c_float * len(flat))(*flat)
    check(_lib.lib().me_yolo_loss(ptr(logits), pitch, n, g, len(anchors), num_classes, arr, float(stride),
                                  ptr(targets) if m else None, m, float(ignore_thres), float(obj_scale), float(noobj_scale),
                                  ptr(workspace), workspace.numel() * workspace.element_size(), ptr(out14), stream_ptr()),
          "me_yolo_loss")
    return out14


class NmsBuffers:
    """Reusable outputs + workspace of filter_nms for one (n, rows, classes) shape."""

    def __init__(self, n, rows, num_classes, max_det, device):
        L = _lib.lib()
        self.n, self.rows, self.nc, self.max_det = n, rows, num_classes, max_det
        # det and count are views of ONE flat fp32 buffer (counts bit-cast): the multi-GPU gather sends it as it is
        # (dist.gather_detections, a single all_gather_into_tensor) and the host read-back is one copy
        cols = 7 + num_classes
        self.flat = torch.zeros((n * max_det * cols + n,), dtype=torch.float32, device=device)
        self.det = self.flat[:n * max_det * cols].view(n, max_det, cols)
        self.count = self.flat[n * max_det * cols:].view(torch.int32)
        self.index = torch.zeros((n, max_det), dtype=torch.int32, device=device)
        self.ws_bytes = L.me_filter_nms_workspace(n, rows, num_classes)
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=device)


def filter_nms(pred, conf_thresh, nms_thresh=0.5, max_det=200, xyxy_inplace=True, buffers=None):
    """pred: (n, rows, 5+C) fp32 cuda, contiguous. Returns the NmsBuffers holding det/count/index."""
    _need_cuda(pred)
    assert pred.dtype == torch.float32 and pred.is_contiguous() and pred.dim() == 3
    n, rows, attrs = pred.shape
    nc = attrs - 5
    if buffers is None:
        buffers = NmsBuffers(n, rows, nc, max_det, pred.device)
    check(_lib.lib().me_filter_nms(ptr(pred), n, rows, nc, float(conf_thresh), float(nms_thresh), max_det,
                                   1 if xyxy_inplace else 0, ptr(buffers.det), ptr(buffers.count), ptr(buffers.index),
                                   ptr(buffers.ws), buffers.ws_bytes, stream_ptr()), "me_filter_nms")
    return buffers


def _count_ptr(counts, index):
    return ctypes.c_void_p(counts.data_ptr() + 4 * index)


def psroi_align(feat, n, h, w, pitch, out_channels, pooled, scale, rois, counts, cap, out, out_pitch):
    _need_cuda(feat, rois, counts, out)
    check(_lib.lib().me_psroi_align(ptr(feat), n, h, w, pitch, out_channels, pooled, float(scale), ptr(rois),
                                    _count_ptr(counts, 1), cap, ptr(out), out_pitch, stream_ptr()), "me_psroi_align")
    return out


def roi_align(feat, n, h, w, pitch, channels, pooled, scale, rois, counts, cap, out, out_pitch):
    _need_cuda(feat, rois, counts, out)
    check(_lib.lib().me_roi_align(ptr(feat), n, h, w, pitch, channels, pooled, float(scale), ptr(rois),
                                  _count_ptr(counts, 1), cap, ptr(out), out_pitch, stream_ptr()), "me_roi_align")
    return out


def roi_gather_bin_major(feat, n, h, w, pitch, channels, pooled, scale, rois, counts, cap, out, out_pitch, position_sensitive):
    """psroi_align / roi_align with output element (ph*P + pw)*C + c (and score-map channel in the same order for the
    position-sensitive gather); see me_roi_gather_bin_major.  bin_major_perm() is the permutation the caller applies."""
    _need_cuda(feat, rois, counts, out)
    check(_lib.lib().me_roi_gather_bin_major(ptr(feat), n, h, w, pitch, channels, pooled, float(scale), ptr(rois),
                                             _count_ptr(counts, 1), cap, ptr(out), out_pitch, int(bool(position_sensitive)),
                                             stream_ptr()), "me_roi_gather_bin_major")
    return out


def bin_major_perm(channels, pooled, device=None):
    """perm[j'] = j: the reference index c*P*P + bin that bin-major position j' = bin*C + c holds."""
    bins = pooled * pooled
    jp = torch.arange(channels * bins, device=device)
    return (jp % channels) * bins + jp // channels


def build_proposals(det, det_count, class_idx, radar_boxes, img_size, img_boxes, rois, counts, cap):
    _need_cuda(det, det_count, img_boxes, rois, counts)
    n, max_det, det_cols = det.shape
    nr = 0 if radar_boxes is None else int(radar_boxes.shape[0])
    check(_lib.lib().me_build_proposals(ptr(det), ptr(det_count), n, max_det, det_cols, class_idx,
                                        ptr(radar_boxes) if nr else None, nr, float(img_size), ptr(img_boxes),
                                        ptr(rois), ptr(counts), cap, stream_ptr()), "me_build_proposals")


def build_proposals_dev(det, det_count, class_idx, radar_boxes, num_radar_dev, img_size, img_boxes, rois, counts, cap):
    """build_proposals with the radar-box count in device memory (radar_boxes: the whole capacity buffer)."""
    _need_cuda(det, det_count, radar_boxes, num_radar_dev, img_boxes, rois, counts)
    n, max_det, det_cols = det.shape
    check(_lib.lib().me_build_proposals_dev(ptr(det), ptr(det_count), n, max_det, det_cols, class_idx, ptr(radar_boxes),
                                            int(radar_boxes.shape[0]), ptr(num_radar_dev), float(img_size), ptr(img_boxes),
                                            ptr(rois), ptr(counts), cap, stream_ptr()), "me_build_proposals_dev")


def fusion_heads(hidden, hidden_pitch, crop, crop_pitch, head_weights, img_boxes, counts, cap, regress, refine, mask):
    _need_cuda(hidden, crop, img_boxes, counts, regress, refine, mask)
    check(_lib.lib().me_fusion_heads(ptr(hidden), hidden_pitch, ptr(crop), crop_pitch, byref(head_weights),
                                     ptr(img_boxes), ptr(counts), cap, ptr(regress), ptr(refine), ptr(mask),
                                     stream_ptr()), "me_fusion_heads")


def stage2_heads(hidden, hidden_pitch, weights, boxes, box_pitch, num_vec, counts, cap, regress, mask, refine=None):
    _need_cuda(hidden, boxes, counts, regress, mask, refine)
    check(_lib.lib().me_stage2_heads(ptr(hidden), hidden_pitch, byref(weights), ptr(boxes), box_pitch, num_vec,
                                     ptr(counts), cap, ptr(regress), ptr(mask), ptr(refine), stream_ptr()), "me_stage2_heads")


def stage2_loss(boxes, box_pitch, rois, refine, regress, mask, counts, cap, iou_labels, target_location, sample_filter, pos_ws,
                out10, iou_hi, alpha, lambda0, lambda1, thr):
    """Losses + counters of reference module2_mixed/my_models.py:399-445 -> out10 (see include/millieye_b200.h)."""
    _need_cuda(boxes, rois, refine, regress, mask, counts, iou_labels, target_location, sample_filter, pos_ws, out10)
    cfg = _lib.Stage2LossCfg(float(iou_hi), float(alpha), float(lambda0), float(lambda1), float(thr))
    check(_lib.lib().me_stage2_loss(ptr(boxes), box_pitch, ptr(rois), ptr(refine), refine.shape[1], ptr(regress), ptr(mask),
                                    ptr(counts), cap, ptr(iou_labels), ptr(target_location), ptr(sample_filter), byref(cfg),
                                    ptr(pos_ws), ptr(out10), stream_ptr()), "me_stage2_loss")


def make_stage2_weights(tensors):
    hw = _lib.Stage2Weights()
    for name, _ in _lib.Stage2Weights._fields_:
        t = tensors[name]
        _need_cuda(t)
        assert t.dtype == torch.float32 and t.is_contiguous()
        setattr(hw, name, t.data_ptr())
    return hw


def finalize_output(img_boxes, rois, refine, regress, mask, counts, cap, thr_img, thr_radar, regress_boxes, out,
                    out_count, ws, box_pitch=9):
    _need_cuda(img_boxes, rois, refine, regress, mask, counts, out, out_count, ws)
    check(_lib.lib().me_finalize_output(ptr(img_boxes), ptr(rois), ptr(refine), ptr(regress), ptr(mask), ptr(counts),
                                        cap, box_pitch, float(thr_img), float(thr_radar), 1 if regress_boxes else 0, ptr(out),
                                        ptr(out_count), ptr(ws), ws.numel() * ws.element_size(), stream_ptr()),
          "me_finalize_output")


def make_head_weights(tensors):
    """tensors: dict name -> fp32 cuda tensor for every field of me_head_weights."""
    hw = HeadWeights()
    for name, _ in HeadWeights._fields_:
        t = tensors[name]
        _need_cuda(t)
        assert t.dtype == torch.float32 and t.is_contiguous()
        setattr(hw, name, t.data_ptr())
    return hw


def stage3_labels(img_boxes, rois, counts, cap, targets_xyxy, num_targets, iou_labels, target_location):
    """obtain_iou_labels on the proposal buffers (reference my_models.py:317-375); targets_xyxy (>= num_targets, 6) fp32 cuda."""
    _need_cuda(img_boxes, rois, counts, iou_labels, target_location)
    if num_targets:
        _need_cuda(targets_xyxy)
    check(_lib.lib().me_stage3_labels(ptr(img_boxes), img_boxes.shape[1], ptr(rois), ptr(counts), cap,
                                      ptr(targets_xyxy) if num_targets else None, num_targets, ptr(iou_labels),
                                      ptr(target_location), stream_ptr()), "me_stage3_labels")


def stage3_loss(rois, refine, regress, mask, counts, cap, iou_labels, target_location, sample_filter, out10, iou_hi, alpha,
                lambda_conf, thr_img, thr_radar):
    """Losses + counters of reference my_models.py:586-635 -> out10 (see include/millieye_b200.h)."""
    _need_cuda(rois, refine, regress, mask, counts, iou_labels, target_location, sample_filter, out10)
    assert sample_filter.dtype == torch.uint8 and out10.dtype == torch.float32 and out10.numel() >= 10
    cfg = _lib.Stage3LossCfg(float(iou_hi), float(alpha), float(lambda_conf), float(thr_img), float(thr_radar))
    check(_lib.lib().me_stage3_loss(ptr(rois), ptr(refine), ptr(regress), ptr(mask), ptr(counts), cap, ptr(iou_labels),
                                    ptr(target_location), ptr(sample_filter), byref(cfg), ptr(out10), stream_ptr()),
          "me_stage3_loss")

