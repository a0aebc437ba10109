"""Oracle restatement of the box utilities and NMS the reference uses.

Test infrastructure only (see oracle/__init__.py).

  xywh2xyxy / xyxy2xywh        utils/utils.py:58-74
  non_max_suppression_cpp      utils/utils.py:337-378
  batched_nms / nms            third-party: torchvision 0.26.0 ops/boxes.py:51-120 and the CPU kernel
                               csrc/ops/cpu/nms_kernel.cpp (greedy loop over a stable descending
                               score order, IoU = inter / (area_i + area_j - inter), strict '>').
All arithmetic is float32 numpy in the reference's operation order so survivor sets are bit-exact.
"""
import numpy as np

F32 = np.float32


def xywh2xyxy(x):
    x = np.asarray(x, dtype=F32)
    y = np.empty_like(x)
    y[..., 0] = x[..., 0] - x[..., 2] / F32(2)
    y[..., 1] = x[..., 1] - x[..., 3] / F32(2)
    y[..., 2] = x[..., 0] + x[..., 2] / F32(2)
    y[..., 3] = x[..., 1] + x[..., 3] / F32(2)
    return y


def xyxy2xywh(x):
    x = np.asarray(x, dtype=F32)
    y = np.zeros_like(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / F32(2)
    y[..., 1] = (x[..., 1] + x[..., 3]) / F32(2)
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def nms(boxes, scores, iou_threshold):
    """Greedy NMS, torchvision CPU kernel order. Returns kept indices, score-descending."""
    boxes = np.asarray(boxes, dtype=F32)
    scores = np.asarray(scores, dtype=F32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores, kind="stable")  # scores.sort(stable=True, descending=True)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        if rest.size == 0:
            continue
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1)
        h = np.maximum(F32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.batched_nms on CPU tensors (both code paths, boxes.py:80-120)."""
    boxes = np.asarray(boxes, dtype=F32)
    scores = np.asarray(scores, dtype=F32)
    idxs = np.asarray(idxs)
    if boxes.size > 4000:
        # _batched_nms_vanilla: per-class nms on the raw boxes, survivors re-sorted by score
        keep_mask = np.zeros(scores.shape[0], dtype=bool)
        for cid in np.unique(idxs):
            cur = np.where(idxs == cid)[0]
            k = nms(boxes[cur], scores[cur], iou_threshold)
            keep_mask[cur[k]] = True
        kept = np.where(keep_mask)[0]
        return kept[np.argsort(-scores[kept], kind="stable")]
    # _batched_nms_coordinate_trick
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    max_coordinate = boxes.max()
    offsets = idxs.astype(F32) * (max_coordinate + F32(1))
    return nms(boxes + offsets[:, None], scores, iou_threshold)


def non_max_suppression_cpp(prediction, conf_thresh, nms_thresh=0.5, detections_per_img=200, use_torchvision=False):
    """prediction: (N, B, 5+C) float32, rewritten in place to xyxy like the reference (:354).

    Returns (detections, source_rows): per image an (k, 7+C) array [x1,y1,x2,y2,conf,class_conf,
    class_pred,cls...] or None, and the prediction-row index of every survivor."""
    prediction[..., :4] = xywh2xyxy(prediction[..., :4])
    out = [None] * len(prediction)
    rows_out = [None] * len(prediction)
    for i, pred in enumerate(prediction):
        sel = np.where(pred[:, 4] >= F32(conf_thresh))[0]
        pred = pred[sel]
        if pred.shape[0] == 0:
            continue
        class_preds = np.argmax(pred[:, 5:], axis=1)            # first maximum, like torch.max
        class_confs = pred[np.arange(pred.shape[0]), 5 + class_preds]
        det = np.concatenate((pred[:, :5], class_confs[:, None], class_preds[:, None].astype(F32), pred[:, 5:]), 1)
        if use_torchvision:
            import torch
            from torchvision.ops import boxes as box_ops
            keep = box_ops.batched_nms(torch.from_numpy(det[:, :4].copy()), torch.from_numpy(det[:, 4].copy()),
                                       torch.from_numpy(det[:, 6].copy()), nms_thresh).numpy()
        else:
            keep = batched_nms(det[:, :4], det[:, 4], det[:, 6], nms_thresh)
        keep = keep[:detections_per_img]
        if len(keep) > 0:
            out[i] = det[keep]
            rows_out[i] = sel[keep]
    return out, rows_out
