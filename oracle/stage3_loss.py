"""Oracle restatement of milliEye's stage-3 labelling + loss branch (reference module3_our_dataset/my_models.py:545-640).

Test infrastructure only (see oracle/__init__.py).  fp32 numpy, python loops (the row counts are a few hundred).
Pinned by tests/golden/stage3_loss_tiny12_192.npz, generated from the unmodified reference by
tests/golden/make_golden_stage3_loss.py (heads in eval mode; `random` seeded for the negative sub-sampling).

Covered: targets rewrite (:548-549), obtain_iou_labels (:317-375, bbox_iou utils/utils.py:255-281 with the +1 pixel
convention), pos/neg filters (:556-557), metric (:562-583), balanced sampling (:592-604, python random.sample),
FocalLoss (:287-314), confidence BCE (:614-619), regression_loss (:394-408), category BCE (:628-633), total (:635).
Not covered: the b.txt appends (:350,369) and the prints.
"""
import random

import numpy as np

F32 = np.float32


def targets_to_pixels(targets, img_size):
    """(m,6) [image, class, cx, cy, w, h] in 0..1 -> [image, class, x1, y1, x2, y2] in pixels, on a copy
    (the reference rewrites the caller's tensor in place, :548-549; utils.xywh2xyxy :68-74)."""
    t = np.array(targets, dtype=F32, copy=True)
    cx, cy, w, h = (t[:, k].copy() for k in (2, 3, 4, 5))
    t[:, 2] = cx - w / F32(2)
    t[:, 3] = cy - h / F32(2)
    t[:, 4] = cx + w / F32(2)
    t[:, 5] = cy + h / F32(2)
    t[:, 2:] *= F32(img_size)
    return t


def bbox_iou_xyxy(box, others):
    """utils/utils.py:255-281 with x1y1x2y2=True: one box (4,) against (m,4); every step rounded to fp32."""
    b1x1, b1y1, b1x2, b1y2 = (F32(v) for v in box)
    b2x1, b2y1, b2x2, b2y2 = (others[:, k].astype(F32) for k in range(4))
    ix1, iy1 = np.maximum(b1x1, b2x1), np.maximum(b1y1, b2y1)
    ix2, iy2 = np.minimum(b1x2, b2x2), np.minimum(b1y2, b2y2)
    one = F32(1)
    inter = np.clip(ix2 - ix1 + one, F32(0), None).astype(F32) * np.clip(iy2 - iy1 + one, F32(0), None).astype(F32)
    a1 = (b1x2 - b1x1 + one) * (b1y2 - b1y1 + one)
    a2 = (b2x2 - b2x1 + one) * (b2y2 - b2y1 + one)
    return (inter / (a1 + a2 - inter + F32(1e-16))).astype(F32)


def obtain_iou_labels(boxes, targets, multi_boxes=True):
    """my_models.py:317-375.  boxes (p,6) [image, class, x1,y1,x2,y2]; targets (q,6) same layout.
    Returns iou_labels (p,1), target_location (p,4).  `detected` holds indices INTO THE FILTERED target list,
    as the reference does (:362-366); Network.forward passes a tuple as multi_boxes, i.e. True (:555)."""
    p = len(boxes)
    labels = np.zeros((p, 1), dtype=F32)
    locs = np.zeros((p, 4), dtype=F32)
    detected = []
    for i in range(p):
        keep = (targets[:, 0] == boxes[i, 0]) & (targets[:, 1] == boxes[i, 1])
        if not keep.any():
            continue
        cand = targets[keep][:, 2:]
        ious = bbox_iou_xyxy(boxes[i, 2:], cand)
        j = int(np.argmax(ious))           # first maximum, like torch.max on CPU
        iou = ious[j]
        if (j not in detected) or multi_boxes:
            labels[i, 0] = iou
            locs[i] = cand[j]
            if iou > 0.7:
                detected.append(j)
    return labels, locs


def focal_loss_sum(inputs, labels_onehot, alpha=0.75, gamma=2):
    """FocalLoss.forward :296-314, reduction='sum'."""
    if len(inputs) == 0:
        return F32(0)
    a = np.where(labels_onehot[:, 1:2] == 1, F32(alpha), F32(1 - alpha)).astype(F32)
    probs = (inputs * labels_onehot).sum(1, dtype=F32).reshape(-1, 1)
    batch = -a * ((F32(1) - probs) ** gamma).astype(F32) * np.log(probs).astype(F32)
    return F32(batch.sum(dtype=F32))


def bce_sum(x, y):
    """nn.BCELoss(reduction='sum'): log terms clamped at -100."""
    if len(x) == 0:
        return F32(0)
    x = x.astype(F32)
    with np.errstate(divide="ignore"):
        lx = np.maximum(np.log(x), F32(-100)).astype(F32)
        l1 = np.maximum(np.log(F32(1) - x), F32(-100)).astype(F32)
    return F32((-(y * lx + (F32(1) - y) * l1)).sum(dtype=F32))


def smooth_l1_sum(a, b):
    d = np.abs(a.astype(F32) - b.astype(F32))
    return F32(np.where(d < 1, F32(0.5) * d * d, d - F32(0.5)).sum(dtype=F32))


def xyxy2xywh(b):
    return np.stack(((b[:, 0] + b[:, 2]) / F32(2), (b[:, 1] + b[:, 3]) / F32(2), b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]), 1).astype(F32)


def regression_loss(reg, target_xyxy, roi_xyxy):
    """my_models.py:394-408."""
    if len(reg) == 0:
        return F32(0), F32(0)
    x, y, w, h = xyxy2xywh(roi_xyxy).T
    xt, yt, wt, ht = xyxy2xywh(target_xyxy).T
    eps = F32(1e-16)
    p01 = np.stack(((xt - x) / (w + eps), (yt - y) / (h + eps)), -1).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        p23 = np.stack((np.log(wt / w + eps), np.log(ht / h + eps)), -1).astype(F32)
    return smooth_l1_sum(p01, reg[:, :2]), smooth_l1_sum(p23, reg[:, 2:])


def sample_filter_reference(pos, neg, balance_factor=5):
    """:592-604 - every positive plus min(5 * #pos, #neg) negatives drawn with python's random.sample
    (the caller seeds `random` to reproduce a run)."""
    pos_idx = np.where(pos)[0]
    neg_idx = np.where(neg)[0]
    top_k = min(len(pos_idx) * balance_factor, len(neg_idx))
    keep = pos.copy()
    sel = neg_idx[random.sample(range(len(neg_idx)), k=top_k)] if top_k > 0 else np.zeros((0,), dtype=np.int64)
    keep[sel] = True
    return keep


def stage3_losses(boxes, masks, ref_vec, reg, n_img, positive_masks, targets_px, iou_thresh=(0.3, 0.7), alpha=0.75,
                  balance_factor=5, loss_lambda=(6, 1), sample_filter=None):
    """The training branch after the forward (:551-635) on numpy inputs:
    boxes (R, 8+c) rows [image, x1,y1,x2,y2, conf, class_conf, class_pred, cls...], masks (R,2) = [1-m, m],
    ref_vec (R,2), reg (R,4), positive_masks (R,) bool, targets_px (q,6) pixels xyxy.
    Returns dict(loss, masks_loss, conf_loss, loss_xy, loss_wh, category_loss, iou_labels, target_location,
    sample_filter, metric)."""
    boxes = boxes.astype(F32)
    R = len(boxes)
    boxes6 = np.concatenate((boxes[:, :1], boxes[:, 7:8], boxes[:, 1:5]), 1)
    iou_labels, target_location = obtain_iou_labels(boxes6, targets_px, True)
    flat = iou_labels.reshape(-1)
    pos = flat > iou_thresh[1]
    neg = flat < iou_thresh[0]
    if sample_filter is None:
        sample_filter = sample_filter_reference(pos, neg, balance_factor)
    pos_idx = np.where(pos)[0]
    onehot = np.tile(np.array([1.0, 0.0], dtype=F32), (R, 1))
    onehot[pos_idx] = np.array([0.0, 1.0], dtype=F32)
    sel_img = sample_filter[:n_img]
    masks_loss = focal_loss_sum(masks[:n_img][sel_img].astype(F32), onehot[:n_img][sel_img], alpha)
    conf_label = np.zeros(R, dtype=F32)
    conf_label[pos_idx] = 1
    conf_loss = bce_sum(ref_vec[sample_filter, 0], conf_label[sample_filter])
    loss_xy, loss_wh = regression_loss(reg[pos], target_location[pos], boxes[pos, 1:5])
    class_num = ref_vec.shape[1] - 1
    class_label = np.zeros((R, class_num), dtype=F32)
    for i, idx in enumerate(pos_idx):            # row i (not idx) is set: the reference's indexing (:629-630)
        class_label[i, int(boxes6[idx, 1])] = 1
    category_loss = bce_sum(ref_vec[pos, 1:].reshape(-1), class_label[pos].reshape(-1))
    loss = F32(masks_loss + conf_loss / F32(loss_lambda[0]))
    conf_1, conf_2 = boxes[:, 5], masks[:, 1]
    metric = dict(total=R, true=int(pos.sum()), positive=int(positive_masks.sum()),
                  tp=float((positive_masks & pos).sum()),
                  conf=dict(conf_1_pos=conf_1[flat > 0.5], conf_1_neg=conf_1[flat < 0.5],
                            conf_2_pos=conf_2[flat > 0.5], conf_2_neg=conf_2[flat < 0.5]))
    return dict(loss=loss, masks_loss=masks_loss, conf_loss=conf_loss, loss_xy=loss_xy, loss_wh=loss_wh,
                category_loss=category_loss, iou_labels=iou_labels, target_location=target_location,
                sample_filter=sample_filter, metric=metric)
