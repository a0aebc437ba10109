"""Oracle of the YOLO training loss of Darknet.forward(x, targets) (reference module3_our_dataset/yolov3/models.py:132-232,
247-267 and utils/utils.py:238-246, 381-440).  SURVEY.md §8(f4): no reference script calls this branch and the product
does not implement it (Darknet.forward raises with targets); this restatement pins the semantics for a later round.

Test infrastructure only (see oracle/__init__.py).  fp32 torch-CPU elementwise operators (sigmoid, exp, log, the
MSE / BCE means), the target assignment restated.  Pinned by tests/golden/yolo_loss_tiny12_160.npz
(tests/golden/make_golden_yolo_loss.py runs the unmodified reference).
"""
import torch
import torch.nn.functional as F

from .darknet import darknet_forward, layer_plan


def bbox_wh_iou(anchor_wh, gwh):
    """utils.py:238-246: IoU of boxes that share a corner; anchor (2,), gwh (m,2) in grid units."""
    w1, h1 = anchor_wh[0], anchor_wh[1]
    w2, h2 = gwh[:, 0], gwh[:, 1]
    inter = torch.min(w1, w2) * torch.min(h1, h2)
    return inter / ((w1 * h1 + 1e-16) + w2 * h2 - inter)


def bbox_iou_cxcywh(b1, b2):
    """utils.py:249-281 with x1y1x2y2=False, row by row (the +1 pixel convention of the intersection / areas)."""
    b1x1, b1x2 = b1[:, 0] - b1[:, 2] / 2, b1[:, 0] + b1[:, 2] / 2
    b1y1, b1y2 = b1[:, 1] - b1[:, 3] / 2, b1[:, 1] + b1[:, 3] / 2
    b2x1, b2x2 = b2[:, 0] - b2[:, 2] / 2, b2[:, 0] + b2[:, 2] / 2
    b2y1, b2y2 = b2[:, 1] - b2[:, 3] / 2, b2[:, 1] + b2[:, 3] / 2
    iw = torch.clamp(torch.min(b1x2, b2x2) - torch.max(b1x1, b2x1) + 1, min=0)
    ih = torch.clamp(torch.min(b1y2, b2y2) - torch.max(b1y1, b2y1) + 1, min=0)
    inter = iw * ih
    a1 = (b1x2 - b1x1 + 1) * (b1y2 - b1y1 + 1)
    a2 = (b2x2 - b2x1 + 1) * (b2y2 - b2y1 + 1)
    return inter / (a1 + a2 - inter + 1e-16)


def yolo_layer_loss(logits, anchors, num_classes, img_dim, targets, ignore_thres=0.5, obj_scale=1, noobj_scale=100):
    """YOLOLayer.forward with targets (models.py:132-232).  logits (N, A*(5+C), G, G); targets (m,6)
    [image, class, cx, cy, w, h] in 0..1.  Returns (total_loss, metrics dict)."""
    n, g = logits.shape[0], logits.shape[2]
    na = len(anchors)
    pred = logits.view(n, na, num_classes + 5, g, g).permute(0, 1, 3, 4, 2).contiguous()
    x, y = torch.sigmoid(pred[..., 0]), torch.sigmoid(pred[..., 1])
    w, h = pred[..., 2], pred[..., 3]
    conf, cls = torch.sigmoid(pred[..., 4]), torch.sigmoid(pred[..., 5:])
    stride = img_dim / g
    scaled = torch.tensor([(aw / stride, ah / stride) for aw, ah in anchors], dtype=torch.float32)
    gx_ = torch.arange(g).repeat(g, 1).view(1, 1, g, g).float()
    gy_ = torch.arange(g).repeat(g, 1).t().view(1, 1, g, g).float()
    boxes = torch.stack((x + gx_, y + gy_, torch.exp(w) * scaled[:, 0].view(1, na, 1, 1),
                         torch.exp(h) * scaled[:, 1].view(1, na, 1, 1)), -1)

    # build_targets (utils.py:381-440): the anchor whose SHAPE fits the box best owns it, at the cell of its centre
    obj = torch.zeros(n, na, g, g, dtype=torch.bool)
    noobj = torch.ones(n, na, g, g, dtype=torch.bool)
    class_mask, iou_scores = torch.zeros(n, na, g, g), torch.zeros(n, na, g, g)
    tx, ty, tw, th = (torch.zeros(n, na, g, g) for _ in range(4))
    tcls = torch.zeros(n, na, g, g, num_classes)
    tb = targets[:, 2:6].float() * g
    gxy, gwh = tb[:, :2], tb[:, 2:]
    ious = torch.stack([bbox_wh_iou(a, gwh) for a in scaled])          # (A, m)
    best_n = ious.max(0)[1]
    b, labels = targets[:, 0].long(), targets[:, 1].long()
    gi, gj = gxy[:, 0].long(), gxy[:, 1].long()
    obj[b, best_n, gj, gi] = True
    noobj[b, best_n, gj, gi] = False
    for i in range(len(targets)):                                      # other anchors that also fit are not penalised
        noobj[b[i], ious[:, i] > ignore_thres, gj[i], gi[i]] = False
    tx[b, best_n, gj, gi] = gxy[:, 0] - gxy[:, 0].floor()
    ty[b, best_n, gj, gi] = gxy[:, 1] - gxy[:, 1].floor()
    tw[b, best_n, gj, gi] = torch.log(gwh[:, 0] / scaled[best_n][:, 0] + 1e-16)
    th[b, best_n, gj, gi] = torch.log(gwh[:, 1] / scaled[best_n][:, 1] + 1e-16)
    tcls[b, best_n, gj, gi, labels] = 1
    class_mask[b, best_n, gj, gi] = (cls[b, best_n, gj, gi].argmax(-1) == labels).float()
    iou_scores[b, best_n, gj, gi] = bbox_iou_cxcywh(boxes[b, best_n, gj, gi], tb)
    tconf = obj.float()

    mse, bce = F.mse_loss, F.binary_cross_entropy
    loss_x, loss_y = mse(x[obj], tx[obj]), mse(y[obj], ty[obj])
    loss_w, loss_h = mse(w[obj], tw[obj]), mse(h[obj], th[obj])
    loss_conf = obj_scale * bce(conf[obj], tconf[obj]) + noobj_scale * bce(conf[noobj], tconf[noobj])
    loss_cls = bce(cls[obj], tcls[obj])
    total = loss_x + loss_y + loss_w + loss_h + loss_conf + loss_cls
    conf50, iou50, iou75 = (conf > 0.5).float(), (iou_scores > 0.5).float(), (iou_scores > 0.75).float()
    detected = conf50 * class_mask * tconf
    metrics = dict(loss=total.item(), x=loss_x.item(), y=loss_y.item(), w=loss_w.item(), h=loss_h.item(),
                   conf=loss_conf.item(), cls=loss_cls.item(), cls_acc=(100 * class_mask[obj].mean()).item(),
                   recall50=(torch.sum(iou50 * detected) / (obj.sum() + 1e-16)).item(),
                   recall75=(torch.sum(iou75 * detected) / (obj.sum() + 1e-16)).item(),
                   precision=(torch.sum(iou50 * detected) / (conf50.sum() + 1e-16)).item(),
                   conf_obj=conf[obj].mean().item(), conf_noobj=conf[noobj].mean().item(), grid_size=g)
    return total, metrics


def darknet_loss(module_defs, sd, images, targets, prefix=""):
    """Darknet.forward(x, targets) (models.py:247-267): the sum of the YOLO layers' losses, with the inference
    outputs.  Returns (loss, featuremap, yolo_outputs, [metrics per YOLO layer])."""
    feat, yolo_out, outs = darknet_forward(module_defs, sd, images, prefix=prefix, collect=True)
    _, blocks = layer_plan(module_defs)
    loss, metrics = 0, []
    for i, blk in enumerate(blocks):
        if blk["type"] == "yolo":
            layer_loss, m = yolo_layer_loss(outs[i - 1], blk["anchors"], blk["classes"], images.shape[2], targets)
            loss = loss + layer_loss
            metrics.append(m)
    return loss, feat, yolo_out, metrics
