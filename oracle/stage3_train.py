"""Oracle of one stage-3 TRAINING step of milliEye's fusion model, the way train.py runs it (reference
module3_our_dataset/train.py:169-186: model.train(), base_detector.eval(), forward with targets, loss.backward()).

Test infrastructure only (see oracle/__init__.py), and ahead of the product: the backward pass is not built yet
(DESIGN.md §7).  This restatement exists so that the kernels of the next round have a pinned checker from day one.
fp32 on CPU; forward wiring restated from my_models.py:433-539 with the heads in train mode (BatchNorm uses batch
statistics and updates its running statistics, momentum 0.1), labels / sampling from oracle/stage3_loss.py, the two
loss terms of :610-635 written with differentiable torch operators, gradients by torch.autograd.  RoI ops are the
installed torchvision's (the library the reference calls, :495-496), which also provides their backward.

Pinned by tests/golden/stage3_grads_tiny12_192.npz (tests/golden/make_golden_stage3_grads.py runs the unmodified
reference): loss, the gradient of every head parameter, BatchNorm running statistics after the step.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import boxes as obox
from . import stage3_loss as s3
from .darknet import darknet_forward


def _bn_train(x, params, buffers, p, momentum=0.1):
    return F.batch_norm(x, buffers[p + "running_mean"], buffers[p + "running_var"], params[p + "weight"], params[p + "bias"],
                        training=True, momentum=momentum, eps=1e-5)


def train_step(module_defs, sd, images, maps, radar_boxes, conf_thresh, targets, class_idx=0, class_num=1,
               iou_thresh=(0.3, 0.7), alpha=0.75, balance_factor=5, loss_lambda=(6, 1), sample_filter=None):
    """One forward + backward.  sd: the Network state_dict (fp32 tensors); targets (m,6) [image, class, cx, cy, w, h]
    in 0..1 (a copy is converted).  Returns dict(loss, grads {name: tensor}, buffers {name: tensor after the step},
    n_img, n_all, sample_filter)."""
    from torchvision.ops import ps_roi_align, roi_align
    params, buffers = {}, {}
    for k, v in sd.items():
        if k.startswith("base_detector.") or not v.is_floating_point():
            continue
        if "running_" in k:
            buffers[k] = v.clone().float()
        else:
            params[k] = v.clone().float().requires_grad_(True)
    with torch.no_grad():
        feat, yolo_out = darknet_forward(module_defs, sd, images, prefix="base_detector.")
    dets, _ = obox.non_max_suppression_cpp(yolo_out.clone().numpy(), conf_thresh)
    rows = []
    for i, d in enumerate(dets):
        if d is None:
            continue
        d = d[d[:, 6] == class_idx]
        if len(d) > 0:
            b = np.zeros((len(d), 8 + class_num), dtype=np.float32)
            b[:, 0] = i
            b[:, 1:] = d[:, :7 + class_num]
            rows.append(b)
    img_boxes = torch.from_numpy(np.concatenate(rows, 0)) if rows else torch.empty((0, 8 + class_num))
    n_img = len(img_boxes)

    p = params
    x = F.conv2d(feat, p["img_cnn_layers.net.conv_0.weight"], p["img_cnn_layers.net.conv_0.bias"])
    roi_score_map = F.leaky_relu(_bn_train(x, p, buffers, "img_cnn_layers.net.batch_norm_0."), 0.1)
    x = maps
    for name in ("conv1", "conv2", "conv3"):
        x = F.conv2d(x, p[f"radar_cnn_layers.{name}.0.weight"], p[f"radar_cnn_layers.{name}.0.bias"], padding=1)
        x = F.leaky_relu(_bn_train(x, p, buffers, f"radar_cnn_layers.{name}.1."), 0.1)
    radar_score_map = torch.sigmoid(F.conv2d(x, p["radar_cnn_layers.conv3.3.weight"], p["radar_cnn_layers.conv3.3.bias"]))

    radar_px = radar_boxes.clone().float()
    if len(radar_px) > 0:
        radar_px[:, 1:] *= images.shape[-1]
    box_locations = torch.cat((img_boxes[:, :5], radar_px), 0)
    crop_img = ps_roi_align(roi_score_map, box_locations, (7, 7), spatial_scale=1. / 16)
    crop_radar = roi_align(radar_score_map, box_locations, (7, 7), spatial_scale=1. / 16)

    h = "refinement_head."
    t = F.leaky_relu(F.linear(crop_img.flatten(start_dim=1), p[h + "net0.0.weight"], p[h + "net0.0.bias"]), 0.1)
    reg = F.linear(t, p[h + "net1.0.weight"], p[h + "net1.0.bias"])
    cls = torch.sigmoid(F.linear(t, p[h + "net2.0.weight"], p[h + "net2.0.bias"]))
    r = F.conv2d(crop_radar, p[h + "radar_net.0.weight"], p[h + "radar_net.0.bias"])
    r = F.leaky_relu(_bn_train(r, p, buffers, h + "radar_net.1."), 0.1)
    r = torch.sigmoid(F.conv2d(r, p[h + "radar_net.3.weight"], p[h + "radar_net.3.bias"]))
    conf = torch.sigmoid(r.squeeze(-1).squeeze(-1) + cls[:, :1])
    ref_vec = torch.cat((conf, cls[:, 1:2]), -1)

    e = "ensemble_head."
    yolo_vec = torch.cat((img_boxes[:, 5:6], img_boxes[:, 8:]), 1)
    z = torch.stack((ref_vec[:n_img], yolo_vec), -1)
    z = F.leaky_relu(F.linear(z, p[e + "fc1.0.weight"], p[e + "fc1.0.bias"]), 0.1).flatten(start_dim=1)
    masks_img = torch.softmax(F.linear(z, p[e + "fc2.0.weight"], p[e + "fc2.0.bias"]), dim=1)
    m = torch.cat((masks_img[:, :1], ref_vec[n_img:, :1]), 0)
    masks = torch.cat((1 - m, m), -1)

    # labels and the balanced sample (numpy, no gradient) - my_models.py:545-604
    radar_rows = torch.cat((radar_px, ref_vec[n_img:].detach(), torch.zeros((len(radar_px), 1)), ref_vec[n_img:, 1:].detach()), -1)
    all_boxes = torch.cat((img_boxes, radar_rows), 0).numpy()
    boxes6 = np.concatenate((all_boxes[:, :1], all_boxes[:, 7:8], all_boxes[:, 1:5]), 1)
    tpx = s3.targets_to_pixels(np.asarray(targets, dtype=np.float32), images.shape[3])
    iou_labels, _ = s3.obtain_iou_labels(boxes6, tpx, True)
    flat = iou_labels.reshape(-1)
    pos, neg = flat > iou_thresh[1], flat < iou_thresh[0]
    if sample_filter is None:
        sample_filter = s3.sample_filter_reference(pos, neg, balance_factor)
    sel = torch.from_numpy(sample_filter)
    pos_t = torch.from_numpy(pos)

    # FocalLoss on the sampled image proposals (:287-314, :611) and the confidence BCE on the whole sample (:614-619)
    sel_img = sel[:n_img]
    onehot = torch.stack(((~pos_t).float(), pos_t.float()), 1)
    inputs, labels = masks[:n_img][sel_img], onehot[:n_img][sel_img]
    a = torch.where(labels[:, 1:2] == 1, torch.tensor(alpha), torch.tensor(1 - alpha))
    probs = (inputs * labels).sum(1).view(-1, 1)
    masks_loss = (-a * torch.pow(1 - probs, 2) * probs.log()).sum()
    conf_loss = F.binary_cross_entropy(ref_vec[sel, 0], pos_t.float()[sel], reduction="sum")
    loss = masks_loss + conf_loss / loss_lambda[0]
    loss.backward()
    grads = {k: v.grad for k, v in params.items() if v.grad is not None}
    return dict(loss=float(loss.item()), grads=grads, buffers=buffers, n_img=n_img, n_all=len(all_boxes), reg=reg.detach(),
                sample_filter=sample_filter, true=int(pos.sum()), pos=pos, box_locations=box_locations.detach(),
                yolo_vec=yolo_vec.detach(), cls=cls.detach(), feat=feat.detach())
