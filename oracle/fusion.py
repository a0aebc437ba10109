"""Oracle restatement of milliEye's fusion forward (reference my_models.py), inference branch.

Test infrastructure only (see oracle/__init__.py).  fp32 on CPU.  The conv / linear / batch-norm
arithmetic goes through the torch CPU operators the reference's modules call; the wiring, the
proposal assembly, the masks, thresholds, box regression and ordering are restated.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import boxes as obox
from . import roi as oroi
from .darknet import darknet_forward


def _bn(x, sd, p, momentum=0.1):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, momentum=momentum, eps=1e-5)


def img_cnn_layers(feat, sd, prefix="img_cnn_layers."):
    """cnn_layers_1((256, 490)) - my_models.py:47-77: 1x1 conv (bias) + BN + LeakyReLU(0.1)."""
    x = F.conv2d(feat, sd[prefix + "net.conv_0.weight"], sd[prefix + "net.conv_0.bias"])
    return F.leaky_relu(_bn(x, sd, prefix + "net.batch_norm_0."), 0.1)


def radar_cnn_layers(maps, sd, prefix="radar_cnn_layers."):
    """cnn_layers_3 - my_models.py:130-157."""
    x = maps
    for name in ("conv1", "conv2", "conv3"):
        x = F.conv2d(x, sd[f"{prefix}{name}.0.weight"], sd[f"{prefix}{name}.0.bias"], padding=1)
        x = F.leaky_relu(_bn(x, sd, f"{prefix}{name}.1."), 0.1)
    x = F.conv2d(x, sd[prefix + "conv3.3.weight"], sd[prefix + "conv3.3.bias"])
    return torch.sigmoid(x)


def refinement_head(radar_maps, img_maps, sd, prefix="refinement_head."):
    """refinement_head.forward - my_models.py:260-284 (net3 / fusion_head are unused there)."""
    flat = img_maps.flatten(start_dim=1)
    t = F.leaky_relu(F.linear(flat, sd[prefix + "net0.0.weight"], sd[prefix + "net0.0.bias"]), 0.1)
    reg = F.linear(t, sd[prefix + "net1.0.weight"], sd[prefix + "net1.0.bias"])
    cls = torch.sigmoid(F.linear(t, sd[prefix + "net2.0.weight"], sd[prefix + "net2.0.bias"]))
    r = F.conv2d(radar_maps, sd[prefix + "radar_net.0.weight"], sd[prefix + "radar_net.0.bias"])
    r = F.leaky_relu(_bn(r, sd, prefix + "radar_net.1."), 0.1)
    r = torch.sigmoid(F.conv2d(r, sd[prefix + "radar_net.3.weight"], sd[prefix + "radar_net.3.bias"]))
    radar_conf = r.squeeze(-1).squeeze(-1)
    conf = torch.sigmoid(radar_conf + cls[:, :1])
    return reg, torch.cat((conf, cls[:, 1:2]), -1)


def ensemble_head(ref_vec, yolo_vec, sd, prefix="ensemble_head."):
    """ensemble_head.forward - my_models.py:202-210."""
    x = torch.stack((ref_vec, yolo_vec), -1)
    x = F.leaky_relu(F.linear(x, sd[prefix + "fc1.0.weight"], sd[prefix + "fc1.0.bias"]), 0.1)
    x = x.flatten(start_dim=1)
    x = F.linear(x, sd[prefix + "fc2.0.weight"], sd[prefix + "fc2.0.bias"])
    return torch.softmax(x, dim=1)


def box_regress(reg, roi_xyxy):
    """my_models.py:378-391."""
    xywh = torch.from_numpy(obox.xyxy2xywh(roi_xyxy.numpy()))
    x, y, w, h = xywh.t()
    out = torch.stack((reg[:, 0] * w + x, reg[:, 1] * h + y, torch.exp(reg[:, 2]) * w, torch.exp(reg[:, 3]) * h), 1)
    return torch.from_numpy(obox.xywh2xyxy(out.numpy()))


def network_forward(module_defs, sd, images, maps, radar_boxes, conf_thresh, model_mode=0, refine_threshold_img=0.0,
                    refine_threshold_radar=0.0, class_idx=0, class_num=1, use_torchvision=False, feature_tap=None,
                    return_intermediates=False):
    """Network.forward with targets=None (my_models.py:433-539).

    radar_boxes (n,5) is taken in the caller's 0..1 scale and, like the reference (:491), scaled by the
    image size - on a copy; the in-place side effect on the caller's tensor is the wrapper's concern.
    Returns output (K,8) = [image_i, x1, y1, x2, y2, new_conf, class_score, class_pred].
    """
    feat, yolo_out = darknet_forward(module_defs, sd, images, prefix="base_detector.", feature_tap=feature_tap)
    pred = yolo_out.clone().numpy()
    dets, _ = obox.non_max_suppression_cpp(pred, conf_thresh, use_torchvision=use_torchvision)
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
    if model_mode == 1:
        return img_boxes[:, :8]
    if model_mode == 2:
        refine_threshold_img = 1

    roi_score_map = img_cnn_layers(feat, sd)
    radar_score_map = radar_cnn_layers(maps, sd)
    radar_boxes = radar_boxes.clone().float()
    if len(radar_boxes) > 0:
        radar_boxes[:, 1:] *= images.shape[-1]
    box_locations = torch.cat((img_boxes[:, :5], radar_boxes), 0)
    if use_torchvision:
        from torchvision.ops import ps_roi_align, roi_align
        crop_img = ps_roi_align(roi_score_map, box_locations, (7, 7), spatial_scale=1. / 16)
        crop_radar = roi_align(radar_score_map, box_locations, (7, 7), spatial_scale=1. / 16)
    else:
        crop_img = torch.from_numpy(oroi.ps_roi_align(roi_score_map.numpy(), box_locations.numpy(), 7, 1. / 16))
        crop_radar = torch.from_numpy(oroi.roi_align(radar_score_map.numpy(), box_locations.numpy(), 7, 1. / 16))
    if len(box_locations) == 0:
        crop_img = crop_img.reshape(0, 10, 7, 7)
        crop_radar = crop_radar.reshape(0, 10, 7, 7)
    reg, ref_vec = refinement_head(crop_radar, crop_img, sd)

    radar_rows = torch.cat((radar_boxes, ref_vec[n_img:], torch.zeros((len(radar_boxes), 1)), ref_vec[n_img:, 1:]), -1)
    all_boxes = torch.cat((img_boxes, radar_rows), 0)
    yolo_vec = torch.cat((img_boxes[:, 5:6], img_boxes[:, 8:]), 1)
    masks_img = ensemble_head(ref_vec[:n_img], yolo_vec, sd)
    m = torch.cat((masks_img[:, :1], ref_vec[n_img:, :1]), 0)
    masks = torch.cat((1 - m, m), -1)
    positive = torch.cat((masks[:n_img, 1] > refine_threshold_img, masks[n_img:, 1] > refine_threshold_radar), 0)
    if model_mode != 2:
        new_xyxy = box_regress(reg[positive], all_boxes[positive, 1:5])
    else:
        new_xyxy = all_boxes[positive, 1:5]
    output = torch.cat((all_boxes[positive, :1], new_xyxy, masks[positive, 1:], all_boxes[positive, 6:8]), -1)
    pri = masks.clone()
    pri[n_img:, 1] /= 5
    order = torch.sort(pri[positive, 1], descending=True, stable=True).indices
    output = output[order]
    if return_intermediates:
        return output, dict(feat=feat, yolo_out=yolo_out, img_boxes=img_boxes, box_locations=box_locations,
                            roi_score_map=roi_score_map, radar_score_map=radar_score_map, crop_img=crop_img,
                            crop_radar=crop_radar, reg=reg, ref_vec=ref_vec, masks=masks, positive=positive,
                            all_boxes=all_boxes, n_img=n_img)
    return output


def network_forward_train(module_defs, sd, images, maps, radar_boxes, conf_thresh, targets, sample_filter=None, **kw):
    """Network.forward with targets (my_models.py:545-640), heads in eval mode: the inference forward followed by
    the labelling + loss branch (oracle/stage3_loss.py).  targets (m,6) [image, class, cx, cy, w, h] in 0..1 (a copy is
    converted; the reference rewrites the caller's tensor).  Returns (loss, output, metric, radar_attention, aux)."""
    from . import stage3_loss as s3
    output, im = network_forward(module_defs, sd, images, maps, radar_boxes, conf_thresh, return_intermediates=True, **kw)
    tpx = s3.targets_to_pixels(np.asarray(targets, dtype=np.float32), images.shape[3])
    aux = s3.stage3_losses(im["all_boxes"].numpy(), im["masks"].numpy(), im["ref_vec"].numpy(), im["reg"].numpy(),
                           im["n_img"], im["positive"].numpy(), tpx, sample_filter=sample_filter)
    attention = im["radar_score_map"][:, :1].numpy()
    return aux["loss"], output, aux["metric"], attention, aux


def network_forward_stage2(module_defs, sd, images, conf_thresh, refine_threshold=0.0, class_num=12,
                           use_torchvision=False, return_intermediates=False):
    """Stage-2 Network.forward with targets=None (reference module2_mixed/my_models.py:299-361): every class is
    kept after NMS (:325-330), fcn_layers (= cnn_layers_1), ps_roi_align (:344), refinement_head (:121-126, Dropout
    is the identity in eval), ensemble_head with a LeakyReLU before the softmax (:149-161), new confidence =
    masks[:, 1] (:352), box regression, sort by the new confidence (:358), result on the CPU."""
    feat, yolo_out = darknet_forward(module_defs, sd, images, prefix="base_detector.")
    dets, _ = obox.non_max_suppression_cpp(yolo_out.clone().numpy(), conf_thresh, use_torchvision=use_torchvision)
    rows = []
    for i, d in enumerate(dets):
        if d is not None:
            b = np.zeros((len(d), 8 + class_num), dtype=np.float32)
            b[:, 0] = i
            b[:, 1:] = d
            rows.append(b)
    boxes = torch.from_numpy(np.concatenate(rows, 0)) if rows else torch.empty((0, 8 + class_num))
    roi_score_map = img_cnn_layers(feat, sd, prefix="fcn_layers.")
    if use_torchvision:
        from torchvision.ops import ps_roi_align
        crop = ps_roi_align(roi_score_map, boxes[:, :5], (7, 7), spatial_scale=1. / 16)
    else:
        crop = torch.from_numpy(oroi.ps_roi_align(roi_score_map.numpy(), boxes[:, :5].numpy(), 7, 1. / 16))
    if len(boxes) == 0:
        crop = crop.reshape(0, 10, 7, 7)
    p = "refinement_head."
    t = F.leaky_relu(F.linear(crop.flatten(start_dim=1), sd[p + "net0.0.weight"], sd[p + "net0.0.bias"]), 0.1)
    reg = F.linear(t, sd[p + "net1.0.weight"], sd[p + "net1.0.bias"])
    ref_vec = torch.sigmoid(F.linear(t, sd[p + "net2.0.weight"], sd[p + "net2.0.bias"]))
    yolo_vec = torch.cat((boxes[:, 5:6], boxes[:, 8:]), 1)
    e = "ensemble_head."
    x = torch.stack((ref_vec, yolo_vec), -1)
    x = F.leaky_relu(F.linear(x, sd[e + "fc1.0.weight"], sd[e + "fc1.0.bias"]), 0.1).flatten(start_dim=1)
    x = F.leaky_relu(F.linear(x, sd[e + "fc2.0.weight"], sd[e + "fc2.0.bias"]), 0.1)
    masks = torch.softmax(x, dim=1)
    positive = masks[:, 1] > refine_threshold
    out = torch.cat((boxes[positive, :1], box_regress(reg[positive], boxes[positive, 1:5]), masks[positive, 1:],
                     boxes[positive, 6:8]), -1)
    out = out[torch.sort(out[:, 5], descending=True, stable=True).indices]
    if return_intermediates:
        return out, dict(boxes=boxes, masks=masks, ref_vec=ref_vec, reg=reg, positive=positive)
    return out


def network_forward_stage2_train(module_defs, sd, images, conf_thresh, targets, sample_filter=None, refine_threshold=0.0,
                                 iou_thresh=(0.3, 0.7), alpha=0.75, balance_fac=5, loss_lambda=(15, 5), **kw):
    """Stage-2 Network.forward with targets (reference module2_mixed/my_models.py:363-461), heads in eval mode: labels
    (obtain_iou_labels :195-244, the same routine as stage 3's with multi_boxes=True), pos / neg filters, metric
    (:387-397), balanced sample with python's random (:399-414), FocalLoss over the sampled proposals' masks (:420),
    confidence BCE (:424-429), regression_loss over the positives (:432-435), category BCE over the 12 class scores of the
    positives with the reference's row-i indexing (:438-443), and the total (:445)
        loss = masks + (conf + category) / lambda[0] + (xy + wh) / lambda[1].
    targets (m,6) [image, class, cx, cy, w, h] in 0..1 (a copy is converted).  Returns (output, loss, metric, aux)."""
    from . import stage3_loss as s3
    output, im = network_forward_stage2(module_defs, sd, images, conf_thresh, refine_threshold=refine_threshold,
                                        return_intermediates=True, **kw)
    boxes, masks, ref_vec, reg = (im[k].numpy().astype(np.float32) for k in ("boxes", "masks", "ref_vec", "reg"))
    positive = im["positive"].numpy()
    R = len(boxes)
    tpx = s3.targets_to_pixels(np.asarray(targets, dtype=np.float32), images.shape[-1])
    boxes6 = np.concatenate((boxes[:, :1], boxes[:, 7:8], boxes[:, 1:5]), 1)
    iou_labels, target_location = s3.obtain_iou_labels(boxes6, tpx, True)
    flat = iou_labels.reshape(-1)
    pos, neg = flat > iou_thresh[1], flat < iou_thresh[0]
    if sample_filter is None:
        sample_filter = s3.sample_filter_reference(pos, neg, balance_fac)
    pos_idx = np.where(pos)[0]
    onehot = np.tile(np.array([1.0, 0.0], dtype=np.float32), (R, 1))
    onehot[pos_idx] = np.array([0.0, 1.0], dtype=np.float32)
    masks_loss = s3.focal_loss_sum(masks[sample_filter], onehot[sample_filter], alpha)
    conf_label = np.zeros(R, dtype=np.float32)
    conf_label[pos_idx] = 1
    conf_loss = s3.bce_sum(ref_vec[sample_filter, 0], conf_label[sample_filter])
    loss_xy, loss_wh = s3.regression_loss(reg[pos], target_location[pos], boxes[pos, 1:5])
    class_num = ref_vec.shape[1] - 1
    class_label = np.zeros((R, class_num), dtype=np.float32)
    for i, idx in enumerate(pos_idx):            # row i (not idx) is set: the reference's indexing (:440-441)
        class_label[i, int(boxes6[idx, 1])] = 1
    category_loss = s3.bce_sum(ref_vec[pos, 1:].reshape(-1), class_label[pos].reshape(-1))
    f32 = np.float32
    loss = f32(masks_loss + f32(conf_loss + category_loss) / f32(loss_lambda[0]) + f32(loss_xy + loss_wh) / f32(loss_lambda[1]))
    conf_1, conf_2 = boxes[:, 5], masks[:, 1]
    metric = dict(total=R, true=int(pos.sum()), positive=int(positive.sum()), tp=float((positive & pos).sum()),
                  conf=dict(conf_1_pos=conf_1[flat > 0.5], conf_1_neg=conf_1[flat < 0.5],
                            conf_2_pos=conf_2[flat > 0.5], conf_2_neg=conf_2[flat < 0.5]))
    aux = dict(masks_loss=masks_loss, conf_loss=conf_loss, loss_xy=loss_xy, loss_wh=loss_wh, category_loss=category_loss,
               iou_labels=iou_labels, target_location=target_location, sample_filter=sample_filter, boxes=boxes)
    return output, loss, metric, aux
