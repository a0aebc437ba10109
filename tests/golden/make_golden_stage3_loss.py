"""Generates tests/golden/stage3_loss_tiny12_192.npz: the UNMODIFIED reference's stage-3 labelling + loss branch
(module3_our_dataset/my_models.py:545-640 with obtain_iou_labels :317-375, FocalLoss :287-314, regression_loss
:394-408) run on seeded inputs, heads in eval mode (running BatchNorm statistics), python `random` seeded so the
negative sub-sampling (:600) is reproducible.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_stage3_loss.py
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import M3, import_reference  # noqa: E402
from oracle import synth  # noqa: E402

SEED_SAMPLING = 7


def make_targets(img_rows, radar_px, size, seed):
    """Ground truth built from the detector's own proposals so that every label class occurs: exact copies and
    small jitters (IoU > 0.7), medium shifts (0.3..0.7), and unrelated boxes; class 0; pixel x1y1x2y2."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for k, b in enumerate(img_rows):
        if k % 9 == 0:
            i, x1, y1, x2, y2 = [float(v) for v in b[:5]]
            w, h = x2 - x1, y2 - y1
            shift = (0.0, 0.03, 0.25)[(k // 9) % 3] * float(0.8 + 0.4 * torch.rand(1, generator=g))
            rows.append([i, 0.0, x1 + shift * w, y1 + shift * h, x2 + shift * w, y2 + shift * h])
    for k, b in enumerate(radar_px):
        if k % 2 == 0:
            rows.append([float(b[0]), 0.0, float(b[1]) + 1.0, float(b[2]) + 1.0, float(b[3]) + 2.0, float(b[4]) + 1.0])
    for i in range(3):
        c = torch.rand(2, generator=g) * 0.6 + 0.2
        rows.append([float(i), 0.0, float(c[0] * size - 9), float(c[1] * size - 7), float(c[0] * size + 9), float(c[1] * size + 7)])
    rows.append([1.0, 3.0, 10.0, 10.0, 60.0, 60.0])   # a class no proposal predicts
    return torch.tensor(rows, dtype=torch.float32)


def normalise(t, size):
    out = t.clone()
    out[:, 2] = (t[:, 2] + t[:, 4]) / 2 / size
    out[:, 3] = (t[:, 3] + t[:, 5]) / 2 / size
    out[:, 4] = (t[:, 4] - t[:, 2]) / size
    out[:, 5] = (t[:, 5] - t[:, 3]) / size
    return out


def main():
    torch.set_num_threads(4)
    _, my_models, ref_utils = import_reference()
    cfg = os.path.join(M3, "config", "yolov3-tiny-12.cfg")
    n, size = 4, 192
    model = my_models.Network(my_models.define_yolo(cfg), conf_thresh=0.02).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=2.0))
    imgs, maps = synth.synth_images(n, size, seed=6), synth.synth_maps(n, size, seed=6)
    rb = synth.synth_radar_boxes(n, seed=5)
    with torch.no_grad():
        img_rows = model(imgs, maps, rb.clone(), 1)           # yolo-only rows [i,x1,y1,x2,y2,conf,cls_conf,cls_pred]
        radar_px = rb.clone()
        radar_px[:, 1:] *= size
        boxes_cpu = torch.cat((torch.cat((img_rows[:, :1], img_rows[:, 7:8], img_rows[:, 1:5]), 1),
                               torch.cat((radar_px[:, :1], torch.zeros(len(radar_px), 1), radar_px[:, 1:]), 1)), 0)
        # The GPU path computes its proposals from fp16 convolutions, so a label that sits on a threshold could fall
        # on the other side there: pick the first target seed whose labels all keep 0.015 clear of 0.3 / 0.5 / 0.7.
        for seed in range(11, 3000):
            targets = normalise(make_targets(img_rows, radar_px, size, seed), size)
            t_px = targets.clone()
            t_px[:, 2:] = ref_utils.xywh2xyxy(t_px[:, 2:])
            t_px[:, 2:] *= size
            lab, _ = my_models.obtain_iou_labels(boxes_cpu, t_px, model.iou_thresh)
            margin = min(float((lab - thr).abs().min()) for thr in (0.3, 0.5, 0.7))
            if margin > 0.015:
                break
        else:
            raise SystemExit("no target seed gives labels clear of the thresholds")
        print("target seed", seed, "margin", margin)
        random.seed(SEED_SAMPLING)
        t_in = targets.clone()
        loss, output, metric, attention = model(imgs, maps, rb.clone(), 0, t_in)
        # the labels themselves (not returned by forward): same call the forward makes (:555)
        iou_labels, target_location = my_models.obtain_iou_labels(boxes_cpu, t_in, model.iou_thresh)
    conf = metric["conf"]
    np.savez_compressed(
        os.path.join(HERE, "stage3_loss_tiny12_192.npz"), targets=targets.numpy(), targets_after=t_in.numpy(),
        loss=np.float32(loss.item()), output=output.numpy(), total=np.int64(metric["total"]),
        true=np.int64(int(metric["true"])), positive=np.int64(int(metric["positive"])), tp=np.float32(float(metric["tp"])),
        conf_1_pos=conf["conf_1_pos"].numpy(), conf_1_neg=conf["conf_1_neg"].numpy(),
        conf_2_pos=conf["conf_2_pos"].numpy(), conf_2_neg=conf["conf_2_neg"].numpy(),
        radar_attention=attention.numpy(), iou_labels=iou_labels.numpy(), target_location=target_location.numpy(),
        sampling_seed=np.int64(SEED_SAMPLING))
    pos = int((iou_labels > 0.7).sum())
    mid = int(((iou_labels >= 0.3) & (iou_labels <= 0.7)).sum())
    print("stage3 loss golden: loss", float(loss), "rows", len(iou_labels), "pos", pos, "mid", mid, "targets", len(targets),
          "output", tuple(output.shape), "metric", metric["total"], int(metric["true"]), int(metric["positive"]), float(metric["tp"]))


if __name__ == "__main__":
    main()
