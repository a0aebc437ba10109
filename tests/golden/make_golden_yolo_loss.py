"""Generates tests/golden/yolo_loss_tiny12_160.npz: the UNMODIFIED reference's Darknet.forward(x, targets)
(module3_our_dataset/yolov3/models.py:247-267 with the YOLOLayer loss branch :180-232 and build_targets,
utils/utils.py:381-440) on seeded inputs - total loss, the per-layer metrics dictionaries and the outputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_yolo_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import M3, import_reference  # noqa: E402
from oracle import synth  # noqa: E402

METRIC_KEYS = ("loss", "x", "y", "w", "h", "conf", "cls", "cls_acc", "recall50", "recall75", "precision", "conf_obj",
               "conf_noobj", "grid_size")


def make_targets(n, seed):
    """(m,6) [image, class, cx, cy, w, h] in 0..1: a few boxes per image, one duplicated cell, sizes from tiny to large."""
    rng = np.random.RandomState(seed)
    rows = []
    for i in range(n):
        for _ in range(3 + i):
            wh = rng.uniform(0.04, 0.7, 2)
            c = rng.uniform(wh / 2, 1 - wh / 2)
            rows.append([i, rng.randint(0, 12), c[0], c[1], wh[0], wh[1]])
    rows.append(list(rows[1]))          # two targets in the same cell with the same best anchor: the last one wins
    rows[-1][1] = (rows[-1][1] + 5) % 12
    return torch.tensor(rows, dtype=torch.float32)


def main():
    torch.set_num_threads(4)
    Darknet, _, _ = import_reference()
    cfg = os.path.join(M3, "config", "yolov3-tiny-12.cfg")
    n, size = 3, 160
    with torch.no_grad():
        net = Darknet(cfg).eval()
        net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=8, obj_bias=-1.0))
        x = synth.synth_images(n, size, seed=8)
        targets = make_targets(n, seed=9)
        loss, feat, yolo = net(x, targets.clone())
        metrics = [m[0].metrics for m in net.module_list if hasattr(m[0], "metrics")]
    # (featuremap / decoded outputs of this model are pinned by darknet_tiny12_96.npz; here every 97th value is enough)
    out = dict(targets=targets.numpy(), loss=np.float32(loss.item()), yolo_sample=yolo.numpy().reshape(-1)[::97].copy())
    for li, m in enumerate(metrics):
        out[f"metrics{li}"] = np.array([float(m[k]) for k in METRIC_KEYS], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "yolo_loss_tiny12_160.npz"), metric_keys=np.array(METRIC_KEYS), **out)
    print("yolo loss golden: loss", float(loss), "layers", len(metrics), [round(m["loss"], 4) for m in metrics], "targets", len(targets))


if __name__ == "__main__":
    main()
