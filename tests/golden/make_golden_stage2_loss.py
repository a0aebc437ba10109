"""Generates tests/golden/stage2_loss_tiny12_160.npz: the UNMODIFIED stage-2 reference's training branch
(/root/reference/module2_mixed/my_models.py:363-461: obtain_iou_labels, FocalLoss, confidence / category BCE,
regression_loss, total loss and metric) on seeded inputs, model in eval mode (running BatchNorm statistics, Dropout off),
python `random` seeded so the negative sub-sampling (:411) is reproducible.

    python tests/golden/make_golden_stage2_loss.py        (build container only)
"""
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
M2 = os.path.join(os.environ.get("MILLIEYE_REFERENCE", "/root/reference"), "module2_mixed")
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402

SEED_SAMPLING = 11
CONF = 0.6


def make_targets(boxes, size, seed):
    """Ground truth built from the detector's own proposals (same class): copies / small jitters (IoU > 0.7), medium
    shifts (0.3..0.7), unrelated boxes and a class nobody predicts.  Pixel x1y1x2y2 rows [image, class, ...]."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for k, b in enumerate(boxes):
        if k % 9 == 0:
            i, x1, y1, x2, y2 = [float(v) for v in b[:5]]
            w, h = x2 - x1, y2 - y1
            shift = (0.0, 0.04, 0.22)[(k // 9) % 3] * float(0.6 + 0.8 * torch.rand(1, generator=g))
            rows.append([i, float(b[7]), x1 + shift * w, y1 + shift * h, x2 + shift * w, y2 + shift * h])
    for i in range(2):
        c = torch.rand(2, generator=g) * 0.6 + 0.2
        rows.append([float(i), 2.0, float(c[0] * size - 9), float(c[1] * size - 7), float(c[0] * size + 9), float(c[1] * size + 7)])
    return torch.tensor(rows, dtype=torch.float32)


def normalise(t, size):
    out = t.clone()
    out[:, 2] = (t[:, 2] + t[:, 4]) / 2 / size
    out[:, 3] = (t[:, 3] + t[:, 5]) / 2 / size
    out[:, 4] = (t[:, 4] - t[:, 2]) / size
    out[:, 5] = (t[:, 5] - t[:, 3]) / size
    return out


def main():
    global CONF
    torch.set_num_threads(4)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        mod = types.ModuleType(name)
        mod.close = lambda *a, **k: None
        sys.modules.setdefault(name, mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.path.insert(0, M2)
    os.chdir(tempfile.mkdtemp())
    import my_models
    import utils.utils as ref_utils
    cfg = os.path.join(M2, "config", "yolov3-tiny-12.cfg")
    n, size = 3, 160
    model = my_models.Network(my_models.define_yolo(cfg), conf_thresh=CONF).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=-1.0, head_gain=1.0))
    imgs = synth.synth_images(n, size, seed=6)
    with torch.no_grad():
        # the proposals exactly as forward builds them (:317-335)
        feat, yolo = model.base_detector(imgs)
        # The GPU path's detector computes in fp16: keep the confidence threshold clear of every decoded objectness so the
        # same rows pass the filter there (the candidate set, hence NMS and the proposal list, must be identical).
        conf_all = yolo[..., 4].reshape(-1)
        best = max((float((conf_all - c).abs().min()), c) for c in [round(0.55 + 0.005 * k, 3) for k in range(50)])
        print("confidence threshold", best[1], "margin", best[0])
        model.conf_thresh = best[1]
        CONF = best[1]
        dets = ref_utils.non_max_suppression_cpp(yolo.cpu(), conf_thresh=model.conf_thresh)
        rows = []
        for i, d in enumerate(dets):
            if d is not None:
                b = torch.zeros((len(d), 8 + model.class_num))
                b[:, 0] = i
                b[:, 1:] = d
                rows.append(b)
        boxes = torch.cat(rows, 0)
        boxes_cpu = torch.cat((boxes[:, :1], boxes[:, 7:8], boxes[:, 1:5]), 1)
        for seed in range(3, 3000):
            targets = normalise(make_targets(boxes, size, seed), size)
            t_px = targets.clone()
            t_px[:, 2:] = ref_utils.xywh2xyxy(t_px[:, 2:])
            t_px[:, 2:] *= size
            lab, _ = my_models.obtain_iou_labels(boxes_cpu, t_px)
            margin = min(float((lab - thr).abs().min()) for thr in (0.3, 0.5, 0.7))
            if seed < 6:
                print("seed", seed, "boxes", len(boxes), "targets", len(targets), "margin", margin, "pos", int((lab > 0.7).sum()))
            if margin > 0.012 and int((lab > 0.7).sum()) >= 4:
                break
        else:
            raise SystemExit("no target seed gives labels clear of the thresholds")
        random.seed(SEED_SAMPLING)
        t_in = targets.clone()
        output, loss, metric = model(imgs, t_in)
        iou_labels, target_location = my_models.obtain_iou_labels(boxes_cpu, t_in)
    conf = metric["conf"]
    np.savez_compressed(
        os.path.join(HERE, "stage2_loss_tiny12_160.npz"), conf_thresh=np.float32(CONF), targets=targets.numpy(), targets_after=t_in.numpy(),
        loss=np.float32(loss.item()), output=output.numpy(), total=np.int64(metric["total"]),
        true=np.int64(int(metric["true"])), positive=np.int64(int(metric["positive"])), tp=np.float32(float(metric["tp"])),
        conf_1_pos=conf["conf_1_pos"].numpy(), conf_1_neg=conf["conf_1_neg"].numpy(),
        conf_2_pos=conf["conf_2_pos"].numpy(), conf_2_neg=conf["conf_2_neg"].numpy(),
        iou_labels=iou_labels.numpy(), target_location=target_location.numpy(), sampling_seed=np.int64(SEED_SAMPLING))
    print("stage2 loss golden: loss", float(loss), "rows", len(iou_labels), "pos", int((iou_labels > 0.7).sum()),
          "targets", len(targets), "seed", seed, "margin", margin, "output", tuple(output.shape))


if __name__ == "__main__":
    main()
