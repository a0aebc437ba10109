"""Generates tests/golden/*.npz by running the UNMODIFIED reference (read-only, from
/root/reference) and the installed torchvision on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference imports matplotlib at module scope (utils/utils.py:12-13); it is absent here and is
stubbed.  Darknet-53: the reference's forward raises on yolov3.cfg because no module is named
conv_8, so `featuremap` is pre-seeded and only yolo_outputs is recorded.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MILLIEYE_REFERENCE", "/root/reference")
M3 = os.path.join(REF, "module3_our_dataset")
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        mod = types.ModuleType(name)
        mod.close = lambda *a, **k: None
        sys.modules.setdefault(name, mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.path.insert(0, M3)
    os.chdir(tempfile.mkdtemp())  # the training branch appends to ./b.txt (my_models.py:350)
    from yolov3.models import Darknet
    import my_models
    import utils.utils as ref_utils
    return Darknet, my_models, ref_utils


def main():
    torch.set_num_threads(4)
    Darknet, my_models, ref_utils = import_reference()
    cfg_tiny = os.path.join(M3, "config", "yolov3-tiny-12.cfg")
    cfg_full = os.path.join(M3, "config", "yolov3.cfg")

    # ---- G1: tiny-12 Darknet forward
    with torch.no_grad():
        net = Darknet(cfg_tiny).eval()
        net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=1))
        x = synth.synth_images(2, 96, seed=1)
        feat, yolo = net(x)
        np.savez_compressed(os.path.join(HERE, "darknet_tiny12_96.npz"), featuremap=feat.numpy(), yolo=yolo.numpy(),
                            keys=np.array(list(net.state_dict().keys())))
        print("G1", feat.shape, yolo.shape, float(yolo[..., 4].max()))

    # ---- G2: Darknet-53 forward (featuremap pre-seeded, F1)
    with torch.no_grad():
        net = Darknet(cfg_full).eval()
        net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=2, conv_gain=0.6))
        net.featuremap = torch.empty(0)
        x = synth.synth_images(1, 64, seed=2)
        _, yolo = net(x)
        np.savez_compressed(os.path.join(HERE, "darknet53_64.npz"), yolo=yolo.numpy())
        print("G2", yolo.shape, float(yolo[..., 4].max()), float(yolo[..., :4].abs().max()))

    # ---- G3: fusion Network forward, tiny-12, modes 0 / 1 / 2
    with torch.no_grad():
        out = {}
        for mode in (0, 1, 2):
            model = my_models.Network(my_models.define_yolo(cfg_tiny), conf_thresh=0.05).eval()
            model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=3, obj_bias=-0.5))
            imgs = synth.synth_images(3, 160, seed=3)
            maps = synth.synth_maps(3, 160, seed=3)
            rb = synth.synth_radar_boxes(3, seed=5)
            rb_in = rb.clone()
            res = model(imgs, maps, rb_in, mode)
            out[f"mode{mode}"] = res.numpy()
            if mode == 0:
                out["radar_boxes_after"] = rb_in.numpy()  # the in-place *= image size (my_models.py:491)
                out["keys"] = np.array(list(model.state_dict().keys()))
            print("G3 mode", mode, res.shape)
        np.savez_compressed(os.path.join(HERE, "fusion_tiny12_160.npz"), **out)

    # ---- G4: non_max_suppression_cpp on synthetic decoded tensors (both batched_nms paths)
    out = {}
    for tag, mu, thr in (("trick", -6.0, 0.01), ("vanilla", -1.0, 0.2)):
        pred = synth.synth_predictions(2, 2535, 12, seed=7, conf_mu=mu)
        dets = ref_utils.non_max_suppression_cpp(pred.clone(), conf_thresh=thr)
        for i, d in enumerate(dets):
            out[f"{tag}_{i}"] = np.zeros((0, 19), np.float32) if d is None else d.numpy()
        ncand = [(pred[i, :, 4] >= thr).sum().item() for i in range(2)]
        out[f"{tag}_ncand"] = np.array(ncand)
        print("G4", tag, ncand, [len(out[f"{tag}_{i}"]) for i in range(2)])
    np.savez_compressed(os.path.join(HERE, "nms_cpp.npz"), **out)

    # ---- G5: torchvision RoI ops (third-party arithmetic pinned to the installed 0.26.0)
    from torchvision.ops import ps_roi_align, roi_align
    rng = np.random.RandomState(11)
    feat = torch.from_numpy(rng.randn(2, 490, 26, 26).astype(np.float32))
    rfeat = torch.from_numpy(rng.rand(2, 10, 26, 26).astype(np.float32))
    R = 48
    x1, y1 = rng.uniform(-30, 380, R), rng.uniform(-30, 380, R)
    w, h = rng.uniform(1, 220, R), rng.uniform(1, 220, R)
    rois = torch.from_numpy(np.stack([rng.randint(0, 2, R), x1, y1, x1 + w, y1 + h], 1).astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "roi_ops.npz"), rois=rois.numpy(),
                        ps=ps_roi_align(feat, rois, (7, 7), 1. / 16).numpy(),
                        ra=roi_align(rfeat, rois, (7, 7), 1. / 16).numpy())
    print("G5 done")


if __name__ == "__main__":
    main()
