"""YOLO training loss on a B200 (`-m gpu`): me_yolo_loss on the oracle's own fp32 head logits (loss
and every metric to 1e-5, including two targets that land on the same cell with the same anchor), and
Darknet.forward(x, targets) end to end against the reference-generated fixture (fp16 detector: 1e-2)."""
import os

import numpy as np
import pytest
import torch

from millieye_b200 import configs, ops
from millieye_b200.models import Darknet
from millieye_b200.parse_config import parse_model_config
from oracle import darknet as odark
from oracle import synth
from oracle import yolo_loss as yl

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _setup(golden_dir):
    g = np.load(os.path.join(golden_dir, "yolo_loss_tiny12_160.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Darknet(cfg).state_dict(), seed=8, obj_bias=-1.0)
    sdf = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    return g, cfg, sd, sdf, synth.synth_images(3, 160, seed=8), torch.from_numpy(g["targets"])


def test_yolo_loss_kernel_on_identical_logits(golden_dir):
    g, cfg, sd, sdf, x, targets = _setup(golden_dir)
    md = parse_model_config(cfg)
    with torch.no_grad():
        _, _, outs = odark.darknet_forward(md, sdf, x, collect=True)
    _, blocks = odark.layer_plan(md)
    keys = [str(k) for k in g["metric_keys"]]
    li = 0
    t_dev = targets.float().contiguous().to(DEV)
    for i, blk in enumerate(blocks):
        if blk["type"] != "yolo":
            continue
        logits = outs[i - 1]                                          # (N, A*(5+C), G, G) fp32
        n, c, gs, _ = logits.shape
        nhwc = logits.permute(0, 2, 3, 1).contiguous().to(DEV)
        out = torch.zeros(14, dtype=torch.float32, device=DEV)
        ws = torch.empty(int(ops._lib.lib().me_yolo_loss_workspace(n, gs, len(blk["anchors"]))), dtype=torch.uint8, device=DEV)
        for _ in range(2):                                            # the workspace is reset by every call
            ops.yolo_loss(nhwc, c, n, gs, blk["anchors"], blk["classes"], 160 / gs, t_dev, out, ws)
        got = dict(zip(ops.METRIC_KEYS, out.cpu().tolist()))
        with torch.no_grad():
            ref_total, ref = yl.yolo_layer_loss(logits, blk["anchors"], blk["classes"], 160, targets)
        for k, r in zip(keys, g[f"metrics{li}"]):
            assert abs(got[k] - r) <= 1e-5 * max(1.0, abs(r)), (li, k, got[k], r)      # the reference's own numbers
            assert abs(got[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), (li, k)
        li += 1
    assert li == 2


def test_darknet_forward_with_targets(golden_dir):
    g, cfg, sd, sdf, x, targets = _setup(golden_dir)
    net = Darknet(cfg).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    loss, feat, yolo = net(x.to(DEV), targets.clone())
    assert feat.shape == (3, 256, 10, 10) and yolo.shape[0] == 3
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    assert rel <= 1e-2, rel
    keys = [str(k) for k in g["metric_keys"]]
    for li, layer in enumerate(net.yolo_layers):
        for k, r in zip(keys, g[f"metrics{li}"]):
            tol = 2e-2 * max(1.0, abs(r)) if k not in ("grid_size",) else 0
            assert abs(layer.metrics[k] - r) <= tol, (li, k, layer.metrics[k], r)
    # no targets: the reference's inference signature
    feat2, yolo2 = net(x.to(DEV))
    assert torch.equal(yolo2, yolo)
    # an empty target list gives NaN means like torch's mean over an empty selection, not an exception
    loss0, _, _ = net(x.to(DEV), torch.zeros((0, 6)))
    assert torch.isnan(loss0)
