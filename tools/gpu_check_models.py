"""Bring-up check of the model-level path on a B200: golden parity + first timings."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from millieye_b200 import configs  # noqa: E402
from millieye_b200.models import Darknet  # noqa: E402
from millieye_b200.my_models import Network, define_yolo  # noqa: E402
from oracle import synth  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def section(name, fn):
    t0 = time.time()
    try:
        out = fn()
        print("CHECK " + json.dumps(dict(name=name, ok=True, secs=round(time.time() - t0, 2), **out)), flush=True)
    except Exception as e:  # noqa: BLE001
        print("CHECK " + json.dumps(dict(name=name, ok=False, err=repr(e)[:400])), flush=True)
        traceback.print_exc()


def tiny_golden():
    g = np.load(os.path.join(G, "darknet_tiny12_96.npz"))
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
    assert list(net.state_dict().keys()) == list(g["keys"]), "state_dict keys differ from the reference"
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=1))
    net.to(dev)
    x = synth.synth_images(2, 96, seed=1).to(dev)
    feat, yolo = net(x)
    torch.cuda.synchronize()
    yo, yg = yolo.cpu().numpy(), g["yolo"]
    return dict(feat_rel=rel(feat.cpu().numpy(), g["featuremap"]), box_rel=rel(yo[..., :4], yg[..., :4]),
                score_abs=float(np.abs(yo[..., 4:] - yg[..., 4:]).max()), launches=net._plans[next(iter(net._plans))].launches)


def d53_golden():
    g = np.load(os.path.join(G, "darknet53_64.npz"))
    net = Darknet(configs.cfg_path("yolov3")).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=2, conv_gain=0.6))
    net.to(dev)
    x = synth.synth_images(1, 64, seed=2).to(dev)
    _, yolo = net(x)
    torch.cuda.synchronize()
    yo, yg = yolo.cpu().numpy(), g["yolo"]
    relbox = np.abs(yo[..., :4] - yg[..., :4]) / (np.abs(yg[..., :4]) + 1e-6)
    return dict(box_rel_global=rel(yo[..., :4], yg[..., :4]), box_rel_elem_max=float(relbox.max()),
                box_rel_elem_p99=float(np.percentile(relbox, 99)), score_abs=float(np.abs(yo[..., 4:] - yg[..., 4:]).max()))


def fusion_golden():
    g = np.load(os.path.join(G, "fusion_tiny12_160.npz"))
    res = {}
    for mode in (0, 1, 2):
        model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.05).eval()
        if mode == 0:
            assert list(model.state_dict().keys()) == list(g["keys"]), "Network state_dict keys differ"
        model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=3, obj_bias=-0.5))
        model.to(dev)
        imgs = synth.synth_images(3, 160, seed=3).to(dev)
        maps = synth.synth_maps(3, 160, seed=3).to(dev)
        rb = synth.synth_radar_boxes(3, seed=5).to(dev)
        out = model(imgs, maps, rb, mode).cpu().numpy()
        ref = g[f"mode{mode}"]
        res[f"mode{mode}_shape"] = [list(out.shape), list(ref.shape)]
        if out.shape == ref.shape and out.size:
            res[f"mode{mode}_idx_equal"] = bool((out[:, 0] == ref[:, 0]).all() and (out[:, 7] == ref[:, 7]).all())
            res[f"mode{mode}_box_rel"] = rel(out[:, 1:5], ref[:, 1:5])
            res[f"mode{mode}_score_abs"] = float(np.abs(out[:, 5:7] - ref[:, 5:7]).max())
        if mode == 0:
            res["radar_scaled_inplace"] = bool(np.allclose(rb.cpu().numpy(), g["radar_boxes_after"]))
    return res


def timing(cfg, n, size, gain, iters=10):
    net = Darknet(configs.cfg_path(cfg)).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=4, conv_gain=gain))
    net.to(dev)
    x = torch.rand(n, 3, size, size, device=dev)
    out = {}
    for graph in (False, True):
        net.use_cuda_graph = graph
        net._plans = {}
        for _ in range(3):
            net.forward_device(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            net.forward_device(x)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / iters * 1e3
        ms = e0.elapsed_time(e1) / iters
        out["graph" if graph else "eager"] = dict(ms=round(ms, 3), wall_ms=round(wall, 3), fps=round(n / ms * 1e3, 1))
    plan = net._plans[next(iter(net._plans))]
    out["launches"] = plan.launches
    out["mem_gb"] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
    return out


if __name__ == "__main__":
    section("tiny_golden", tiny_golden)
    section("d53_golden", d53_golden)
    section("fusion_golden", fusion_golden)
    section("time_tiny_b32", lambda: timing("yolov3-tiny-12", 32, 416, 1.0))
    section("time_d53_b32", lambda: timing("yolov3", 32, 416, 0.6))
