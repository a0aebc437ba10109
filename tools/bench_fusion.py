"""Secondary measurements (BASELINE.json configs[0] and configs[2]): YOLOv3-tiny-12 forward and the milliEye fusion
forward (YOLO + R-CNN + radar MLP), batch 32, 416x416, 64 radar points per frame.  Not the headline bench.

    python tools/bench_fusion.py [steps]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from millieye_b200 import configs, radar  # noqa: E402
from millieye_b200.my_models import Network, define_yolo  # noqa: E402
from oracle import synth  # noqa: E402

N, S, PTS = 32, 416, 64


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) / steps * 1e3


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda:0")
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.2).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=0, obj_bias=-2.0, head_gain=0.3)   # bench.FUSION_WEIGHTS)
    model.to(dev)
    imgs = torch.rand(N, 3, S, S, device=dev)
    host_imgs = imgs.cpu().pin_memory()
    rng = np.random.RandomState(0)
    # 64 radar points per frame drawn from the fixture's empirical ranges
    pts = np.stack([rng.uniform(-3, 3, (N, PTS)), rng.uniform(1, 10, (N, PTS)), rng.uniform(-1.5, 1.5, (N, PTS)),
                    rng.uniform(-3, 3, (N, PTS))], -1).astype(np.float32)
    pts_d = torch.from_numpy(pts).to(dev)
    cnt_d = torch.full((N,), PTS, dtype=torch.int32, device=dev)
    cfg = radar.make_cfg(out_size=S // 16)
    boxes = synth.synth_radar_boxes(N, seed=1).to(dev)

    def darknet_only():
        model.base_detector.forward_device(imgs)

    def fusion():
        maps = radar.radar_maps(pts_d, cnt_d, cfg)
        return model(imgs, maps, boxes.clone(), 0)

    def fusion_e2e():
        maps = radar.radar_maps(pts_d, cnt_d, cfg)
        return model(host_imgs, maps, boxes.clone(), 0).cpu()

    ms_t, _ = timed(darknet_only, steps)
    out = fusion()
    ms_f, wall_f = timed(fusion, steps)
    ms_e, wall_e = timed(fusion_e2e, steps)
    print("FUSION " + json.dumps(dict(
        tiny12_forward=dict(ms=round(ms_t, 4), fps=round(N / ms_t * 1e3, 1)),
        fusion_forward=dict(ms=round(ms_f, 4), wall_ms=round(wall_f, 4), fps=round(N / wall_f * 1e3, 1), rows=int(out.shape[0])),
        fusion_e2e_host_images=dict(ms=round(ms_e, 4), wall_ms=round(wall_e, 4), fps=round(N / wall_e * 1e3, 1)),
        batch=N, size=S, radar_points=PTS)), flush=True)


if __name__ == "__main__":
    main()
