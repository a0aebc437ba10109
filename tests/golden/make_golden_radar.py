"""Generates tests/golden/radar_maps.npz by running the UNMODIFIED reference radar chain on frames of the
recording shipped with it (module3_our_dataset/data_collection/data/20210305-000127/pointcloud.pkl):
from_3d_to_2d (data_collection/utils/utils.py) -> the FOV/depth/velocity filter of prepare_data.py:108 ->
plot_radar_heatmap + pad_to_square (utils/datasets.py) -> the bilinear resize of collate_fn (:320-322).

    python tests/golden/make_golden_radar.py        (build container only)
"""
import importlib.util
import os
import pickle
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
M3 = os.path.join(os.environ.get("MILLIEYE_REFERENCE", "/root/reference"), "module3_our_dataset")
sys.path.insert(0, ROOT)

from millieye_b200.radar import CALIB_FOV90  # noqa: E402  (the yaml's numbers + load_calib's translation)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        mod = types.ModuleType(name)
        mod.close = lambda *a, **k: None
        sys.modules.setdefault(name, mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for name in ("filterpy", "filterpy.kalman", "serial"):   # tracking.py / ReadRadar.py imports, unused here (F9)
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["filterpy.kalman"].KalmanFilter = object
    sys.path.insert(0, M3)
    import importlib
    dc_utils = importlib.import_module("data_collection.utils.utils")
    import utils.datasets as ref_ds
    from torchvision import transforms

    with open(os.path.join(M3, "data_collection", "data", "20210305-000127", "pointcloud.pkl"), "rb") as fh:
        rec = pickle.load(fh)
    calib = np.array(CALIB_FOV90)
    w, h, max_depth, min_velocity = 640, 480, 50, 0.1
    cap = 128
    frames, counts, maps26, maps32, clouds, kept = [], [], [], [], [], []
    for start in range(0, 400, 17):
        overlay = 1 + (start // 17) % 4                    # 1..4 radar frames overlaid, like prepare_data.py:95-103
        x = y = z = v = np.array([])
        for i in range(start, min(start + overlay, len(rec))):
            d = rec[i]["Data"]
            x, y, z, v = np.append(x, d["x"]), np.append(y, d["y"]), np.append(z, d["z"]), np.append(v, d["velocity"])
        pts = np.array([x, y, z, v])
        uv, xyzv = dc_utils.from_3d_to_2d(pts, calib)
        filt = [0 <= i[0] < w and 0 <= i[1] < h and j[2] < max_depth and abs(j[3]) >= min_velocity for i, j in zip(uv, xyzv)]
        uv, xyzv = uv[filt], xyzv[filt]
        cloud = np.concatenate((uv, xyzv[..., 2:]), -1)
        rmap = transforms.ToTensor()(ref_ds.plot_radar_heatmap(cloud.transpose(), (w, h))).float()
        rmap, _ = ref_ds.pad_to_square(rmap, 0)
        r26 = F.interpolate(rmap.unsqueeze(0), 26, mode="bilinear", align_corners=True).squeeze(0)
        raw = np.zeros((cap, 4), np.float32)
        raw[:pts.shape[1]] = pts.T.astype(np.float32)
        assert np.array_equal(raw[:pts.shape[1]].astype(np.float64), pts.T), "recording is float32 data"
        c = np.zeros((cap, 4), np.float32)
        c[:len(cloud)] = cloud
        frames.append(raw); counts.append(pts.shape[1]); maps26.append(r26.numpy()); maps32.append(rmap.numpy())
        clouds.append(c); kept.append(len(cloud))
    np.savez_compressed(os.path.join(HERE, "radar_maps.npz"), points=np.stack(frames), counts=np.array(counts, np.int32),
                        maps26=np.stack(maps26), maps32=np.stack(maps32), clouds=np.stack(clouds),
                        kept=np.array(kept, np.int32))
    print("radar golden:", len(frames), "frames, points/frame", min(counts), "..", max(counts), "kept", min(kept), "..", max(kept))


if __name__ == "__main__":
    main()
