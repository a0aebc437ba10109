"""Radar pre-processing of the reference's data path, on the device (SURVEY.md §8f, rows f1 + f2).

The reference spreads this over three places, all numpy on the host, per frame:
  data_collection/utils/utils.py:105-120  from_3d_to_2d            radar xyz -> pixel (u, v), float64, int truncation
  data_collection/prepare_data.py:108     FOV / depth / velocity filter
  utils/datasets.py:56-106, 16-26, 320    plot_radar_heatmap -> pad_to_square -> bilinear resize to S/16
`radar_maps` does the whole chain for a batch of frames in one kernel launch (me_radar_maps) and returns the
(N, 3, S/16, S/16) tensor `Network.forward` takes as `maps`.
"""
import numpy as np
import torch

from . import _lib
from ._lib import RadarCfg, check, ptr, stream_ptr

# calib_FOV90.yaml of the reference (camera_matrix fx, cx, fy, cy + plumb-bob k1, k2, t1, t2, k3) followed by the
# radar->camera translation hard-coded in load_calib (data_collection/utils/utils.py:69)
CALIB_FOV90 = (456.937531, 313.081521, 458.889439, 243.930116, 0.096449, -0.103986, 0.000836, 0.001810, 0.0,
               -0.07, -0.05, 0.0)


def load_calib(filename):
    """[fx, cx, fy, cy, k1, k2, t1, t2, k3, tx, ty, tz] from a ROS camera-calibration yaml
    (same vector as the reference's load_calib, utils.py:63-76; uses a safe yaml loader)."""
    import yaml
    with open(filename, "r") as fh:
        y = yaml.safe_load(fh)
    cam = np.resize(y["camera_matrix"]["data"], (3, 3))
    dist = list(y["distortion_coefficients"]["data"])
    return np.array([cam[0, 0], cam[0, 2], cam[1, 1], cam[1, 2], *dist, -0.07, -0.05, 0.0])


def make_cfg(calib=CALIB_FOV90, img_size=(640, 480), max_depth=50.0, min_velocity=0.1, radar_maps_size=32, out_size=26):
    """Host-side constants of plot_radar_heatmap: bin counts use Python's round(), edges are np.linspace in float64
    exactly as np.histogram2d builds them."""
    w, h = img_size
    scale = max(img_size) / radar_maps_size
    bin_w, bin_h = round(w / scale), round(h / scale)
    if max(bin_w, bin_h) > 32:
        raise _lib.MeError("radar maps larger than 32 bins are not supported")
    cfg = RadarCfg()
    for i, v in enumerate(calib):
        cfg.calib[i] = float(v)
    cfg.img_w, cfg.img_h = int(w), int(h)
    cfg.max_depth, cfg.min_velocity = float(max_depth), float(min_velocity)
    cfg.bin_w, cfg.bin_h = int(bin_w), int(bin_h)
    for i, v in enumerate(np.linspace(0, w, bin_w + 1)):
        cfg.edges_w[i] = float(v)
    for i, v in enumerate(np.linspace(0, h, bin_h + 1)):
        cfg.edges_h[i] = float(v)
    cfg.out_size = int(out_size)
    return cfg


def radar_maps(points, counts, cfg=None, return_points=False, **cfg_kwargs):
    """points: (N, P, 4) fp32 CUDA tensor of radar (x, y, z, velocity); counts: (N,) int32 live points per frame.
    Returns maps (N, 3, out, out) fp32 and, with return_points, the filtered (u, v, depth, velocity) rows + counts."""
    if not points.is_cuda:
        raise _lib.MeError("radar_maps runs on the GPU; pass CUDA tensors (no CPU fallback)")
    assert points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 3 and points.shape[2] == 4
    assert counts.dtype == torch.int32 and counts.is_cuda
    cfg = cfg or make_cfg(**cfg_kwargs)
    n, cap, _ = points.shape
    maps = torch.empty((n, 3, cfg.out_size, cfg.out_size), dtype=torch.float32, device=points.device)
    pts = kept = None
    if return_points:
        pts = torch.zeros((n, cap, 4), dtype=torch.float32, device=points.device)
        kept = torch.zeros((n,), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().me_radar_maps(ptr(points), ptr(counts), n, cap, _lib.byref(cfg), ptr(maps), ptr(pts), ptr(kept),
                                       stream_ptr()), "me_radar_maps")
    return (maps, pts, kept) if return_points else maps
