"""Oracle restatement of the reference's radar pre-processing chain (numpy float64, like the reference).

Test infrastructure only (see oracle/__init__.py).
  projection_xyr_to_uv / from_3d_to_2d   data_collection/utils/utils.py:81-120
  FOV / depth / velocity filter          data_collection/prepare_data.py:108
  plot_radar_heatmap                     utils/datasets.py:56-106 (np.histogram2d, the same library call)
  pad_to_square                          utils/datasets.py:16-26
  bilinear resize (align_corners=True)   utils/datasets.py:320-322 (torch F.interpolate, the same library call)
"""
import numpy as np
import torch
import torch.nn.functional as F


def projection_xyr_to_uv(points, calib):
    fx, cx, fy, cy, k1, k2, t1, t2, k3, tx, ty, tz = calib
    x, y = (points[0] + tx) / (points[2] + tz), (points[1] + ty) / (points[2] + tz)
    x_2, y_2 = x ** 2, y ** 2
    r_2 = x_2 + y_2
    r_4 = r_2 ** 2
    r_6 = r_2 ** 3
    tmp = 1 + k1 * r_2 + k2 * r_4 + k3 * r_6
    xu = x * tmp + 2 * t1 * x * y + t2 * (r_2 + 2 * x_2)
    yu = y * tmp + 2 * t2 * x * y + t1 * (r_2 + 2 * y_2)
    return xu * fx + cx, yu * fy + cy


def from_3d_to_2d(points, calib):
    """points (4, n) radar x, y, z, v -> uv (n, 2) int64, xyzV (n, 4) camera coordinates."""
    x, y, z, vel = points[0], -points[2], points[1], points[3]
    u, v = projection_xyr_to_uv([x, y, z], calib)
    uv = np.stack((u, v), 1).astype(np.int64).reshape(-1, 2)
    xyzv = np.stack((x, y, z + calib[-1], vel), 1).reshape(-1, 4)
    return uv, xyzv


def filter_points(uv, xyzv, img_size=(640, 480), max_depth=50.0, min_velocity=0.1):
    keep = np.array([0 <= i[0] < img_size[0] and 0 <= i[1] < img_size[1] and j[2] < max_depth and abs(j[3]) >= min_velocity
                     for i, j in zip(uv, xyzv)], dtype=bool).reshape(-1)
    return np.concatenate((uv[keep], xyzv[keep][:, 2:]), -1)  # rows (u, v, depth, velocity)


def plot_radar_heatmap(points, img_size, radar_maps_size=32):
    """points (4, k): u, v, depth, velocity -> (bin_h, bin_w, 3) float64 in 0..1."""
    scale = max(img_size) / radar_maps_size
    bin_w, bin_h = round(img_size[0] / scale), round(img_size[1] / scale)
    rng = [[0, img_size[0]], [0, img_size[1]]]
    h0 = np.histogram2d(x=points[0, :], y=points[1, :], bins=[bin_w, bin_h], range=rng)[0].T
    h1 = np.histogram2d(x=points[0, :], y=points[1, :], bins=[bin_w, bin_h], range=rng, weights=points[2, :])[0].T
    h1 /= (h0 + 1e-6)
    h1 = np.where(h1 < 1, 100, h1)
    h2 = np.histogram2d(x=points[0, :], y=points[1, :], bins=[bin_w, bin_h], range=rng, weights=points[3, :])[0].T
    h2 = np.absolute(h2 / (h0 + 1e-6))
    maps = np.stack((h0, h1, h2), axis=-1)
    for i, (lo, hi) in enumerate(((0, 5), (12, 0), (0, 4))):
        maps[..., i] = np.clip((maps[..., i] - lo) / (hi - lo), 0, 1)
    return maps


def pad_to_square(img, pad_value=0.0):
    c, h, w = img.shape
    diff = abs(h - w)
    p1, p2 = diff // 2, diff - diff // 2
    pad = (0, 0, p1, p2) if h <= w else (p1, p2, 0, 0)
    return F.pad(img, pad, "constant", value=pad_value)


def radar_maps(points_xyzv, calib, img_size=(640, 480), max_depth=50.0, min_velocity=0.1, out_size=26):
    """One frame: radar points (n, 4) float -> (3, out_size, out_size) float32 network input + the filtered list."""
    pts = np.asarray(points_xyzv, dtype=np.float64).T.reshape(4, -1)
    uv, xyzv = from_3d_to_2d(pts, np.asarray(calib, dtype=np.float64))
    cloud = filter_points(uv, xyzv, img_size, max_depth, min_velocity)
    hm = plot_radar_heatmap(cloud.transpose().astype(np.float64), img_size)
    t = torch.from_numpy(hm.transpose(2, 0, 1).copy()).float()   # transforms.ToTensor() on a float HWC array
    t = pad_to_square(t, 0)
    if t.shape[-1] != out_size:
        t = F.interpolate(t.unsqueeze(0), out_size, mode="bilinear", align_corners=True).squeeze(0)
    return t, cloud
