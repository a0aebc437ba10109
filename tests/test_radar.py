"""Radar pre-processing (SURVEY.md §8f rows f1 + f2): oracle vs the reference-generated fixture (CPU) and the
device kernel vs both (GPU).  The fixture comes from the recording shipped with the reference
(tests/golden/make_golden_radar.py)."""
import os

import numpy as np
import pytest
import torch

from millieye_b200.radar import CALIB_FOV90, make_cfg
from oracle import radar as orad


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "radar_maps.npz"))


def test_oracle_matches_reference(golden_dir):
    g = _golden(golden_dir)
    assert len(g["counts"]) == 24 and g["counts"].max() > 64 and g["kept"].min() >= 1
    for i, n in enumerate(g["counts"]):
        t26, cloud = orad.radar_maps(g["points"][i, :n], CALIB_FOV90)
        assert len(cloud) == g["kept"][i]
        np.testing.assert_array_equal(cloud.astype(np.float32), g["clouds"][i, :len(cloud)])   # integer pixels: exact
        np.testing.assert_array_equal(t26.numpy(), g["maps26"][i])
        t32, _ = orad.radar_maps(g["points"][i, :n], CALIB_FOV90, out_size=32)
        np.testing.assert_array_equal(t32.numpy(), g["maps32"][i])


def test_cfg_matches_numpy_histogram_edges():
    cfg = make_cfg()
    assert (cfg.bin_w, cfg.bin_h) == (32, 24)
    assert list(cfg.edges_w)[:33] == list(np.linspace(0, 640, 33))
    assert list(cfg.edges_h)[:25] == list(np.linspace(0, 480, 25))
    cfg = make_cfg(img_size=(1600, 900))
    assert (cfg.bin_w, cfg.bin_h) == (32, 18)


@pytest.mark.gpu
@pytest.mark.parametrize("out_size", [26, 32, 20])
def test_radar_maps_kernel(golden_dir, out_size):
    from millieye_b200.radar import radar_maps
    g = _golden(golden_dir)
    dev = torch.device("cuda:0")
    pts = torch.from_numpy(g["points"]).to(dev)
    cnt = torch.from_numpy(g["counts"]).to(dev)
    maps, cloud, kept = radar_maps(pts, cnt, return_points=True, out_size=out_size)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(kept.cpu().numpy(), g["kept"])
    for i, k in enumerate(g["kept"]):
        np.testing.assert_array_equal(cloud[i, :k].cpu().numpy(), g["clouds"][i, :k])          # exact pixel coordinates
    m = maps.cpu().numpy()
    if out_size == 32:
        np.testing.assert_array_equal(m, g["maps32"])                                          # fp64 histogram path: exact
    else:
        ref = g["maps26"] if out_size == 26 else np.stack(
            [orad.radar_maps(g["points"][i, :n], CALIB_FOV90, out_size=out_size)[0].numpy() for i, n in enumerate(g["counts"])])
        np.testing.assert_allclose(m, ref, rtol=0, atol=2e-6)                                  # fp32 bilinear weights


@pytest.mark.gpu
def test_radar_maps_empty_and_out_of_view():
    from millieye_b200.radar import radar_maps
    dev = torch.device("cuda:0")
    pts = torch.zeros((3, 16, 4), device=dev)
    pts[1, :4] = torch.tensor([[0.0, 5.0, 0.0, 1.0], [100.0, 1.0, 0.0, 1.0], [0.0, 60.0, 0.0, 1.0], [0.2, 4.0, 0.1, 0.01]])
    cnt = torch.tensor([0, 4, 16], dtype=torch.int32, device=dev)
    maps, cloud, kept = radar_maps(pts, cnt, return_points=True)
    ref = [orad.radar_maps(pts[i, :int(cnt[i])].cpu().numpy(), CALIB_FOV90) for i in range(3)]
    for i in range(2):   # frame 2 is all-zero points: depth 0 -> division by zero -> NaN pixels, dropped by the filter
        assert int(kept[i]) == len(ref[i][1])
        np.testing.assert_allclose(maps[i].cpu().numpy(), ref[i][0].numpy(), atol=2e-6)
    assert int(kept[0]) == 0 and float(maps[0, 0].abs().max()) == 0.0
    assert float(maps[0, 1].min()) == 0.0 and int(kept[1]) == 1   # only the in-view, moving, near point survives
