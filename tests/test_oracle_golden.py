"""Pins the oracle (oracle/) against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from the unmodified reference) and against the installed torchvision for
the third-party box ops.  CPU only."""
import os

import numpy as np
import pytest
import torch

from millieye_b200 import configs
from oracle import boxes as obox
from oracle import darknet as odark
from oracle import fusion as ofus
from oracle import roi as oroi
from oracle import synth
from oracle.parse_config import parse_model_config


def _tiny_sd(seed, **kw):
    from millieye_b200.models import Darknet
    net = Darknet(configs.cfg_path("yolov3-tiny-12"))
    return synth.fill_state_dict(net.state_dict(), seed=seed, **kw)


def test_darknet_tiny_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "darknet_tiny12_96.npz"))
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    sd = _tiny_sd(1)
    assert list(sd.keys()) == list(g["keys"])
    with torch.no_grad():
        feat, yolo = odark.darknet_forward(md, sd, synth.synth_images(2, 96, seed=1))
    np.testing.assert_allclose(feat.numpy(), g["featuremap"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(yolo.numpy(), g["yolo"], rtol=1e-5, atol=1e-5)


def test_darknet53_matches_reference(golden_dir):
    from millieye_b200.models import Darknet
    g = np.load(os.path.join(golden_dir, "darknet53_64.npz"))
    md = parse_model_config(configs.cfg_path("yolov3"))
    sd = synth.fill_state_dict(Darknet(configs.cfg_path("yolov3")).state_dict(), seed=2, conv_gain=0.6)
    with torch.no_grad():
        feat, yolo = odark.darknet_forward(md, sd, synth.synth_images(1, 64, seed=2))
    assert feat is None  # block 8 of yolov3.cfg is a shortcut: no reference featuremap
    np.testing.assert_allclose(yolo.numpy(), g["yolo"], rtol=2e-5, atol=2e-5)


def test_yolo_row_order_is_anchor_major():
    """Rows are anchor-major, then gy, then gx (reference models.py:142-177)."""
    x = torch.zeros(1, 3 * 17, 2, 2)
    x[0, 17 * 1 + 0, 1, 0] = 10.0  # anchor 1, tx at gy=1, gx=0
    out = odark.yolo_decode(x, [(10, 14), (23, 27), (37, 58)], 12, 64)
    row = 1 * 4 + 1 * 2 + 0
    assert out[0, row, 0] == pytest.approx((1 / (1 + np.exp(-10.0)) + 0) * 32, rel=1e-6)
    assert out[0, row, 1] == pytest.approx((0.5 + 1) * 32)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fusion_matches_reference(golden_dir, mode):
    from millieye_b200.my_models import Network, define_yolo
    g = np.load(os.path.join(golden_dir, "fusion_tiny12_160.npz"))
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.05)
    sd = synth.fill_state_dict(model.state_dict(), seed=3, obj_bias=-0.5)
    assert list(sd.keys()) == list(g["keys"])
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    with torch.no_grad():
        out = ofus.network_forward(md, sd, synth.synth_images(3, 160, seed=3), synth.synth_maps(3, 160, seed=3),
                                   synth.synth_radar_boxes(3, seed=5), conf_thresh=0.05, model_mode=mode)
    ref = g[f"mode{mode}"]
    assert tuple(out.shape) == ref.shape
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("tag,mu,thr", [("trick", -6.0, 0.01), ("vanilla", -1.0, 0.2)])
def test_nms_cpp_matches_reference(golden_dir, tag, mu, thr):
    g = np.load(os.path.join(golden_dir, "nms_cpp.npz"))
    pred = synth.synth_predictions(2, 2535, 12, seed=7, conf_mu=mu).numpy()
    ncand = [(pred[i, :, 4] >= np.float32(thr)).sum() for i in range(2)]
    assert list(g[f"{tag}_ncand"]) == ncand
    assert (ncand[0] > 1000) == (tag == "vanilla")  # the two batched_nms code paths
    dets, _ = obox.non_max_suppression_cpp(pred.copy(), thr)
    for i in range(2):
        np.testing.assert_array_equal(dets[i], g[f"{tag}_{i}"])  # bit-exact rows, same order


@pytest.mark.parametrize("n", [0, 1, 17, 300, 1001, 1500])
def test_batched_nms_matches_torchvision(n):
    from torchvision.ops import batched_nms
    rng = np.random.RandomState(n)
    xy = rng.uniform(0, 400, (n, 2))
    wh = rng.uniform(4, 150, (n, 2))
    b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    s = rng.rand(n).astype(np.float32)
    lab = rng.randint(0, 12, n).astype(np.float32)
    mine = obox.batched_nms(b, s, lab, 0.5)
    theirs = batched_nms(torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(lab), 0.5).numpy()
    np.testing.assert_array_equal(mine, theirs)


def test_nms_ties_and_threshold_edges():
    from torchvision.ops import nms
    # equal scores: stable order keeps the lower index first; IoU exactly 0.5 is NOT suppressed (strict >)
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [0, 0, 10, 5], [20, 20, 30, 30]], dtype=np.float32)
    s = np.array([0.5, 0.5, 0.5, 0.9], dtype=np.float32)
    mine = obox.nms(b, s, 0.5)
    theirs = nms(torch.from_numpy(b), torch.from_numpy(s), 0.5).numpy()
    np.testing.assert_array_equal(mine, theirs)
    assert list(mine) == [3, 0, 2]


def test_roi_ops_match_torchvision(golden_dir):
    g = np.load(os.path.join(golden_dir, "roi_ops.npz"))
    rng = np.random.RandomState(11)
    feat = rng.randn(2, 490, 26, 26).astype(np.float32)
    rfeat = rng.rand(2, 10, 26, 26).astype(np.float32)
    np.testing.assert_allclose(oroi.ps_roi_align(feat, g["rois"]), g["ps"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(oroi.roi_align(rfeat, g["rois"]), g["ra"], rtol=0, atol=2e-6)


def test_roi_empty_and_degenerate():
    from torchvision.ops import roi_align
    feat = np.random.RandomState(0).rand(1, 10, 26, 26).astype(np.float32)
    assert oroi.roi_align(feat, np.zeros((0, 5), np.float32)).shape == (0, 10, 7, 7)
    rois = np.array([[0, 50, 50, 50, 50], [0, 400, 400, 500, 500], [0, -40, -40, -20, -20]], dtype=np.float32)
    ref = roi_align(torch.from_numpy(feat), torch.from_numpy(rois), (7, 7), 1 / 16.).numpy()
    np.testing.assert_allclose(oroi.roi_align(feat, rois), ref, atol=2e-6)


def test_conv_flops_match_survey():
    md = parse_model_config(configs.cfg_path("yolov3"))
    assert odark.conv_flops(md, 416) / 1e9 == pytest.approx(65.864, abs=0.01)
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    assert odark.conv_flops(md, 416) / 1e9 == pytest.approx(5.459, abs=0.01)


def test_stage2_matches_reference(golden_dir):
    """module2_mixed Network.forward (YOLO + R-CNN refinement, all 12 classes kept)."""
    from millieye_b200.my_models_stage2 import Network, define_yolo
    g = np.load(os.path.join(golden_dir, "stage2_tiny12_160.npz"))
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.3)
    sd = synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=-1.0, head_gain=1.0)
    assert list(sd.keys()) == list(g["keys"])
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    with torch.no_grad():
        out = ofus.network_forward_stage2(md, sd, synth.synth_images(2, 160, seed=6), 0.3)
    assert tuple(out.shape) == g["out"].shape and out.shape[0] > 100
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=2e-4, atol=2e-4)


def test_stage2_loss_matches_reference(golden_dir):
    """Stage-2 training branch (module2_mixed/my_models.py:363-461): the oracle's loss, metric and labels equal the
    unmodified reference's on the fixture of tests/golden/make_golden_stage2_loss.py (random seeded like there)."""
    import random
    from millieye_b200.my_models_stage2 import Network as Network2
    from millieye_b200.my_models import define_yolo
    g = np.load(os.path.join(golden_dir, "stage2_loss_tiny12_160.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Network2(define_yolo(cfg), conf_thresh=float(g["conf_thresh"])).state_dict(), seed=6, obj_bias=-1.0,
                               head_gain=1.0)
    random.seed(int(g["sampling_seed"]))
    with torch.no_grad():
        out, loss, metric, aux = ofus.network_forward_stage2_train(
            parse_model_config(cfg), {k: v.float() if v.is_floating_point() else v for k, v in sd.items()},
            synth.synth_images(3, 160, seed=6), float(g["conf_thresh"]), g["targets"], use_torchvision=True)
    assert np.array_equal(aux["iou_labels"], g["iou_labels"]) and np.array_equal(aux["target_location"], g["target_location"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert (metric["total"], metric["true"], metric["positive"], metric["tp"]) == (int(g["total"]), int(g["true"]),
                                                                                  int(g["positive"]), float(g["tp"]))
    for k in ("conf_1_pos", "conf_1_neg", "conf_2_pos", "conf_2_neg"):
        assert np.abs(metric["conf"][k] - g[k]).max() <= 1e-5
    assert out.shape == g["output"].shape and np.abs(out.numpy() - g["output"]).max() <= 1e-3


def test_stage3_loss_matches_reference(golden_dir):
    """Labelling + loss branch (my_models.py:545-640): oracle == reference on the fixture generated by
    tests/golden/make_golden_stage3_loss.py - labels bit-exact, loss / output / attention to fp32 round-off."""
    import random
    from millieye_b200.my_models import Network, define_yolo
    from oracle import stage3_loss as s3
    g = np.load(os.path.join(golden_dir, "stage3_loss_tiny12_192.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Network(define_yolo(cfg), conf_thresh=0.02).state_dict(), seed=6, obj_bias=2.0)
    imgs, maps = synth.synth_images(4, 192, seed=6), synth.synth_maps(4, 192, seed=6)
    random.seed(int(g["sampling_seed"]))
    with torch.no_grad():
        loss, out, metric, att, aux = ofus.network_forward_train(
            parse_model_config(cfg), {k: v.float() for k, v in sd.items()}, imgs, maps, synth.synth_radar_boxes(4, seed=5),
            0.02, g["targets"])
    assert np.array_equal(s3.targets_to_pixels(g["targets"], 192), g["targets_after"])
    assert np.array_equal(aux["iou_labels"], g["iou_labels"]) and np.array_equal(aux["target_location"], g["target_location"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert (metric["total"], metric["true"], metric["positive"], metric["tp"]) == (int(g["total"]), int(g["true"]),
                                                                                  int(g["positive"]), float(g["tp"]))
    assert np.abs(out.numpy() - g["output"]).max() <= 1e-4 and np.abs(att - g["radar_attention"]).max() <= 1e-6
    for k in ("conf_1_pos", "conf_1_neg", "conf_2_pos", "conf_2_neg"):
        assert np.allclose(metric["conf"][k], g[k], atol=1e-6)


def test_obtain_iou_labels_rules():
    """Same image AND class, +1 pixel IoU, first maximum, zeros when nothing matches; the multi_boxes=False path keeps
    indices into the FILTERED target list like the reference (my_models.py:362-366)."""
    from oracle import stage3_loss as s3
    t = np.array([[0, 0, 10, 10, 50, 50], [0, 0, 10, 10, 50, 50], [0, 1, 10, 10, 50, 50], [1, 0, 0, 0, 20, 20]], np.float32)
    b = np.array([[0, 0, 10, 10, 50, 50], [0, 1, 12, 10, 50, 50], [2, 0, 10, 10, 50, 50], [1, 0, 0, 0, 9, 9],
                  [0, 0, 11, 10, 50, 50]], np.float32)
    lab, loc = s3.obtain_iou_labels(b, t, True)
    assert lab[0, 0] == 1.0 and np.array_equal(loc[0], t[0, 2:])
    assert 0.9 < lab[1, 0] < 1.0 and np.array_equal(loc[1], t[2, 2:])
    assert lab[2, 0] == 0.0 and not loc[2].any()
    assert abs(lab[3, 0] - 100.0 / 441.0) < 1e-6
    lab2, _ = s3.obtain_iou_labels(b, t, False)
    assert lab2[0, 0] == 1.0 and lab2[4, 0] == 0.0          # target 0 of the filtered list was claimed by box 0



def test_stage3_train_step_matches_reference(golden_dir):
    """Oracle of the (not yet built) stage-3 backward pass: one train-mode forward + backward (train.py:169-186) gives
    the reference's loss, the gradient of every head parameter and the BatchNorm running statistics after the step
    (fixture from tests/golden/make_golden_stage3_grads.py; the three large matrices are sub-sampled there)."""
    import random
    from millieye_b200.my_models import Network, define_yolo
    from oracle import stage3_train as st
    g = np.load(os.path.join(golden_dir, "stage3_grads_tiny12_192.npz"))
    gl = np.load(os.path.join(golden_dir, "stage3_loss_tiny12_192.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Network(define_yolo(cfg), conf_thresh=0.02).state_dict(), seed=6, obj_bias=2.0)
    random.seed(int(gl["sampling_seed"]))
    res = st.train_step(parse_model_config(cfg), {k: v.float() if v.is_floating_point() else v for k, v in sd.items()},
                        synth.synth_images(4, 192, seed=6), synth.synth_maps(4, 192, seed=6),
                        synth.synth_radar_boxes(4, seed=5), 0.02, gl["targets"])
    assert abs(res["loss"] - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert (res["n_all"], res["true"]) == (int(g["total"]), int(g["true"]))
    names = [str(n) for n in g["names"]]
    assert sorted(res["grads"]) == sorted(names)                      # the same parameters receive a gradient
    for name in names:
        got = res["grads"][name].numpy()
        if "grad/" + name in g.files:
            ref = g["grad/" + name]
            # (a conv bias in front of a train-mode BatchNorm has a zero gradient: only round-off noise, hence the atol)
            assert np.abs(got - ref).max() <= 1e-6 + 1e-4 * np.abs(ref).max(), name
        else:
            ref, sums = g["gsample/" + name], g["gsum/" + name]
            assert np.abs(got.reshape(-1)[::37] - ref).max() <= 1e-6 + 1e-4 * np.abs(ref).max(), name
            assert abs(got.astype(np.float64).sum() - sums[0]) <= 1e-6 + 1e-4 * sums[1], name
    for k in g.files:
        if k.startswith("buf/"):
            assert np.abs(res["buffers"][k[4:]].numpy() - g[k]).max() <= 1e-6, k


def test_yolo_loss_matches_reference(golden_dir):
    """Oracle of Darknet.forward(x, targets) (the product raises for it): total loss and every entry of
    the per-layer metrics dictionaries equal the reference's on the fixture of tests/golden/make_golden_yolo_loss.py,
    which includes two targets that land in the same cell with the same anchor (the last one wins)."""
    from millieye_b200.models import Darknet
    from oracle import yolo_loss as yl
    g = np.load(os.path.join(golden_dir, "yolo_loss_tiny12_160.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Darknet(cfg).state_dict(), seed=8, obj_bias=-1.0)
    with torch.no_grad():
        loss, feat, yolo, metrics = yl.darknet_loss(parse_model_config(cfg), {k: v.float() if v.is_floating_point() else v
                                                                              for k, v in sd.items()},
                                                    synth.synth_images(3, 160, seed=8), torch.from_numpy(g["targets"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    keys = [str(k) for k in g["metric_keys"]]
    assert len(metrics) == 2
    for li, m in enumerate(metrics):
        for k, r in zip(keys, g[f"metrics{li}"]):
            assert abs(m[k] - r) <= 1e-5 * max(1.0, abs(r)), (li, k)
    assert np.abs(yolo.numpy().reshape(-1)[::97] - g["yolo_sample"]).max() <= 1e-4 and feat.shape == (3, 256, 10, 10)


def test_manual_backward_matches_autograd(golden_dir):
    """oracle/stage3_backward.py (hand-derived backward of the parameters train.py updates: BatchNorm on batch
    statistics, 3x3 conv wgrad / dgrad by im2col, the RoIAlign adjoint, the radar_net and ensemble heads, focal + BCE
    loss; with the image path also the PS-RoIAlign adjoint, FC 490->256 / 256->13 and the 1x1 conv + BatchNorm) against
    torch.autograd on the same step: every one of the 24 / 32 gradients to 1e-4 (or 1e-6 absolute for the conv biases in
    front of a BatchNorm, whose gradient is zero)."""
    import random
    from millieye_b200.my_models import Network, define_yolo
    from oracle import stage3_backward as sb
    from oracle import stage3_train as st
    gl = np.load(os.path.join(golden_dir, "stage3_loss_tiny12_192.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Network(define_yolo(cfg), conf_thresh=0.02).state_dict(), seed=6, obj_bias=2.0)
    sdf = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    maps = synth.synth_maps(4, 192, seed=6)
    random.seed(int(gl["sampling_seed"]))
    res = st.train_step(parse_model_config(cfg), sdf, synth.synth_images(4, 192, seed=6), maps,
                        synth.synth_radar_boxes(4, seed=5), 0.02, gl["targets"])
    with torch.no_grad():
        cache = sb.forward_train(sdf, maps, res["box_locations"], res["n_img"], res["yolo_vec"], res["cls"],
                                 torch.from_numpy(res["pos"]), torch.from_numpy(res["sample_filter"]))
        grads = sb.backward(cache)                       # the set train.py updates with --pretrained_module2
        full = sb.forward_train(sdf, maps, res["box_locations"], res["n_img"], res["yolo_vec"], None,
                                torch.from_numpy(res["pos"]), torch.from_numpy(res["sample_filter"]), feat=res["feat"])
        grads_all = sb.backward(full, image_path=True)   # training from scratch: the image path too
    assert abs(cache["loss"] - res["loss"]) <= 1e-5 * res["loss"] and abs(full["loss"] - res["loss"]) <= 1e-5 * res["loss"]
    assert len(grads) == 24 and sorted(grads_all) == sorted(res["grads"]) and len(grads_all) == 32
    for got in (grads, grads_all):
        for k, v in got.items():
            ref = res["grads"][k]
            assert float((v - ref).abs().max()) <= 1e-6 + 1e-4 * float(ref.abs().max()), k
