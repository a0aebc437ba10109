"""Model-level parity on a B200 (`-m gpu`): Darknet / Network drop-ins against the reference-generated
golden fixtures and against the oracle on larger seeded inputs.

Tolerances.  The kernels compute in fp16 (operands, stored activations) with fp32 accumulation; the
reference is fp32.  north_star asks for boxes/scores within 1e-3 relative on Darknet-53 and bit-exact
NMS survivor sets given the same decoded tensor (tested in test_gpu_ops.py).  Box coordinates are
compared relative to the image size (the scale NMS and IoU work at), scores absolutely.
"""
import os

import numpy as np
import pytest
import torch

from millieye_b200 import configs
from millieye_b200._lib import MeError
from millieye_b200.models import Darknet
from millieye_b200.my_models import Network, define_yolo
from millieye_b200.parse_config import parse_model_config
from oracle import darknet as odark
from oracle import fusion as ofus
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _decode_err(got, ref, size):
    """(centre error / image size, worst and 99th-percentile relative w/h error, worst score error)."""
    centre = float(np.abs(got[..., :2] - ref[..., :2]).max() / size)
    wh = np.abs(got[..., 2:4] - ref[..., 2:4]) / np.maximum(np.abs(ref[..., 2:4]), 1.0)
    return centre, float(wh.max()), float(np.percentile(wh, 99)), float(np.abs(got[..., 4:] - ref[..., 4:]).max())


def _record(name, **metrics):
    """Parity numbers are appended to gpurun_out/parity_metrics.jsonl so a GPU run leaves evidence behind."""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_metrics.jsonl"), "a") as fh:
            fh.write(json.dumps(dict(test=name, **metrics)) + "\n")


def test_darknet_tiny_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "darknet_tiny12_96.npz"))
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
    assert list(net.state_dict().keys()) == list(g["keys"])
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=1))
    net.to(DEV)
    feat, yolo = net(synth.synth_images(2, 96, seed=1).to(DEV))
    assert feat.shape == (2, 256, 6, 6) and yolo.shape == (2, 135, 17)
    f, y = feat.cpu().numpy(), yolo.cpu().numpy()
    assert np.abs(f - g["featuremap"]).max() <= 2e-3 * np.abs(g["featuremap"]).max()
    # these synthetic weights drive logits to +-15, where exp() amplifies the fp16 activation error
    assert np.abs(y[..., :4] - g["yolo"][..., :4]).max() <= 5e-3 * np.abs(g["yolo"][..., :4]).max()
    assert np.abs(y[..., 4:] - g["yolo"][..., 4:]).max() <= 1e-2
    # second call replays the captured CUDA graph and must reproduce the first bit for bit
    feat2, yolo2 = net(synth.synth_images(2, 96, seed=1).to(DEV))
    assert torch.equal(yolo, yolo2) and torch.equal(feat, feat2)


def test_darknet53_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "darknet53_64.npz"))
    net = Darknet(configs.cfg_path("yolov3")).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=2, conv_gain=0.6))
    net.to(DEV)
    feat, yolo = net(synth.synth_images(1, 64, seed=2).to(DEV))
    assert feat.numel() == 0  # no conv_8 in yolov3.cfg: the reference has no featuremap either (F1)
    y, yg = yolo.cpu().numpy(), g["yolo"]
    rel = np.abs(y[..., :4] - yg[..., :4]) / (np.abs(yg[..., :4]) + 1e-6)
    assert rel.max() <= 2e-3          # element-wise relative, worst element
    assert np.percentile(rel, 99) <= 1e-3
    assert np.abs(y[..., 4:] - yg[..., 4:]).max() <= 1e-3


@pytest.mark.parametrize("cfg,n,size,gain", [("yolov3-tiny-12", 3, 416, 1.0), ("yolov3-tiny-12", 2, 320, 1.0),
                                              ("yolov3", 2, 416, 0.6), ("yolov3", 1, 512, 0.6),
                                              ("yolov3", 32, 416, 0.6)])     # BASELINE config 2's own shape
def test_darknet_vs_oracle(cfg, n, size, gain):
    net = Darknet(configs.cfg_path(cfg)).eval()
    sd = synth.fill_state_dict(net.state_dict(), seed=9, conv_gain=gain, head_gain=0.5)
    net.load_state_dict(sd)
    net.to(DEV)
    if cfg == "yolov3":
        net.feature_tap = 91  # a 256-channel stride-16 block (undefined in the reference)
    x = synth.synth_images(n, size, seed=size)
    feat, yolo = net(x.to(DEV))
    md = parse_model_config(configs.cfg_path(cfg))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        rfeat, ryolo = odark.darknet_forward(md, {k: v.float() for k, v in sd.items()}, x, feature_tap=net.feature_tap)
    y, yr = yolo.cpu().numpy(), ryolo.numpy()
    assert y.shape == yr.shape
    centre, wh_max, wh_p99, score = _decode_err(y, yr, size)
    feat_err = float(np.abs(feat.cpu().numpy() - rfeat.numpy()).max() / np.abs(rfeat.numpy()).max())
    _record("darknet_vs_oracle", cfg=cfg, n=n, size=size, centre=centre, wh_max=wh_max, wh_p99=wh_p99, score=score,
            feat=feat_err)
    if cfg == "yolov3":
        # north_star's bar on its headline config: boxes and scores within 1e-3 (measured ~8e-4 / 3e-4)
        assert centre <= 1e-3 and wh_max <= 1e-3 and score <= 1e-3
    else:
        # tiny has no residual trunk to damp the fp16 rounding of 13 chained convs: measured 5.5e-3 worst
        # element, 2.8e-3 at the 99th percentile on exp()-decoded sizes, 1.3e-3 on scores
        assert centre <= 1e-3 and wh_p99 <= 4e-3 and wh_max <= 1e-2 and score <= 2.5e-3
    assert feat_err <= 3e-3


def test_darknet_input_contract():
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval().to(DEV)
    with pytest.raises(MeError):
        net(torch.rand(1, 3, 96, 128, device=DEV))      # non-square
    out3 = net(torch.rand(1, 3, 96, 96, device=DEV), targets=torch.tensor([[0, 1, 0.5, 0.5, 0.2, 0.3]]))
    assert len(out3) == 3 and out3[0].dim() == 0         # (loss, featuremap, yolo_outputs) like models.py:267
    # weights changed in place -> refresh_weights() rebuilds the packed copies
    x = torch.rand(1, 3, 96, 96, device=DEV)
    _, y0 = net(x)
    with torch.no_grad():
        net.module_list[0][0].weight.mul_(0.0)
    net.refresh_weights()
    _, y1 = net(x)
    assert not torch.equal(y0, y1)


def test_darknet_weights_file_roundtrip(tmp_path):
    a = Darknet(configs.cfg_path("yolov3-tiny-12"))
    a.load_state_dict(synth.fill_state_dict(a.state_dict(), seed=5))
    path = str(tmp_path / "t.weights")
    a.save_darknet_weights(path, cutoff=len(a.module_list))
    b = Darknet(configs.cfg_path("yolov3-tiny-12"))
    b.load_darknet_weights(path)
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        if va.is_floating_point():
            assert torch.equal(va, vb), k


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fusion_golden(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "fusion_tiny12_160.npz"))
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.05).eval()
    assert list(model.state_dict().keys()) == list(g["keys"])
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=3, obj_bias=-0.5))
    model.to(DEV)
    rb = synth.synth_radar_boxes(3, seed=5).to(DEV)
    out = model(synth.synth_images(3, 160, seed=3).to(DEV), synth.synth_maps(3, 160, seed=3).to(DEV), rb, mode)
    ref = g[f"mode{mode}"]
    o = out.cpu().numpy()
    assert o.shape == ref.shape
    np.testing.assert_array_equal(o[:, 0], ref[:, 0])      # image index, i.e. same proposals in the same order
    np.testing.assert_array_equal(o[:, 7], ref[:, 7])      # class prediction
    # the synthetic heads decode boxes far larger than the image; compare relative to the box scale
    # (tiny-12 in fp16: measured 4.1e-3 worst element, see the tiny tolerances in test_darknet_vs_oracle)
    assert (np.abs(o[:, 1:5] - ref[:, 1:5]) / np.maximum(np.abs(ref[:, 1:5]), 160)).max() <= 8e-3
    assert np.abs(o[:, 5:7] - ref[:, 5:7]).max() <= 5e-3
    if mode == 0:
        np.testing.assert_allclose(rb.cpu().numpy(), g["radar_boxes_after"], rtol=1e-6)  # in-place scaling (F6)
    if mode == 2:
        assert model.refine_threshold_img == 1                 # persistent side effect (my_models.py:480)


@pytest.mark.parametrize("n,size,thr", [(4, 416, 0.2), (2, 320, 0.05)])
def test_fusion_vs_oracle(n, size, thr):
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=thr).eval()
    sd = synth.fill_state_dict(model.state_dict(), seed=21, obj_bias=-1.0, head_gain=0.4)
    model.load_state_dict(sd)
    model.to(DEV)
    imgs, maps = synth.synth_images(n, size, seed=21), synth.synth_maps(n, size, seed=21)
    rb = synth.synth_radar_boxes(n, seed=22)
    out = model(imgs.to(DEV), maps.to(DEV), rb.clone().to(DEV), 0).cpu().numpy()
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    with torch.no_grad():
        ref = ofus.network_forward(md, {k: v.float() for k, v in sd.items()}, imgs, maps, rb, thr).numpy()
    # fp16 activations can flip a box across the confidence / IoU thresholds, so compare as sets of
    # (image, class) rows matched by position, requiring near-total agreement
    assert abs(len(out) - len(ref)) <= max(2, len(ref) // 50)
    matched = 0
    for r in ref:
        cand = out[(out[:, 0] == r[0]) & (out[:, 7] == r[7])]
        scale = max(size, float(np.abs(r[1:5]).max()))
        if len(cand) and (np.abs(cand[:, 1:5] - r[1:5]).max(1) / scale).min() <= 3e-3:
            j = (np.abs(cand[:, 1:5] - r[1:5]).max(1)).argmin()
            if abs(cand[j, 5] - r[5]) <= 5e-3 and abs(cand[j, 6] - r[6]) <= 5e-3:
                matched += 1
    assert matched >= 0.97 * len(ref)
    assert len(ref) > 10  # the case must actually exercise the heads


def test_fusion_batch32_radar_points_vs_oracle():
    """BASELINE config 3's own shape: batch 32, 416 x 416, 64 radar points per frame.  The heat-maps come from the device
    kernel (me_radar_maps, checked against the oracle in tests/test_radar.py) and feed both the model under test and the
    oracle; rows are matched like in test_fusion_vs_oracle and the match statistics are recorded."""
    from millieye_b200 import radar
    n, size, thr = 32, 416, 0.2
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=thr).eval()
    sd = synth.fill_state_dict(model.state_dict(), seed=31, obj_bias=-2.0, head_gain=0.4)
    model.load_state_dict(sd)
    model.to(DEV)
    rng = np.random.RandomState(7)
    pts = np.stack([rng.uniform(-3, 3, (n, 64)), rng.uniform(1, 10, (n, 64)), rng.uniform(-1.5, 1.5, (n, 64)),
                    rng.uniform(-3, 3, (n, 64))], -1).astype(np.float32)
    maps = radar.radar_maps(torch.from_numpy(pts).to(DEV), torch.full((n,), 64, dtype=torch.int32, device=DEV),
                            radar.make_cfg(out_size=size // 16))
    assert maps.shape == (n, 3, size // 16, size // 16)
    imgs = synth.synth_images(n, size, seed=31)
    rb = synth.synth_radar_boxes(n, seed=32)
    out = model(imgs.to(DEV), maps, rb.clone().to(DEV), 0).cpu().numpy()
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = ofus.network_forward(md, {k: v.float() for k, v in sd.items()}, imgs, maps.cpu(), rb, thr).numpy()
    assert len(ref) > 100
    assert abs(len(out) - len(ref)) <= max(2, len(ref) // 50)
    matched, box_err, score_err = 0, 0.0, 0.0
    for r in ref:
        cand = out[(out[:, 0] == r[0]) & (out[:, 7] == r[7])]
        if not len(cand):
            continue
        scale = max(size, float(np.abs(r[1:5]).max()))
        err = np.abs(cand[:, 1:5] - r[1:5]).max(1) / scale
        j = int(err.argmin())
        if err[j] <= 3e-3 and abs(cand[j, 5] - r[5]) <= 5e-3 and abs(cand[j, 6] - r[6]) <= 5e-3:
            matched += 1
            box_err, score_err = max(box_err, float(err[j])), max(score_err, float(np.abs(cand[j, 5:7] - r[5:7]).max()))
    _record("fusion_batch32_vs_oracle", rows_ref=len(ref), rows_gpu=len(out), matched=matched, box_err=box_err, score_err=score_err)
    assert matched >= 0.97 * len(ref)


def test_fusion_pipeline_equals_forward():
    """FusionPipeline.submit (tail kernels of batch i on a second stream under the backbone of batch i+1) returns the rows
    of the blocking Network.forward bit for bit, for a stream of different batches, uint8 host frames included, with
    records consumed late (two batches in flight) and an empty batch (no radar boxes, threshold nothing passes) in between."""
    from millieye_b200 import radar
    from millieye_b200.my_models import FusionPipeline
    n, size, thr = 8, 416, 0.2
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=thr).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=31, obj_bias=-2.0, head_gain=0.4))
    model.to(DEV)
    cfg = radar.make_cfg(out_size=size // 16)
    batches = []
    for k in range(6):
        rng = np.random.RandomState(50 + k)
        pts = np.stack([rng.uniform(-3, 3, (n, 64)), rng.uniform(1, 10, (n, 64)), rng.uniform(-1.5, 1.5, (n, 64)),
                        rng.uniform(-3, 3, (n, 64))], -1).astype(np.float32)
        maps = radar.radar_maps(torch.from_numpy(pts).to(DEV), torch.full((n,), 64, dtype=torch.int32, device=DEV), cfg)
        u8 = torch.randint(0, 256, (n, 3, size, size), generator=torch.Generator().manual_seed(60 + k), dtype=torch.uint8)
        imgs = u8.pin_memory() if k % 2 else (u8.float() / 255.0).to(DEV)
        rb = synth.synth_radar_boxes(n, seed=70 + k) if k != 3 else torch.zeros((0, 5))
        batches.append((imgs, maps, rb))
    want = [model(imgs, maps, rb.clone().to(DEV), 0).cpu() for imgs, maps, rb in batches]
    assert sum(len(w) for w in want) > 50
    # 18 submits over the 6 batches: every (fusion plan, detector output slot) pair is used three times - as is, under
    # capture, and as a graph replay (FusionPipeline._run)
    for graphs in (True, False):
        pipe = FusionPipeline(model, depth=3, use_cuda_graph=graphs)
        got, recs = [], []
        for k in range(18):
            imgs, maps, rb = batches[k % 6]
            recs.append(pipe.submit(imgs, maps, rb.clone().to(DEV), 0, readback=(k % 6 != 4)))
            if len(recs) == 3:                      # two further batches are in flight when a record is read
                got.append(recs.pop(0).wait().cpu().clone())
        got += [r.wait().cpu().clone() for r in recs]
        assert len(got) == 18
        for k, a in enumerate(got):
            b = want[k % 6]
            assert a.shape == b.shape and torch.equal(a, b), f"graphs={graphs} submit {k}"
    # mode 2 (radar only) keeps the reference's persistent side effect, mode 1 is refused
    imgs, maps, rb = batches[0]
    w2 = model(imgs, maps, rb.clone().to(DEV), 2).cpu()
    g2 = pipe.submit(imgs, maps, rb.clone().to(DEV), 2).wait()
    assert torch.equal(g2, w2) and model.refine_threshold_img == 1
    with pytest.raises(MeError):
        pipe.submit(imgs, maps, rb.clone().to(DEV), 1)


def test_fusion_empty_and_errors():
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.999999).eval().to(DEV)
    imgs = torch.rand(2, 3, 96, 96, device=DEV)
    maps = torch.rand(2, 3, 6, 6, device=DEV)
    out = model(imgs, maps, torch.zeros((0, 5), device=DEV), 1)
    assert out.shape == (0, 8)                        # empty proposals give a (0,8) tensor, never raise
    out = model(imgs, maps, torch.zeros((0, 5), device=DEV), 0)
    assert out.shape[1] == 8
    with pytest.raises(MeError):      # train() mode is only defined for the training step (with targets): no silent inference
        model.train()
        model(imgs, maps, torch.zeros((0, 5), device=DEV), 0)
    model.eval()
    loss, out, metric, att = model(imgs, maps, torch.zeros((0, 5), device=DEV), 0, targets=torch.zeros(0, 6))
    assert float(loss) == 0.0 and out.shape[1] == 8 and att.shape == (2, 1, 6, 6) and metric["true"] == 0


def test_stage2_golden(golden_dir):
    """Stage-2 model (module2_mixed/my_models.py) against the reference-generated fixture."""
    from millieye_b200.my_models_stage2 import Network as Network2
    g = np.load(os.path.join(golden_dir, "stage2_tiny12_160.npz"))
    model = Network2(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.3).eval()
    assert list(model.state_dict().keys()) == list(g["keys"])
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=-1.0, head_gain=1.0))
    model.to(DEV)
    out = model(synth.synth_images(2, 160, seed=6).to(DEV))
    assert not out.is_cuda                                  # the reference returns .cpu() (:361)
    o, ref = out.numpy(), g["out"]
    # fp16 activations can move a box across the confidence threshold or swap near-equal confidences:
    # match rows by (image, class, box) instead of by position
    assert abs(len(o) - len(ref)) <= max(2, len(ref) // 50)
    matched = 0
    for r in ref:
        cand = o[(o[:, 0] == r[0]) & (o[:, 7] == r[7])]
        if not len(cand):
            continue
        scale = max(160.0, float(np.abs(r[1:5]).max()))
        err = np.abs(cand[:, 1:5] - r[1:5]).max(1) / scale
        j = err.argmin()
        if err[j] <= 8e-3 and abs(cand[j, 5] - r[5]) <= 5e-3 and abs(cand[j, 6] - r[6]) <= 5e-3:
            matched += 1
    assert matched >= 0.97 * len(ref)
    assert np.all(np.diff(o[:, 5]) <= 0)                     # sorted by the new confidence


def test_stage2_loss_golden(golden_dir):
    """Stage-2 Network.forward(images, targets) in eval mode against the reference-generated fixture (labels, balanced
    sample, FocalLoss + confidence / category BCE + box regression, metric; module2_mixed/my_models.py:363-461)."""
    import random
    from millieye_b200.my_models_stage2 import Network as Network2
    g = np.load(os.path.join(golden_dir, "stage2_loss_tiny12_160.npz"))
    model = Network2(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=float(g["conf_thresh"])).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=-1.0, head_gain=1.0))
    model.to(DEV)
    targets = torch.from_numpy(g["targets"].copy())
    random.seed(int(g["sampling_seed"]))
    out, loss, metric = model(synth.synth_images(3, 160, seed=6).to(DEV), targets)
    assert not out.is_cuda and out.shape[1] == 8
    assert np.allclose(targets.numpy(), g["targets_after"], atol=1e-4)            # rewritten in place like the reference
    assert metric["total"] == int(g["total"])
    plan = next(iter(model._plans.values()))
    lab, ref_lab = plan.iou_labels[:metric["total"]].cpu().numpy(), g["iou_labels"].reshape(-1)
    assert np.array_equal(lab > 0.7, ref_lab > 0.7) and np.array_equal(lab < 0.3, ref_lab < 0.3)
    assert int(metric["true"]) == int(g["true"]) and int(metric["positive"]) == int(g["positive"])
    assert float(metric["tp"]) == float(g["tp"])
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    _record("stage2_loss_golden", loss=float(loss), ref=float(g["loss"]), rel=rel, label_err=float(np.abs(lab - ref_lab).max()))
    assert rel <= 2e-2
    model.train()
    with pytest.raises(MeError):
        model(synth.synth_images(3, 160, seed=6).to(DEV), torch.from_numpy(g["targets"].copy()))


def test_detect_pipeline_matches_sequential():
    """DetectPipeline (NMS of batch i on a second stream while batch i+1 runs) returns, for every batch of a
    stream, exactly the detections of the sequential forward -> filter_nms path, device and host copies alike."""
    from millieye_b200 import ops
    from millieye_b200.models import DetectPipeline
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=2, obj_bias=-1.0))
    net.to(DEV)
    batches = [synth.synth_images(3, 160, seed=10 + i) for i in range(5)]
    want = []
    for x in batches:
        plan = net.forward_device(x.to(DEV))
        b = ops.filter_nms(plan.yolo_out, 0.1, 0.5, 200, xyxy_inplace=True)
        torch.cuda.synchronize()
        want.append((b.det.cpu().clone(), b.count.cpu().clone()))
    assert sum(int(c.sum()) for _, c in want) > 0
    pipe = DetectPipeline(net, 0.1)
    pinned = [x.pin_memory() for x in batches]
    got = []
    for x in pinned:                      # host batches: H2D on the copy stream, NMS + readback on the side stream
        rec = pipe.submit(x, readback=True).wait()
        got.append((rec.host_det.clone(), rec.host_cnt.clone()))
    recs = [pipe.submit(x.to(DEV), readback=True) for x in batches[:2]]   # two batches in flight, both slots
    for (d, c), (wd, wc) in zip(got, want):
        assert torch.equal(c, wc)
        for i in range(d.shape[0]):
            assert torch.equal(d[i, :int(c[i])], wd[i, :int(c[i])])
    for rec, (wd, wc) in zip(recs, want[:2]):
        rec.wait()
        assert torch.equal(rec.host_cnt, wc)


def test_stage3_loss_golden(golden_dir):
    """Network.forward(..., targets) in eval mode against the reference-generated fixture (labelling + losses +
    metric + radar attention, my_models.py:545-640).  The proposals come out of fp16 convolutions, so labels are
    compared with a tolerance and the pos / neg split must agree wherever the reference label is not within 0.015 of
    a threshold (none is in this fixture); with the same split python's seeded random.sample draws the same rows."""
    import random
    g = np.load(os.path.join(golden_dir, "stage3_loss_tiny12_192.npz"))
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.02).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=2.0))
    model.to(DEV)
    imgs, maps = synth.synth_images(4, 192, seed=6), synth.synth_maps(4, 192, seed=6)
    rb = synth.synth_radar_boxes(4, seed=5)
    targets = torch.from_numpy(g["targets"].copy())
    random.seed(int(g["sampling_seed"]))
    loss, out, metric, att = model(imgs.to(DEV), maps.to(DEV), rb.clone().to(DEV), 0, targets)
    torch.cuda.synchronize()
    assert np.allclose(targets.numpy(), g["targets_after"], atol=1e-4)          # rewritten in place like the reference
    ref_lab = g["iou_labels"].reshape(-1)
    for thr in (0.3, 0.5, 0.7):
        assert np.abs(ref_lab - thr).min() > 0.015   # the fixture generator guarantees it
    plan = next(iter(model._plans.values()))
    lab = plan.iou_labels[:len(ref_lab)].cpu().numpy()
    assert metric["total"] == int(g["total"]) == len(ref_lab)
    # sub-pixel differences of the fp16 boxes move the +1-pixel IoU of the smallest boxes by more than a percent:
    # nearly all rows agree to 2e-2, a few small boxes may not - the pos / neg split below is what the loss depends on
    err = np.abs(lab - ref_lab)
    bad = np.where(err > 2e-2)[0]
    wh = (plan.rois[:len(ref_lab), 3:5] - plan.rois[:len(ref_lab), 1:3]).cpu().numpy()
    assert len(bad) <= 0.03 * len(ref_lab), (bad, lab[bad], ref_lab[bad], wh[bad])
    assert err.max() <= 0.25 and all(wh[i].min() < 24 for i in bad), (bad, lab[bad], ref_lab[bad], wh[bad])
    assert np.array_equal(lab > 0.7, ref_lab > 0.7) and np.array_equal(lab < 0.3, ref_lab < 0.3)
    assert int(metric["true"]) == int(g["true"]) and int(metric["positive"]) == int(g["positive"])
    assert float(metric["tp"]) == float(g["tp"])
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    _record("stage3_loss_golden", loss=float(loss), ref=float(g["loss"]), rel=rel, label_err=float(np.abs(lab - ref_lab).max()))
    assert rel <= 2e-2
    assert out.shape == g["output"].shape
    assert np.abs(att.cpu().numpy() - g["radar_attention"]).max() <= 5e-3
    for k in ("conf_1_pos", "conf_1_neg", "conf_2_pos", "conf_2_neg"):
        got = metric["conf"][k].numpy()           # split at label 0.5: a small box whose label moved may change sides
        assert abs(len(got) - len(g[k])) <= len(bad)
        if len(got) == len(g[k]):
            assert np.abs(got - g[k]).max() <= 2e-2



def test_next_input_zero_copy():
    """A batch written straight into DarknetPlan.next_input() and handed to forward_device() gives the same
    result as a batch that is copied in; the two slots alternate."""
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=4))
    net.to(DEV)
    xs = [synth.synth_images(2, 96, seed=20 + i).to(DEV) for i in range(3)]
    want = [net.forward_device(x).yolo_out.clone() for x in xs]
    plan = net.plan_for(2, 96, DEV)
    seen = set()
    for x, w in zip(xs, want):
        buf = plan.next_input()
        seen.add(buf.data_ptr())
        buf.copy_(x)
        assert net.forward_device(buf) is plan
        assert torch.equal(plan.yolo_out, w)
    assert len(seen) == 2


def test_fused_decode_equals_separate_kernels(monkeypatch):
    """The plan with the decode fused into the head convs reproduces the plan with separate decode kernels exactly."""
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ME_FUSE_DECODE", flag)
        net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
        net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=5))
        net.to(DEV)
        plan = net.forward_device(synth.synth_images(3, 224, seed=5).to(DEV))
        assert (len(plan.post_ops) > 0) == (flag == "0")
        outs.append(plan.yolo_out.clone())
    assert torch.equal(outs[0], outs[1])


def test_uint8_frames_equal_totensor():
    """Frames uploaded as bytes ((N,3,S,S) uint8, pinned host or device) give exactly the result of the reference's
    contract input, ToTensor's x / 255 in fp32 - through Darknet.forward, DetectPipeline and Network.forward."""
    from millieye_b200.models import DetectPipeline
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=4, obj_bias=-1.0))
    net.to(DEV)
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (3, 3, 160, 160), generator=g, dtype=torch.uint8)
    want = net(u8.float().div(255).to(DEV))[1]
    assert torch.equal(net(u8.to(DEV))[1], want)                 # bytes already on the device
    assert torch.equal(net(u8.pin_memory())[1], want)            # bytes in pinned host memory (copy stream)
    pipe = DetectPipeline(net, 0.1)
    a = pipe.submit(u8.pin_memory(), readback=True).wait()
    b = pipe.submit(u8.float().div(255).pin_memory(), readback=True).wait()
    assert torch.equal(a.host_cnt, b.host_cnt) and torch.equal(a.host_det, b.host_det)
    with pytest.raises(MeError):
        net(u8.permute(0, 2, 3, 1).contiguous().to(DEV))          # NHWC bytes are not the contract layout
