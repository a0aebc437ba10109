"""CPU-only checks of the host side: cfg handling, state_dict compatibility, the C-ABI surface (the
library loads and exports every symbol include/millieye_b200.h declares - no compute calls), loud
failure without CUDA / without the library, plan description, weights-file format, sharding maths."""
import os
import re

import numpy as np
import pytest
import torch

from millieye_b200 import _lib, configs
from millieye_b200.dist import shard_bounds, shard_rows_by_frame
from millieye_b200.engine import describe_blocks
from millieye_b200.models import Darknet
from millieye_b200.my_models import Network, define_yolo
from millieye_b200.parse_config import parse_data_config, parse_model_config
from oracle import synth
from oracle.parse_config import parse_model_config as oracle_parse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CFG = "/root/reference/module3_our_dataset/config"


@pytest.mark.parametrize("name", ["yolov3-tiny-12", "yolov3-tiny-coco", "yolov3"])
def test_cfg_parsers_agree(name):
    path = configs.cfg_path(name)
    assert parse_model_config(path) == oracle_parse(path)
    if os.path.isdir(REF_CFG):  # build container only: generated cfg == reference cfg, block by block
        ref = oracle_parse(os.path.join(REF_CFG, name + ".cfg"))
        mine = parse_model_config(path)
        assert parse_model_config(os.path.join(REF_CFG, name + ".cfg")) == ref
        assert len(ref) == len(mine)
        keys = ("type", "batch_normalize", "filters", "size", "stride", "activation", "layers", "from", "mask",
                "anchors", "classes")
        for a, b in zip(mine[1:], ref[1:]):
            for k in keys:
                assert str(a.get(k)).replace(" ", "") == str(b.get(k)).replace(" ", "")


def test_cfg_parser_rules(tmp_path):
    p = tmp_path / "x.cfg"
    p.write_text("[net]\nchannels=3\n# comment\n\n[convolutional]\nfilters = 8 \nsize=3\nstride=1\nactivation=leaky\n")
    blocks = parse_model_config(str(p))
    assert blocks[1]["batch_normalize"] == 0 and blocks[1]["filters"] == "8" and blocks[0]["channels"] == "3"
    d = tmp_path / "x.data"
    d.write_text("classes= 12\ntrain=a b\n# c\n\nnames=n.txt\n")
    opts = parse_data_config(str(d))
    assert opts["gpus"] == "0,1,2,3" and opts["train"] == ["a", "b"] and opts["names"] == "n.txt"


def test_plan_description_matches_reference_shapes():
    _, tiny = describe_blocks(parse_model_config(configs.cfg_path("yolov3-tiny-12")))
    assert [b["type"] for b in tiny].count("convolutional") == 13 and len(tiny) == 24
    assert tiny[20]["out_c"] == 384 and tiny[20]["layers"] == [19, 8] and tiny[17]["layers"] == [13]
    assert tiny[15]["filters"] == 51 and not tiny[15]["bn"] and not tiny[15]["leaky"]
    _, full = describe_blocks(parse_model_config(configs.cfg_path("yolov3")))
    kinds = [b["type"] for b in full]
    assert (len(full), kinds.count("convolutional"), kinds.count("shortcut"), kinds.count("route")) == (107, 75, 23, 4)
    assert full[86]["out_c"] == 768 and full[98]["out_c"] == 384 and full[4]["src"] == 1
    assert full[82]["anchors"] == [(116, 90), (156, 198), (373, 326)]


def test_state_dict_keys_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "fusion_tiny12_160.npz"))
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.2)
    assert list(model.state_dict().keys()) == list(g["keys"])
    assert len(g["keys"]) == 123 and sum(k.startswith("base_detector.") for k in g["keys"]) == 70
    for attr in ("base_detector", "refinement_head", "refine_threshold_radar", "seen", "conf_thresh", "class_idx"):
        assert hasattr(model, attr)


def test_header_symbols_exported(built_lib):
    text = open(os.path.join(ROOT, "include", "millieye_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t)\s+(me_\w+)\s*\(", text, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), f"{name} missing from {built_lib}"
    assert handle.me_version() >= 100
    assert [handle.me_conv_k_block(c) for c in (3, 16, 17, 32, 33, 64, 490)] == [16, 16, 32, 32, 64, 64, 64]
    assert handle.me_conv_cin_pad(490) == 512 and handle.me_conv_cin_pad(16) == 16


def test_no_cpu_fallback():
    net = Darknet(configs.cfg_path("yolov3-tiny-12")).eval()
    with pytest.raises(_lib.MeError):
        net(torch.rand(1, 3, 96, 96))
    from millieye_b200 import ops, utils
    with pytest.raises(_lib.MeError):
        utils.non_max_suppression_cpp(torch.rand(1, 10, 17), 0.1)
    with pytest.raises(_lib.MeError):
        ops.maxpool2(torch.zeros(1, 2, 2, 8, dtype=torch.float16), torch.zeros(1, 1, 1, 8, dtype=torch.float16), 1, 2, 2, 8, 8, 8, 2)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib._build, "LIB", "/nonexistent/libmillieye_b200.so")
    with pytest.raises(_lib.MeError, match="only implementation"):
        _lib.lib()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "millieye_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_darknet_weights_file_roundtrip(tmp_path):
    a = Darknet(configs.cfg_path("yolov3-tiny-12"))
    a.load_state_dict(synth.fill_state_dict(a.state_dict(), seed=5))
    a.seen = 1234
    path = str(tmp_path / "t.weights")
    a.save_darknet_weights(path, cutoff=len(a.module_list))
    raw = np.fromfile(path, dtype=np.int32, count=5)
    assert raw[3] == 1234
    b = Darknet(configs.cfg_path("yolov3-tiny-12"))
    b.load_darknet_weights(path)
    assert b.seen == 1234
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        if va.is_floating_point():
            assert torch.equal(va, vb), k
    n_floats = sum(v.numel() for k, v in a.state_dict().items() if v.is_floating_point())
    assert os.path.getsize(path) == 20 + 4 * n_floats


def test_shard_bounds():
    for total in (1, 7, 32, 33, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    rows = torch.tensor([[0, 1.0], [3, 2.0], [4, 3.0], [7, 4.0]])
    got = shard_rows_by_frame(rows, 8, 2, 1)
    assert got.tolist() == [[0, 3.0], [3, 4.0]]


def test_plan_launch_list_on_cpu(monkeypatch):
    """DarknetPlan's wiring without a GPU: the library calls are replaced by recorders and the plan is built on CPU
    tensors, so the launch list of Darknet-53 can be checked for what the kernels rely on - residuals fused into the
    producing conv and pointing at the shortcut source, route concats written in place through pitch / channel offset,
    fp32 linear head convs, decode kernels on the post-processing list with the reference's row offsets (models.py:266)."""
    import torch
    from millieye_b200 import engine, ops
    from millieye_b200._lib import ME_ACT_LEAKY, ME_ACT_LINEAR
    calls = []

    def fake_pack(weight, conv_bias=None, bn=None, cout_pad=None):
        cout, cin, k, _ = weight.shape
        pad = cout_pad or ops.round_up(cout, 32)
        return ops.PackedConv(torch.zeros(1), torch.zeros(1), cin, cout, pad, k)

    monkeypatch.setattr(ops, "pack_conv", fake_pack)
    monkeypatch.setattr(ops, "pack_first_conv", lambda w, b=None, bn=None: ops.FirstConv.__new__(ops.FirstConv))
    monkeypatch.setattr(ops, "conv_workspace", lambda dev: torch.zeros(1))
    monkeypatch.setattr(ops, "conv_first", lambda x, f, out, pitch, act, pool=False: calls.append(("first", out.data_ptr(), pitch, act)))
    monkeypatch.setattr(ops, "conv_gemm", lambda x, p, n, h, w, in_pitch, out, out_pitch, **kw: calls.append(
        ("conv", x.data_ptr(), out.data_ptr(), h, in_pitch, out_pitch, kw)))
    monkeypatch.setattr(ops, "upsample2", lambda x, y, n, h, w, c, ip, op_: calls.append(("up", x.data_ptr(), y.data_ptr(), c, ip, op_)))
    monkeypatch.setattr(ops, "yolo_decode", lambda logits, pitch, out, n, g, anchors, nc, stride, rows, off: calls.append(
        ("decode", logits.data_ptr(), g, rows, off, stride)))
    monkeypatch.delenv("ME_FUSE_DECODE", raising=False)

    net = __import__("millieye_b200.models", fromlist=["Darknet"]).Darknet(configs.cfg_path("yolov3"))
    tensors = {k: v.detach().float() for k, v in net.state_dict().items() if v.is_floating_point()}
    plan = engine.DarknetPlan(net._blocks, tensors, 1, 64, torch.device("cpu"), None)
    assert plan.op_kinds.count("conv") == 75 and plan.op_kinds.count("upsample") == 2 and len(plan.post_ops) == 3
    assert plan.rows_total == 3 * (2 * 2 + 4 * 4 + 8 * 8)
    plan.enqueue()
    plan.enqueue_post()
    convs = [c for c in calls if c[0] == "conv"]
    assert len(convs) == 74 and calls[0][0] == "first"
    out_of = {}                      # output pointer of the conv that produces each cfg block's tensor
    conv_blocks = [b for b, k in zip(plan.op_blocks, plan.op_kinds) if k == "conv"]
    for blk, c in zip(conv_blocks[1:], convs):
        out_of[blk] = c[2]
    out_of[0] = calls[0][1]
    n_res = 0
    for blk, c in zip(conv_blocks[1:], convs):
        kw = c[6]
        nxt = plan.blocks[blk + 1] if blk + 1 < len(plan.blocks) else {"type": ""}
        if nxt["type"] == "shortcut":            # the add is fused: residual = output of block `from` (an alias chain)
            n_res += 1
            src = nxt["src"]
            while plan.blocks[src]["type"] == "shortcut":
                src -= 1                          # a shortcut's tensor is the tensor of the conv in front of it
            assert kw["residual"] is not None and kw["residual"].data_ptr() == out_of[src]
            assert kw["res_pitch"] >= kw["cout"] and c[5] >= kw["cout"]   # (the output may sit in a wider concat buffer)
        else:
            assert kw.get("residual") is None
        if nxt["type"] == "yolo":
            assert kw["out_f32"] and kw["act"] == ME_ACT_LINEAR and kw["cout"] == 256
        else:
            assert not kw["out_f32"]
    assert n_res == 23
    # route -1,61 (block 86): the upsample of block 85 and the tensor of block 61 share one 768-channel buffer
    ups = [c for c in calls if c[0] == "up"]
    assert ups[0][5] == 768 and ups[1][5] == 384                     # upsample writes with the concat buffer's pitch
    conv60 = convs[conv_blocks[1:].index(60)]
    assert conv60[5] == 768 and conv60[2] == ups[0][2] + 2 * 256      # 256 fp16 channels after the upsampled slice
    conv87 = convs[conv_blocks[1:].index(87)]
    assert conv87[4] == 768 and conv87[1] == ups[0][2]                # the conv after the route reads the whole buffer
    dec = [c for c in calls if c[0] == "decode"]
    assert [(d[2], d[4]) for d in dec] == [(2, 0), (4, 12), (8, 60)] and all(d[3] == plan.rows_total for d in dec)
    assert [d[5] for d in dec] == [32.0, 16.0, 8.0]


def test_bin_major_permutation():
    """ops.bin_major_perm: position bin * C + c of the bin-major order holds reference element c * P * P + bin (the flatten
    order of my_models.py:262), and permuting a producer's rows and a consumer's columns with it leaves the product
    unchanged - what _FusionPlan relies on when it packs img_cnn_layers / refinement_head.net0 / radar_net."""
    from millieye_b200 import ops
    perm = ops.bin_major_perm(10, 7)
    assert perm.shape == (490,) and sorted(perm.tolist()) == list(range(490))
    for c in (0, 3, 9):
        for b in (0, 17, 48):
            assert int(perm[b * 10 + c]) == c * 49 + b
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(5, 490, generator=g)            # a crop in reference order
    w = torch.randn(8, 490, generator=g)               # a layer that reads it
    assert torch.allclose(feat[:, perm] @ w[:, perm].t(), feat @ w.t(), atol=1e-4)


def test_tiny_plan_fuses_first_pool_on_cpu(monkeypatch):
    """yolov3-tiny-12 at 64 x 64 without a GPU: block 1 (MaxPool 2/2) has no launch of its own - the first conv is asked to
    pool and writes the 32 x 32 tensor; the stride-1 pool of block 11 and the other four pools stay separate launches;
    ME_FUSE_POOL=0 restores the six pools; a 48 x 48 input (width % 32 != 0) is not fused either."""
    import torch
    from millieye_b200 import engine, ops
    first_calls = []

    def fake_pack(weight, conv_bias=None, bn=None, cout_pad=None):
        cout, cin, k, _ = weight.shape
        return ops.PackedConv(torch.zeros(1), torch.zeros(1), cin, cout, cout_pad or ops.round_up(cout, 32), k)

    monkeypatch.setattr(ops, "pack_conv", fake_pack)
    monkeypatch.setattr(ops, "pack_first_conv", lambda w, b=None, bn=None: ops.FirstConv.__new__(ops.FirstConv))
    monkeypatch.setattr(ops, "conv_workspace", lambda dev: torch.zeros(1))
    monkeypatch.setattr(ops, "conv_first", lambda x, f, out, pitch, act, pool=False: first_calls.append((tuple(out.shape), pitch, pool)))
    for name in ("conv_gemm", "maxpool2", "upsample2", "yolo_decode"):
        monkeypatch.setattr(ops, name, lambda *a, **k: None)
    net = __import__("millieye_b200.models", fromlist=["Darknet"]).Darknet(configs.cfg_path("yolov3-tiny-12"))
    tensors = {k: v.detach().float() for k, v in net.state_dict().items() if v.is_floating_point()}

    def build(size):
        first_calls.clear()
        plan = engine.DarknetPlan(net._blocks, tensors, 1, size, torch.device("cpu"), 8)
        plan.enqueue()
        return plan

    monkeypatch.delenv("ME_FUSE_POOL", raising=False)
    plan = build(64)
    assert plan.op_kinds.count("maxpool") == 5 and plan.op_kinds.count("conv") == 13
    assert first_calls[0][2] is True and first_calls[0][1] == 16
    assert plan.feature_view is not None and plan.feature_view.real_c == 256 and plan.feature_view.h == 4
    plan = build(48)
    assert plan.op_kinds.count("maxpool") == 6 and first_calls[0][2] is False
    monkeypatch.setenv("ME_FUSE_POOL", "0")
    plan = build(64)
    assert plan.op_kinds.count("maxpool") == 6 and first_calls[0][2] is False


def test_chain_formation_on_cpu(monkeypatch):
    """DarknetPlan._form_chains without a GPU: Darknet-53's launch list becomes 8 per-layer launches (the 416^2 .. 104^2
    layers with fewer than 128 filters, and the 3x3 layers between them), three chains (the trunk from the last 104^2
    layer down plus the 13^2 head, the 26^2 head, the 52^2 head - each including its fp32 head conv) separated by the two
    upsamples; every chain
    layer names the layer that produces its input / residual, or -1 when that tensor is complete before the chain
    starts; chains that write head logits exist once per output slot."""
    import torch
    from millieye_b200 import engine, ops

    def fake_pack(weight, conv_bias=None, bn=None, cout_pad=None):
        cout, cin, k, _ = weight.shape
        return ops.PackedConv(torch.zeros(1), torch.zeros(1), cin, cout, cout_pad or ops.round_up(cout, 32), k)

    chains = []

    class FakeChain:
        def __init__(self, layers, device):
            self.layers = layers
            chains.append(self)

        def run(self):
            pass

    def eligible(d):
        if d.ksize not in (1, 3) or d.cin % 64:
            return False
        return (d.cout % 128 == 0 and d.res_pitch == 0) if d.out_f32 else d.cout % 64 == 0

    monkeypatch.setattr(ops, "pack_conv", fake_pack)
    monkeypatch.setattr(ops, "pack_first_conv", lambda w, b=None, bn=None: ops.FirstConv.__new__(ops.FirstConv))
    monkeypatch.setattr(ops, "ConvChain", FakeChain)
    monkeypatch.setattr(ops, "conv_chain_eligible", eligible)
    net = __import__("millieye_b200.models", fromlist=["Darknet"]).Darknet(configs.cfg_path("yolov3"))
    tensors = {k: v.detach().float() for k, v in net.state_dict().items() if v.is_floating_point()}
    plan = engine.DarknetPlan(net._blocks, tensors, 1, 64, torch.device("cpu"), None)
    assert not plan.use_chains and len(plan.ops) == 77          # no GPU: one launch per layer
    plan._form_chains()
    kinds = plan.op_kinds
    assert kinds.count("upsample") == 2 and len(chains) == 6     # three chains x two output slots
    sizes = [len(c.layers) for c in chains[::2]]
    assert sizes == [52, 8, 7] and sum(sizes) + 8 == 75          # every conv is launched exactly once
    # launch order: 8 early layers, chain, upsample, chain, upsample, chain
    assert [isinstance(b, list) for b in plan.op_blocks] == [False] * 8 + [True, False, True, False, True]
    for c0, c1 in zip(chains[::2], chains[1::2]):
        heads = [j for j, l in enumerate(c0.layers) if l["desc"].out_f32]
        assert len(heads) == 1
        for j, (l0, l1) in enumerate(zip(c0.layers, c1.layers)):  # the slots differ in the head's output buffer only
            assert (l0["y"].data_ptr() != l1["y"].data_ptr()) == (j in heads)
            assert l0["x"].data_ptr() == l1["x"].data_ptr() and l0["dep"] == l1["dep"]
    for c in chains:
        for j, l in enumerate(c.layers):
            assert l["dep"] < j and l["res"] < j
            if l["dep"] >= 0:                                    # the producer writes exactly what this layer reads
                assert c.layers[l["dep"]]["y"].data_ptr() == l["x"].data_ptr()
            if l["residual"] is not None and l["res"] >= 0:
                assert c.layers[l["res"]]["y"].data_ptr() == l["residual"].data_ptr()
    first = chains[0].layers
    assert first[0]["dep"] == -1 and first[0]["res"] == -1 and first[0]["residual"] is not None   # 104^2 3x3 64->128 + shortcut
    assert first[1]["desc"].stride == 2 and first[1]["desc"].cout == 256
    assert [l["dep"] for l in first[1:50]] == list(range(49))    # a straight line down to the last 13^2 trunk conv
    assert first[50]["desc"].out_f32 == 1 and first[50]["dep"] == 49     # the 13^2 head conv reads the 3x3 before it
    assert first[51]["dep"] == 48 and first[51]["desc"].cout == 256      # route -4: the 1x1 in front of the upsample
    assert sum(1 for l in first if l["residual"] is not None) == 21      # 1 + 8 + 8 + 4 residual blocks
    assert all(l["dep"] == -1 for l in (chains[2].layers[0], chains[4].layers[0]))   # they read concat buffers


def test_bench_detection_matcher():
    """bench.match_detections (the in-bench parity check): equal rows match; a row at the confidence threshold or an NMS
    order flip between two overlapping, equally scored rows is explained; anything else fails the check."""
    import bench
    base = np.array([[10, 10, 60, 60, 0.90, 0.8, 3], [100, 100, 180, 150, 0.50, 0.9, 1], [200, 40, 260, 90, 0.2004, 0.7, 5]],
                    np.float32)

    def run(ref_rows, got_rows):
        det = np.zeros((1, 8, 7), np.float32)
        det[0, :len(got_rows)] = got_rows
        return bench.match_detections([ref_rows], det, np.array([len(got_rows)]), 1, 0.2)

    r = run(base, base + np.array([0.01, -0.01, 0.02, 0.0, 1e-4, -1e-4, 0], np.float32))
    assert r["rows_matched"] == 3 and r["within_tolerance"] and r["max_box_err_rel"] < 1e-3
    r = run(base, base[:2])                              # the third row's confidence is within 1e-3 of the threshold
    assert r["rows_only_one_side"]["conf_within_tol_of_threshold"] == 1 and r["within_tolerance"]
    flip_a = np.array([[300, 300, 360, 360, 0.6000, 0.5, 2]], np.float32)     # two overlapping boxes, scores 1e-4 apart:
    flip_b = np.array([[310, 300, 370, 360, 0.6001, 0.5, 2]], np.float32)     # each side kept the other one
    r = run(np.concatenate([base, flip_a]), np.concatenate([base, flip_b]))
    assert r["rows_only_one_side"]["nms_order_flip_within_tol"] == 2 and r["within_tolerance"]
    r = run(base, np.concatenate([base, np.array([[5, 300, 50, 380, 0.7, 0.9, 4]], np.float32)]))
    assert r["rows_only_one_side"]["unexplained"] == 1 and not r["within_tolerance"]
    r = run(base, base + np.array([0.5, 0, 0, 0, 0, 0, 0], np.float32))   # 0.5 px on a 50 px box: 1e-2 > tolerance
    assert r["rows_matched"] == 3 and not r["within_tolerance"]


def test_chain_schedules_complete_on_cpu(monkeypatch, built_lib):
    """The work lists me_conv_chain_build would write for Darknet-53 at batch 32 / 416^2 (three chains) and for a small
    ragged case, produced WITHOUT a GPU (me_conv_chain_plan: no tensor maps) and replayed by me_conv_chain_verify under the
    kernel's waiting rules: every tile listed exactly once, no pair ever blocked for good, every counter at its target.
    A corrupted list (two items of one pair swapped so that a tile precedes the tile it reads) is reported, not run."""
    import ctypes
    import torch
    from millieye_b200 import engine, ops

    def fake_pack(weight, conv_bias=None, bn=None, cout_pad=None):
        cout, cin, k, _ = weight.shape
        return ops.PackedConv(torch.zeros(1), torch.zeros(1), cin, cout, cout_pad or ops.round_up(cout, 32), k)

    chains = []

    class FakeChain:
        def __init__(self, layers, device):
            self.layers = layers
            chains.append(self)

        def run(self):
            pass

    L = _lib.lib()
    monkeypatch.setattr(ops, "pack_conv", fake_pack)
    monkeypatch.setattr(ops, "pack_first_conv", lambda w, b=None, bn=None: ops.FirstConv.__new__(ops.FirstConv))
    monkeypatch.setattr(ops, "ConvChain", FakeChain)
    monkeypatch.setattr(ops, "_need_cuda", lambda *a: None)
    net = __import__("millieye_b200.models", fromlist=["Darknet"]).Darknet(configs.cfg_path("yolov3"))
    tensors = {k: v.detach().float() for k, v in net.state_dict().items() if v.is_floating_point()}
    plan = engine.DarknetPlan(net._blocks, tensors, 2, 416, torch.device("cpu"), None)   # buffers for 2 frames ...
    plan._form_chains()
    assert [len(c.layers) for c in chains[::2]] == [52, 8, 7]
    BATCH = 32                                                                            # ... schedules for BASELINE's 32

    def plan_blob(layers):
        n = len(layers)
        arr = (_lib.ChainLayer * n)()
        for a, l in zip(arr, layers):
            a.d = l["desc"]
            a.d.n = BATCH
            a.x = a.w_packed = a.bias = a.y = 0x1000          # never dereferenced without tensor maps
            a.residual = 0x1000 if l.get("residual") is not None else None
            a.dep_layer, a.res_layer = l.get("dep", -1), l.get("res", -1)
        nbytes = L.me_conv_chain_blob_bytes(arr, n)
        assert nbytes > 0
        raw = (ctypes.c_ubyte * (nbytes + 128))()
        addr = ctypes.addressof(raw)
        base = addr + (-addr) % 128
        _lib.check(L.me_conv_chain_plan(arr, n, ctypes.c_void_p(base), nbytes), "me_conv_chain_plan")
        return raw, base

    for c in chains[::2]:
        raw, base = plan_blob(c.layers)
        assert L.me_conv_chain_verify(ctypes.c_void_p(base)) == 0
        # load balance of the static lists in the builder's own cost model (K blocks x 512 clocks + 1500 per tile)
        hdr = (ctypes.c_int * 8).from_address(base + 8)
        npairs, stride = hdr[1], hdr[2]
        offs = (ctypes.c_longlong * 4).from_address(base + 8 + 6 * 4)
        work = (ctypes.c_int * (npairs * stride)).from_address(base + offs[1])
        cost_of = [(9 if l["desc"].ksize == 3 else 1) * (l["desc"].cin // 64) * 512 + 1500 for l in c.layers]
        loads = []
        for q in range(npairs):
            t, k = 0, 0
            while work[q * stride + k] >= 0:
                t += cost_of[work[q * stride + k] >> 20]
                k += 1
            loads.append(t)
        assert npairs == 74 and max(loads) <= 1.08 * (sum(loads) / npairs), (max(loads), sum(loads) / npairs)
    # other batch sizes (ragged last m tiles, fewer tiles than pairs) complete too
    for BATCH in (1, 5, 16):
        for c in chains[::2]:
            raw, base = plan_blob(c.layers)
            assert L.me_conv_chain_verify(ctypes.c_void_p(base)) == 0
    BATCH = 32
    # corrupt the first chain's schedule: header = magic u64, then n_layers, npairs, work_stride (ints), ..., work_off
    raw, base = plan_blob(chains[0].layers)
    hdr = (ctypes.c_int * 8).from_address(base + 8)
    n_layers, npairs, stride = hdr[0], hdr[1], hdr[2]
    offs = (ctypes.c_longlong * 4).from_address(base + 8 + 6 * 4)       # layers_off, work_off, counters_off, total_bytes
    work = (ctypes.c_int * (npairs * stride)).from_address(base + offs[1])
    # find a pair that owns a layer-1 tile Y (3x3 / stride 2: 104^2 -> 52^2) AND one of the layer-0 m tiles Y reads, and
    # swap the two items: the pair then reaches Y before the tile Y waits for - a schedule that can never complete
    d0, d1 = chains[0].layers[0]["desc"], chains[0].layers[1]["desc"]
    assert (d1.ksize, d1.stride, d1.h, d1.w) == (3, 2, 104, 104) and d0.cout // 128 >= 1
    tiles_n0, tiles_n1 = (d.cout // (256 if d.cout % 256 == 0 else 128) for d in (d0, d1))   # chain_bn(): 256- or 128-column tiles
    swapped = False
    for q in range(npairs):
        lst, k = [], 0
        while work[q * stride + k] >= 0:
            lst.append(work[q * stride + k])
            k += 1
        assert lst == sorted(lst)                                     # layer-major lists
        own0 = {(it & 0xFFFFF) // tiles_n0: j for j, it in enumerate(lst) if it >> 20 == 0}
        for j, it in enumerate(lst):
            if it >> 20 != 1:
                continue
            tm = (it & 0xFFFFF) // tiles_n1
            m0, m1 = tm * 256, min(tm * 256 + 255, BATCH * 52 * 52 - 1)
            n0, y0, n1, y1 = m0 // 2704, (m0 % 2704) // 52, m1 // 2704, (m1 % 2704) // 52
            lo = (n0 * 104 + max(2 * y0 - 1, 0)) * 104
            hi = (n1 * 104 + min(2 * y1 + 1, 103)) * 104 + 103
            hit = [own0[t] for t in range(lo // 256, hi // 256 + 1) if t in own0]
            if hit:
                i0 = hit[0]
                work[q * stride + i0], work[q * stride + j] = work[q * stride + j], work[q * stride + i0]
                swapped = True
                break
        if swapped:
            break
    assert swapped
    assert L.me_conv_chain_verify(ctypes.c_void_p(base)) != 0
    buf = ctypes.create_string_buffer(512)
    L.me_last_error(buf, 512)
    assert b"cannot complete" in buf.value and b"blocked" in buf.value
