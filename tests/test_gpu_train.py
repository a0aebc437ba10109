"""Stage-3 training step on a B200 (`-m gpu`): the fp32 train-mode forward / backward kernels of the heads against the
reference-generated gradient fixture (tests/golden/stage3_grads_tiny12_192.npz) on identical fp32 inputs, the Adam
kernel against torch.optim.Adam, and the whole `Network.forward(..., targets)` -> `loss.backward()` ->
`optimizer.step()` sequence of train.py:185-190 through the drop-in module."""
import os
import random

import numpy as np
import pytest
import torch

from millieye_b200 import configs
from millieye_b200 import stage3_train as st
from millieye_b200.my_models import Network, define_yolo
from millieye_b200.parse_config import parse_model_config
from oracle import stage3_train as ost
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _oracle_step(golden_dir):
    gl = np.load(os.path.join(golden_dir, "stage3_loss_tiny12_192.npz"))
    cfg = configs.cfg_path("yolov3-tiny-12")
    sd = synth.fill_state_dict(Network(define_yolo(cfg), conf_thresh=0.02).state_dict(), seed=6, obj_bias=2.0)
    sdf = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    random.seed(int(gl["sampling_seed"]))
    maps = synth.synth_maps(4, 192, seed=6)
    res = ost.train_step(parse_model_config(cfg), sdf, synth.synth_images(4, 192, seed=6), maps,
                         synth.synth_radar_boxes(4, seed=5), 0.02, gl["targets"])
    return sd, sdf, maps, res, gl


def _check_grads(grads, g, tol=1e-4):
    names = [str(n) for n in g["names"]]
    assert sorted(grads) == sorted(names)
    worst = 0.0
    for name in names:
        got = grads[name].detach().cpu().numpy()
        if "grad/" + name in g.files:
            ref = g["grad/" + name]
            # (a conv / linear bias in front of a train-mode BatchNorm has a zero gradient: only round-off noise of the
            # row sums, a few 1e-6 with atomics and two ranks - hence the floor of the denominator)
            err = np.abs(got - ref).max() / max(np.abs(ref).max(), 5e-2)
        else:
            ref, sums = g["gsample/" + name], g["gsum/" + name]
            err = np.abs(got.reshape(-1)[::37] - ref).max() / max(np.abs(ref).max(), 5e-2)
            assert abs(got.astype(np.float64).sum() - sums[0]) <= 1e-6 + tol * sums[1], name
        assert err <= tol, (name, err)
        worst = max(worst, float(err))
    return worst


def test_head_training_kernels_match_reference_gradients(golden_dir):
    """HeadTrainer on the reference's own fp32 inputs (feature map, proposals, labels, sample): loss terms, all 32
    parameter gradients and the BatchNorm running statistics equal the reference's autograd step to 1e-4."""
    sd, sdf, maps, res, gl = _oracle_step(golden_dir)
    g = np.load(os.path.join(golden_dir, "stage3_grads_tiny12_192.npz"))
    n, gsz = 4, 12
    P = n * gsz * gsz
    feat_rows = res["feat"].permute(0, 2, 3, 1).reshape(P, 256).contiguous().to(DEV)
    maps_rows = maps.permute(0, 2, 3, 1).reshape(P, 3).contiguous().to(DEV)
    rois = res["box_locations"].float().contiguous().to(DEV)
    n_img, n_all = res["n_img"], res["n_all"]
    img_boxes = torch.zeros((max(n_img, 1), 9), dtype=torch.float32)
    img_boxes[:n_img, 5] = res["yolo_vec"][:, 0]
    img_boxes[:n_img, 8] = res["yolo_vec"][:, 1]
    params = {k: v.clone().to(DEV).contiguous() for k, v in sdf.items() if v.is_floating_point()
              and not k.startswith("base_detector.") and "running_" not in k}
    buffers = {k: v.clone().to(DEV) for k, v in sdf.items() if "running_" in k and not k.startswith("base_detector.")}
    tr = st.HeadTrainer(DEV)
    cache = tr.forward(params, buffers, feat_rows, maps_rows, n, gsz, rois, img_boxes.to(DEV), n_img, n_all)
    torch.cuda.synchronize()
    # forward values: regression parameters and the loss terms recomputed from refine / mask
    assert np.abs(cache["regress"][:n_all].cpu().numpy() - res["reg"].numpy()).max() <= 1e-4
    pos = torch.from_numpy(res["pos"])
    sel = torch.from_numpy(res["sample_filter"])
    conf = cache["refine"][:n_all, 0].cpu().double()
    m = cache["mask"][:n_img].cpu().double()
    prob = torch.where(pos[:n_img], m, 1 - m)
    a_f = torch.where(pos[:n_img], torch.tensor(0.75, dtype=torch.float64), torch.tensor(0.25, dtype=torch.float64))
    focal = (-a_f * (1 - prob) ** 2 * prob.log())[sel[:n_img]].sum()
    y = pos.double()
    bce = -(y * conf.log() + (1 - y) * (1 - conf).log())[sel].sum()
    loss = float(focal + bce / 6.0)
    assert abs(loss - float(g["loss"])) <= 1e-4 * float(g["loss"])
    grads = tr.backward(params, cache, pos.to(torch.uint8).to(DEV), sel.to(torch.uint8).to(DEV), 0.75, 6.0, image_path=True)
    torch.cuda.synchronize()
    worst = _check_grads(grads, g)
    for k in g.files:
        if k.startswith("buf/"):
            assert np.abs(buffers[k[4:]].cpu().numpy() - g[k]).max() <= 1e-5, k
    # frozen image path (train.py --pretrained_module2): the 24 radar / fusion gradients are unchanged, the rest absent
    grads2 = tr.backward(params, cache, pos.to(torch.uint8).to(DEV), sel.to(torch.uint8).to(DEV), 0.75, 6.0, image_path=False)
    assert len(grads2) == 24 and all(torch.equal(grads2[k], grads[k]) for k in grads2 if "radar_net" in k or "ensemble" in k)
    print("worst relative gradient error", worst)


@pytest.mark.parametrize("m,n,k", [(21632, 490, 256), (5000, 131, 27), (2400, 256, 490), (300, 70, 40)])
def test_gemm_f32_all_paths_vs_torch(m, n, k):
    """me_gemm_f32 through its three layouts (x W^T with bias + LeakyReLU, dZ W, and the split-K weight gradient dZ^T x) on
    shapes that take the 128 x 128 tiles, the 64 x 64 / 32 x 32 tiles and the split-K paths, against torch in float64;
    column sums (cluster-split reduction) on the same matrices."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(m + n)
    x = torch.randn(m, k, generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(DEV)
    b = torch.randn(n, generator=g).to(DEV)
    out = torch.full((m, n), float("nan"), device=DEV)
    st.gemm_nt(x, w, out, b, 1)
    ref = torch.nn.functional.leaky_relu(x.double() @ w.double().t() + b.double(), 0.1)
    assert float((out.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    dz = torch.randn(m, n, generator=g).to(DEV)
    dx = torch.full((m, k), float("nan"), device=DEV)
    st.gemm_nn(dz, w, dx)
    ref = dz.double() @ w.double()
    assert float((dx.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    dw = torch.full((n, k), float("nan"), device=DEV)
    st.gemm_tn(dz, x, dw)
    ref = dz.double().t() @ x.double()
    assert float((dw.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    cs = torch.full((n,), float("nan"), device=DEV)
    st.colsum(dz, cs)
    assert float((cs.double() - dz.double().sum(0)).abs().max()) <= 1e-5 * max(1.0, float(dz.double().sum(0).abs().max()))
    cs2 = torch.full((n,), float("nan"), device=DEV)
    st.colsum(dz, cs2, y=out)
    ref = (dz.double() * out.double()).sum(0)
    assert float((cs2.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def test_adam_kernel_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(5000, device=DEV)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4)
    flat, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    from millieye_b200 import _lib
    for step in range(1, 6):
        gr = torch.randn(5000, device=DEV) * (10.0 ** (step - 3))
        ref.grad = gr.clone()
        opt.step()
        _lib.check(_lib.lib().me_adam_step(_lib.ptr(flat), _lib.ptr(gr), _lib.ptr(m), _lib.ptr(v), 5000, 5e-4, 0.9, 0.999, 1e-8,
                                           step, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert float((flat - ref.detach()).abs().max()) <= 2e-6


def test_training_step_through_the_drop_in(golden_dir):
    """train.py:169-190 on the drop-in: model.train(); base_detector.eval(); loss, out, metric, att = model(..., targets);
    loss.backward(); optimizer.step().  The detector runs in fp16, so the proposals differ from the fp32 reference in the
    last bits: the loss agrees to 2 %, the gradients point the same way (cosine >= 0.95 per tensor, norms within 25 %), and a step of
    the flat Adam (Stage3Optimizer) equals torch.optim.Adam's on the same gradients."""
    g = np.load(os.path.join(golden_dir, "stage3_grads_tiny12_192.npz"))
    gl = np.load(os.path.join(golden_dir, "stage3_loss_tiny12_192.npz"))
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.02)
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=2.0))
    model.to(DEV)
    model.train()
    model.base_detector.eval()
    imgs, maps = synth.synth_images(4, 192, seed=6), synth.synth_maps(4, 192, seed=6)
    rb = synth.synth_radar_boxes(4, seed=5)
    random.seed(int(gl["sampling_seed"]))
    targets = torch.from_numpy(gl["targets"].copy())
    loss, out, metric, att = model(imgs.to(DEV), maps.to(DEV), rb.clone().to(DEV), 0, targets)
    assert loss.requires_grad and att.shape == (4, 1, 12, 12) and out.shape[1] == 8
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    assert rel <= 2e-2, rel
    assert metric["total"] == int(g["total"])
    before = {k: v.detach().clone() for k, v in model.named_parameters() if not k.startswith("base_detector.")}
    loss.backward()
    torch.cuda.synchronize()
    names = [str(n) for n in g["names"]]
    got = {k: v.grad for k, v in model.named_parameters() if v.grad is not None}
    assert sorted(got) == sorted(names)                       # same set as the reference (net1 / net3 / fusion_head: none)
    for name in names:
        if "grad/" + name not in g.files:
            continue
        ref = g["grad/" + name].reshape(-1).astype(np.float64)
        mine = got[name].cpu().numpy().reshape(-1).astype(np.float64)
        if np.abs(ref).max() < 1e-5:                          # conv bias in front of a train-mode BatchNorm: zero gradient
            assert np.abs(mine).max() < 1e-4
            continue
        cos = float(mine @ ref / (np.linalg.norm(mine) * np.linalg.norm(ref) + 1e-30))
        assert cos >= 0.95, (name, cos)       # measured 0.978 .. 1.000 (the early radar convs are the most sensitive)
        assert 0.8 <= np.linalg.norm(mine) / np.linalg.norm(ref) <= 1.25, name
    # running statistics moved like the reference's (momentum 0.1)
    for k in g.files:
        if k.startswith("buf/"):
            cur = dict(model.named_buffers())[k[4:]].cpu().numpy()
            assert np.abs(cur - g[k]).max() <= 5e-3, k
    assert int(model.img_cnn_layers.net.batch_norm_0.num_batches_tracked) == 1
    # optimizer step: torch.optim.Adam on the module (what train.py does) vs the flat Adam kernel on a copy
    grads = {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    opt.step()
    after_torch = {k: v.detach().clone() for k, v in model.named_parameters() if k in grads}
    with torch.no_grad():
        for k, v in model.named_parameters():
            if k in before:
                v.copy_(before[k])
    flat_opt = st.Stage3Optimizer(model, lr=5e-4)
    for k, gr in grads.items():
        flat_opt.grads[k].copy_(gr)
    flat_opt.step()
    torch.cuda.synchronize()
    for k, v in model.named_parameters():
        if k in after_torch:
            assert float((v.detach() - after_torch[k]).abs().max()) <= 2e-6, k
    # a second step through the flat optimizer without autograd: forward, backward_into(flat gradients), step
    flat_opt.zero_grad()
    random.seed(1)
    loss2, *_ = model(imgs.to(DEV), maps.to(DEV), rb.clone().to(DEV), 0, torch.from_numpy(gl["targets"].copy()))
    model.backward_into(flat_opt.grads)
    flat_opt.all_reduce()
    flat_opt.step()
    torch.cuda.synchronize()
    assert torch.isfinite(flat_opt.flat).all() and float(loss2) > 0
    # back to inference: eval() drops the stale plans and the forward works on the updated weights
    model.eval()
    out_eval = model(imgs.to(DEV), maps.to(DEV), rb.clone().to(DEV), 0)
    assert out_eval.shape[1] == 8


# ------------------------------------------------------------------------------------------------ two ranks
def _shard_worker(rank, world, port, golden_dir, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        sd, sdf, maps, res, gl = _oracle_step(golden_dir)
        n, gsz = 4, 12
        per = n // world
        lo = rank * per
        hw = gsz * gsz
        feat_rows = res["feat"].permute(0, 2, 3, 1).reshape(n * hw, 256)[lo * hw:(lo + per) * hw].contiguous().to(DEV)
        maps_rows = maps.permute(0, 2, 3, 1).reshape(n * hw, 3)[lo * hw:(lo + per) * hw].contiguous().to(DEV)
        rois_all = res["box_locations"].float()
        n_img_all = res["n_img"]
        frame = rois_all[:, 0].long()
        mine = (frame >= lo) & (frame < lo + per)
        idx = torch.nonzero(mine).reshape(-1)                       # image proposals first, then radar: order kept
        n_img = int((idx < n_img_all).sum())
        rois = rois_all[idx].clone()
        rois[:, 0] -= lo
        img_boxes = torch.zeros((max(n_img, 1), 9))
        img_boxes[:n_img, 5] = res["yolo_vec"][idx[:n_img], 0]
        img_boxes[:n_img, 8] = res["yolo_vec"][idx[:n_img], 1]
        pos = torch.from_numpy(res["pos"])[idx].to(torch.uint8).to(DEV)
        sel = torch.from_numpy(res["sample_filter"])[idx].to(torch.uint8).to(DEV)
        params = {k: v.clone().to(DEV).contiguous() for k, v in sdf.items() if v.is_floating_point()
                  and not k.startswith("base_detector.") and "running_" not in k}
        buffers = {k: v.clone().to(DEV) for k, v in sdf.items() if "running_" in k and not k.startswith("base_detector.")}
        tr = st.HeadTrainer(DEV)                                    # two ranks -> synchronised BatchNorm statistics
        assert tr.sync_bn
        cache = tr.forward(params, buffers, feat_rows, maps_rows, per, gsz, rois.contiguous().to(DEV), img_boxes.to(DEV), n_img,
                           len(idx))
        grads = tr.backward(params, cache, pos, sel, 0.75, 6.0, image_path=True)
        flat = torch.cat([grads[k].reshape(-1) for k in st.TRAINABLE])
        st._all_reduce(flat, None)                                  # the gradient all-reduce (SUM: the losses are sums)
        torch.cuda.synchronize()
        out, off = {}, 0
        for k in st.TRAINABLE:
            cnt = grads[k].numel()
            out[k] = flat[off:off + cnt].reshape(grads[k].shape).cpu().numpy()
            off += cnt
        q.put((rank, out, {k: v.cpu().numpy() for k, v in buffers.items()}, len(idx)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_equals_reference_single_process(golden_dir):
    """The batch of the gradient fixture split over two ranks (two frames and their proposals each; both ranks on this
    GPU, gloo for the collectives): with synchronised BatchNorm statistics and a SUM all-reduce of the gradients every
    rank ends up with the gradients - and running statistics - of the reference's single-process step on the whole
    batch, to the same 1e-4 as the single-rank test."""
    import socket
    import torch.multiprocessing as mp
    g = np.load(os.path.join(golden_dir, "stage3_grads_tiny12_192.npz"))
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, golden_dir, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sum(r[3] for r in results) == int(g["total"])
    for rank, grads, bufs, _ in results:
        _check_grads({k: torch.from_numpy(v) for k, v in grads.items()}, g)
        for k in g.files:
            if k.startswith("buf/"):
                assert np.abs(bufs[k[4:]] - g[k]).max() <= 1e-5, (rank, k)
