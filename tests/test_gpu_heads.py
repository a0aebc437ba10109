"""Per-kernel parity of the fusion tail on a B200 (`-m gpu`): me_build_proposals, me_fusion_heads, me_stage2_heads and
me_finalize_output on IDENTICAL inputs against an fp32 torch / numpy restatement of the reference lines they replace
(my_models.py:459-539, module2_mixed/my_models.py:121-161,325-358).  With the same numbers going in, the only differences
left are fp32 summation order and exp / sigmoid implementations: values to 2e-6, indices, counts and row order exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from millieye_b200 import ops

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _close(a, b, tol=2e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), float(np.abs(a - b).max())


# ------------------------------------------------------------------------------------------------ proposals
@pytest.mark.parametrize("class_idx,n_radar", [(0, 5), (3, 0), (-1, 0)])
def test_build_proposals_exact(class_idx, n_radar):
    rng = np.random.RandomState(3)
    n, max_det, nc = 4, 200, 12
    cols = 7 + nc
    det = rng.rand(n, max_det, cols).astype(np.float32) * 100
    det[..., 6] = rng.randint(0, 5, (n, max_det))
    counts = np.array([200, 0, 37, 120], dtype=np.int32)
    radar = rng.rand(max(n_radar, 1), 5).astype(np.float32)
    radar[:, 0] = rng.randint(0, n, len(radar))
    cap = n * max_det + 8
    pitch = 9 if class_idx >= 0 else 1 + cols
    img_boxes = torch.zeros((cap, pitch), device=DEV)
    rois = torch.zeros((cap, 5), device=DEV)
    cnt = torch.zeros(2, dtype=torch.int32, device=DEV)
    ops.build_proposals(torch.from_numpy(det).to(DEV), torch.from_numpy(counts).to(DEV), class_idx,
                        torch.from_numpy(radar[:n_radar]).to(DEV) if n_radar else None, 416.0, img_boxes, rois, cnt, cap)
    torch.cuda.synchronize()
    # reference: my_models.py:459-473 (image-major order, class filter) + :490-492 (radar rows scaled by the image size)
    rows = []
    for i in range(n):
        d = det[i, :counts[i]]
        if class_idx >= 0:
            d = d[d[:, 6] == class_idx]
            b = np.concatenate((np.full((len(d), 1), i, np.float32), d[:, :7], d[:, 7 + class_idx:8 + class_idx]), 1)
        else:
            b = np.concatenate((np.full((len(d), 1), i, np.float32), d), 1)
        rows.append(b)
    ref = np.concatenate(rows, 0)
    n_img, n_all = (int(v) for v in cnt.tolist())
    assert n_img == len(ref) and n_all == len(ref) + n_radar
    np.testing.assert_array_equal(img_boxes[:n_img].cpu().numpy(), ref)
    np.testing.assert_array_equal(rois[:n_img].cpu().numpy(), ref[:, :5])
    if n_radar:
        rr = radar[:n_radar].copy()
        rr[:, 1:] *= np.float32(416.0)
        np.testing.assert_array_equal(rois[n_img:n_all].cpu().numpy(), rr)


# ------------------------------------------------------------------------------------------------ heads
def _head_params(seed):
    g = torch.Generator().manual_seed(seed)

    def r(*shape, s=1.0):
        return (torch.randn(*shape, generator=g) * s).contiguous()
    return dict(net1_w=r(4, 256, s=0.06), net1_b=r(4, s=0.1), net2_w=r(13, 256, s=0.06), net2_b=r(13, s=0.1),
                radar_w=r(10, 490, s=0.05), radar_b=r(10, s=0.1), radar2_w=r(10, s=0.4), radar2_b=r(1, s=0.1),
                fc1_w=r(32, 2, s=0.7), fc1_b=r(32, s=0.1), fc2_w=r(2, 64, s=0.3), fc2_b=r(2, s=0.1))


def test_fusion_heads_on_identical_inputs():
    torch.manual_seed(0)
    n_img, n_rad = 700, 45
    R = n_img + n_rad
    cap = R + 11
    hidden = (torch.randn(cap, 256) * 0.8).half()
    crop = torch.zeros(cap, 496, dtype=torch.float16)
    crop[:, :490] = torch.rand(cap, 490).half()
    img_boxes = torch.rand(cap, 9)
    w = _head_params(1)
    wd = {k: v.to(DEV) for k, v in w.items()}
    hw = ops.make_head_weights(wd)
    regress = torch.full((cap, 4), float("nan"), device=DEV)
    refine = torch.full((cap, 2), float("nan"), device=DEV)
    mask = torch.full((cap,), float("nan"), device=DEV)
    counts = torch.tensor([n_img, R], dtype=torch.int32, device=DEV)
    ops.fusion_heads(hidden.to(DEV), 256, crop.to(DEV), 496, hw, img_boxes.to(DEV), counts, cap, regress, refine, mask)
    torch.cuda.synchronize()
    # reference (my_models.py:264-284, 202-210, 513-514) in fp64 on the same numbers
    t = hidden[:R].double()
    reg = t @ w["net1_w"].double().t() + w["net1_b"].double()
    cls = torch.sigmoid(t @ w["net2_w"].double().t() + w["net2_b"].double())
    r1 = F.leaky_relu(crop[:R, :490].double() @ w["radar_w"].double().t() + w["radar_b"].double(), 0.1)
    rc = torch.sigmoid(r1 @ w["radar2_w"].double() + w["radar2_b"].double())
    conf = torch.sigmoid(rc + cls[:, 0])
    ref_vec = torch.stack((conf, cls[:, 1]), 1)
    yolo_vec = torch.stack((img_boxes[:n_img, 5], img_boxes[:n_img, 8]), 1).double()
    u = torch.stack((ref_vec[:n_img], yolo_vec), -1)
    h = F.leaky_relu(u @ w["fc1_w"].double().t() + w["fc1_b"].double(), 0.1).flatten(1)
    p = torch.softmax(h @ w["fc2_w"].double().t() + w["fc2_b"].double(), 1)
    m = torch.cat((p[:, 0], conf[n_img:]))
    _close(regress[:R].cpu().numpy(), reg.numpy())
    _close(refine[:R].cpu().numpy(), ref_vec.numpy())
    _close(mask[:R].cpu().numpy(), m.numpy())
    assert torch.isnan(mask[R:]).all() and torch.isnan(regress[R:]).all()          # rows past the live count untouched


def test_stage2_heads_on_identical_inputs():
    torch.manual_seed(1)
    R, cap, nv = 333, 350, 13
    hidden = (torch.randn(cap, 256) * 0.8).half()
    boxes = torch.rand(cap, 8 + 12)
    g = torch.Generator().manual_seed(5)
    w = dict(net1_w=torch.randn(4, 256, generator=g) * 0.06, net1_b=torch.randn(4, generator=g) * 0.1,
             net2_w=torch.randn(13, 256, generator=g) * 0.06, net2_b=torch.randn(13, generator=g) * 0.1,
             fc1_w=torch.randn(32, 2, generator=g) * 0.7, fc1_b=torch.randn(32, generator=g) * 0.1,
             fc2_w=torch.randn(2, 32 * nv, generator=g) * 0.1, fc2_b=torch.randn(2, generator=g) * 0.1)
    wd = {k: v.contiguous().to(DEV) for k, v in w.items()}
    weights = ops.make_stage2_weights(wd)
    regress = torch.zeros((cap, 4), device=DEV)
    mask = torch.zeros((cap,), device=DEV)
    refine = torch.zeros((cap, nv), device=DEV)
    counts = torch.tensor([R, R], dtype=torch.int32, device=DEV)
    ops.stage2_heads(hidden.to(DEV), 256, weights, boxes.to(DEV), 20, nv, counts, cap, regress, mask, refine)
    torch.cuda.synchronize()
    t = hidden[:R].double()
    reg = t @ w["net1_w"].double().t() + w["net1_b"].double()
    ref_vec = torch.sigmoid(t @ w["net2_w"].double().t() + w["net2_b"].double())
    yolo_vec = torch.cat((boxes[:R, 5:6], boxes[:R, 8:]), 1).double()
    x = torch.stack((ref_vec, yolo_vec), -1)
    x = F.leaky_relu(x @ w["fc1_w"].double().t() + w["fc1_b"].double(), 0.1).flatten(1)
    x = F.leaky_relu(x @ w["fc2_w"].double().t() + w["fc2_b"].double(), 0.1)
    p = torch.softmax(x, 1)
    _close(regress[:R].cpu().numpy(), reg.numpy())
    _close(refine[:R].cpu().numpy(), ref_vec.numpy())
    _close(mask[:R].cpu().numpy(), p[:, 1].numpy())


# ------------------------------------------------------------------------------------------------ output stage
@pytest.mark.parametrize("thr_img,thr_radar,do_regress", [(0.0, 0.0, True), (0.4, 0.56, True), (1.0, 0.3, False)])
def test_finalize_output_rows_and_order(thr_img, thr_radar, do_regress):
    rng = np.random.RandomState(11)
    n_img, n_rad = 900, 60
    R = n_img + n_rad
    cap = R + 5
    img_boxes = rng.rand(cap, 9).astype(np.float32)
    img_boxes[:, 0] = rng.randint(0, 32, cap)
    rois = np.zeros((cap, 5), np.float32)
    rois[:, 0] = img_boxes[:, 0]
    xy = rng.rand(cap, 2).astype(np.float32) * 300
    wh = rng.rand(cap, 2).astype(np.float32) * 100 + 4
    rois[:, 1:3], rois[:, 3:5] = xy, xy + wh
    refine = rng.rand(cap, 2).astype(np.float32)
    regress = (rng.randn(cap, 4) * 0.2).astype(np.float32)
    mask = rng.rand(cap).astype(np.float32)
    mask[5] = mask[17] = mask[400]                       # ties: the lower row index comes first (stable sort)
    out = torch.zeros((cap, 8), device=DEV)
    out_count = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = torch.zeros(int(ops._lib.lib().me_finalize_workspace(cap)), dtype=torch.uint8, device=DEV)
    dev = [torch.from_numpy(a).to(DEV) for a in (img_boxes, rois, refine, regress, mask)]
    counts = torch.tensor([n_img, R], dtype=torch.int32, device=DEV)
    ops.finalize_output(dev[0], dev[1], dev[2], dev[3], dev[4], counts, cap, thr_img, thr_radar, do_regress, out, out_count, ws)
    torch.cuda.synchronize()
    # reference: my_models.py:516-539 + box_regress :378-391
    m = torch.from_numpy(mask[:R])
    positive = torch.cat((m[:n_img] > thr_img, m[n_img:] > thr_radar))
    pri = m.clone()
    pri[n_img:] /= 5
    idx = torch.nonzero(positive).reshape(-1)
    order = idx[torch.sort(pri[idx], descending=True, stable=True).indices].numpy()
    k = int(out_count.item())
    assert k == len(order)
    got = out[:k].cpu().numpy()
    r = torch.from_numpy(rois[order, 1:5]).double()
    if do_regress:
        g = torch.from_numpy(regress[order]).double()
        cx, cy, w, h = (r[:, 0] + r[:, 2]) / 2, (r[:, 1] + r[:, 3]) / 2, r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]
        nx, ny, nw, nh = g[:, 0] * w + cx, g[:, 1] * h + cy, torch.exp(g[:, 2]) * w, torch.exp(g[:, 3]) * h
        r = torch.stack((nx - nw / 2, ny - nh / 2, nx + nw / 2, ny + nh / 2), 1)
    np.testing.assert_array_equal(got[:, 0], rois[order, 0])                    # the same proposals in the same order
    _close(got[:, 1:5], r.numpy(), tol=2e-6)
    np.testing.assert_array_equal(got[:, 5], mask[order])
    is_img = order < n_img
    np.testing.assert_array_equal(got[is_img, 6], img_boxes[order[is_img], 6])
    np.testing.assert_array_equal(got[is_img, 7], img_boxes[order[is_img], 7])
    np.testing.assert_array_equal(got[~is_img, 6], refine[order[~is_img], 1])
    assert (got[~is_img, 7] == 0).all()
