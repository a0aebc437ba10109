"""Per-kernel parity tests through the C-ABI on a B200 (`-m gpu`): every kernel against the oracle
(oracle/) or the torch fp32 operator on the same seeded inputs, plus the committed golden fixtures."""
import os

import numpy as np
import pytest
from millieye_b200._lib import MeError
import torch
import torch.nn.functional as F

from millieye_b200 import ops
from oracle import boxes as obox
from oracle import darknet as odark
from oracle import roi as oroi
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ------------------------------------------------------------------------------- conv GEMM
_WS = []


def _conv_ws():
    if not _WS:
        _WS.append(ops.conv_workspace(DEV))
    return _WS[0]


def _conv_case(n, h, w, cin, cout, k, s, act, bn=True, res=False, f32=False, seed=0):
    torch.manual_seed(seed)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    bias = None if bn else torch.randn(cout) * 0.1
    bnp = (torch.rand(cout) + 0.5, torch.randn(cout) * 0.1, torch.randn(cout) * 0.1, torch.rand(cout) + 0.5, 1e-5) if bn else None
    in_pitch, cout_pad = ops.round_up(cin, 8), ops.round_up(cout, 32)
    xh = torch.zeros(n, h, w, in_pitch, dtype=torch.float16)
    xh[..., :cin] = x.permute(0, 2, 3, 1).half()
    x_used = xh[..., :cin].float().permute(0, 3, 1, 2).contiguous()
    pad = (k - 1) // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    resid = torch.randn(n, ho, wo, cout_pad).half() if res else None
    packed = ops.pack_conv(wt.to(DEV), None if bias is None else bias.to(DEV),
                           None if bnp is None else tuple(t.to(DEV) if torch.is_tensor(t) else t for t in bnp), cout_pad=cout_pad)
    cin_pad = packed.w.shape[1] // (k * k)
    wq = packed.w.float().cpu().view(cout_pad, k * k, cin_pad)[:cout, :, :cin].permute(0, 2, 1).reshape(cout, cin, k, k)
    ref = F.conv2d(x_used, wq, packed.bias.cpu()[:cout], stride=s, padding=pad)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else (torch.sigmoid(ref) if act == 2 else ref)
    if res:
        ref = ref + resid[..., :cout].float().permute(0, 3, 1, 2)
    out = torch.full((n, ho, wo, cout_pad), float("nan"), dtype=torch.float32 if f32 else torch.float16, device=DEV)
    # like the engine: the workspace lets the pair kernel split the tail tiles along K
    ops.conv_gemm(xh.to(DEV), packed, n, h, w, in_pitch, out, cout_pad, stride=s, act=act,
                  residual=None if resid is None else resid.to(DEV), res_pitch=cout_pad, out_f32=f32, workspace=_conv_ws())
    torch.cuda.synchronize()
    got = out.float().cpu()[..., :cout].permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    # fp32 accumulate over exactly representable fp16 products: only the output rounding differs
    tol = (2e-5 if f32 else 1.5e-3) * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol
    if cout_pad > cout and act != 2:  # padded channels carry zeros
        assert float(out.float().cpu()[..., cout:].abs().max()) == 0.0


CONV_CASES = {
    "1x1_64_64_linear_bias": dict(n=1, h=16, w=16, cin=64, cout=64, k=1, s=1, act=0, bn=False),
    "1x1_128_128": dict(n=2, h=32, w=32, cin=128, cout=128, k=1, s=1, act=1),
    "1x1_256_512": dict(n=2, h=13, w=13, cin=256, cout=512, k=1, s=1, act=1),
    "1x1_ragged_m": dict(n=1, h=13, w=13, cin=64, cout=32, k=1, s=1, act=1),
    "3x3_64_64_13": dict(n=2, h=13, w=13, cin=64, cout=64, k=3, s=1, act=1),
    "3x3_128_256_26": dict(n=3, h=26, w=26, cin=128, cout=256, k=3, s=1, act=1),
    "3x3_s2_64_128": dict(n=2, h=16, w=16, cin=64, cout=128, k=3, s=2, act=1),
    "3x3_s2_32_64": dict(n=2, h=32, w=32, cin=32, cout=64, k=3, s=2, act=1),
    "3x3_bk32": dict(n=2, h=26, w=26, cin=32, cout=64, k=3, s=1, act=1),
    "3x3_bk16": dict(n=2, h=26, w=26, cin=16, cout=32, k=3, s=1, act=1),
    "3x3_residual": dict(n=2, h=13, w=13, cin=64, cout=128, k=3, s=1, act=1, res=True),
    "1x1_head_f32_255": dict(n=2, h=13, w=13, cin=1024, cout=255, k=1, s=1, act=0, bn=False, f32=True),
    "1x1_head_f32_51": dict(n=2, h=26, w=26, cin=256, cout=51, k=1, s=1, act=0, bn=False, f32=True),
    "1x1_490": dict(n=2, h=26, w=26, cin=256, cout=490, k=1, s=1, act=1),
    "3x3_cin384": dict(n=2, h=26, w=26, cin=384, cout=256, k=3, s=1, act=1),
    "1x1_sigmoid_10": dict(n=2, h=26, w=26, cin=128, cout=10, k=1, s=1, act=2, bn=False),
    "fc_490_256": dict(n=300, h=1, w=1, cin=490, cout=256, k=1, s=1, act=1, bn=False),
    "3x3_single_pixel_rows": dict(n=5, h=1, w=7, cin=64, cout=32, k=3, s=1, act=1),
    "3x3_many_tiles": dict(n=4, h=52, w=52, cin=64, cout=128, k=3, s=1, act=1),
    # several tiles per CTA: both epilogue groups, residual prefetch into the staging tile, N=256 tiles
    "3x3_res_thin_multi": dict(n=8, h=104, w=104, cin=32, cout=64, k=3, s=1, act=1, res=True),
    "3x3_res_pair_multi": dict(n=8, h=52, w=52, cin=128, cout=256, k=3, s=1, act=1, res=True),
    "1x1_256_128_multi": dict(n=8, h=52, w=52, cin=256, cout=128, k=1, s=1, act=1),
    "1x1_512_256_res": dict(n=8, h=26, w=26, cin=512, cout=256, k=1, s=1, act=1, res=True),
    # 88 pair tiles on 74 pairs: the 14 tail tiles are split along K (4 slices), partial sums reduced by the last arriver
    "3x3_pair_split_13_res": dict(n=32, h=13, w=13, cin=128, cout=1024, k=3, s=1, act=1, res=True),
    "3x3_pair_split_13": dict(n=32, h=13, w=13, cin=192, cout=1024, k=3, s=1, act=1),
}


@pytest.mark.parametrize("name", sorted(CONV_CASES))
def test_conv_gemm(name):
    _conv_case(**CONV_CASES[name])


def test_conv_gemm_concat_slice():
    """Output written into a channel slice of a wider buffer, input read from another slice (route fusion)."""
    torch.manual_seed(3)
    n, h, w, cin, cout = 2, 13, 13, 64, 64
    wide_in = torch.randn(n, h, w, 192).half()
    wt = torch.randn(cout, cin, 1, 1) / 8
    packed = ops.pack_conv(wt.to(DEV))
    wide_out = torch.zeros(n, h, w, 160, dtype=torch.float16, device=DEV)
    xin = wide_in.to(DEV)
    ops.conv_gemm(xin.view(-1)[64:], packed, n, h, w, 192, wide_out.view(-1)[96:], 160, act=0, cin=64, cout=64)
    torch.cuda.synchronize()
    ref = F.conv2d(wide_in[..., 64:128].float().permute(0, 3, 1, 2), packed.w.float().cpu().view(cout, cin, 1, 1))
    got = wide_out.cpu().float()
    assert float((got[..., 96:160].permute(0, 3, 1, 2) - ref).abs().max()) < 5e-3
    assert float(got[..., :96].abs().max()) == 0.0  # neighbours untouched


def test_conv_rejects_bad_arguments():
    from millieye_b200._lib import MeError
    packed = ops.pack_conv(torch.randn(32, 64, 1, 1, device=DEV))
    x = torch.zeros(1, 4, 4, 64, dtype=torch.float16, device=DEV)
    y = torch.zeros(1, 4, 4, 32, dtype=torch.float16, device=DEV)
    with pytest.raises(MeError):
        ops.conv_gemm(x, packed, 1, 4, 4, 64, y, 32, cout=24)       # cout not a multiple of 32
    with pytest.raises(MeError):
        ops.conv_gemm(x, packed, 1, 4, 4, 60, y, 32)                # pitch < cin
    with pytest.raises(MeError):
        ops.conv_gemm(x.cpu(), packed, 1, 4, 4, 64, y, 32)          # CPU tensor: no fallback


# ------------------------------------------------------------------------------- SIMT glue
@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("cout,act,n,h,w", [(16, 1, 2, 20, 28), (32, 1, 2, 20, 28), (32, 2, 1, 26, 26), (64, 1, 3, 17, 9),
                                             (32, 1, 2, 416, 416)])
def test_conv_first(cout, act, n, h, w, tensor_cores):
    torch.manual_seed(1)
    x = torch.rand(n, 3, h, w)
    wt = torch.randn(cout, 3, 3, 3) * 0.3
    bn = (torch.rand(cout) + 0.5, torch.randn(cout) * 0.1, torch.randn(cout) * 0.1, torch.rand(cout) + 0.5, 1e-5)
    first = ops.pack_first_conv(wt.to(DEV), None, tuple(t.to(DEV) if torch.is_tensor(t) else t for t in bn))
    out = torch.zeros(n, h, w, cout, dtype=torch.float16, device=DEV)
    ops.conv_first(x.to(DEV), first, out, cout, act=act, tensor_cores=tensor_cores)
    ref = F.batch_norm(F.conv2d(x, wt, padding=1), bn[2], bn[3], bn[0], bn[1], False, 0.9, 1e-5)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else torch.sigmoid(ref)
    got = out.float().cpu().permute(0, 3, 1, 2)
    # SIMT path: fp32 math, fp16 output rounding; tensor-core path also rounds inputs/weights to fp16
    tol = (4e-3 if tensor_cores else 2e-3) * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol


@pytest.mark.parametrize("cout,n,h,w", [(16, 2, 36, 96), (32, 3, 64, 32), (16, 8, 416, 416)])
def test_conv_first_pool_equals_two_calls(cout, n, h, w):
    """me_conv_first_tc_pool (2x2 / stride-2 max-pool in the first conv's epilogue) == me_conv_first_tc + me_maxpool2,
    bit for bit; shapes the tile mapping does not cover are refused, not approximated."""
    torch.manual_seed(5)
    x = torch.rand(n, 3, h, w).to(DEV)
    wt = torch.randn(cout, 3, 3, 3) * 0.3
    bn = (torch.rand(cout) + 0.5, torch.randn(cout) * 0.1, torch.randn(cout) * 0.1, torch.rand(cout) + 0.5, 1e-5)
    first = ops.pack_first_conv(wt.to(DEV), None, tuple(t.to(DEV) if torch.is_tensor(t) else t for t in bn))
    full = torch.zeros(n, h, w, cout, dtype=torch.float16, device=DEV)
    ops.conv_first(x, first, full, cout, act=1)
    want = torch.zeros(n, h // 2, w // 2, cout, dtype=torch.float16, device=DEV)
    ops.maxpool2(full, want, n, h, w, cout, cout, cout, 2)
    got = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=torch.float16, device=DEV)
    ops.conv_first(x, first, got, cout, act=1, pool=True)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    if (cout, h) == (16, 36):
        bad = torch.rand(1, 3, 36, 48).to(DEV)      # w % 32 != 0
        with pytest.raises(MeError):
            ops.conv_first(bad, first, torch.zeros(1, 18, 24, cout, dtype=torch.float16, device=DEV), cout, act=1, pool=True)


@pytest.mark.parametrize("cin,cout,n,h,w", [(16, 32, 2, 48, 48), (32, 64, 3, 32, 24), (16, 32, 4, 208, 208), (32, 64, 4, 104, 104),
                                             (16, 64, 1, 16, 8), (32, 32, 2, 10, 16), (16, 32, 1, 6, 24)])
def test_conv_pool_equals_two_calls(cin, cout, n, h, w):
    """me_conv_pool (thin 3x3 conv with the 2x2 / stride-2 max-pool in its epilogue) == me_conv_gemm + me_maxpool2 bit for
    bit, for both tile-block widths (16 columns x 8 rows when w % 16 == 0, else 8 x 16; the last block row of an image may
    be partial); other shapes are refused."""
    torch.manual_seed(7)
    x = (torch.randn(n, h, w, cin) * 0.5).half().to(DEV)
    wt = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    bn = (torch.rand(cout) + 0.5, torch.randn(cout) * 0.1, torch.randn(cout) * 0.1, torch.rand(cout) + 0.5, 1e-5)
    packed = ops.pack_conv(wt.to(DEV), None, tuple(t.to(DEV) if torch.is_tensor(t) else t for t in bn), cout_pad=cout)
    assert ops.conv_pool_supported(packed, n, h, w, cin, cout)
    full = torch.zeros(n, h, w, cout, dtype=torch.float16, device=DEV)
    ops.conv_gemm(x, packed, n, h, w, cin, full, cout, stride=1, act=1)
    want = torch.zeros(n, h // 2, w // 2, cout, dtype=torch.float16, device=DEV)
    ops.maxpool2(full, want, n, h, w, cout, cout, cout, 2)
    got = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=torch.float16, device=DEV)
    ops.conv_pool(x, packed, n, h, w, cin, got, cout, act=1)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    if (h, w) == (48, 48):
        assert not ops.conv_pool_supported(packed, n, 20, 20, cin, cout)
        with pytest.raises(MeError):
            ops.conv_pool(x[:, :20, :20].contiguous(), packed, n, 20, 20, cin, got, cout, act=1)


@pytest.mark.parametrize("stride", [1, 2])
def test_maxpool(stride):
    torch.manual_seed(2)
    n, h, w, c = 2, 13 if stride == 1 else 26, 13 if stride == 1 else 26, 64
    x = torch.randn(n, h, w, c).half()
    ho = h if stride == 1 else h // 2
    out = torch.zeros(n, ho, ho, c, dtype=torch.float16, device=DEV)
    ops.maxpool2(x.to(DEV), out, n, h, w, c, c, c, stride)
    xr = x.float().permute(0, 3, 1, 2)
    if stride == 1:
        xr = F.pad(xr, (0, 1, 0, 1))   # reference models.py:46-49
    ref = F.max_pool2d(xr, 2, stride)
    assert torch.equal(out.float().cpu().permute(0, 3, 1, 2), ref)


def test_upsample_into_slice_and_layout_converts():
    torch.manual_seed(4)
    n, h, w, c = 2, 13, 13, 128
    x = torch.randn(n, h, w, c).half()
    wide = torch.zeros(n, 2 * h, 2 * w, 384, dtype=torch.float16, device=DEV)
    ops.upsample2(x.to(DEV), wide, n, h, w, c, c, 384)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    got = wide.float().cpu()
    assert torch.equal(got[..., :c].permute(0, 3, 1, 2), ref)
    assert float(got[..., c:].abs().max()) == 0.0
    nchw = ops.nhwc_to_nchw_f32(wide.view(-1)[0:], n, 2 * h, 2 * w, c, 384)
    assert torch.equal(nchw.cpu(), ref)
    back = ops.nchw_f32_to_nhwc(nchw.contiguous(), out_pitch=c)
    assert torch.equal(back.float().cpu().permute(0, 3, 1, 2), ref)


# ------------------------------------------------------------------------------- decode
@pytest.mark.parametrize("g,anchors,classes,size", [
    (13, [(81, 82), (135, 169), (344, 319)], 12, 416),
    (26, [(23, 27), (37, 58), (81, 82)], 12, 416),
    (52, [(10, 13), (16, 30), (33, 23)], 80, 416),
    (10, [(116, 90), (156, 198), (373, 326)], 80, 320),
])
def test_yolo_decode(g, anchors, classes, size):
    torch.manual_seed(5)
    n, attrs = 2, 5 + classes
    ch = 3 * attrs
    pitch = ops.round_up(ch, 32)
    logits = torch.randn(n, ch, g, g) * 2
    nhwc = torch.zeros(n, g, g, pitch)
    nhwc[..., :ch] = logits.permute(0, 2, 3, 1)
    rows = 3 * g * g
    out = torch.zeros(n, rows + 7, attrs, device=DEV)
    ops.yolo_decode(nhwc.to(DEV), pitch, out, n, g, anchors, classes, size / g, rows + 7, 7)
    ref = odark.yolo_decode(logits, anchors, classes, size)
    got = out.cpu()
    assert float(got[:, :7].abs().max()) == 0.0
    # same formula in fp32; only exp/sigmoid implementations differ (<= 2 ulp)
    np.testing.assert_allclose(got[:, 7:].numpy(), ref.numpy(), rtol=3e-6, atol=1e-6)


# ------------------------------------------------------------------------------- NMS
def _check_nms(pred, thr, golden=None):
    ref_dets, ref_rows = obox.non_max_suppression_cpp(pred.numpy().copy(), thr)
    dev_pred = pred.clone().to(DEV)
    buf = ops.filter_nms(dev_pred, thr, 0.5, 200, xyxy_inplace=True)
    torch.cuda.synchronize()
    counts = buf.count.cpu().numpy()
    for i in range(pred.shape[0]):
        k = int(counts[i])
        if ref_dets[i] is None:
            assert k == 0
            continue
        assert k == len(ref_dets[i])
        # survivor rows and order are bit-exact: indices, then every float of the row
        np.testing.assert_array_equal(buf.index[i, :k].cpu().numpy(), ref_rows[i])
        np.testing.assert_array_equal(buf.det[i, :k].cpu().numpy(), ref_dets[i])
        if golden is not None:
            np.testing.assert_array_equal(buf.det[i, :k].cpu().numpy(), golden[i])
    # in-place xywh -> xyxy of the whole tensor, like utils.py:354
    np.testing.assert_array_equal(dev_pred[..., :4].cpu().numpy(), obox.xywh2xyxy(pred[..., :4].numpy()))


@pytest.mark.parametrize("tag,mu,thr", [("trick", -6.0, 0.01), ("vanilla", -1.0, 0.2)])
def test_filter_nms_golden(golden_dir, tag, mu, thr):
    g = np.load(os.path.join(golden_dir, "nms_cpp.npz"))
    pred = synth.synth_predictions(2, 2535, 12, seed=7, conf_mu=mu)
    _check_nms(pred, thr, golden=[g[f"{tag}_0"], g[f"{tag}_1"]])


@pytest.mark.parametrize("rows,classes,mu,thr", [
    (2535, 12, -6.0, 0.01), (2535, 12, -2.0, 0.25), (10647, 80, -5.0, 0.01), (10647, 80, -2.5, 0.2),
    (375, 12, 3.0, 0.01), (135, 1, -1.0, 0.5), (10647, 80, 4.0, 0.01),
    # YOLOv3 at 448 / 480 / 512 / 608: 12 257..16 384 rows used to fail the shared-memory limit (keys + fixed areas),
    # 22 743 rows sort in the global workspace
    (12348, 80, -4.0, 0.05), (14175, 12, -3.0, 0.2), (16128, 80, -4.5, 0.01), (22743, 12, -4.0, 0.1),
])
def test_filter_nms_vs_oracle(rows, classes, mu, thr):
    _check_nms(synth.synth_predictions(3, rows, classes, seed=rows + classes, conf_mu=mu), thr)


def test_filter_nms_edges():
    # nothing above threshold; exactly-at-threshold rows are kept (>=); duplicate boxes / equal scores
    pred = synth.synth_predictions(2, 300, 12, seed=1, conf_mu=-9.0)
    pred[0, :, 4] = 0.001
    pred[1, 5, 4] = 0.25
    pred[1, 9, :] = pred[1, 5, :]
    pred[1, 200, :4] = pred[1, 5, :4]
    pred[1, 200, 4] = 0.25
    _check_nms(pred, 0.25)


# ------------------------------------------------------------------------------- RoI gathers
def _roi_inputs(seed, n=2, g=26):
    rng = np.random.RandomState(seed)
    feat = rng.randn(n, 490, g, g).astype(np.float32)
    rfeat = rng.rand(n, 10, g, g).astype(np.float32)
    return feat, rfeat


def _run_roi(feat, rfeat, rois, cap):
    n, _, g, _ = feat.shape
    f = torch.zeros(n, g, g, 512, dtype=torch.float16)
    f[..., :490] = torch.from_numpy(feat).permute(0, 2, 3, 1).half()
    r = torch.zeros(n, g, g, 32, dtype=torch.float16)
    r[..., :10] = torch.from_numpy(rfeat).permute(0, 2, 3, 1).half()
    rois_d = torch.zeros(cap, 5)
    rois_d[:len(rois)] = torch.from_numpy(rois)
    counts = torch.tensor([0, len(rois)], dtype=torch.int32)
    out_ps = torch.full((cap, 512), 7.0, dtype=torch.float16, device=DEV)
    out_ra = torch.full((cap, 496), 7.0, dtype=torch.float16, device=DEV)
    ops.psroi_align(f.to(DEV), n, g, g, 512, 10, 7, 1 / 16., rois_d.to(DEV), counts.to(DEV), cap, out_ps, 512)
    ops.roi_align(r.to(DEV), n, g, g, 32, 10, 7, 1 / 16., rois_d.to(DEV), counts.to(DEV), cap, out_ra, 496)
    # bin-major variant (the layout Network's plan uses): the same numbers under ops.bin_major_perm, bit for bit
    perm = ops.bin_major_perm(10, 7)
    f_bm = torch.zeros_like(f)
    f_bm[..., :490] = f[..., :490][..., perm]
    bm_ps = torch.full((cap, 512), 7.0, dtype=torch.float16, device=DEV)
    bm_ra = torch.full((cap, 496), 7.0, dtype=torch.float16, device=DEV)
    ops.roi_gather_bin_major(f_bm.to(DEV), n, g, g, 512, 10, 7, 1 / 16., rois_d.to(DEV), counts.to(DEV), cap, bm_ps, 512, True)
    ops.roi_gather_bin_major(r.to(DEV), n, g, g, 32, 10, 7, 1 / 16., rois_d.to(DEV), counts.to(DEV), cap, bm_ra, 496, False)
    torch.cuda.synchronize()
    for std, bm in ((out_ps, bm_ps), (out_ra, bm_ra)):
        a, b = std[:, :490][:, perm.to(DEV)], bm[:, :490]
        assert torch.equal(torch.nan_to_num(a.float(), nan=-77.0), torch.nan_to_num(b.float(), nan=-77.0))
        assert torch.equal(torch.nan_to_num(std[:, 490:].float()), torch.nan_to_num(bm[:, 490:].float()))
    # oracle on the same fp16-rounded maps
    ref_ps = oroi.ps_roi_align(f[..., :490].float().permute(0, 3, 1, 2).numpy(), rois)
    ref_ra = oroi.roi_align(r[..., :10].float().permute(0, 3, 1, 2).numpy(), rois)
    return out_ps.float().cpu().numpy(), out_ra.float().cpu().numpy(), ref_ps, ref_ra


def test_roi_gathers(golden_dir):
    g = np.load(os.path.join(golden_dir, "roi_ops.npz"))
    rng = np.random.RandomState(11)
    feat = rng.randn(2, 490, 26, 26).astype(np.float32)
    rfeat = rng.rand(2, 10, 26, 26).astype(np.float32)
    rois = g["rois"]
    cap = 64
    ps, ra, ref_ps, ref_ra = _run_roi(feat, rfeat, rois, cap)
    k = len(rois)
    tol = 2e-3  # fp16 storage of maps and outputs
    np.testing.assert_allclose(ps[:k, :490], ref_ps.reshape(k, -1), atol=tol * np.abs(ref_ps).max())
    np.testing.assert_allclose(ra[:k, :490], ref_ra.reshape(k, -1), atol=tol)
    # and against torchvision's own output on the fp32 maps (fp16 rounding of the inputs included)
    np.testing.assert_allclose(ps[:k, :490], g["ps"].reshape(k, -1), atol=4e-3 * np.abs(g["ps"]).max())
    np.testing.assert_allclose(ra[:k, :490], g["ra"].reshape(k, -1), atol=4e-3)
    assert np.all(ps[:k, 490:] == 0) and np.all(ps[k:] == 0) and np.all(ra[k:] == 0)  # padding / dead rows zeroed


def test_roi_gathers_empty_and_degenerate():
    feat, rfeat = _roi_inputs(3)
    rois = np.array([[0, 50, 50, 50, 50], [1, 400, 400, 500, 500], [0, -40, -40, -20, -20], [1, 0, 0, 416, 416]],
                    dtype=np.float32)
    ps, ra, ref_ps, ref_ra = _run_roi(feat, rfeat, rois, 8)
    good = ~np.isnan(ref_ps.reshape(4, -1))
    np.testing.assert_allclose(ps[:4, :490][good], ref_ps.reshape(4, -1)[good], atol=1e-2)
    np.testing.assert_allclose(ra[:4, :490], ref_ra.reshape(4, -1), atol=2e-3)
    ps, ra, _, _ = _run_roi(feat, rfeat, np.zeros((0, 5), np.float32), 8)
    assert np.all(ps == 0) and np.all(ra == 0)


# ------------------------------------------------------------------------------- stage-3 labelling + loss
def _stage3_case(seed, n_img=500, n_radar=200, frames=8, n_t=60, cap=1024):
    rng = np.random.default_rng(seed)
    R = n_img + n_radar

    def boxes(k):
        c = rng.uniform(20, 396, (k, 2)).astype(np.float32)
        wh = rng.uniform(8, 120, (k, 2)).astype(np.float32)
        return np.concatenate((c - wh / 2, c + wh / 2), 1).astype(np.float32)

    tg = np.zeros((n_t, 6), np.float32)
    tg[:, 0] = rng.integers(0, frames - 1, n_t)          # the last frame has no ground truth at all
    tg[:, 1] = rng.integers(0, 3, n_t)
    tg[:, 2:] = boxes(n_t)
    tg[5] = tg[4]                                         # identical targets: the first maximum must win
    rows = np.zeros((cap, 9), np.float32)
    rois = np.zeros((cap, 5), np.float32)
    rois[:R, 0] = np.sort(rng.integers(0, frames, R))
    rois[:R, 1:] = boxes(R)
    for k in range(0, R, 3):                              # every third proposal sits on / near a target of its frame
        cand = np.where((tg[:, 0] == rois[k, 0]) & (tg[:, 1] == 0))[0]   # class_num is 1: positives are class 0 (:629-630)
        if len(cand):
            j = cand[k % len(cand)]
            rois[k, 1:] = tg[j, 2:] + (rng.normal(0, (0.0, 2.0, 9.0)[(k // 3) % 3], 4)).astype(np.float32)
    far = [k for k in range(n_img) if k % 3]                 # unrelated proposals of other classes: filtered by class
    rows[far, 7] = rng.integers(0, 3, len(far))
    rows[:n_img, 0] = rois[:n_img, 0]
    rows[:n_img, 1:5] = rois[:n_img, 1:]
    rows[:n_img, 5] = rng.uniform(0.02, 0.99, n_img)
    rows[n_img:, 7] = 0
    refine = rng.uniform(0.001, 0.999, (cap, 2)).astype(np.float32)
    refine[3, 0] = 1.0                                    # log(1 - x) = -inf -> clamped at -100 like BCELoss
    regress = rng.normal(0, 0.5, (cap, 4)).astype(np.float32)
    mask = rng.uniform(0.001, 0.999, cap).astype(np.float32)
    return dict(rows=rows, rois=rois, refine=refine, regress=regress, mask=mask, targets=tg, n_img=n_img, R=R, cap=cap)


@pytest.mark.parametrize("seed,n_img,n_radar,n_t", [(0, 500, 200, 60), (1, 0, 37, 12), (2, 300, 0, 0), (3, 1024 - 24, 24, 150)])
def test_stage3_labels_and_loss_vs_oracle(seed, n_img, n_radar, n_t):
    """me_stage3_labels bit-exact against the oracle's obtain_iou_labels (python loop, fp32 steps), me_stage3_loss
    against the oracle's losses (fp32 terms; sums differ only in order: 2e-6 relative)."""
    import random
    from oracle import stage3_loss as s3
    c = _stage3_case(seed, n_img, n_radar, n_t=max(n_t, 6))
    tg = c["targets"][:n_t]
    R, cap = c["R"], c["cap"]
    dev = lambda a, dt=None: torch.from_numpy(a).to(DEV) if dt is None else torch.from_numpy(a).to(DEV, dt)
    rows, rois, counts = dev(c["rows"]), dev(c["rois"]), torch.tensor([n_img, R], dtype=torch.int32, device=DEV)
    lab = torch.full((cap,), -1.0, device=DEV)
    loc = torch.full((cap, 4), -1.0, device=DEV)
    tdev = dev(tg) if n_t else torch.zeros((1, 6), device=DEV)
    ops.stage3_labels(rows, rois, counts, cap, tdev, n_t, lab, loc)
    boxes6 = np.concatenate((c["rois"][:R, :1], c["rows"][:R, 7:8], c["rois"][:R, 1:]), 1)
    want_lab, want_loc = s3.obtain_iou_labels(boxes6, tg.reshape(-1, 6), True)
    got_lab, got_loc = lab.cpu().numpy(), loc.cpu().numpy()
    assert np.array_equal(got_lab[:R], want_lab.reshape(-1)) and np.array_equal(got_loc[:R], want_loc)
    assert not got_lab[R:].any() and not got_loc[R:].any()
    if n_t:
        assert (want_lab > 0.7).sum() > 0 or n_img == 0

    all_boxes = np.zeros((R, 9), np.float32)
    all_boxes[:, 0], all_boxes[:, 1:5], all_boxes[:, 7] = c["rois"][:R, 0], c["rois"][:R, 1:], c["rows"][:R, 7]
    all_boxes[:n_img, 5] = c["rows"][:n_img, 5]
    m = c["mask"][:R]
    masks = np.stack((np.float32(1) - m, m), 1)
    positive = np.concatenate((m[:n_img] > 0.3, m[n_img:] > 0.56))
    random.seed(5)
    want = s3.stage3_losses(all_boxes, masks, c["refine"][:R], c["regress"][:R], n_img, positive, tg.reshape(-1, 6))
    keep = np.zeros(cap, np.uint8)
    keep[:R] = want["sample_filter"]
    out = torch.zeros(16, device=DEV)
    ops.stage3_loss(rois, dev(c["refine"]), dev(c["regress"]), dev(c["mask"]), counts, cap, lab, loc, dev(keep), out,
                    0.7, 0.75, 6, 0.3, 0.56)
    got = out.cpu().numpy()
    for i, k in enumerate(("masks_loss", "conf_loss", "loss_xy", "loss_wh", "category_loss", "loss")):
        assert abs(got[i] - want[k]) <= 2e-6 * max(1.0, abs(float(want[k]))), (k, got[i], want[k])
    assert (got[6], got[7], got[8], got[9]) == (want["metric"]["true"], want["metric"]["positive"], want["metric"]["tp"], R)



# ------------------------------------------------------------------------------- head conv + fused decode
@pytest.mark.parametrize("n,g,cin,classes,size", [(2, 13, 1024, 80, 416), (3, 26, 256, 12, 416), (2, 52, 256, 80, 416),
                                                   (5, 10, 512, 12, 320)])
def test_conv_gemm_yolo_matches_unfused(n, g, cin, classes, size):
    """me_conv_gemm_yolo (decode in the head conv's epilogue) == me_conv_gemm (fp32 logits) + me_yolo_decode, bit for bit."""
    torch.manual_seed(g)
    anchors = [(116, 90), (156, 198), (373, 326)]
    attrs, cout = 5 + classes, 3 * (5 + classes)
    cout_pad = ops.round_up(cout, 32)
    x = (torch.randn(n, g, g, cin, device=DEV) * 0.5).half()
    packed = ops.pack_conv(torch.randn(cout, cin, 1, 1, device=DEV) / cin ** 0.5, torch.randn(cout, device=DEV) * 0.5, None,
                           cout_pad=cout_pad)
    rows_total, row_off = 3 * g * g + 77, 40
    logits = torch.zeros(n, g, g, cout_pad, device=DEV)
    ops.conv_gemm(x, packed, n, g, g, cin, logits, cout_pad, act=ops.ME_ACT_LINEAR, out_f32=True)
    want = torch.full((n, rows_total, attrs), -7.0, device=DEV)
    ops.yolo_decode(logits, cout_pad, want, n, g, anchors, classes, size / g, rows_total, row_off)
    got = torch.full((n, rows_total, attrs), -7.0, device=DEV)
    ops.conv_gemm_yolo(x, packed, n, g, g, cin, got, g, anchors, classes, size / g, rows_total, row_off)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert float(got[:, row_off:row_off + 3 * g * g, 4].min()) >= 0.0          # decoded rows were all written


def test_pair_tail_split_is_deterministic_and_reusable():
    """The split-K tail of the pair kernel: repeated launches (counters reset by the last arriver) give identical
    bits, and equal the unsplit kernel to fp32 summation-order round-off."""
    torch.manual_seed(3)
    n, g, cin, cout = 32, 13, 256, 1024
    x = (torch.randn(n, g, g, cin, device=DEV) * 0.5).half()
    packed = ops.pack_conv(torch.randn(cout, cin, 3, 3, device=DEV) / (cin * 9) ** 0.5, None,
                           (torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV), torch.zeros(cout, device=DEV),
                            torch.ones(cout, device=DEV), 1e-5))
    outs = []
    for k in range(4):
        out = torch.zeros(n, g, g, cout, dtype=torch.float16, device=DEV)
        # the last run has no workspace: unsplit kernel
        ops.conv_gemm(x, packed, n, g, g, cin, out, cout, workspace=_conv_ws() if k < 3 else None)
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert float(outs[0].float().abs().max()) > 0.1
    assert float((outs[0].float() - outs[3].float()).abs().max()) <= 2e-3 * float(outs[3].float().abs().max())


@pytest.mark.parametrize("name", ["3x3_bk32", "3x3_s2_32_64", "3x3_res_thin_multi", "3x3_bk16"])
def test_conv_thin_kernel_all_shapes(name):
    """The cp.async-fed thin kernel also serves 32-channel inputs under ME_CONV_THIN=2 (off by default: no faster than
    the TMA kernel there); run the conv cases that qualify in a child process with it forced on."""
    import subprocess
    import sys
    here = os.path.abspath(__file__)
    code = ("import sys, importlib.util; sys.path.insert(0, %r); "
            "spec = importlib.util.spec_from_file_location('gpu_ops_cases', %r); t = importlib.util.module_from_spec(spec); "
            "spec.loader.exec_module(t); t._conv_case(**t.CONV_CASES[%r]); print('THIN-OK')"
            % (os.path.dirname(os.path.dirname(here)), here, name))
    env = dict(os.environ, ME_CONV_THIN="2")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "THIN-OK" in r.stdout, r.stderr[-2000:]
