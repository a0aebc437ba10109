"""me_conv_chain_* (one persistent kernel for a run of conv layers, tile-level dependencies) against the per-layer
me_conv_gemm launches on the same buffers, through the C-ABI on a B200 (`-m gpu`)."""
import pytest
import torch

from millieye_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _packed(cout, cin, k, seed, head=False):
    g = torch.Generator().manual_seed(seed)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    if head:   # YOLO head conv: bias, no BatchNorm
        return ops.pack_conv(wt.to(DEV), (torch.randn(cout, generator=g) * 0.3).to(DEV), None, cout_pad=cout)
    bn = (torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1, torch.randn(cout, generator=g) * 0.1,
          torch.rand(cout, generator=g) + 0.5, 1e-5)
    return ops.pack_conv(wt.to(DEV), None, tuple(t.to(DEV) if torch.is_tensor(t) else t for t in bn), cout_pad=cout)


def _run_case(n, size, cin0, spec):
    """spec: list of (k, stride, cout, res_from[, "head"]) with res_from = index of the layer whose output is added (or
    None); layer j reads layer j-1 (layer 0 reads the input), or the layer before a head; "head" marks a linear fp32
    layer (YOLO head conv) whose output nothing reads."""
    torch.manual_seed(0)
    x0 = (torch.randn(n, size, size, cin0) * 0.5).half().to(DEV)
    layers, shapes = [], []
    h, c = size, cin0
    src_of = []
    prev = -1
    for j, entry in enumerate(spec):
        k, s, cout, res_from = entry[:4]
        head = len(entry) > 4
        pad = (k - 1) // 2
        ho = (h + 2 * pad - k) // s + 1
        shapes.append((h, c, ho, cout))
        layers.append(dict(k=k, s=s, packed=_packed(cout, c, k, 100 + j, head), res_from=res_from, head=head))
        src_of.append(prev)
        if not head:
            h, c, prev = ho, cout, j

    def outputs():
        return [torch.full((n, ho, ho, cout), float("nan"), dtype=torch.float32 if L["head"] else torch.float16, device=DEV)
                for (_, _, ho, cout), L in zip(shapes, layers)]

    # per-layer reference launches
    ref = outputs()
    for j, L in enumerate(layers):
        h_in, c_in, ho, cout = shapes[j]
        src = x0 if src_of[j] < 0 else ref[src_of[j]]
        res = None if L["res_from"] is None else ref[L["res_from"]]
        ops.conv_gemm(src, L["packed"], n, h_in, h_in, c_in, ref[j], cout, stride=L["s"], act=0 if L["head"] else 1, residual=res,
                      res_pitch=0 if res is None else cout, out_f32=L["head"])
    torch.cuda.synchronize()

    got = outputs()
    chain_layers = []
    for j, L in enumerate(layers):
        h_in, c_in, ho, cout = shapes[j]
        res = None if L["res_from"] is None else got[L["res_from"]]
        desc = ops.conv_desc(L["packed"], n, h_in, h_in, c_in, cout, stride=L["s"], act=0 if L["head"] else 1,
                             res_pitch=0 if res is None else cout, out_f32=L["head"])
        assert ops.conv_chain_eligible(desc)
        chain_layers.append(dict(desc=desc, x=x0 if src_of[j] < 0 else got[src_of[j]], packed=L["packed"], y=got[j], residual=res,
                                 dep=src_of[j], res=-1 if L["res_from"] is None else L["res_from"]))
    chain = ops.ConvChain(chain_layers, torch.device(DEV))
    for _ in range(3):   # counters are reset by every run
        chain.run()
    torch.cuda.synchronize()
    for j in range(len(layers)):
        a, b = got[j].float(), ref[j].float()
        assert not torch.isnan(a).any(), f"layer {j}: unwritten output"
        err = float((a - b).abs().max())
        # same tiles, same K order: identical up to the fp16 rounding of upstream layers (the per-layer path may
        # use other tile shapes / split-K for some layers)
        assert err <= 2e-3 * max(1.0, float(b.abs().max())), f"layer {j}: max abs err {err}"


def test_chain_small_ragged():
    # 20x20x2 = 800 rows: 4 m tiles, the last one with 32 live rows in CTA 0 and none in CTA 1
    _run_case(2, 40, 64, [(3, 2, 128, None), (1, 1, 256, None), (3, 1, 128, 0), (3, 2, 256, None), (1, 1, 256, None)])


def test_chain_residual_blocks_52():
    # Darknet-53's 52^2 stage at batch 8: 85 m tiles, several work items per SM pair, N = 128 and N = 256 tiles
    _run_case(8, 104, 128, [(3, 2, 256, None), (1, 1, 128, None), (3, 1, 256, 0), (1, 1, 128, None), (3, 1, 256, 2),
                            (1, 1, 128, None), (3, 1, 256, 4)])


def test_chain_13_stage_long_k():
    # 13^2 at batch 32: 22 m tiles x 4 n tiles per 3x3 layer, 72 K blocks per tile, tiles of consecutive layers
    # overlap in time across pairs
    _run_case(32, 26, 512, [(3, 2, 1024, None), (1, 1, 512, None), (3, 1, 1024, 0), (1, 1, 512, None), (3, 1, 1024, 2)])


def test_chain_thin_layers_and_fp32_head():
    # 64-column tiles (Darknet-53's 104^2 bottlenecks: 1x1 128 -> 64), a residual across them, and a linear fp32 head conv
    # in the middle of the run (its output feeds nothing; the next layer reads the layer in front of it)
    _run_case(4, 104, 64, [(3, 2, 128, None), (1, 1, 64, None), (3, 1, 128, 0), (1, 1, 64, None), (3, 1, 128, 2),
                           (1, 1, 256, None, "head"), (1, 1, 64, None), (3, 1, 256, None), (1, 1, 128, None, "head")])


def test_chain_rejects_ineligible():
    from millieye_b200._lib import MeError
    p = _packed(96, 64, 1, 1)
    d = ops.conv_desc(p, 1, 8, 8, 64, 96)
    assert not ops.conv_chain_eligible(d)
    x = torch.zeros(1, 8, 8, 64, dtype=torch.float16, device=DEV)
    y = torch.zeros(1, 8, 8, 96, dtype=torch.float16, device=DEV)
    with pytest.raises(MeError):
        ops.ConvChain([dict(desc=d, x=x, packed=p, y=y, residual=None, dep=-1, res=-1)] * 2, torch.device(DEV))
