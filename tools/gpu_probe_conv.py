"""Bring-up probe for the tcgen05 conv kernel: runs each case in its own process (a trapped
kernel must not take the rest down) and prints an error summary per case.

    python tools/gpu_probe_conv.py            # all cases
    python tools/gpu_probe_conv.py CASE_NAME   # one case, in-process
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    # name: dict(n,h,w,cin,cout,k,stride,act,bn,res,f32)
    "g1x1_64_64": dict(n=1, h=16, w=16, cin=64, cout=64, k=1, s=1, act=0, bn=False),
    "g1x1_128_128_leaky": dict(n=2, h=32, w=32, cin=128, cout=128, k=1, s=1, act=1, bn=True),
    "g1x1_256_512": dict(n=2, h=13, w=13, cin=256, cout=512, k=1, s=1, act=1, bn=True),
    "g1x1_tail_m": dict(n=1, h=13, w=13, cin=64, cout=32, k=1, s=1, act=1, bn=True),
    "c3x3_64_64": dict(n=2, h=13, w=13, cin=64, cout=64, k=3, s=1, act=1, bn=True),
    "c3x3_128_256": dict(n=3, h=26, w=26, cin=128, cout=256, k=3, s=1, act=1, bn=True),
    "c3x3_s2_64_128": dict(n=2, h=16, w=16, cin=64, cout=128, k=3, s=2, act=1, bn=True),
    "c3x3_s2_32_64": dict(n=2, h=32, w=32, cin=32, cout=64, k=3, s=2, act=1, bn=True),
    "c3x3_bk32": dict(n=2, h=26, w=26, cin=32, cout=64, k=3, s=1, act=1, bn=True),
    "c3x3_bk16": dict(n=2, h=26, w=26, cin=16, cout=32, k=3, s=1, act=1, bn=True),
    "c3x3_res": dict(n=2, h=13, w=13, cin=64, cout=128, k=3, s=1, act=1, bn=True, res=True),
    "g1x1_f32_255": dict(n=2, h=13, w=13, cin=1024, cout=255, k=1, s=1, act=0, bn=False, f32=True),
    "g1x1_490": dict(n=2, h=26, w=26, cin=256, cout=490, k=1, s=1, act=1, bn=True),
    "c3x3_sigmoid_cin384": dict(n=2, h=26, w=26, cin=384, cout=256, k=3, s=1, act=1, bn=True),
    "big_3x3_256_512": dict(n=8, h=26, w=26, cin=256, cout=512, k=3, s=1, act=1, bn=True, time=True),
    "big_1x1_64_32_208": dict(n=8, h=208, w=208, cin=64, cout=32, k=1, s=1, act=1, bn=True, time=True),
    "big_3x3_32_64_208": dict(n=8, h=208, w=208, cin=32, cout=64, k=3, s=1, act=1, bn=True, time=True),
    # Darknet-53 layer shapes at batch 32 (timing, GPU reference would be too slow on CPU -> spot check only)
    "d53_104_3x3_64_128_res": dict(n=32, h=104, w=104, cin=64, cout=128, k=3, s=1, act=1, bn=True, res=True, time=True, spot=True),
    "d53_52_1x1_256_128": dict(n=32, h=52, w=52, cin=256, cout=128, k=1, s=1, act=1, bn=True, time=True, spot=True),
    "d53_52_3x3_128_256_res": dict(n=32, h=52, w=52, cin=128, cout=256, k=3, s=1, act=1, bn=True, res=True, time=True, spot=True),
    "d53_26_1x1_512_256": dict(n=32, h=26, w=26, cin=512, cout=256, k=1, s=1, act=1, bn=True, time=True, spot=True),
    "d53_26_3x3_256_512_res": dict(n=32, h=26, w=26, cin=256, cout=512, k=3, s=1, act=1, bn=True, res=True, time=True, spot=True),
    "d53_13_1x1_1024_512": dict(n=32, h=13, w=13, cin=1024, cout=512, k=1, s=1, act=1, bn=True, time=True, spot=True),
    "d53_13_3x3_512_1024_res": dict(n=32, h=13, w=13, cin=512, cout=1024, k=3, s=1, act=1, bn=True, res=True, time=True, spot=True),
    "e53_208_3x3_s2_32_64": dict(n=32, h=416, w=416, cin=32, cout=64, k=3, s=2, act=1, bn=True, time=True, noref=True),
    "e53_208_3x3_32_64_res": dict(n=32, h=208, w=208, cin=32, cout=64, k=3, s=1, act=1, bn=True, res=True, time=True, noref=True),
    "e53_208_1x1_64_32": dict(n=32, h=208, w=208, cin=64, cout=32, k=1, s=1, act=1, bn=True, time=True, noref=True),
    "e53_104_3x3_s2_64_128": dict(n=32, h=208, w=208, cin=64, cout=128, k=3, s=2, act=1, bn=True, time=True, noref=True),
    "t12_208_3x3_16_32": dict(n=32, h=208, w=208, cin=16, cout=32, k=3, s=1, act=1, bn=True, time=True, noref=True),
    "t12_104_3x3_32_64": dict(n=32, h=104, w=104, cin=32, cout=64, k=3, s=1, act=1, bn=True, time=True, noref=True),
    "d53_26_3x3_s2_256_512": dict(n=32, h=52, w=52, cin=256, cout=512, k=3, s=2, act=1, bn=True, time=True, spot=True),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    from millieye_b200 import _lib, ops

    c = CASES[name]
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    n, h, w, cin, cout, k, s = c["n"], c["h"], c["w"], c["cin"], c["cout"], c["k"], c["s"]
    if c.get("noref"):  # timing only (too large for a CPU reference): everything generated on the device
        pad = (k - 1) // 2
        ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        xd = torch.randn(n, h, w, cin, device=dev).half()
        packed = ops.pack_conv(torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5, None,
                               (torch.ones(cout, device=dev), torch.zeros(cout, device=dev), torch.zeros(cout, device=dev),
                                torch.ones(cout, device=dev), 1e-5))
        out = torch.empty(n, ho, wo, cout, dtype=torch.float16, device=dev)
        resd = torch.randn(n, ho, wo, cout, device=dev).half() if c.get("res") else None
        run = lambda: ops.conv_gemm(xd, packed, n, h, w, cin, out, cout, stride=s, act=c["act"], residual=resd, res_pitch=cout)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        flops = 2.0 * n * ho * wo * cout * cin * k * k
        byts = xd.numel() * 2 + out.numel() * 2 * (2 if resd is not None else 1)
        print("PROBE " + json.dumps(dict(case=name, ok=True, max_err=0.0, ms=ms, tflops=flops / ms / 1e9, gbs=byts / ms / 1e6,
                                         debug_word=hex(_lib.debug_status()))), flush=True)
        return
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    bias = None if c.get("bn") else torch.randn(cout) * 0.1
    bn = None
    if c.get("bn"):
        bn = (torch.rand(cout) + 0.5, torch.randn(cout) * 0.1, torch.randn(cout) * 0.1, torch.rand(cout) + 0.5, 1e-5)
    in_pitch = ops.round_up(cin, 8)
    cout_pad = ops.round_up(cout, 32)
    xh = torch.zeros(n, h, w, in_pitch, dtype=torch.float16)
    xh[..., :cin] = x.permute(0, 2, 3, 1).half()
    x_used = xh[..., :cin].float().permute(0, 3, 1, 2).contiguous()
    pad = (k - 1) // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    res = None
    if c.get("res"):
        res = torch.randn(n, ho, wo, cout_pad).half()

    packed = ops.pack_conv(wt.to(dev), None if bias is None else bias.to(dev), None if bn is None else tuple(
        t.to(dev) if torch.is_tensor(t) else t for t in bn), cout_pad=cout_pad)
    # reference with the same fp16-rounded folded weights
    cin_pad = packed.w.shape[1] // (k * k)
    wq = packed.w.float().cpu().view(cout_pad, k * k, cin_pad)[:cout, :, :cin].permute(0, 2, 1).reshape(cout, cin, k, k)
    bq = packed.bias.cpu()[:cout]
    ref = F.conv2d(x_used, wq, bq, stride=s, padding=pad)
    if c["act"] == 1:
        ref = F.leaky_relu(ref, 0.1)
    elif c["act"] == 2:
        ref = torch.sigmoid(ref)
    if res is not None:
        ref = ref + res[..., :cout].float().permute(0, 3, 1, 2)

    f32 = bool(c.get("f32"))
    out = torch.full((n, ho, wo, cout_pad), float("nan"), dtype=torch.float32 if f32 else torch.float16, device=dev)
    xd = xh.to(dev)
    resd = None if res is None else res.to(dev)
    ops.conv_gemm(xd, packed, n, h, w, in_pitch, out, cout_pad, stride=s, act=c["act"], residual=resd,
                  res_pitch=cout_pad, out_f32=f32)
    torch.cuda.synchronize()
    got = out.float().cpu()[..., :cout].permute(0, 3, 1, 2)
    err = (got - ref).abs()
    tol = 2e-2 * ref.abs().max().item() if not f32 else 2e-3 * ref.abs().max().item()
    bad = (err > tol) | torch.isnan(got)
    info = dict(case=name, max_err=float(err.nan_to_num(1e9).max()), ref_max=float(ref.abs().max()),
                bad_frac=float(bad.float().mean()), nan=int(torch.isnan(got).sum()), ok=bool(bad.sum() == 0))
    if not info["ok"]:
        b = bad.permute(0, 2, 3, 1).reshape(-1, cout)  # rows = pixels
        rows = b.any(1).nonzero().flatten()
        cols = b.any(0).nonzero().flatten()
        info["bad_rows"] = [int(v) for v in rows[:24]]
        info["n_bad_rows"] = int(rows.numel())
        info["bad_cols"] = [int(v) for v in cols[:24]]
        info["n_bad_cols"] = int(cols.numel())
        info["sample_got"] = [float(v) for v in got.permute(0, 2, 3, 1).reshape(-1, cout)[rows[0], :8]]
        info["sample_ref"] = [float(v) for v in ref.permute(0, 2, 3, 1).reshape(-1, cout)[rows[0], :8]]
    if c.get("time"):
        for _ in range(3):
            ops.conv_gemm(xd, packed, n, h, w, in_pitch, out, cout_pad, stride=s, act=c["act"], out_f32=f32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 20
        for _ in range(iters):
            ops.conv_gemm(xd, packed, n, h, w, in_pitch, out, cout_pad, stride=s, act=c["act"], out_f32=f32)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * n * ho * wo * cout * cin * k * k
        byts = xd.numel() * 2 + out.numel() * out.element_size() + packed.w.numel() * 2
        info.update(ms=ms, tflops=flops / ms / 1e9, gbs=byts / ms / 1e6)
    info["debug_word"] = hex(_lib.debug_status())
    print("PROBE " + json.dumps(info), flush=True)


def main():
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
        return
    flt = os.environ.get("PROBE_FILTER", "")
    for name in CASES:
        if flt and not any(name.startswith(f) for f in flt.split(",")):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=180)
            lines = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            if lines:
                print(lines[-1])
            else:
                print(f"PROBE-FAIL {name} rc={r.returncode} stderr_tail={r.stderr[-600:]!r}")
        except subprocess.TimeoutExpired:
            print(f"PROBE-TIMEOUT {name}")
        print(f"  ({name}: {time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
