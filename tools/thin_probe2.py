"""Attribution of conv_thin (ME_THIN_DBG bits: 1 no output stores, 2 no input copies, 4 no MMAs); one process per setting."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from millieye_b200 import ops
    cin, cout, size = (int(v) for v in sys.argv[1:4])
    n = 32
    x = (torch.randn(n, size, size, cin, device="cuda") * 0.5).half()
    wt = torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
    packed = ops.pack_conv(wt, torch.zeros(cout, device="cuda"), None, cout_pad=cout)
    y = torch.zeros(n, size, size, cout, dtype=torch.float16, device="cuda")
    run = lambda: ops.conv_gemm(x, packed, n, size, size, cin, y, cout, stride=1, act=1)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    print(f"{cin}->{cout} @{size}  DBG={os.environ.get('ME_THIN_DBG','0')} THIN={os.environ.get('ME_CONV_THIN','1')}: {e0.elapsed_time(e1)/20*1e3:7.1f} us", flush=True)
else:
    for shape in (("16", "32", "208"),):
        for dbg in ("0", "1", "2", "3", "4", "7"):
            env = dict(os.environ, ME_THIN_DBG=dbg, ME_CONV_THIN="2")
            subprocess.run([sys.executable, __file__, *shape], env=env)
