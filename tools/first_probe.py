"""Attribution of the first conv (plain / pooled): ME_FIRST_DBG bits (1 no stores, 2 no image loads, 4 no tile build),
ME_FIRST_EPI=3 (three epilogue groups).  Each configuration runs in its own process (the switches are read once)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from millieye_b200 import ops
    pool = sys.argv[1] == "pool"
    n, s, cout = 32, 416, int(os.environ.get("FIRST_COUT", "16"))
    x = torch.rand(n, 3, s, s, device="cuda")
    wt = torch.randn(cout, 3, 3, 3, device="cuda") * 0.3
    first = ops.pack_first_conv(wt, torch.zeros(cout, device="cuda"), None)
    so = s // 2 if pool else s
    out = torch.zeros(n, so, so, cout, dtype=torch.float16, device="cuda")
    for _ in range(3):
        ops.conv_first(x, first, out, cout, act=1, pool=pool)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_first(x, first, out, cout, act=1, pool=pool)
    e1.record(); torch.cuda.synchronize()
    print(f"{sys.argv[1]:5s} DBG={os.environ.get('ME_FIRST_DBG','0')} COUT={os.environ.get('FIRST_COUT','16')} EPI={os.environ.get('ME_FIRST_EPI_PLAIN','2')}: {e0.elapsed_time(e1)/20*1e3:7.1f} us")
else:
    for cout in ("16", "32"):
        for dbg in ("0", "7", "31"):
            for epi in ("2",):
                env = dict(os.environ, ME_FIRST_DBG=dbg, ME_FIRST_EPI_PLAIN=epi, FIRST_COUT=cout)
                subprocess.run([sys.executable, __file__, "plain"], env=env)
