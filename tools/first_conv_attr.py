"""Times the first conv (3->32, 416^2, batch 32) alone; ME_FIRST_DBG selects an attribution mode."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from millieye_b200 import ops
dev = torch.device("cuda:0")
n, s = 32, 416
x = torch.rand(n, 3, s, s, device=dev)
w = torch.randn(32, 3, 3, 3, device=dev) * 0.2
bn = (torch.ones(32, device=dev), torch.zeros(32, device=dev), torch.zeros(32, device=dev), torch.ones(32, device=dev), 1e-5)
f = ops.pack_first_conv(w, None, bn)
out = torch.empty(n, s, s, 32, dtype=torch.float16, device=dev)
for _ in range(3):
    ops.conv_first(x, f, out, 32, act=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv_first(x, f, out, 32, act=1)
e1.record()
torch.cuda.synchronize()
print("FIRST dbg=%s us=%.1f" % (os.environ.get("ME_FIRST_DBG", "0"), e0.elapsed_time(e1) / 20 * 1e3))
