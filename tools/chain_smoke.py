"""One Darknet-53 forward through the conv chains with the watchdog word decoded on failure:
    python tools/chain_smoke.py [batch] [size]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from millieye_b200 import _lib, configs  # noqa: E402
from millieye_b200.models import Darknet  # noqa: E402
from oracle import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
dev = torch.device("cuda:0")
net = Darknet(configs.cfg_path("yolov3")).eval()
net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=0, conv_gain=0.6))
net.to(dev)
net.use_cuda_graph = False
x = torch.rand(n, 3, size, size, device=dev)
try:
    for it in range(3):
        _, y = net(x)
        torch.cuda.synchronize()
    print("chain smoke ok", tuple(y.shape), float(y.abs().max()))
    plan = net.plan_for(n, size, dev)
    print("ops:", [k if not isinstance(b, list) else f"chain{len(b)}" for k, b in zip(plan.op_kinds, plan.op_blocks)])
except Exception as e:  # noqa: BLE001
    w = _lib.debug_status()
    print("FAILED:", str(e).splitlines()[0])
    print("debug word 0x%016x: tag 0x%x cta %d aux 0x%x" % (w, w >> 32, (w >> 8) & 0xffffff, w & 0xff))
    sys.exit(1)
