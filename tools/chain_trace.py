"""Per-launch device time of a Darknet plan (CUDA events) and the in-kernel wait totals of its conv chains
(me_conv_set_trace): where the persistent multi-layer kernels lose time.

    python tools/chain_trace.py [cfg] [batch] [size]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from millieye_b200 import _lib, configs  # noqa: E402
from millieye_b200.models import Darknet  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 416
    dev = torch.device("cuda:0")
    net = Darknet(configs.cfg_path(cfg)).eval()
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), seed=0, conv_gain=0.6 if cfg == "yolov3" else 1.0))
    net.to(dev)
    plan = net.plan_for(n, size, dev)
    plan.load_input(torch.rand(n, 3, size, size, device=dev))
    for _ in range(2):
        plan.enqueue()
    torch.cuda.synchronize()
    lib = _lib.lib()
    trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    total = 0.0
    for k, (kind, blk, fn) in enumerate(zip(plan.op_kinds, plan.op_blocks, plan.ops)):
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(0, n)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn(0, n)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        total += us
        row = dict(op=k, kind=kind, blocks=blk if not isinstance(blk, list) else [blk[0], blk[-1], len(blk)], us=round(us, 1))
        if isinstance(blk, list):
            trace.zero_()
            lib.me_conv_set_trace(trace.data_ptr())
            fn(0, n)
            torch.cuda.synchronize()
            lib.me_conv_set_trace(None)
            t = trace.view(148, 16).cpu().double()
            lead = t[0::2]
            span = (t[:, 13] - t[:, 0])
            row.update(cta_cycles_mean=int(span.mean()), cta_cycles_max=int(span.max()),
                       setup=int((t[:, 1] - t[:, 0]).mean()),
                       prod_dep_wait=int(t[:, 2].mean()), prod_dep_wait_max=int(t[:, 2].max()),
                       prod_empty_wait=int(t[:, 3].mean()),
                       prod_end=int((t[:, 4] - t[:, 0]).mean()),
                       mma_first=int((lead[:, 7] - lead[:, 0]).mean()), mma_span=int((lead[:, 8] - lead[:, 7]).mean()),
                       mma_full_wait=int(lead[:, 5].mean()), mma_acc_wait=int(lead[:, 6].mean()),
                       items_mean=round(float(lead[:, 14].mean()), 1), items_min=int(lead[:, 14].min()), items_max=int(lead[:, 14].max()),
                       epi_tfull_wait=int(t[:, 9].mean()), epi_stg_wait=int(t[:, 10].mean()),
                       epi_end=int((t[:, 12] - t[:, 0]).mean()))
        print("OP " + json.dumps(row), flush=True)
    print("SUM_US", round(total, 1))


if __name__ == "__main__":
    main()
