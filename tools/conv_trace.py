"""In-kernel phase timeline of every conv GEMM launch of one Darknet plan (me_conv_set_trace).

For each conv layer: cycles per CTA (mean over CTAs; pair kernels: leader CTAs for the MMA columns) of
  setup    kernel entry -> barriers/TMEM ready
  pdl      griddepcontrol.wait
  fill     first operand stage landed (after pdl)
  mma      first MMA -> last commit issued        (w_full / w_acc: cycles of it spent waiting on operands / accumulator)
  tail     last commit issued -> CTA end           (last epilogue + store drain)
  p_empty  producer cycles waiting for a free stage
  e_tfull / e_other / e_work   epilogue: waiting for accumulators / staging+residual / doing the work
and the event-timed duration of the launch.

    python tools/conv_trace.py [cfg] [batch] [size]
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
    plan.splits = 1
    plan.load_input(torch.rand(n, 3, size, size, device=dev))
    plan.enqueue()
    torch.cuda.synchronize()
    kinds = [(i, plan.blocks[i]) for i in plan.op_blocks]
    trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib = _lib.lib()
    only = [int(v) for v in os.environ.get("ONLY_BLOCKS", "").split(",") if v]
    for (i, b), fn in zip(kinds, plan.ops):
        if b["type"] != "convolutional" or i == 0 or (only and i not in only):
            continue
        fn(0, n)
        torch.cuda.synchronize()
        trace.zero_()
        lib.me_conv_set_trace(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(0, n)
        e1.record()
        torch.cuda.synchronize()
        lib.me_conv_set_trace(None)
        t = trace.view(148, 16).cpu().double()
        live = t[:, 0] > 0
        t = t[live]
        mma = t[t[:, 7] > 0]
        hw = plan.hw[i]
        row = dict(block=i, hw=hw, k=b["size"], s=b["stride"], cin=b["cin"], cout=b["filters"],
                   res=(i + 1 < len(plan.blocks) and plan.blocks[i + 1]["type"] == "shortcut"),
                   us=round(e0.elapsed_time(e1) * 1e3, 1), ctas=int(live.sum()), tiles=round(float(mma[:, 14].mean()), 2),
                   total=int((t[:, 13] - t[:, 0]).mean()), total_max=int((t[:, 13] - t[:, 0]).max()),
                   setup=int((t[:, 1] - t[:, 0]).mean()), pdl=int((t[:, 2] - t[:, 1]).mean()),
                   fill=int((mma[:, 7] - mma[:, 2]).mean()), mma=int((mma[:, 8] - mma[:, 7]).mean()),
                   w_full=int(mma[:, 5].mean()), w_acc=int(mma[:, 6].mean()),
                   tail=int((mma[:, 13] - mma[:, 8]).mean()), p_empty=int(t[:, 3].mean()),
                   e_tfull=int(t[:, 9].mean()), e_other=int(t[:, 10].mean()), e_work=int(t[:, 11].mean()))
        print("TRACE " + json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
