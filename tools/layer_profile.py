"""Per-op device times of one Darknet plan (CUDA events around repeated launches of each op).
Back-to-back repeats keep small layers L2-warm, so read the numbers as per-layer ceilings and use the
ncu launch list (profiles/) for the in-sequence shares.

    python tools/layer_profile.py [cfg] [batch] [size]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from millieye_b200 import configs  # noqa: E402
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
    total = 0.0
    rows = []
    for (i, b), fn in zip(kinds, plan.ops):
        for _ in range(2):
            fn(0, n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            fn(0, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        total += ms
        row = dict(block=i, type=b["type"], ms=round(ms, 4))
        hw = plan.hw[i]
        if b["type"] == "convolutional":
            flops = 2.0 * n * hw * hw * b["filters"] * b["cin"] * b["size"] ** 2
            hin = hw * b["stride"]
            byts = 2.0 * n * (hin * hin * b["cin"] + hw * hw * b["filters"]) + 2.0 * b["filters"] * b["cin"] * b["size"] ** 2
            nxt = plan.blocks[i + 1]["type"] if i + 1 < len(plan.blocks) else ""
            if nxt == "shortcut":
                byts += 2.0 * n * hw * hw * b["filters"]
            row.update(k=b["size"], s=b["stride"], cin=b["cin"], cout=b["filters"], hw=hw, gflop=round(flops / 1e9, 2),
                       tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 1), res=(nxt == "shortcut"))
        rows.append(row)
        print("LAYER " + json.dumps(row), flush=True)
    conv_ms = sum(r["ms"] for r in rows if r["type"] == "convolutional")
    print("TOTAL " + json.dumps(dict(cfg=cfg, n=n, size=size, sum_ms=round(total, 3), conv_ms=round(conv_ms, 3),
                                     ops=len(rows))), flush=True)


if __name__ == "__main__":
    main()
