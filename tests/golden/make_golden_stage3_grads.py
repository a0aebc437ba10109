"""Generates tests/golden/stage3_grads_tiny12_192.npz: the UNMODIFIED reference's stage-3 training step in the mode
train.py runs it (module3_our_dataset/train.py:169-186: model.train(), base_detector.eval(), forward with targets,
loss.backward()) on seeded inputs - the loss, the gradient of every parameter that receives one, and the BatchNorm
running statistics after the step.  It pins the oracle of the not-yet-built backward pass (oracle/stage3_train.py).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_stage3_grads.py
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import M3, import_reference  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    torch.set_num_threads(4)
    _, my_models, _ = import_reference()
    g = np.load(os.path.join(HERE, "stage3_loss_tiny12_192.npz"))     # same inputs and targets as the loss fixture
    cfg = os.path.join(M3, "config", "yolov3-tiny-12.cfg")
    n, size = 4, 192
    model = my_models.Network(my_models.define_yolo(cfg), conf_thresh=0.02)
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=2.0))
    model.train()
    model.base_detector.eval()
    imgs, maps = synth.synth_images(n, size, seed=6), synth.synth_maps(n, size, seed=6)
    rb = synth.synth_radar_boxes(n, seed=5)
    random.seed(int(g["sampling_seed"]))
    loss, output, metric, attention = model(imgs, maps, rb.clone(), 0, torch.from_numpy(g["targets"].copy()))
    loss.backward()
    out = dict(loss=np.float32(loss.item()), output=output.detach().numpy(), total=np.int64(metric["total"]),
               true=np.int64(int(metric["true"])))
    names = []
    for name, p in model.named_parameters():
        if p.grad is not None and not name.startswith("base_detector."):
            gr = p.grad.numpy()
            if gr.size <= 20000:
                out["grad/" + name] = gr
            else:   # the three large matrices are kept as every 37th element plus (sum, sum of magnitudes) in float64
                out["gsample/" + name] = gr.reshape(-1)[::37].copy()
                out["gsum/" + name] = np.array([gr.astype(np.float64).sum(), np.abs(gr.astype(np.float64)).sum()])
            names.append(name)
    for name, b in model.named_buffers():
        if not name.startswith("base_detector.") and ("running_" in name):
            out["buf/" + name] = b.numpy()
    np.savez_compressed(os.path.join(HERE, "stage3_grads_tiny12_192.npz"), names=np.array(names), **out)
    print("stage3 grads golden: loss", float(loss), "params with grad", len(names),
          "rows", int(metric["total"]), int(metric["true"]))


if __name__ == "__main__":
    main()
