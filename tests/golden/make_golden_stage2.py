"""Generates tests/golden/stage2_tiny12_160.npz from the UNMODIFIED stage-2 reference
(/root/reference/module2_mixed/my_models.py, Network.forward inference branch).  Separate from
make_golden.py because module2_mixed and module3_our_dataset both define top-level `my_models` / `utils`.

    python tests/golden/make_golden_stage2.py        (build container only)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
M2 = os.path.join(os.environ.get("MILLIEYE_REFERENCE", "/root/reference"), "module2_mixed")
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        mod = types.ModuleType(name)
        mod.close = lambda *a, **k: None
        sys.modules.setdefault(name, mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.path.insert(0, M2)
    os.chdir(tempfile.mkdtemp())
    import my_models
    cfg = os.path.join(M2, "config", "yolov3-tiny-12.cfg")
    with torch.no_grad():
        model = my_models.Network(my_models.define_yolo(cfg), conf_thresh=0.3).eval()
        model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=6, obj_bias=-1.0, head_gain=1.0))
        out = model(synth.synth_images(2, 160, seed=6))
    np.savez_compressed(os.path.join(HERE, "stage2_tiny12_160.npz"), out=out.numpy(),
                        keys=np.array(list(model.state_dict().keys())))
    print("stage2 golden", tuple(out.shape))


if __name__ == "__main__":
    main()
