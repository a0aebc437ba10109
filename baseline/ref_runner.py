"""Runs the UNMODIFIED reference (baseline/_ref, staged by baseline/stage_reference.py) for bench.py's reference arm and
CPU baseline.  Import shims only (matplotlib is imported at module scope by utils/utils.py:12-13 and is
not installed) - no reference code is changed, wrapped or replaced."""
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
M3 = os.path.join(REF, "module3_our_dataset")


def available():
    return os.path.exists(os.path.join(M3, "yolov3", "models.py"))


_LOADED = None


def load():
    """-> (Darknet class, utils.utils module) of the reference."""
    global _LOADED
    if _LOADED is None:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
            mod = types.ModuleType(name)
            mod.close = lambda *a, **k: None
            sys.modules.setdefault(name, mod)
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
        sys.path.insert(0, M3)
        cwd = os.getcwd()
        os.chdir(tempfile.mkdtemp())   # the reference writes ./b.txt from its training branch (my_models.py:350)
        try:
            from yolov3.models import Darknet
            import utils.utils as ref_utils
        finally:
            os.chdir(cwd)
        _LOADED = (Darknet, ref_utils)
    return _LOADED


def darknet53_step_fn(state_dict, images, conf_thresh, threads):
    """One step of the bench workload on the reference: Darknet(yolov3.cfg).forward + non_max_suppression_cpp
    (yolov3/models.py:247-267, utils/utils.py:337-378), fp32, torch CPU.  `featuremap` is pre-seeded because the
    reference's forward raises on yolov3.cfg otherwise (no module is named conv_8)."""
    import torch
    Darknet, ref_utils = load()
    torch.set_num_threads(threads)
    net = Darknet(os.path.join(M3, "config", "yolov3.cfg")).eval()
    net.load_state_dict(state_dict)
    net.featuremap = torch.empty(0)

    def step():
        with torch.no_grad():
            _, y = net(images)
            return ref_utils.non_max_suppression_cpp(y, conf_thresh)
    return step
