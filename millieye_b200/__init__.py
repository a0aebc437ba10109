"""millieye_b200 - B200 (sm_100a) implementation of milliEye's detection-and-fusion hot path.

Public surface mirrors the reference's modules for this path:
    millieye_b200.models      <-> module3_our_dataset/yolov3/models.py   (Darknet, create_modules, YOLOLayer)
    millieye_b200.my_models   <-> module3_our_dataset/my_models.py       (Network, define_yolo, init_yolo)
    millieye_b200.utils       <-> module3_our_dataset/utils/utils.py     (non_max_suppression_cpp, box converts)
    millieye_b200.parse_config<-> module3_our_dataset/utils/parse_config.py
The arithmetic lives in millieye_b200/csrc (CUDA, C-ABI declared in include/millieye_b200.h); importing
this package does not load it - the first forward does, and fails loudly if it has not been built.
"""
__version__ = "0.1.0"

__all__ = ["Darknet", "Network", "define_yolo", "init_yolo", "non_max_suppression_cpp", "parse_model_config"]


def __getattr__(name):
    if name == "Darknet":
        from .models import Darknet
        return Darknet
    if name in ("Network", "define_yolo", "init_yolo"):
        from . import my_models
        return getattr(my_models, name)
    if name == "non_max_suppression_cpp":
        from .utils import non_max_suppression_cpp
        return non_max_suppression_cpp
    if name == "parse_model_config":
        from .parse_config import parse_model_config
        return parse_model_config
    raise AttributeError(name)
