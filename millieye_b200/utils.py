"""Drop-ins for the helpers of the reference's utils/utils.py that sit on the hot path.

`non_max_suppression_cpp` keeps the reference's interface (utils.py:337-378: prediction tensor in,
list of per-image (k, 7+C) tensors or None out, prediction[..., :4] rewritten in place to x1y1x2y2)
but runs me_filter_nms on the device instead of a Python loop over torchvision.batched_nms on CPU.
"""
import torch

from . import ops
from ._lib import MeError


def to_cpu(tensor):
    return tensor.detach().cpu()


def weights_init_normal(m):
    """Same initialisation rule as the reference (utils.py:29-37)."""
    name = m.__class__.__name__
    if "Conv" in name:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif "BatchNorm2d" in name:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)
    elif "Linear" in name:
        torch.nn.init.kaiming_normal_(m.weight.data)


def xyxy2xywh(x):
    """[x1, y1, x2, y2] -> [cx, cy, w, h] (utils.py:58-65)."""
    y = torch.zeros_like(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def xywh2xyxy(x):
    """[cx, cy, w, h] -> [x1, y1, x2, y2] (utils.py:68-74)."""
    y = x.new(x.shape)
    y[..., 0] = x[..., 0] - x[..., 2] / 2
    y[..., 1] = x[..., 1] - x[..., 3] / 2
    y[..., 2] = x[..., 0] + x[..., 2] / 2
    y[..., 3] = x[..., 1] + x[..., 3] / 2
    return y


def non_max_suppression_cpp(prediction, conf_thresh, nms_thresh=0.5, detections_per_img=200):
    """prediction: (N, B, 5+C) fp32 CUDA tensor [cx,cy,w,h,conf,cls...]; returns a list with, per image,
    a (k, 7+C) tensor [x1,y1,x2,y2,conf,class_conf,class_pred,cls...] (k <= detections_per_img) or None."""
    if not prediction.is_cuda:
        raise MeError("non_max_suppression_cpp runs on the GPU; pass a CUDA tensor (no CPU fallback)")
    if prediction.dtype != torch.float32 or not prediction.is_contiguous():
        raise MeError("prediction must be a contiguous float32 tensor")
    with torch.cuda.device(prediction.device):
        buf = ops.filter_nms(prediction, conf_thresh, nms_thresh, detections_per_img, xyxy_inplace=True)
        counts = buf.count.cpu().tolist()   # the one host sync: result sizes are data dependent
    return [buf.det[i, :k].clone() if k > 0 else None for i, k in enumerate(counts)]
