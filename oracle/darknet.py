"""Oracle restatement of the Darknet/YOLOv3 forward (reference yolov3/models.py).

Test infrastructure only (see oracle/__init__.py).  fp32 on CPU; conv / batch-norm / pooling go
through the same torch operators the reference's nn.Modules dispatch to, everything else
(layer wiring, padding rules, decode arithmetic, output order) is restated here.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .parse_config import parse_model_config


def layer_plan(module_defs):
    """Static description of every block (models.py:12-79): returns (hyperparams, blocks) where
    blocks[i] has type, out channels and the keys the forward needs."""
    hyper = module_defs[0]
    out_filters = [int(hyper["channels"])]
    blocks = []
    for i, d in enumerate(module_defs[1:]):
        b = {"type": d["type"], "index": i}
        if d["type"] == "convolutional":
            b["bn"] = int(d["batch_normalize"])
            b["filters"] = filters = int(d["filters"])
            b["size"] = int(d["size"])
            b["stride"] = int(d["stride"])
            b["pad"] = (b["size"] - 1) // 2          # models.py:25 (cfg 'pad' key is ignored)
            b["leaky"] = d["activation"] == "leaky"  # models.py:40
            b["cin"] = out_filters[-1]
        elif d["type"] == "maxpool":
            b["size"] = int(d["size"])
            b["stride"] = int(d["stride"])
            filters = out_filters[-1]
        elif d["type"] == "upsample":
            b["stride"] = int(d["stride"])
            filters = out_filters[-1]
        elif d["type"] == "route":
            b["layers"] = [int(v) for v in d["layers"].split(",")]
            filters = sum(out_filters[1:][j] for j in b["layers"])
        elif d["type"] == "shortcut":
            b["from"] = int(d["from"])
            filters = out_filters[1:][b["from"]]
        elif d["type"] == "yolo":
            idxs = [int(v) for v in d["mask"].split(",")]
            flat = [int(v) for v in d["anchors"].split(",")]
            pairs = [(flat[j], flat[j + 1]) for j in range(0, len(flat), 2)]
            b["anchors"] = [pairs[j] for j in idxs]
            b["classes"] = int(d["classes"])
            filters = out_filters[-1]
        else:
            raise ValueError(f"unknown block type {d['type']}")
        b["out_channels"] = filters
        blocks.append(b)
        out_filters.append(filters)
    return hyper, blocks


def yolo_decode(x, anchors, num_classes, img_dim):
    """YOLOLayer.forward, inference branch (models.py:132-179). x: (N, A*(5+C), G, G)."""
    n, _, g, _ = x.shape
    a = len(anchors)
    pred = x.view(n, a, num_classes + 5, g, g).permute(0, 1, 3, 4, 2).contiguous()
    stride = img_dim / g                                                  # :124 python float
    grid_x = torch.arange(g).repeat(g, 1).view(1, 1, g, g).float()        # :126
    grid_y = torch.arange(g).repeat(g, 1).t().view(1, 1, g, g).float()    # :127
    scaled = torch.tensor([(aw / stride, ah / stride) for aw, ah in anchors], dtype=torch.float32)
    aw = scaled[:, 0:1].view(1, a, 1, 1)
    ah = scaled[:, 1:2].view(1, a, 1, 1)
    boxes = torch.empty(pred[..., :4].shape, dtype=torch.float32)
    boxes[..., 0] = torch.sigmoid(pred[..., 0]) + grid_x
    boxes[..., 1] = torch.sigmoid(pred[..., 1]) + grid_y
    boxes[..., 2] = torch.exp(pred[..., 2]) * aw
    boxes[..., 3] = torch.exp(pred[..., 3]) * ah
    conf = torch.sigmoid(pred[..., 4])
    cls = torch.sigmoid(pred[..., 5:])
    return torch.cat((boxes.view(n, -1, 4) * stride, conf.view(n, -1, 1), cls.view(n, -1, num_classes)), -1)


def conv_block(x, b, sd, prefix, i):
    """conv -> eval-mode BN -> LeakyReLU(0.1) (models.py:22-41)."""
    w = sd[f"{prefix}module_list.{i}.conv_{i}.weight"]
    bias = sd.get(f"{prefix}module_list.{i}.conv_{i}.bias")
    x = F.conv2d(x, w, bias, stride=b["stride"], padding=b["pad"])
    if b["bn"]:
        p = f"{prefix}module_list.{i}.batch_norm_{i}."
        x = F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                         training=False, momentum=0.9, eps=1e-5)
    if b["leaky"]:
        x = F.leaky_relu(x, 0.1)
    return x


def darknet_forward(module_defs, sd, x, prefix="", feature_tap=None, collect=False):
    """Darknet.forward without targets (models.py:247-267).

    sd: state_dict-like mapping (fp32 CPU tensors) with the reference's key names.
    feature_tap: index of the block whose output is the RoI feature map.  The reference taps the
      module *named* conv_8 (:254-255), i.e. block 8 when it is convolutional; on yolov3.cfg
      block 8 is a shortcut and the reference raises - pass an explicit tap there.
    Returns (featuremap or None, yolo_outputs) and, with collect=True, every block output too.
    """
    _, blocks = layer_plan(module_defs)
    img_dim = x.shape[2]
    if feature_tap is None:
        feature_tap = 8 if len(blocks) > 8 and blocks[8]["type"] == "convolutional" else None
    outs, yolo_outs, feat = [], [], None
    for i, b in enumerate(blocks):
        t = b["type"]
        if t == "convolutional":
            x = conv_block(x, b, sd, prefix, i)
        elif t == "maxpool":
            if b["size"] == 2 and b["stride"] == 1:
                x = F.pad(x, (0, 1, 0, 1))                                # ZeroPad2d((0,1,0,1)) :47
            x = F.max_pool2d(x, b["size"], b["stride"], padding=(b["size"] - 1) // 2)
        elif t == "upsample":
            x = F.interpolate(x, scale_factor=b["stride"], mode="nearest")
        elif t == "route":
            x = torch.cat([outs[j] for j in b["layers"]], 1)
        elif t == "shortcut":
            x = outs[-1] + outs[b["from"]]
        elif t == "yolo":
            x = yolo_decode(x, b["anchors"], b["classes"], img_dim)
            yolo_outs.append(x)
        if i == feature_tap:
            feat = x
        outs.append(x)
    y = torch.cat(yolo_outs, 1)
    return (feat, y, outs) if collect else (feat, y)


def conv_flops(module_defs, img_size):
    """2*MAC count of all conv blocks for one frame (65.864 G for yolov3.cfg@416)."""
    _, blocks = layer_plan(module_defs)
    sizes, total = [], 0
    s = img_size
    for b in blocks:
        t = b["type"]
        if t == "convolutional":
            s_in = s
            s = (s_in + 2 * b["pad"] - b["size"]) // b["stride"] + 1
            total += 2 * s * s * b["filters"] * b["cin"] * b["size"] ** 2
        elif t == "maxpool":
            s = s // 2 if b["stride"] == 2 else s
        elif t == "upsample":
            s = s * b["stride"]
        elif t == "route":
            s = sizes[b["layers"][0]]
        elif t == "shortcut":
            s = sizes[-1]
        sizes.append(s)
    return total


__all__ = ["parse_model_config", "layer_plan", "yolo_decode", "darknet_forward", "conv_flops", "np"]
