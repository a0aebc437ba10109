"""Generators for the darknet .cfg files the reference ships under module3_our_dataset/config/
(yolov3-tiny-12.cfg, yolov3-tiny-coco.cfg, yolov3.cfg).  The networks are described here as compact
layer programs and written out as cfg text; tests check that parsing the generated text gives the
same block list as parsing the reference's files, so `Darknet(cfg_path(...))` builds the same model.
"""
import os

_NET = dict(batch=1, subdivisions=1, width=416, height=416, channels=3, momentum=0.9, decay=0.0005, angle=0,
            saturation=1.5, exposure=1.5, hue=.1, learning_rate=0.001, burn_in=1000, max_batches=500200,
            policy="steps", steps="400000,450000", scales=".1,.1")

_TINY_ANCHORS = "10,14,  23,27,  37,58,  81,82,  135,169,  344,319"
_FULL_ANCHORS = "10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326"


def _conv(filters, size, stride=1, bn=True, act="leaky"):
    b = [("type", "convolutional")]
    if bn:
        b.append(("batch_normalize", 1))
    b += [("filters", filters), ("size", size), ("stride", stride), ("pad", 1), ("activation", act)]
    return b


def _yolo(mask, anchors, classes, num):
    return [("type", "yolo"), ("mask", mask), ("anchors", anchors), ("classes", classes), ("num", num),
            ("jitter", .3), ("ignore_thresh", .7), ("truth_thresh", 1), ("random", 1)]


def tiny_blocks(classes=12):
    """YOLOv3-tiny: 24 blocks (13 conv, 6 maxpool, 2 route, 1 upsample, 2 yolo)."""
    head = 3 * (classes + 5)
    b = []
    for f in (16, 32, 64, 128, 256):
        b.append(_conv(f, 3))
        b.append([("type", "maxpool"), ("size", 2), ("stride", 2)])
    b.append(_conv(512, 3))
    b.append([("type", "maxpool"), ("size", 2), ("stride", 1)])
    b += [_conv(1024, 3), _conv(256, 1), _conv(512, 3), _conv(head, 1, bn=False, act="linear"),
          _yolo("3,4,5", _TINY_ANCHORS, classes, 6),
          [("type", "route"), ("layers", "-4")], _conv(128, 1), [("type", "upsample"), ("stride", 2)],
          [("type", "route"), ("layers", "-1, 8")], _conv(256, 3), _conv(head, 1, bn=False, act="linear"),
          _yolo("1,2,3", _TINY_ANCHORS, classes, 6)]
    return b


def full_blocks(classes=80):
    """YOLOv3 / Darknet-53: 107 blocks (75 conv, 23 shortcut, 4 route, 2 upsample, 3 yolo)."""
    head = 3 * (classes + 5)
    b = [_conv(32, 3)]

    def stage(filters, repeats):
        b.append(_conv(filters, 3, stride=2))
        for _ in range(repeats):
            b.append(_conv(filters // 2, 1))
            b.append(_conv(filters, 3))
            b.append([("type", "shortcut"), ("from", -3), ("activation", "linear")])

    for f, r in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        stage(f, r)

    def neck(filters):
        for _ in range(3):
            b.append(_conv(filters, 1))
            b.append(_conv(filters * 2, 3))
        b.append(_conv(head, 1, bn=False, act="linear"))

    neck(512)
    b.append(_yolo("6,7,8", _FULL_ANCHORS, classes, 9))
    b += [[("type", "route"), ("layers", "-4")], _conv(256, 1), [("type", "upsample"), ("stride", 2)],
          [("type", "route"), ("layers", "-1, 61")]]
    neck(256)
    b.append(_yolo("3,4,5", _FULL_ANCHORS, classes, 9))
    b += [[("type", "route"), ("layers", "-4")], _conv(128, 1), [("type", "upsample"), ("stride", 2)],
          [("type", "route"), ("layers", "-1, 36")]]
    neck(128)
    b.append(_yolo("0,1,2", _FULL_ANCHORS, classes, 9))
    return b


def cfg_text(blocks, net=None):
    lines = ["[net]"] + [f"{k}={v}" for k, v in (net or _NET).items()] + [""]
    for i, blk in enumerate(blocks):
        lines.append(f"# {i}")
        lines.append(f"[{blk[0][1]}]")
        lines += [f"{k}={v}" for k, v in blk[1:]]
        lines.append("")
    return "\n".join(lines)


_BUILDERS = {
    "yolov3-tiny-12": lambda: tiny_blocks(12),
    "yolov3-tiny-coco": lambda: tiny_blocks(80),
    "yolov3": lambda: full_blocks(80),
}


def cfg_path(name):
    """Writes (once) and returns millieye_b200/config/<name>.cfg."""
    if name.endswith(".cfg"):
        name = name[:-4]
    if name not in _BUILDERS:
        raise KeyError(f"unknown cfg '{name}' (have {sorted(_BUILDERS)})")
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name + ".cfg")
    text = cfg_text(_BUILDERS[name]())
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as fh:
            fh.write(text)
    return path
