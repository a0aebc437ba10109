"""Deterministic synthetic weights and inputs shared by the golden generator and the tests.

Test infrastructure only (see oracle/__init__.py).  Values come from numpy's RandomState (stable
across numpy versions and machines), filled in state_dict key order, so the reference model built
in the build container and the model under test on the GPU box get identical parameters without
shipping a checkpoint.  Scales are chosen so activations stay O(1) through 75 conv layers (the
reference's weights_init_normal, utils.py:29-37, makes them vanish) and so YOLO objectness has a
realistic spread instead of sitting at 0.5.
"""
import numpy as np
import torch


def fill_state_dict(sd, seed=0, obj_bias=-3.0, head_gain=3.0, conv_gain=1.0):
    """Overwrites every tensor of `sd` (an nn.Module.state_dict()) in place, returns sd.

    conv_gain scales the He-initialised weights of every conv that is followed by batch norm; the
    residual stages of Darknet-53 double the variance per shortcut, 0.6 keeps them O(1)."""
    rng = np.random.RandomState(seed)
    keys = list(sd.keys())
    for k in keys:
        t = sd[k]
        shape = tuple(t.shape)
        if k.endswith("num_batches_tracked"):
            t.zero_()
            continue
        is_bn = ("batch_norm" in k) or _is_bn_key(k, sd)
        if k.endswith("running_mean"):
            v = rng.normal(0.0, 0.1, shape)
        elif k.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif is_bn and k.endswith("weight"):
            v = rng.uniform(0.8, 1.2, shape)
        elif is_bn and k.endswith("bias"):
            v = rng.normal(0.0, 0.1, shape)
        elif k.endswith("weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            v = rng.normal(0.0, np.sqrt(2.0 / (1.01 * fan_in)), shape)
            if _has_bn_sibling(k, sd):
                v = v * conv_gain
        elif k.endswith("weight") and len(shape) == 2:
            v = rng.normal(0.0, np.sqrt(1.0 / shape[1]), shape)
        elif k.endswith("bias"):
            v = rng.normal(0.0, 0.1, shape)
        else:
            v = rng.normal(0.0, 1.0, shape)
        t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
    # YOLO head convs (no batch norm, bias): widen logits and push objectness down so that the
    # confidence filter and NMS see a realistic candidate count.
    for k in keys:
        if k.endswith(".bias") and "conv_" in k and "batch_norm" not in k and "module_list" in k:
            wkey = k[:-4] + "weight"
            cout = sd[k].shape[0]
            if wkey in sd and sd[wkey].dim() == 4 and sd[wkey].shape[2] == 1:
                bn_key = k.replace("conv_", "batch_norm_")
                if bn_key in sd:
                    continue
                sd[wkey].mul_(head_gain)
                for attrs in (85, 17):
                    if cout % attrs == 0:
                        sd[k][4::attrs] += obj_bias
                        break
    return sd


def _has_bn_sibling(k, sd):
    """True for module_list.{i}.conv_{i}.weight when module_list.{i}.batch_norm_{i}.* exists."""
    return "conv_" in k and k.replace("conv_", "batch_norm_") in sd


def _is_bn_key(k, sd):
    base = k.rsplit(".", 1)[0]
    return (base + ".running_mean") in sd


def synth_images(n, size, seed=0):
    rng = np.random.RandomState(1000 + seed)
    return torch.from_numpy(rng.rand(n, 3, size, size).astype(np.float32))


def synth_maps(n, size, seed=0):
    rng = np.random.RandomState(2000 + seed)
    return torch.from_numpy(rng.rand(n, 3, size // 16, size // 16).astype(np.float32))


def synth_radar_boxes(n, seed=0, max_per_frame=3):
    """(m,5) [frame, x1, y1, x2, y2] in 0..1, 0..max_per_frame boxes per frame."""
    rng = np.random.RandomState(3000 + seed)
    rows = []
    for i in range(n):
        for _ in range(rng.randint(0, max_per_frame + 1)):
            x1, y1 = rng.uniform(0, 0.7, 2)
            w, h = rng.uniform(0.05, 0.3, 2)
            rows.append([i, x1, y1, min(x1 + w, 1.0), min(y1 + h, 1.0)])
    if not rows:
        return torch.zeros((0, 5), dtype=torch.float32)
    return torch.tensor(rows, dtype=torch.float32)


def synth_predictions(n, rows, num_classes, seed=0, size=416.0, conf_mu=-6.0, conf_sigma=2.0):
    """Decoded-YOLO-like tensor (n, rows, 5+C) with conf = sigmoid(N(mu, sigma))."""
    rng = np.random.RandomState(4000 + seed)
    p = np.empty((n, rows, 5 + num_classes), dtype=np.float32)
    p[..., 0:2] = rng.uniform(0, size, (n, rows, 2))
    p[..., 2:4] = np.exp(rng.normal(3.5, 0.8, (n, rows, 2)))
    p[..., 4] = 1.0 / (1.0 + np.exp(-rng.normal(conf_mu, conf_sigma, (n, rows))))
    p[..., 5:] = 1.0 / (1.0 + np.exp(-rng.normal(-1.0, 2.0, (n, rows, num_classes))))
    return torch.from_numpy(p)
