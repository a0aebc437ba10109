"""Drop-in for the reference's yolov3/models.py: `Darknet(config_path)` with the same constructor,
forward signature, return structure and state_dict key names
(module_list.{i}.conv_{i}.weight, module_list.{i}.batch_norm_{i}.*), executed by the sm_100a kernels
of libmillieye_b200 instead of nn.Module calls.

The nn.Conv2d / nn.BatchNorm2d objects below only own parameters (checkpoint compatibility,
`load_state_dict`, `load_darknet_weights`); their forward is never used.  There is no CPU or
eager-PyTorch fallback: a forward on a non-CUDA tensor, or without the built library, raises.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import MeError
from .engine import DarknetPlan, describe_blocks
from .parse_config import parse_model_config


def _lib_bytes(n, g, na):
    from . import _lib
    return int(_lib.lib().me_yolo_loss_workspace(n, g, na))


class _Holder(nn.Module):
    """Parameter-free placeholder keeping the reference's module names (route_, shortcut_, yolo_ ...)."""

    def __init__(self, **attrs):
        super().__init__()
        for k, v in attrs.items():
            setattr(self, k, v)


class YOLOLayer(_Holder):
    """Detection layer placeholder; the decode runs in me_yolo_decode (reference models.py:102-179)."""

    def __init__(self, anchors, num_classes, img_dim=416):
        super().__init__(anchors=anchors, num_anchors=len(anchors), num_classes=num_classes, img_dim=img_dim,
                         metrics={}, ignore_thres=0.5, obj_scale=1, noobj_scale=100, grid_size=0)


def create_modules(module_defs):
    """Parameter containers named exactly like the reference's create_modules (models.py:12-79).
    Consumes module_defs[0] ([net]) like the reference does (pop)."""
    hyperparams = module_defs.pop(0)
    _, blocks = describe_blocks([hyperparams] + module_defs)
    module_list = nn.ModuleList()
    for i, b in enumerate(blocks):
        seq = nn.Sequential()
        kind = b["type"]
        if kind == "convolutional":
            seq.add_module(f"conv_{i}", nn.Conv2d(b["cin"], b["filters"], b["size"], b["stride"], (b["size"] - 1) // 2,
                                                  bias=not b["bn"]))
            if b["bn"]:
                seq.add_module(f"batch_norm_{i}", nn.BatchNorm2d(b["filters"], momentum=0.9, eps=1e-5))
            if b["leaky"]:
                seq.add_module(f"leaky_{i}", nn.LeakyReLU(0.1))
        elif kind == "yolo":
            seq.add_module(f"yolo_{i}", YOLOLayer(b["anchors"], b["classes"], int(hyperparams["height"])))
        else:
            seq.add_module(f"{kind}_{i}", _Holder())
        module_list.append(seq)
    return hyperparams, module_list


class Darknet(nn.Module):
    """YOLOv3 / YOLOv3-tiny detector. forward(x) -> (featuremap, yolo_outputs) like reference models.py:247-267.

    featuremap: (N, C, S/16, S/16) fp32 - output of the block named conv_8 (reference :254-255).  On
      cfgs where block 8 is not a conv (yolov3.cfg: the reference itself raises there)
      set `feature_tap` to a block index, or leave None to get an empty tensor.
    yolo_outputs: (N, sum A*G*G, 5+C) fp32, rows [cx, cy, w, h, conf, cls...] in pixels.
    """

    def __init__(self, config_path, img_size=416):
        super().__init__()
        self.module_defs = parse_model_config(config_path)
        self.hyperparams, self.module_list = create_modules(self.module_defs)
        _, self._blocks = describe_blocks([self.hyperparams] + self.module_defs)
        self.yolo_layers = [seq[0] for seq in self.module_list if isinstance(seq[0], YOLOLayer)]
        self.img_size = img_size
        self.seen = 0
        self.header_info = np.array([0, 0, 0, self.seen, 0], dtype=np.int32)
        self.feature_tap = 8 if len(self._blocks) > 8 and self._blocks[8]["type"] == "convolutional" else None
        self.use_cuda_graph = True
        self._plans = {}
        self._weights_version = 0

    # ------------------------------------------------------------------ engine plumbing
    def _invalidate(self, *_):
        self._plans = {}

    def load_state_dict(self, *args, **kwargs):
        self._invalidate()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def refresh_weights(self):
        """Call after mutating parameters in place (the packed fp16 copies are rebuilt lazily)."""
        self._invalidate()

    def plan_for(self, n, size, device):
        splits = getattr(self, "sub_batches", None)
        if splits is None and os.environ.get("ME_SPLITS"):
            splits = int(os.environ["ME_SPLITS"])
        key = (n, size, device.index, self.feature_tap, splits)
        plan = self._plans.get(key)
        if plan is None:
            if device.type != "cuda":
                raise MeError("Darknet.forward needs CUDA tensors: the sm_100a library is the only implementation")
            tensors = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in self.state_dict().items()
                       if v.is_floating_point()}
            with torch.cuda.device(device):
                plan = DarknetPlan(self._blocks, tensors, n, size, device, self.feature_tap,
                                   in_channels=int(self.hyperparams["channels"]), splits=splits)
            self._plans[key] = plan
        return plan

    def forward_device(self, x, decode=True):
        """Runs the forward and returns the plan (outputs stay in the plan's device buffers).  decode=False stops
        after the head convs; plan.run_decode() finishes the job (DetectPipeline does that on its second stream)."""
        if x.dim() != 4 or x.shape[2] != x.shape[3]:
            raise MeError(f"expected a square (N,C,S,S) batch, got {tuple(x.shape)}")
        dev = x.device if x.is_cuda else next(self.parameters()).device
        plan = self.plan_for(x.shape[0], x.shape[2], dev)
        with torch.cuda.device(dev):
            plan.load_input(x)   # device->device, or pinned host->device on the copy stream
            plan.run(self.use_cuda_graph, decode=decode)
        return plan

    def forward(self, x, targets=None):
        """(featuremap, yolo_outputs), or with targets (m,6) [image, class, cx, cy, w, h] in 0..1 the reference's
        (loss, featuremap, yolo_outputs) (models.py:247-267): loss = sum of the YOLO layers' losses, every layer's
        `metrics` dictionary filled like models.py:212-227.  The loss is a value: the backward pass through the detector
        is not part of the accelerated path (the reference has no stage-1 training script)."""
        plan = self.forward_device(x)
        n = plan.n
        with torch.cuda.device(plan.device):
            if plan.feature_view is not None:
                v = plan.feature_view
                feat = ops.nhwc_to_nchw_f32(v.t, n, v.h, v.w, v.real_c, v.pitch)
            else:
                feat = torch.empty(0, device=plan.device)
            self.featuremap = feat
            if targets is None:
                return feat, plan.yolo_out.clone()
            return self._yolo_loss(plan, targets), feat, plan.yolo_out.clone()

    def _yolo_loss(self, plan, targets):
        heads = plan.head_logits()
        if len(heads) != len(self.yolo_layers):
            raise MeError("the YOLO loss needs the head logits: run without ME_FUSE_DECODE=1")
        dev = plan.device
        t = targets.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = torch.zeros((len(heads), 14), dtype=torch.float32, device=dev)
        for k, (view, blk, g) in enumerate(heads):
            need = _lib_bytes(plan.n, g, len(blk["anchors"]))
            ws = getattr(plan, "_loss_ws", None)
            if ws is None or ws.numel() < need:
                ws = plan._loss_ws = torch.empty((need,), dtype=torch.uint8, device=dev)
            layer = self.yolo_layers[k]
            ops.yolo_loss(view.t, view.pitch, plan.n, g, blk["anchors"], blk["classes"], plan.size / g, t, out[k], ws,
                          layer.ignore_thres, layer.obj_scale, layer.noobj_scale)
        vals = out.cpu()
        for k, layer in enumerate(self.yolo_layers):
            layer.metrics = {key: (int(vals[k, j]) if key == "grid_size" else float(vals[k, j]))
                             for j, key in enumerate(ops.METRIC_KEYS)}
            layer.grid_size = int(vals[k, 13])
        return out[:, 0].sum()

    # ------------------------------------------------------------------ darknet binary weights
    def _conv_bn_pairs(self, cutoff=None):
        for i, (b, seq) in enumerate(zip(self._blocks, self.module_list)):
            if cutoff is not None and i == cutoff:
                return
            if b["type"] == "convolutional":
                yield seq[0], (seq[1] if b["bn"] else None)

    def load_darknet_weights(self, weights_path):
        """Darknet .weights: 5 x int32 header, then per conv [bn bias, bn weight, bn mean, bn var | conv bias],
        conv weight, all float32 (reference models.py:269-326)."""
        with open(weights_path, "rb") as fh:
            header = np.fromfile(fh, dtype=np.int32, count=5)
            flat = np.fromfile(fh, dtype=np.float32)
        self.header_info = header
        self.seen = header[3]
        cutoff = None
        if "darknet53.conv.74" in weights_path:
            cutoff = 75
        if "yolov3-tiny.conv.15" in weights_path:
            cutoff = 15
        pos = 0

        def take(dst):
            nonlocal pos
            cnt = dst.numel()
            dst.data.copy_(torch.from_numpy(flat[pos:pos + cnt]).view_as(dst))
            pos += cnt

        for conv, bn in self._conv_bn_pairs(cutoff):
            if bn is not None:
                for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var):
                    take(t)
            else:
                take(conv.bias)
            take(conv.weight)
        self._invalidate()

    def save_darknet_weights(self, path, cutoff=-1):
        """Inverse of load_darknet_weights (reference models.py:328-352); cutoff=-1 keeps the reference's
        slice semantics (all blocks but the last)."""
        blocks = list(zip(self._blocks, self.module_list))[:cutoff]
        with open(path, "wb") as fh:
            self.header_info[3] = self.seen
            self.header_info.tofile(fh)
            for b, seq in blocks:
                if b["type"] != "convolutional":
                    continue
                conv = seq[0]
                if b["bn"]:
                    bn = seq[1]
                    for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var):
                        t.data.cpu().numpy().tofile(fh)
                else:
                    conv.bias.data.cpu().numpy().tofile(fh)
                conv.weight.data.cpu().numpy().tofile(fh)


class DetectPipeline:
    """Detector forward + non_max_suppression_cpp (utils/utils.py:337-378) as a two-stage device pipeline for
    streams of batches (run_sp.py:213-214 / run_mp.py:314-320 call the two back to back for every frame): the
    confidence filter + NMS, the optional all-gather over ranks and the device->host read of batch i run on a second
    CUDA stream while the main stream already executes the convolutions of batch i+1.  Every submit() returns the
    record of its batch; record.wait() blocks the host until that batch's detections are complete.

    Records live in a ring of `depth` (default 4) per plan: a record's buffers are reused by the depth-th submit after
    it, and that submit first waits (on the host) until the record has completed, so a caller may keep up to `depth`
    batches in flight; a record must be consumed before `depth` further submits are made.

    gather: False | True (one all_gather_into_tensor per batch) | "peer" (dist.PeerGather: the shards are pushed into the
    peers' buffers with copy engines over NVLink, no kernel competes with the convolutions for SMs).  With gather the
    record's det / count are views of the gathered buffer, shaped (world, n_local, max_det, 7 + C) / (world, n_local)
    and read on the pipeline's second stream (record.wait() first when reading from another stream)."""

    class Record:
        def __init__(self, nms):
            self.nms = nms
            self.det, self.count = nms.det, nms.count
            self.host_flat = torch.empty_like(nms.flat, device="cpu").pin_memory()
            body = nms.det.numel()
            self.host_det = self.host_flat[:body].view(nms.det.shape)
            self.host_cnt = self.host_flat[body:].view(torch.int32)
            self.host_all = None          # gathered detections of every rank (gather, readback=True): (det, count) views
            self.host_all_flat = None     # of this pinned (world, rows + counts) buffer
            self.gather_out = None
            self.done = torch.cuda.Event()
            self.nms_done = torch.cuda.Event()   # the detector's output slot is free again (before gather / read-back)
            self.pending = False

        def wait(self):
            self.done.synchronize()
            self.pending = False
            return self

    def __init__(self, net, conf_thresh, nms_thresh=0.5, max_det=200, gather=False, depth=4, host_all=True):
        self.net, self.conf_thresh, self.nms_thresh, self.max_det, self.gather = net, conf_thresh, nms_thresh, max_det, gather
        self.host_all = host_all    # readback=True also copies the gathered batch of every rank to this rank's host
        self.depth = max(2, int(depth))
        self._peer = {}      # plan -> dist.PeerGather (gather="peer")
        self._rings = {}     # plan -> [records], round robin
        self._next = {}
        self._side = None

    def _record(self, plan, dev):
        # plans are rebuilt when weights / devices change: drop the rings of plans the net no longer holds
        live = {id(p) for p in self.net._plans.values()}
        for key in [k for k in self._rings if k not in live]:
            del self._rings[key], self._next[key]
        ring = self._rings.setdefault(id(plan), [])
        k = self._next.get(id(plan), 0)
        self._next[id(plan)] = (k + 1) % self.depth
        if len(ring) <= k:
            ring.append(DetectPipeline.Record(ops.NmsBuffers(plan.n, plan.rows_total, plan.attrs - 5, self.max_det, dev)))
        rec = ring[k]
        if rec.pending:
            rec.done.synchronize()   # its previous occupant is still in flight: wait before its buffers are reused
        rec.pending = True
        return rec

    def _peer_gather(self, plan, numel, dev):
        pg = self._peer.get(id(plan))
        if pg is None or pg.numel != numel:
            from .dist import PeerGather
            pg = self._peer[id(plan)] = PeerGather(numel, self.depth, dev)
        return pg

    def submit(self, x, readback=False):
        plan = self.net.forward_device(x, decode=False)
        dev = plan.device
        with torch.cuda.device(dev):
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            rec = self._record(plan, dev)
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                plan.run_decode(self.net.use_cuda_graph)
                ops.filter_nms(plan.yolo_out, self.conf_thresh, self.nms_thresh, self.max_det, xyxy_inplace=True,
                               buffers=rec.nms)
                rec.det, rec.count = rec.nms.det, rec.nms.count
                rec.nms_done.record()     # later forwards only wait for this: the collective and the host copies read
                if self.gather:           # the record's own buffers and stay off the convolutions' critical path
                    import torch.distributed as tdist
                    world, flat = tdist.get_world_size(), rec.nms.flat
                    if self.gather == "peer":      # copy engines + stream memory operations over peer memory: no kernel
                        g = self._peer_gather(plan, flat.numel(), dev).gather(flat)
                    else:                          # one NCCL / gloo collective into the record's own output buffer
                        if rec.gather_out is None:
                            rec.gather_out = torch.empty((world * flat.numel(),), dtype=torch.float32, device=dev)
                        tdist.all_gather_into_tensor(rec.gather_out, flat)
                        g = rec.gather_out.view(world, flat.numel())
                    body = rec.nms.det.numel()
                    # views of the gathered buffer (rank-major; a rank's rows are followed by its counts): no copy
                    rec.det = g[:, :body].view(world, *rec.nms.det.shape)
                    rec.count = g[:, body:].view(torch.int32)
                if readback:
                    rec.host_flat.copy_(rec.nms.flat, non_blocking=True)      # this rank's shard: one copy
                    if self.gather and self.host_all:
                        if rec.host_all_flat is None:
                            rec.host_all_flat = torch.empty(tuple(g.shape), dtype=torch.float32).pin_memory()
                            rec.host_all = (rec.host_all_flat[:, :body].view(world, *rec.nms.det.shape),
                                            rec.host_all_flat[:, body:].view(torch.int32))
                        rec.host_all_flat.copy_(g, non_blocking=True)          # and the whole batch: one copy
                rec.done.record()
            plan.hold_output(rec.nms_done)
        return rec
