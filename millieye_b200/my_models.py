"""Drop-in for the reference's module3_our_dataset/my_models.py: `Network(base_detector, conf_thresh)`
with the same constructor, attributes, forward signature and state_dict keys, running the whole
detection-and-fusion forward (my_models.py:433-539) on the device through libmillieye_b200:

  Darknet (graph replay) -> me_filter_nms -> me_build_proposals -> conv GEMMs for the image / radar
  score maps -> me_psroi_align / me_roi_align -> GEMM 490->256 -> me_fusion_heads -> me_finalize_output

The reference's device boundary (D2H of every decoded box for CPU NMS, H2D of the survivors,
my_models.py:457-470) disappears; the only host synchronisation is the read of the final row count.
The nn modules below are parameter containers (checkpoint compatibility); their forward is unused.
"""
import os

import torch
from torch import nn

from . import ops
from ._lib import ME_ACT_LEAKY, ME_ACT_SIGMOID, MeError
from .models import Darknet


def define_yolo(model_def):
    """reference my_models.py:13-24"""
    return Darknet(model_def)


def init_yolo(model, weights_path):
    """reference my_models.py:27-44: darknet .weights, ultralytics .pt (positional remap) or a state_dict."""
    if weights_path.endswith(".weights"):
        model.load_darknet_weights(weights_path)
    elif weights_path.endswith(".pt"):
        param = torch.load(weights_path)["model"]
        names = list(param)
        own = model.state_dict()
        for i, name in enumerate(own):
            own[name] = param[names[i]]
        model.load_state_dict(own)
    else:
        model.load_state_dict(torch.load(weights_path))


class cnn_layers_1(nn.Module):
    """1x1 conv + BN + LeakyReLU stack on the feature map (reference :47-77). Parameter container."""

    def __init__(self, channels):
        super().__init__()
        self.net = nn.Sequential()
        for i in range(len(channels) - 1):
            self.net.add_module(f"conv_{i}", nn.Conv2d(channels[i], channels[i + 1], kernel_size=(1, 1), stride=(1, 1)))
            self.net.add_module(f"batch_norm_{i}", nn.BatchNorm2d(channels[i + 1], momentum=0.1))
            self.net.add_module(f"leaky_{i}", nn.LeakyReLU(0.1))


class cnn_layers_3(nn.Module):
    """Radar heat-map CNN 3->32->64->128->10 + sigmoid (reference :130-157). Parameter container."""

    def __init__(self):
        super().__init__()

        def block(cin, cout, *tail):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1),
                                 nn.BatchNorm2d(cout, momentum=0.1), nn.LeakyReLU(0.1), *tail)

        self.conv1 = block(3, 32)
        self.conv2 = block(32, 64)
        self.conv3 = block(64, 128, nn.Conv2d(128, 10, kernel_size=1, stride=1))


class ensemble_head(nn.Module):
    """reference :176-210. Parameter container."""

    def __init__(self, channels, activation_softmax=True):
        super().__init__()
        self.activation_softmax = activation_softmax
        self.fc1 = nn.Sequential(nn.Linear(channels[0], channels[1]), nn.LeakyReLU(0.1))
        self.fc2 = nn.Sequential(nn.Linear(channels[2], channels[3]))
        self.softmax = nn.Softmax(dim=1)


class refinement_head(nn.Module):
    """reference :213-284. Parameter container (net3 / fusion_head exist in checkpoints but are unused, F7)."""

    def __init__(self, channels):
        super().__init__()
        self.count = 0
        tmp = 49
        self.net0 = nn.Sequential(nn.Linear(channels[0], channels[1]), nn.LeakyReLU(0.1))
        self.net1 = nn.Sequential(nn.Linear(channels[1], 4))
        self.net2 = nn.Sequential(nn.Linear(channels[1], 13), nn.Sigmoid())
        self.net3 = nn.Sequential(nn.Linear(channels[1], tmp), nn.Sigmoid())
        self.radar_net = nn.Sequential(nn.Conv2d(10, 10, kernel_size=7, stride=1, padding=0),
                                       nn.BatchNorm2d(10, momentum=0.1), nn.LeakyReLU(0.1),
                                       nn.Conv2d(10, 1, kernel_size=1, stride=1, padding=0), nn.Sigmoid())
        self.fusion_head = nn.Sequential(nn.Linear(2 * tmp, 1), nn.Sigmoid())


class _FusionPlan:
    """Device buffers + packed weights of the fusion tail for one (n, size) shape."""

    MAX_DET = 200  # detections_per_img of non_max_suppression_cpp (utils.py:337)

    def __init__(self, net, base_plan, n, size, device, radar_cap):
        self.n, self.size, self.device = n, size, device
        self.base = base_plan
        fv = base_plan.feature_view
        if fv is None:
            raise MeError("the base detector has no feature tap (set base_detector.feature_tap to a stride-16 "
                          "block with 256 channels; the reference only defines it for the tiny cfgs, SURVEY.md F1)")
        if fv.real_c != 256:
            raise MeError(f"feature map has {fv.real_c} channels, img_cnn_layers expects 256")
        self.g = fv.h
        g = self.g
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in net.state_dict().items()
              if v.is_floating_point() and not k.startswith("base_detector.")}

        def bn(prefix):
            return (sd[prefix + "weight"], sd[prefix + "bias"], sd[prefix + "running_mean"], sd[prefix + "running_var"], 1e-5)

        f16 = dict(dtype=torch.float16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        # score maps.  The 490 position-sensitive channels are produced in bin-major order ((ph, pw, c) instead of the
        # reference's (c, ph, pw)), the two RoI crops are written in that order too, and the layers that read the crops
        # (refinement_head.net0, radar_net) get their input columns permuted the same way: same numbers, but the gathers
        # read 10 contiguous channels per bin (ops.roi_gather_bin_major).
        perm = ops.bin_major_perm(10, 7, device)
        g_, b_, m_, v_, eps = bn("img_cnn_layers.net.batch_norm_0.")
        self.img_conv = ops.pack_conv(sd["img_cnn_layers.net.conv_0.weight"][perm].contiguous(),
                                      sd["img_cnn_layers.net.conv_0.bias"][perm].contiguous(),
                                      (g_[perm].contiguous(), b_[perm].contiguous(), m_[perm].contiguous(), v_[perm].contiguous(), eps),
                                      cout_pad=512)
        self.roi_score = torch.zeros((n, g, g, 512), **f16)
        self.maps_in = torch.zeros((n, 3, g, g), **f32)
        self.r1 = ops.pack_first_conv(sd["radar_cnn_layers.conv1.0.weight"], sd["radar_cnn_layers.conv1.0.bias"],
                                      bn("radar_cnn_layers.conv1.1."))
        self.r2 = ops.pack_conv(sd["radar_cnn_layers.conv2.0.weight"], sd["radar_cnn_layers.conv2.0.bias"],
                                bn("radar_cnn_layers.conv2.1."))
        self.r3 = ops.pack_conv(sd["radar_cnn_layers.conv3.0.weight"], sd["radar_cnn_layers.conv3.0.bias"],
                                bn("radar_cnn_layers.conv3.1."))
        self.r4 = ops.pack_conv(sd["radar_cnn_layers.conv3.3.weight"], sd["radar_cnn_layers.conv3.3.bias"], None,
                                cout_pad=32)
        self.ra = torch.zeros((n, g, g, 32), **f16)
        self.rb = torch.zeros((n, g, g, 64), **f16)
        self.rc = torch.zeros((n, g, g, 128), **f16)
        self.radar_score = torch.zeros((n, g, g, 32), **f16)
        # proposals
        self.radar_cap = radar_cap
        self.cap = n * self.MAX_DET + radar_cap
        cap = self.cap
        self.nms = ops.NmsBuffers(n, base_plan.rows_total, base_plan.attrs - 5, self.MAX_DET, device)
        self.radar_dev = torch.zeros((max(radar_cap, 1), 5), **f32)
        self.radar_n_dev = torch.zeros((1,), **i32)      # radar boxes of the current batch (FusionPipeline's graphs)
        self.img_boxes = torch.zeros((cap, 9), **f32)
        self.rois = torch.zeros((cap, 5), **f32)
        self.counts = torch.zeros((2,), **i32)
        self.crop_img = torch.zeros((cap, 512), **f16)
        self.crop_radar = torch.zeros((cap, 496), **f16)
        self.hidden = torch.zeros((cap, 256), **f16)
        w0 = sd["refinement_head.net0.0.weight"][:, perm].contiguous()
        self.fc0 = ops.pack_conv(w0.view(w0.shape[0], w0.shape[1], 1, 1), sd["refinement_head.net0.0.bias"], None)
        # small heads stay fp32
        rw = sd["refinement_head.radar_net.0.weight"]
        g_, b_, m_, v_, eps = bn("refinement_head.radar_net.1.")
        scale = g_ / torch.sqrt(v_ + eps)
        self._hw_tensors = {
            "net1_w": sd["refinement_head.net1.0.weight"].contiguous(), "net1_b": sd["refinement_head.net1.0.bias"].contiguous(),
            "net2_w": sd["refinement_head.net2.0.weight"].contiguous(), "net2_b": sd["refinement_head.net2.0.bias"].contiguous(),
            "radar_w": (rw.reshape(rw.shape[0], -1) * scale[:, None])[:, perm].contiguous(),
            "radar_b": ((sd["refinement_head.radar_net.0.bias"] - m_) * scale + b_).contiguous(),
            "radar2_w": sd["refinement_head.radar_net.3.weight"].reshape(-1).contiguous(),
            "radar2_b": sd["refinement_head.radar_net.3.bias"].contiguous(),
            "fc1_w": sd["ensemble_head.fc1.0.weight"].contiguous(), "fc1_b": sd["ensemble_head.fc1.0.bias"].contiguous(),
            "fc2_w": sd["ensemble_head.fc2.0.weight"].contiguous(), "fc2_b": sd["ensemble_head.fc2.0.bias"].contiguous(),
        }
        self.head_weights = ops.make_head_weights(self._hw_tensors)
        self.regress = torch.zeros((cap, 4), **f32)
        self.refine = torch.zeros((cap, 2), **f32)
        self.mask = torch.zeros((cap,), **f32)
        # rows and their count in ONE buffer: a host that wants both reads them with a single copy (FusionPipeline)
        self.out_flat = torch.zeros((cap * 8 + 4,), **f32)
        self.out = self.out_flat[:cap * 8].view(cap, 8)
        self.out_count = self.out_flat[cap * 8:cap * 8 + 1].view(torch.int32)
        ws_bytes = 1
        while ws_bytes < cap:
            ws_bytes <<= 1
        self.final_ws = torch.zeros((ws_bytes * 8,), dtype=torch.uint8, device=device)
        self.launches = 0

    def ensure_loss_buffers(self, num_targets):
        """Buffers of the labelling + loss branch (allocated on first use; inference never pays for them)."""
        f32 = dict(dtype=torch.float32, device=self.device)
        if getattr(self, "iou_labels", None) is None:
            self.iou_labels = torch.zeros((self.cap,), **f32)
            self.target_location = torch.zeros((self.cap, 4), **f32)
            self.sample_filter = torch.zeros((self.cap,), dtype=torch.uint8, device=self.device)
            self.loss_out = torch.zeros((16,), **f32)
            self.targets_dev = torch.zeros((64, 6), **f32)
        if self.targets_dev.shape[0] < num_targets:
            self.targets_dev = torch.zeros((1 << (num_targets - 1).bit_length(), 6), **f32)

    def score_maps(self):
        """img_cnn_layers + radar_cnn_layers (my_models.py:486-487)."""
        n, g = self.n, self.g
        fv = self.base.feature_view
        ops.conv_gemm(fv.t, self.img_conv, n, g, g, fv.pitch, self.roi_score, 512, act=ME_ACT_LEAKY, cin=256)
        ops.conv_first(self.maps_in, self.r1, self.ra, 32, act=ME_ACT_LEAKY)
        ops.conv_gemm(self.ra, self.r2, n, g, g, 32, self.rb, 64, act=ME_ACT_LEAKY)
        ops.conv_gemm(self.rb, self.r3, n, g, g, 64, self.rc, 128, act=ME_ACT_LEAKY)
        ops.conv_gemm(self.rc, self.r4, n, g, g, 128, self.radar_score, 32, act=ME_ACT_SIGMOID)
        self.launches += 5

    def proposals(self, conf_thresh, class_idx, num_radar):
        ops.filter_nms(self.base.yolo_out, conf_thresh, 0.5, self.MAX_DET, xyxy_inplace=True, buffers=self.nms)
        ops.build_proposals(self.nms.det, self.nms.count, class_idx, self.radar_dev[:num_radar] if num_radar else None,
                            1.0, self.img_boxes, self.rois, self.counts, self.cap)
        self.launches += 3

    def proposals_dev(self, conf_thresh, class_idx):
        """proposals() with the radar-box count taken from self.radar_n_dev: no per-batch host scalar in any launch."""
        ops.filter_nms(self.base.yolo_out, conf_thresh, 0.5, self.MAX_DET, xyxy_inplace=True, buffers=self.nms)
        ops.build_proposals_dev(self.nms.det, self.nms.count, class_idx, self.radar_dev, self.radar_n_dev, 1.0,
                                self.img_boxes, self.rois, self.counts, self.cap)
        self.launches += 3

    def run_graphed(self, name, key, fn, enabled=True):
        """fn() enqueues kernels of this plan on the current stream.  First use of (name, key): run it as is (one-time
        initialisation inside the library must not happen under capture); second use: capture it into a CUDA graph;
        afterwards: replay.  Every shape is fixed by the plan and row counts live on the device, so a graph serves
        every batch; `key` holds the host scalars baked into the launches (thresholds, detector output slot)."""
        if not enabled:
            return fn()
        cache = self.__dict__.setdefault("_graphs", {})
        state = cache.get((name, key))
        if state is None:
            if len(cache) >= 16:        # thresholds that change every call: no point in capturing
                return fn()
            cache[(name, key)] = "warm"
            return fn()
        if state == "warm":
            from .engine import capture_graph
            torch.cuda.synchronize(self.device)
            state = cache[(name, key)] = capture_graph(fn)
        state.replay()

    def tail(self, conf_thresh, class_idx, thr_img, thr_radar, regress_boxes, graphs=True):
        """NMS -> proposals (radar-box count from self.radar_n_dev) -> RoI gathers -> heads -> output rows, as one graph
        per (detector output slot, thresholds)."""
        key = (self.base._slot, bool(regress_boxes), float(conf_thresh), int(class_idx), float(thr_img), float(thr_radar))

        def body():
            self.proposals_dev(key[2], key[3])
            self.heads(key[4], key[5], key[1])

        self.run_graphed("tail", key, body, graphs)

    def heads(self, thr_img, thr_radar, regress_boxes):
        n, g, cap = self.n, self.g, self.cap
        ops.roi_gather_bin_major(self.roi_score, n, g, g, 512, 10, 7, 1.0 / 16, self.rois, self.counts, cap, self.crop_img, 512, True)
        ops.roi_gather_bin_major(self.radar_score, n, g, g, 32, 10, 7, 1.0 / 16, self.rois, self.counts, cap, self.crop_radar, 496, False)
        ops.conv_gemm(self.crop_img, self.fc0, cap, 1, 1, 512, self.hidden, 256, act=ME_ACT_LEAKY, cin=490)
        ops.fusion_heads(self.hidden, 256, self.crop_radar, 496, self.head_weights, self.img_boxes, self.counts, cap,
                         self.regress, self.refine, self.mask)
        ops.finalize_output(self.img_boxes, self.rois, self.refine, self.regress, self.mask, self.counts, cap,
                            thr_img, thr_radar, regress_boxes, self.out, self.out_count, self.final_ws)
        self.launches += 5


class Network(nn.Module):
    """milliEye fusion model (reference my_models.py:411-640), inference branch on the B200 kernels."""

    def __init__(self, base_detector, conf_thresh):
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.conf_thresh = conf_thresh
        self.seen = 0
        self.iou_thresh = (0.3, 0.7)
        self.alpha = 0.75
        self.balance_factor = 5
        self.loss_lambda = (6, 1)
        self.refine_threshold_img, self.refine_threshold_radar = 0, 0
        self.class_num = 1
        self.class_idx = 0

        self.base_detector = base_detector.eval()
        self.img_cnn_layers = cnn_layers_1((256, 490))
        self.radar_cnn_layers = cnn_layers_3()
        self.refinement_head = refinement_head((490, 256, 128, self.class_num + 1))
        self.ensemble_head = ensemble_head((2, 32, 32 * (1 + self.class_num), 2))
        self._plans = {}

    def _invalidate(self):
        self._plans = {}

    def train(self, mode=True):
        """nn.Module.train; leaving training mode drops the inference plans, whose packed fp16 weights and folded
        BatchNorm statistics were taken before the training steps changed them."""
        was_training = any(m.training for m in (self.img_cnn_layers, self.radar_cnn_layers, self.refinement_head))
        super().train(mode)
        if was_training and not mode:
            self._invalidate()
        return self

    def load_state_dict(self, *args, **kwargs):
        self._invalidate()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def refresh_weights(self):
        self._invalidate()
        self.base_detector.refresh_weights()

    def _plan(self, base_plan, num_radar, slot=0):
        key = (base_plan.n, base_plan.size, base_plan.device.index, id(base_plan), slot)
        plan = self._plans.get(key)
        need = max(num_radar, 1)
        if plan is None or plan.radar_cap < need:
            cap = max(16 * base_plan.n, 1 << (need - 1).bit_length())
            with torch.cuda.device(base_plan.device):
                plan = _FusionPlan(self, base_plan, base_plan.n, base_plan.size, base_plan.device, cap)
            self._plans[key] = plan
        return plan

    def forward(self, images, maps, radar_boxes_location, model_mode=0, targets=None):
        """Same contract as reference my_models.py:433-452.  images (N,3,S,S) fp32 0..1; maps (N,3,S/16,S/16);
        radar_boxes_location (n,5) [frame, x1,y1,x2,y2] in 0..1 - scaled by S IN PLACE like the reference (:491);
        model_mode 0 fusion / 1 YOLO only / 2 radar only.  Returns output (K,8) on the device."""
        heads_training = any(m.training for m in (self.img_cnn_layers, self.radar_cnn_layers, self.refinement_head))
        if targets is not None and model_mode != 0:
            raise MeError("targets are only used with model_mode 0 (reference my_models.py:476-480 returns / "
                          "changes thresholds before the loss branch in the other modes)")
        if heads_training and (targets is None or model_mode != 0):
            raise MeError("in train() mode the heads run on batch statistics, which only the training step "
                          "(forward(images, maps, radar_boxes, 0, targets), train.py:185) defines; call .eval() for inference")
        base = self.base_detector
        plan_b = base.forward_device(images)
        dev = plan_b.device
        n_radar = int(radar_boxes_location.shape[0]) if radar_boxes_location is not None else 0
        plan = self._plan(plan_b, n_radar)
        with torch.cuda.device(dev):
            plan.launches = 0
            if model_mode == 1:
                plan.proposals(self.conf_thresh, self.class_idx, 0)
                n_img = int(plan.counts[0].item())
                return plan.img_boxes[:n_img, :8].clone()
            if model_mode == 2:
                self.refine_threshold_img = 1   # persistent, like the reference (:480)
            if n_radar > 0:
                radar_boxes_location[:, 1:] *= images.shape[-1]          # reference side effect (:491)
                plan.radar_dev[:n_radar].copy_(radar_boxes_location, non_blocking=True)
            if heads_training:
                return self._train_branch(plan, images, maps, n_radar, targets)
            plan.maps_in.copy_(maps, non_blocking=True)
            plan.radar_n_dev.fill_(n_radar)
            # the blocking call is bound by the kernels themselves (0.85 ms at batch 32 with or without graphs,
            # tools/fusion_block_probe.py), so it only replays graphs on request; FusionPipeline always does
            graphs = os.environ.get("ME_FUSION_GRAPHS_BLOCKING", "0") == "1"
            plan.run_graphed("score", (), plan.score_maps, graphs)
            plan.tail(self.conf_thresh, self.class_idx, self.refine_threshold_img, self.refine_threshold_radar, model_mode != 2,
                      graphs)
            self.refinement_head.count += 1
            k = int(plan.out_count.item())                               # the forward's only host sync
            output = plan.out[:k].clone()
            if targets is None:
                return output
            return self._loss_branch(plan, images.shape[3], targets, output)

    def _loss_branch(self, plan, img_size, targets, output, radar_score_rows=None):
        """Reference my_models.py:545-640 on the forward's device buffers: labels (me_stage3_labels), balanced
        sampling with python's `random` on the host like the reference (:600), losses + counters (me_stage3_loss).
        `targets` (m,6) [image, class, cx, cy, w, h] in 0..1 is rewritten IN PLACE to pixel x1y1x2y2 as the reference
        does (:548-549).  Returns (loss, output, metric, radar_attention); loss carries no autograd graph."""
        import random

        import numpy as np

        from .utils import xywh2xyxy
        targets[:, 2:] = xywh2xyxy(targets[:, 2:])
        targets[:, 2:] *= img_size
        t = int(targets.shape[0])
        plan.ensure_loss_buffers(t)
        if t:
            plan.targets_dev[:t].copy_(targets.to(torch.float32), non_blocking=True)
        ops.stage3_labels(plan.img_boxes, plan.rois, plan.counts, plan.cap, plan.targets_dev, t, plan.iou_labels,
                          plan.target_location)
        n_img, n_all = (int(v) for v in plan.counts.tolist())
        iou = plan.iou_labels[:n_all].cpu().numpy()
        pos, neg = iou > self.iou_thresh[1], iou < self.iou_thresh[0]
        pos_idx, neg_idx = np.where(pos)[0], np.where(neg)[0]
        top_k = min(len(pos_idx) * self.balance_factor, len(neg_idx))
        keep = pos.copy()
        if top_k > 0:
            keep[neg_idx[random.sample(range(len(neg_idx)), k=top_k)]] = True
        plan.sample_filter.zero_()
        if n_all:
            plan.sample_filter[:n_all].copy_(torch.from_numpy(keep.astype(np.uint8)), non_blocking=True)
        ops.stage3_loss(plan.rois, plan.refine, plan.regress, plan.mask, plan.counts, plan.cap, plan.iou_labels,
                        plan.target_location, plan.sample_filter, plan.loss_out, self.iou_thresh[1], self.alpha,
                        self.loss_lambda[0], float(self.refine_threshold_img), float(self.refine_threshold_radar))
        vals = plan.loss_out.cpu()
        conf_1 = torch.cat((plan.img_boxes[:n_img, 5], plan.refine[n_img:n_all, 0])).cpu()
        conf_2 = plan.mask[:n_all].cpu()
        lab = torch.from_numpy(iou)
        confs = dict(conf_1_pos=conf_1[lab > 0.5], conf_1_neg=conf_1[lab < 0.5],
                     conf_2_pos=conf_2[lab > 0.5], conf_2_neg=conf_2[lab < 0.5])
        metric = dict(total=n_all, true=torch.tensor(int(vals[6])), positive=torch.tensor(int(vals[7])), tp=vals[8].clone(),
                      conf=confs)
        self.last_losses = dict(masks_loss=float(vals[0]), conf_loss=float(vals[1]), loss_xy=float(vals[2]),
                                loss_wh=float(vals[3]), category_loss=float(vals[4]))
        self._last_sample = (pos, keep)
        g = plan.g
        if radar_score_rows is not None:     # train branch: fp32 rows [n*g*g, 10]
            radar_attention = radar_score_rows[:, 0].reshape(plan.n, 1, g, g).clone()
        else:
            radar_attention = plan.radar_score[..., 0].to(torch.float32).reshape(plan.n, 1, g, g)
        return plan.loss_out[5].clone(), output, metric, radar_attention

    # ------------------------------------------------------------------ training step (train.py:169-191)
    def _head_tensors(self):
        params = {k: v.data for k, v in self.named_parameters() if not k.startswith("base_detector.")}
        buffers = {k: v for k, v in self.named_buffers() if not k.startswith("base_detector.") and "running_" in k}
        return params, buffers

    def _train_branch(self, plan, images, maps, n_radar, targets):
        """model.train() forward with targets: the frozen detector's proposals, then the heads in fp32 on batch
        statistics (stage3_train.HeadTrainer), labels / balanced sample / losses like the eval branch, and a loss tensor
        whose backward() runs the hand-derived backward kernels and fills .grad of every head parameter that requires
        one - so `loss.backward(); optimizer.step()` of train.py:186-190 work unchanged."""
        from . import stage3_train as st
        dev = plan.device
        if getattr(self, "_trainer", None) is None or self._trainer.device != dev:
            self._trainer = st.HeadTrainer(dev)
        tr = self._trainer
        n, g = plan.n, plan.g
        P = n * g * g
        plan.proposals(self.conf_thresh, self.class_idx, n_radar)
        n_img, n_all = (int(v) for v in plan.counts.tolist())
        fv = plan.base.feature_view
        feat_rows = st.half_rows_to_float(fv.t, P, 256, fv.pitch, tr._buf("feat_rows", (P, 256)))
        maps_rows = st.nchw_to_rows(maps.to(device=dev, dtype=torch.float32).contiguous(), tr._buf("maps_rows", (P, 3)))
        params, buffers = self._head_tensors()
        for k, v in params.items():
            if v.dtype != torch.float32 or not v.is_contiguous():
                raise MeError(f"training needs contiguous fp32 parameters ({k} is {v.dtype})")
        cache = tr.forward(params, buffers, feat_rows, maps_rows, n, g, plan.rois, plan.img_boxes, n_img, n_all,
                           regress_out=plan.regress, refine_out=plan.refine, mask_out=plan.mask)
        for m in (self.img_cnn_layers, self.radar_cnn_layers, self.refinement_head):
            for mod in m.modules():
                if isinstance(mod, (nn.BatchNorm2d, nn.BatchNorm1d)) and mod.num_batches_tracked is not None:
                    mod.num_batches_tracked += 1
        ops.finalize_output(plan.img_boxes, plan.rois, plan.refine, plan.regress, plan.mask, plan.counts, plan.cap,
                            float(self.refine_threshold_img), float(self.refine_threshold_radar), True, plan.out, plan.out_count,
                            plan.final_ws)
        self.refinement_head.count += 1
        k = int(plan.out_count.item())
        output = plan.out[:k].clone()
        loss_val, output, metric, radar_attention = self._loss_branch(plan, images.shape[3], targets, output,
                                                                      radar_score_rows=cache["s"])
        pos, keep = self._last_sample
        pos_d = tr._bufs["pos"] = torch.from_numpy(pos.astype("uint8")).to(dev) if n_all else torch.zeros(1, dtype=torch.uint8, device=dev)
        sel_d = tr._bufs["sel"] = torch.from_numpy(keep.astype("uint8")).to(dev) if n_all else torch.zeros(1, dtype=torch.uint8, device=dev)
        named = dict(self.named_parameters())
        image_path = all(named[k].requires_grad for k in st.IMAGE_PATH)
        names = [k for k in st.TRAINABLE if named[k].requires_grad and (image_path or k not in st.IMAGE_PATH)]
        self._train_state = dict(cache=cache, params=params, pos=pos_d, sel=sel_d, image_path=image_path, names=names)
        loss = _Stage3Loss.apply(loss_val, self, *[named[k] for k in names])
        return loss, output, metric, radar_attention

    def backward_into(self, grads_out=None):
        """The backward of the last training forward, without autograd: gradients are written into `grads_out`
        ({name: tensor}, e.g. Stage3Optimizer.grads - overwritten, not accumulated) and returned."""
        s = self._train_state
        return self._trainer.backward(s["params"], s["cache"], s["pos"], s["sel"], self.alpha, self.loss_lambda[0],
                                      image_path=s["image_path"], out=grads_out)


class _Stage3Loss(torch.autograd.Function):
    """loss tensor of the training forward; backward() = HeadTrainer.backward (csrc/train_ops.cu)."""

    @staticmethod
    def forward(ctx, loss_val, model, *params):
        ctx.model = model
        ctx.names = list(model._train_state["names"])
        return loss_val.clone()

    @staticmethod
    def backward(ctx, grad_out):
        grads = ctx.model.backward_into(None)
        return (None, None) + tuple(grads[k] * grad_out for k in ctx.names)


class FusionPipeline:
    """Network.forward (reference my_models.py:433-640, inference branch) for STREAMS of batches: run_mp.py:300-330 and
    test_fusion.py:70-80 call the model once per frame / batch and read the rows back before the next call, so the
    device idles during the proposal + head kernels' short launches and the host read.  submit() enqueues one batch
    and returns its record at once; record.wait() gives the batch's (K, 8) rows.

    Stream layout per batch i:
      main stream  backbone(i) (graph replay) -> score maps(i) (they read the backbone's single feature-map buffer, so
                   they stay in front of backbone(i+1))
      side stream  YOLO decode -> filter + NMS -> proposals -> RoI gathers -> refinement / ensemble heads -> finalize
                   -> ONE device->host copy of rows + count; all of it overlaps backbone(i+1) on the main stream.
    Each in-flight batch owns one fusion plan (score maps, proposal and head buffers) out of a ring of `depth`;
    a plan is reused only after its previous record completed.  The rows equal Network.forward's bit for bit
    (tests/test_gpu_models.py::test_fusion_pipeline_equals_forward)."""

    class Record:
        def __init__(self, plan):
            self.plan = plan
            self.host_flat = torch.empty_like(plan.out_flat, device="cpu").pin_memory()
            self.done = torch.cuda.Event()
            self.pending = False
            self.readback = False

        def wait(self):
            """Blocks until the batch is complete; returns its rows — a view of the record's pinned host buffer when
            the batch was submitted with readback=True (valid until the record is reused, `depth` submits later),
            else a device tensor."""
            self.done.synchronize()
            self.pending = False
            cap = self.plan.cap
            if self.readback:
                k = int(self.host_flat[cap * 8:cap * 8 + 1].view(torch.int32)[0])
                return self.host_flat[:cap * 8].view(cap, 8)[:k]
            return self.plan.out[:int(self.plan.out_count.item())]

    TAIL_LAUNCHES = 13        # 5 score-map convs, filter + NMS (2), proposals, 2 RoI gathers, FC GEMM, heads, finalize

    def __init__(self, model, depth=4, use_cuda_graph=None):
        self.model = model
        # an even ring keeps every fusion plan on ONE of the detector's two output slots: one tail graph per plan
        self.depth = max(2, int(depth) + (int(depth) & 1))
        self._records = {}     # id(fusion plan) -> Record
        self._next = 0
        self._side = None
        self.launches = self.TAIL_LAUNCHES   # fusion-tail kernels per submit (the backbone's are in its plan)
        # the score-map convs and the proposal / head kernels replay as two CUDA graphs per plan (13 launches -> 2):
        # all their shapes are fixed by the plan, row counts live on the device
        self.use_cuda_graph = (os.environ.get("ME_FUSION_GRAPHS", "1") != "0") if use_cuda_graph is None else bool(use_cuda_graph)

    def submit(self, images, maps, radar_boxes_location, model_mode=0, readback=True):
        m = self.model
        if model_mode == 1:
            raise MeError("model_mode 1 is the bare detector + NMS: use millieye_b200.models.DetectPipeline")
        if any(x.training for x in (m.img_cnn_layers, m.radar_cnn_layers, m.refinement_head)):
            raise MeError("FusionPipeline is the inference path: call model.eval() first")
        base = m.base_detector
        plan_b = base.forward_device(images, decode=False)
        dev = plan_b.device
        n_radar = int(radar_boxes_location.shape[0]) if radar_boxes_location is not None else 0
        slot = self._next
        self._next = (slot + 1) % self.depth
        plan = m._plan(plan_b, n_radar, slot=1 + slot)   # slot 0 belongs to the blocking forward()
        with torch.cuda.device(dev):
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            rec = self._records.get(id(plan))
            if rec is None or rec.plan is not plan:
                live = {id(p) for p in m._plans.values()}
                for key in [k for k in self._records if k not in live]:
                    del self._records[key]
                rec = self._records[id(plan)] = FusionPipeline.Record(plan)
            if rec.pending:
                rec.done.synchronize()      # the plan's previous batch still runs: wait before its buffers are reused
            rec.pending, rec.readback = True, bool(readback)
            plan.launches = 0
            if model_mode == 2:
                m.refine_threshold_img = 1   # persistent, like the reference (:480)
            if n_radar > 0:
                radar_boxes_location[:, 1:] *= images.shape[-1]          # reference side effect (:491)
                plan.radar_dev[:n_radar].copy_(radar_boxes_location, non_blocking=True)
            plan.radar_n_dev.fill_(n_radar)
            plan.maps_in.copy_(maps, non_blocking=True)
            plan.run_graphed("score", (), plan.score_maps, self.use_cuda_graph)
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                plan_b.run_decode(base.use_cuda_graph)
                plan.tail(m.conf_thresh, m.class_idx, m.refine_threshold_img, m.refine_threshold_radar, model_mode != 2,
                          self.use_cuda_graph)
                if readback:
                    rec.host_flat.copy_(plan.out_flat, non_blocking=True)
                rec.done.record()
            plan_b.hold_output(rec.done)
            m.refinement_head.count += 1
        return rec
