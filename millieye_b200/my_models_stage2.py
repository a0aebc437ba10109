"""Drop-in for the reference's module2_mixed/my_models.py (stage-2 model: YOLO + R-CNN refinement, 12 classes, no
radar): `Network(base_detector, conf_thresh).forward(images, targets=None)` -> output (K,8) on the CPU, like the
reference (:361).  Same state_dict keys (fcn_layers.*, refinement_head.net{0,1,2}.*, ensemble_head.fc{1,2}.*).

Pipeline on the device: Darknet graph -> me_filter_nms (every class kept) -> me_build_proposals(class_idx=-1) ->
me_conv_gemm 256->490 score map -> me_psroi_align -> me_conv_gemm 490->256 -> me_stage2_heads -> me_finalize_output.
"""
import torch
from torch import nn

from . import ops
from ._lib import ME_ACT_LEAKY, MeError
from .my_models import cnn_layers_1, define_yolo, init_yolo  # noqa: F401  (same helpers as the stage-3 module)


class fcn_layers(cnn_layers_1):
    """module2_mixed/my_models.py:47-77 (identical to stage 3's cnn_layers_1). Parameter container."""


class refinement_head(nn.Module):
    """module2_mixed/my_models.py:96-126. Parameter container (Dropout is the identity in eval mode)."""

    def __init__(self, channels):
        super().__init__()
        self.net0 = nn.Sequential(nn.Linear(channels[0], channels[1]), nn.LeakyReLU(0.1), nn.Dropout(0.5))
        self.net1 = nn.Sequential(nn.Linear(channels[1], 4))
        self.net2 = nn.Sequential(nn.Linear(channels[1], channels[2]), nn.Sigmoid())


class ensemble_head(nn.Module):
    """module2_mixed/my_models.py:129-163. Parameter container."""

    def __init__(self, channels, use_activation=True):
        super().__init__()
        self.use_activation = use_activation
        self.fc1 = nn.Sequential(nn.Linear(channels[0], channels[1]), nn.LeakyReLU(0.1))
        self.fc2 = nn.Sequential(nn.Linear(channels[2], channels[3]), nn.LeakyReLU(0.1))
        self.softmax = nn.Softmax(dim=1)


class _Stage2Plan:
    MAX_DET = 200

    def __init__(self, net, base_plan):
        self.base = base_plan
        n, device = base_plan.n, base_plan.device
        fv = base_plan.feature_view
        if fv is None or fv.real_c != 256:
            raise MeError("stage-2 Network needs a 256-channel stride-16 feature tap on the base detector")
        self.n, self.g, self.device = n, fv.h, device
        g = self.g
        nc = base_plan.attrs - 5
        if nc != net.class_num:
            raise MeError(f"detector has {nc} classes, stage-2 heads expect {net.class_num}")
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in net.state_dict().items()
              if v.is_floating_point() and not k.startswith("base_detector.")}
        f16 = dict(dtype=torch.float16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        p = "fcn_layers.net."
        self.img_conv = ops.pack_conv(sd[p + "conv_0.weight"], sd[p + "conv_0.bias"],
                                      (sd[p + "batch_norm_0.weight"], sd[p + "batch_norm_0.bias"],
                                       sd[p + "batch_norm_0.running_mean"], sd[p + "batch_norm_0.running_var"], 1e-5),
                                      cout_pad=512)
        self.roi_score = torch.zeros((n, g, g, 512), **f16)
        self.cap = cap = n * self.MAX_DET
        self.box_pitch = 8 + nc
        self.nms = ops.NmsBuffers(n, base_plan.rows_total, nc, self.MAX_DET, device)
        self.boxes = torch.zeros((cap, self.box_pitch), **f32)
        self.rois = torch.zeros((cap, 5), **f32)
        self.counts = torch.zeros((2,), **i32)
        self.crop = torch.zeros((cap, 512), **f16)
        self.hidden = torch.zeros((cap, 256), **f16)
        w0 = sd["refinement_head.net0.0.weight"]
        self.fc0 = ops.pack_conv(w0.view(w0.shape[0], w0.shape[1], 1, 1), sd["refinement_head.net0.0.bias"], None)
        self._w = {
            "net1_w": sd["refinement_head.net1.0.weight"].contiguous(), "net1_b": sd["refinement_head.net1.0.bias"].contiguous(),
            "net2_w": sd["refinement_head.net2.0.weight"].contiguous(), "net2_b": sd["refinement_head.net2.0.bias"].contiguous(),
            "fc1_w": sd["ensemble_head.fc1.0.weight"].contiguous(), "fc1_b": sd["ensemble_head.fc1.0.bias"].contiguous(),
            "fc2_w": sd["ensemble_head.fc2.0.weight"].contiguous(), "fc2_b": sd["ensemble_head.fc2.0.bias"].contiguous(),
        }
        self.weights = ops.make_stage2_weights(self._w)
        self.regress = torch.zeros((cap, 4), **f32)
        self.mask = torch.zeros((cap,), **f32)
        self.refine = torch.zeros((cap, 1 + nc), **f32)      # refinement vectors: [conf, class scores] (loss branch)
        self.out = torch.zeros((cap, 8), **f32)
        self.out_count = torch.zeros((1,), **i32)
        ws = 1
        while ws < cap:
            ws <<= 1
        self.final_ws = torch.zeros((ws * 8,), dtype=torch.uint8, device=device)

    def run(self, conf_thresh, refine_threshold):
        n, g, cap = self.n, self.g, self.cap
        fv = self.base.feature_view
        ops.conv_gemm(fv.t, self.img_conv, n, g, g, fv.pitch, self.roi_score, 512, act=ME_ACT_LEAKY, cin=256)
        ops.filter_nms(self.base.yolo_out, conf_thresh, 0.5, self.MAX_DET, xyxy_inplace=True, buffers=self.nms)
        ops.build_proposals(self.nms.det, self.nms.count, -1, None, 1.0, self.boxes, self.rois, self.counts, cap)
        ops.psroi_align(self.roi_score, n, g, g, 512, 10, 7, 1.0 / 16, self.rois, self.counts, cap, self.crop, 512)
        ops.conv_gemm(self.crop, self.fc0, cap, 1, 1, 512, self.hidden, 256, act=ME_ACT_LEAKY, cin=490)
        ops.stage2_heads(self.hidden, 256, self.weights, self.boxes, self.box_pitch, self.box_pitch - 7, self.counts, cap,
                         self.regress, self.mask, self.refine)
        ops.finalize_output(self.boxes, self.rois, self.mask, self.regress, self.mask, self.counts, cap,
                            refine_threshold, refine_threshold, True, self.out, self.out_count, self.final_ws,
                            box_pitch=self.box_pitch)


class Network(nn.Module):
    """Stage-2 fusion-less model (reference module2_mixed/my_models.py:280-461), inference branch."""

    def __init__(self, base_detector, conf_thresh):
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.conf_thresh = conf_thresh
        self.seen = 0
        self.iou_thresh = (0.3, 0.7)
        self.alpha = 0.75
        self.balance_fac = 5
        self.loss_lambda = (15, 5)
        self.refine_threshold = 0
        self.class_num = 12

        self.base_detector = base_detector.eval()
        self.fcn_layers = fcn_layers((256, 490))
        self.refinement_head = refinement_head((490, 256, self.class_num + 1))
        self.ensemble_head = ensemble_head((2, 32, 32 * (self.class_num + 1), 2))
        self._plans = {}

    def _invalidate(self):
        self._plans = {}

    def load_state_dict(self, *args, **kwargs):
        self._invalidate()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def refresh_weights(self):
        self._invalidate()
        self.base_detector.refresh_weights()

    def forward(self, images, targets=None):
        """images (N,3,S,S) fp32 -> output (K,8) [image_i, x1,y1,x2,y2, new_conf, class_score, class_pred] on the
        CPU, sorted by new_conf descending (reference :299-361)."""
        if targets is not None and any(m.training for m in (self.fcn_layers, self.refinement_head)):
            raise MeError("the stage-2 loss branch runs with running BatchNorm statistics and without Dropout: call .eval() "
                          "(training the R-CNN heads - batch statistics, Dropout, backward - is not on the accelerated path)")
        base_plan = self.base_detector.forward_device(images)
        key = id(base_plan)
        plan = self._plans.get(key)
        with torch.cuda.device(base_plan.device):
            if plan is None:
                plan = self._plans[key] = _Stage2Plan(self, base_plan)
            plan.run(self.conf_thresh, float(self.refine_threshold))
            k = int(plan.out_count.item())
            output = plan.out[:k].cpu()
            if targets is None:
                return output
            return self._loss_branch(plan, images.shape[-1], targets, output)

    def _loss_branch(self, plan, img_size, targets, output):
        """Reference module2_mixed/my_models.py:363-461 on the forward's device buffers: labels (me_stage3_labels -
        obtain_iou_labels is the same routine in both modules), balanced sampling with python's `random` on the host
        like the reference (:411), losses + counters (me_stage2_loss).  `targets` (m,6) is rewritten IN PLACE to pixel
        x1y1x2y2 as the reference does (:367-368).  Returns (output, loss, metric); loss carries no autograd graph."""
        import random

        import numpy as np

        from .utils import xywh2xyxy
        targets[:, 2:] = xywh2xyxy(targets[:, 2:])
        targets[:, 2:] *= img_size
        t = int(targets.shape[0])
        dev = plan.device
        f32 = dict(dtype=torch.float32, device=dev)
        if getattr(plan, "iou_labels", None) is None:
            plan.iou_labels = torch.zeros((plan.cap,), **f32)
            plan.target_location = torch.zeros((plan.cap, 4), **f32)
            plan.sample_filter = torch.zeros((plan.cap,), dtype=torch.uint8, device=dev)
            plan.pos_ws = torch.zeros((plan.cap,), dtype=torch.int32, device=dev)
            plan.loss_out = torch.zeros((16,), **f32)
        targets_dev = targets.to(**f32).contiguous() if t else None
        ops.stage3_labels(plan.boxes, plan.rois, plan.counts, plan.cap, targets_dev, t, plan.iou_labels, plan.target_location)
        n_all = int(plan.counts[1].item())
        iou = plan.iou_labels[:n_all].cpu().numpy()
        pos, neg = iou > self.iou_thresh[1], iou < self.iou_thresh[0]
        pos_idx, neg_idx = np.where(pos)[0], np.where(neg)[0]
        top_k = min(len(pos_idx) * self.balance_fac, len(neg_idx))
        keep = pos.copy()
        if top_k > 0:
            keep[neg_idx[random.sample(range(len(neg_idx)), k=top_k)]] = True
        plan.sample_filter.zero_()
        if n_all:
            plan.sample_filter[:n_all].copy_(torch.from_numpy(keep.astype(np.uint8)), non_blocking=True)
        ops.stage2_loss(plan.boxes, plan.box_pitch, plan.rois, plan.refine, plan.regress, plan.mask, plan.counts, plan.cap,
                        plan.iou_labels, plan.target_location, plan.sample_filter, plan.pos_ws, plan.loss_out,
                        self.iou_thresh[1], self.alpha, self.loss_lambda[0], self.loss_lambda[1], float(self.refine_threshold))
        vals = plan.loss_out.cpu()
        conf_1 = plan.boxes[:n_all, 5].cpu()
        conf_2 = plan.mask[:n_all].cpu()
        lab = torch.from_numpy(iou)
        confs = dict(conf_1_pos=conf_1[lab > 0.5], conf_1_neg=conf_1[lab < 0.5],
                     conf_2_pos=conf_2[lab > 0.5], conf_2_neg=conf_2[lab < 0.5])
        metric = dict(total=n_all, true=torch.tensor(int(vals[6])), positive=torch.tensor(int(vals[7])), tp=vals[8].clone(),
                      conf=confs)
        self.last_losses = dict(masks_loss=float(vals[0]), conf_loss=float(vals[1]), loss_xy=float(vals[2]),
                                loss_wh=float(vals[3]), category_loss=float(vals[4]))
        return output, plan.loss_out[5].clone(), metric
