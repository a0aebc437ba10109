"""Stage-3 TRAINING step of the fusion model on the B200 kernels (reference module3_our_dataset/train.py:169-191 and the
autograd backward of my_models.py:486-539 / 545-640): train-mode forward of the heads (batch-statistics BatchNorm with
running-statistics update), the loss gradient pushed back through ensemble_head, refinement_head, RoIAlign /
PS-RoIAlign and the two score-map CNNs, Adam, and the gradient all-reduce of a sharded batch.

The detector is frozen (`base_detector.eval()`, its outputs are detached in the reference, yolov3/models.py:255,266)
and stays on the fp16 tensor-core engine; everything with a gradient runs through the fp32 kernels of
csrc/train_ops.cu so that gradients match the reference's fp32 autograd (oracle/stage3_backward.py is the op-by-op
derivation; tests/test_gpu_train.py checks every parameter gradient against the reference-generated fixture).

  HeadTrainer.forward(...)  -> cache (device buffers of every intermediate the backward needs)
  HeadTrainer.backward(...) -> {parameter name: gradient}
  Stage3Optimizer           -> flat parameter / gradient / Adam-moment buffers: one NCCL all-reduce + one Adam kernel
"""
import ctypes

import torch

from . import _lib
from ._lib import ME_ACT_LEAKY, ME_ACT_LINEAR, ME_ACT_SIGMOID, check, ptr, stream_ptr

EPS = 1e-5
MOMENTUM = 0.1   # nn.BatchNorm2d(momentum=0.1) in cnn_layers_1 / cnn_layers_3 / radar_net (my_models.py:66,137,248)


# ---------------------------------------------------------------------------------------------- kernel front ends
def gemm_nt(x, w, out, bias=None, act=ME_ACT_LINEAR):
    """out[M,N] = x[M,K] @ w[N,K]^T (+ bias, act)"""
    m, k = x.shape
    n = w.shape[0]
    check(_lib.lib().me_gemm_f32(m, n, k, ptr(x), x.stride(0), 1, ptr(w), 1, w.stride(0), ptr(out), out.stride(0), ptr(bias), act, 0,
                                 stream_ptr()), "me_gemm_f32")
    return out


def gemm_nn(dz, w, out):
    """out[M,K] = dz[M,N] @ w[N,K]"""
    m, n = dz.shape
    k = w.shape[1]
    check(_lib.lib().me_gemm_f32(m, k, n, ptr(dz), dz.stride(0), 1, ptr(w), w.stride(0), 1, ptr(out), out.stride(0), None, 0, 0,
                                 stream_ptr()), "me_gemm_f32")
    return out


def gemm_tn(dz, x, out):
    """out[N,K] = dz[M,N]^T @ x[M,K]  (weight gradient)"""
    m, n = dz.shape
    k = x.shape[1]
    check(_lib.lib().me_gemm_f32(n, k, m, ptr(dz), 1, dz.stride(0), ptr(x), x.stride(0), 1, ptr(out), out.stride(0), None, 0, 0,
                                 stream_ptr()), "me_gemm_f32")
    return out


def colsum(x, out, y=None):
    rows, cols = x.shape
    check(_lib.lib().me_colsum_f32(ptr(x), ptr(y), rows, cols, x.stride(0), y.stride(0) if y is not None else 0, ptr(out),
                                   stream_ptr()), "me_colsum_f32")
    return out


def _all_reduce(t, group):
    """SUM over ranks in place (NCCL on GPUs; with the gloo backend - CPU-side tests - a CUDA tensor is staged through
    host memory)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if t.is_cuda and dist.get_backend(group) == "gloo":
            host = t.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
            t.copy_(host)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def bn_train_fwd(z, gamma, beta, running_mean, running_var, sums, mean_ws, inv_std, xhat, a, group=None, sync=False):
    """BatchNorm on batch statistics + LeakyReLU.  sums: float64 [2*cols + 1] (sum, sum of squares, row count) - with
    sync=True the ranks add theirs (ONE small all-reduce), so every rank normalises with the statistics of the whole
    batch and updates its running statistics identically."""
    L = _lib.lib()
    rows, cols = z.shape
    check(L.me_bn_partial_stats(ptr(z), rows, cols, ptr(sums), stream_ptr()), "me_bn_partial_stats")
    if sync:
        _all_reduce(sums, group)
    check(L.me_bn_finalize(ptr(sums), cols, EPS, MOMENTUM, ptr(running_mean), ptr(running_var), ptr(mean_ws), ptr(inv_std),
                           stream_ptr()), "me_bn_finalize")
    check(L.me_bn_apply(ptr(z), rows, cols, ptr(mean_ws), ptr(inv_std), ptr(gamma), ptr(beta), ptr(xhat), ptr(a), stream_ptr()),
          "me_bn_apply")


def bn_train_bwd(da, a, xhat, gamma, inv_std, sums, dgamma, dbeta, dz, scratch, group=None, sync=False):
    """dgamma / dbeta receive THIS rank's sums (the gradient all-reduce adds the ranks later); the input gradient uses
    the sums over the whole batch (all-reduced copy in `scratch`, float32 [2*cols]) and its row count (sums[2*cols])."""
    L = _lib.lib()
    rows, cols = a.shape
    check(L.me_bn_bwd_sums(ptr(da), ptr(a), ptr(xhat), rows, cols, ptr(dgamma), ptr(dbeta), stream_ptr()), "me_bn_bwd_sums")
    tg, tb = dgamma, dbeta
    if sync:
        scratch[:cols].copy_(dgamma.reshape(-1))
        scratch[cols:2 * cols].copy_(dbeta.reshape(-1))
        _all_reduce(scratch, group)
        tg, tb = scratch[:cols], scratch[cols:2 * cols]
    check(L.me_bn_bwd_apply(ptr(da), ptr(xhat), rows, cols, ptr(gamma), ptr(inv_std), ptr(tg), ptr(tb),
                            ctypes.c_void_p(sums.data_ptr() + 16 * cols), ptr(dz), stream_ptr()), "me_bn_bwd_apply")


def roi_align_f32(ps, feat, n, h, w, channels, rois, num_rois, out=None, grad_out=None, dfeat=None):
    backward = grad_out is not None
    chan_total = (feat if feat is not None else dfeat).shape[-1]
    check(_lib.lib().me_roi_align_f32(1 if ps else 0, 1 if backward else 0, ptr(feat), ptr(dfeat), n, h, w, chan_total, channels, 7,
                                      1.0 / 16, ptr(rois), num_rois, ptr(out), ptr(grad_out), stream_ptr()), "me_roi_align_f32")


TRAINABLE = (
    "img_cnn_layers.net.conv_0.weight", "img_cnn_layers.net.conv_0.bias",
    "img_cnn_layers.net.batch_norm_0.weight", "img_cnn_layers.net.batch_norm_0.bias",
    "radar_cnn_layers.conv1.0.weight", "radar_cnn_layers.conv1.0.bias", "radar_cnn_layers.conv1.1.weight", "radar_cnn_layers.conv1.1.bias",
    "radar_cnn_layers.conv2.0.weight", "radar_cnn_layers.conv2.0.bias", "radar_cnn_layers.conv2.1.weight", "radar_cnn_layers.conv2.1.bias",
    "radar_cnn_layers.conv3.0.weight", "radar_cnn_layers.conv3.0.bias", "radar_cnn_layers.conv3.1.weight", "radar_cnn_layers.conv3.1.bias",
    "radar_cnn_layers.conv3.3.weight", "radar_cnn_layers.conv3.3.bias",
    "refinement_head.net0.0.weight", "refinement_head.net0.0.bias", "refinement_head.net2.0.weight", "refinement_head.net2.0.bias",
    "refinement_head.radar_net.0.weight", "refinement_head.radar_net.0.bias", "refinement_head.radar_net.1.weight",
    "refinement_head.radar_net.1.bias", "refinement_head.radar_net.3.weight", "refinement_head.radar_net.3.bias",
    "ensemble_head.fc1.0.weight", "ensemble_head.fc1.0.bias", "ensemble_head.fc2.0.weight", "ensemble_head.fc2.0.bias",
)
IMAGE_PATH = TRAINABLE[:4] + TRAINABLE[18:22]   # the set train.py freezes with --pretrained_module2 (train.py:117-149)


class HeadTrainer:
    """fp32 train-mode forward and backward of the heads.  `params` / `buffers`: {reference key: fp32 cuda tensor}
    (the live nn.Parameter / buffer storage: running statistics are updated in place like nn.BatchNorm2d does)."""

    def __init__(self, device, group=None, sync_bn=None):
        """sync_bn: None = synchronise BatchNorm statistics whenever torch.distributed runs with more than one rank
        (the sharded step then reproduces the single-process step on the whole batch); False = per-shard statistics
        (what DistributedDataParallel does without SyncBatchNorm)."""
        self.device, self.group = device, group
        if sync_bn is None:
            import torch.distributed as dist
            sync_bn = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.sync_bn = bool(sync_bn)
        self._bufs = {}

    def _buf(self, name, shape, zero=False, dtype=torch.float32):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self._bufs[name] = torch.zeros(shape, dtype=dtype, device=self.device)
        elif zero:
            t.zero_()
        return t

    def _rows(self, name, rows, cols, zero=False):
        """[rows, cols] view of a buffer that always has at least one row (kernels get valid pointers for 0 rows)."""
        return self._buf(name, (max(rows, 1), cols), zero=zero)[:rows]

    # ------------------------------------------------------------------ forward
    def forward(self, params, buffers, feat_rows, maps_rows, n, g, rois, img_boxes, n_img, n_all, regress_out=None,
                refine_out=None, mask_out=None, update_running=True):
        """feat_rows [n*g*g, 256] fp32 (detector feature map, NHWC rows), maps_rows [n*g*g, 3], rois [>= n_all, 5] pixels,
        img_boxes [>= n_img, 9].  Returns the cache; refine [n_all,2] / mask [n_all] / regress [n_all,4] are written into
        the given buffers (the fusion plan's) or fresh ones."""
        L = _lib.lib()
        p, P = params, n * g * g
        B = self._buf
        c = dict(n=n, g=g, P=P, n_img=n_img, n_all=n_all, rois=rois, img_boxes=img_boxes, feat=feat_rows)

        def running(prefix):
            if not update_running:
                return None, None
            return buffers[prefix + "running_mean"], buffers[prefix + "running_var"]

        # img_cnn_layers: 1x1 conv 256 -> 490 + BN (batch statistics) + LeakyReLU   (my_models.py:62-77)
        wi = p["img_cnn_layers.net.conv_0.weight"].view(490, 256)
        zi = gemm_nt(feat_rows, wi, B("zi", (P, 490)), p["img_cnn_layers.net.conv_0.bias"])
        rm, rv = running("img_cnn_layers.net.batch_norm_0.")
        c["xhi"], c["ai"], c["istdi"] = B("xhi", (P, 490)), B("ai", (P, 490)), B("istdi", (490,))
        sb = dict(group=self.group, sync=self.sync_bn)
        c["sumsi"] = B("sumsi", (2 * 490 + 1,), dtype=torch.float64)
        bn_train_fwd(zi, p["img_cnn_layers.net.batch_norm_0.weight"], p["img_cnn_layers.net.batch_norm_0.bias"], rm, rv,
                     c["sumsi"], B("meani", (490,)), c["istdi"], c["xhi"], c["ai"], **sb)
        # radar_cnn_layers: three 3x3 conv + BN + LeakyReLU, 1x1 conv + sigmoid   (:133-157)
        a = maps_rows
        c["a0"] = a
        for i, (name, cin, cout) in enumerate((("conv1", 3, 32), ("conv2", 32, 64), ("conv3", 64, 128)), 1):
            cols = B(f"cols{i}", (P, cin * 9))
            check(L.me_im2col3_f32(ptr(a), n, g, g, cin, ptr(cols), stream_ptr()), "me_im2col3_f32")
            w = p[f"radar_cnn_layers.{name}.0.weight"].view(cout, cin * 9)
            z = gemm_nt(cols, w, B(f"z{i}", (P, cout)), p[f"radar_cnn_layers.{name}.0.bias"])
            rm, rv = running(f"radar_cnn_layers.{name}.1.")
            xh, act, istd = B(f"xh{i}", (P, cout)), B(f"a{i}", (P, cout)), B(f"istd{i}", (cout,))
            c[f"sums{i}"] = B(f"sums{i}", (2 * cout + 1,), dtype=torch.float64)
            bn_train_fwd(z, p[f"radar_cnn_layers.{name}.1.weight"], p[f"radar_cnn_layers.{name}.1.bias"], rm, rv,
                         c[f"sums{i}"], B(f"mean{i}", (cout,)), istd, xh, act, **sb)
            c[f"cols{i}"], c[f"xh{i}"], c[f"a{i}"], c[f"istd{i}"] = cols, xh, act, istd
            a = act
        w4 = p["radar_cnn_layers.conv3.3.weight"].view(10, 128)
        c["s"] = gemm_nt(a, w4, B("s", (P, 10)), p["radar_cnn_layers.conv3.3.bias"], ME_ACT_SIGMOID)
        R = n_all   # may be 0 on this rank: every call below then works on empty row sets (the collectives still run)
        RB = self._rows
        # RoI crops (:495-496), flatten order (c, ph, pw)
        c["x_img"] = RB("x_img", R, 490)
        c["x_rad"] = RB("x_rad", R, 490)
        roi_align_f32(True, c["ai"], n, g, g, 10, rois, R, out=c["x_img"])
        roi_align_f32(False, c["s"], n, g, g, 10, rois, R, out=c["x_rad"])
        # refinement_head (:260-284)
        h = "refinement_head."
        c["t"] = gemm_nt(c["x_img"], p[h + "net0.0.weight"], RB("t", R, 256), p[h + "net0.0.bias"], ME_ACT_LEAKY)
        regress = regress_out if regress_out is not None else B("regress", (max(R, 1), 4))
        gemm_nt(c["t"], p[h + "net1.0.weight"], regress[:R], p[h + "net1.0.bias"])
        c["cls"] = gemm_nt(c["t"], p[h + "net2.0.weight"], RB("cls", R, 13), p[h + "net2.0.bias"], ME_ACT_SIGMOID)
        wr = p[h + "radar_net.0.weight"].view(10, 490)
        r1 = gemm_nt(c["x_rad"], wr, RB("r1", R, 10), p[h + "radar_net.0.bias"])
        rm, rv = running(h + "radar_net.1.")
        c["xhr"], c["ar"], c["istdr"] = RB("xhr", R, 10), RB("ar", R, 10), B("istdr", (10,))
        c["sumsr"] = B("sumsr", (21,), dtype=torch.float64)
        bn_train_fwd(r1, p[h + "radar_net.1.weight"], p[h + "radar_net.1.bias"], rm, rv, c["sumsr"], B("meanr", (10,)),
                     c["istdr"], c["xhr"], c["ar"], **sb)
        r2 = gemm_nt(c["ar"], p[h + "radar_net.3.weight"].view(1, 10), RB("r2", R, 1), p[h + "radar_net.3.bias"])
        # confidence, ensemble head, masks (:276-284, 202-210, 513-514)
        e = "ensemble_head."
        c["rc"], c["p"] = B("rc", (max(R, 1),)), B("p", (max(n_img, 1), 2))
        refine = refine_out if refine_out is not None else B("refine", (max(R, 1), 2))
        mask = mask_out if mask_out is not None else B("mask", (max(R, 1),))
        check(L.me_stage3_tail_fwd(ptr(r2), ptr(c["cls"]), 13, ptr(img_boxes), n_img, R, ptr(p[e + "fc1.0.weight"]),
                                   ptr(p[e + "fc1.0.bias"]), ptr(p[e + "fc2.0.weight"]), ptr(p[e + "fc2.0.bias"]), ptr(c["rc"]),
                                   ptr(refine), ptr(mask), ptr(c["p"]), stream_ptr()), "me_stage3_tail_fwd")
        c["refine"], c["mask"], c["regress"] = refine, mask, regress
        return c

    # ------------------------------------------------------------------ backward
    def backward(self, params, c, pos, sel, alpha, lambda_conf, image_path=True, out=None):
        """pos / sel: uint8 [n_all] (label > iou_thresh[1]; the balanced sample).  Returns {name: gradient}; `out`
        ({name: tensor}) receives the gradients in place (flat gradient buffer of Stage3Optimizer)."""
        L = _lib.lib()
        p, B = params, self._buf
        n, g, P, n_img, R = c["n"], c["g"], c["P"], c["n_img"], c["n_all"]
        grads = {}

        def G(name):
            t = out[name] if out is not None and name in out else torch.empty_like(p[name])
            grads[name] = t
            return t

        e, h, q = "ensemble_head.", "refinement_head.", "radar_cnn_layers."
        ni = max(n_img, 1)
        RB = self._rows
        sb = dict(group=self.group, sync=self.sync_bn)
        d_o, hl, dhp, u = B("d_o", (ni, 2)), B("hl", (ni, 64)), B("dhp", (2 * ni, 32)), B("u", (2 * ni, 2))
        dr2, dz2 = RB("dr2", R, 1), RB("dz2", R, 13)
        check(L.me_stage3_tail_bwd(ptr(c["rc"]), ptr(c["refine"]), ptr(c["cls"]), 13, ptr(c["p"]), ptr(c["img_boxes"]), n_img, R,
                                   ptr(pos), ptr(sel), float(alpha), float(lambda_conf), ptr(p[e + "fc1.0.weight"]),
                                   ptr(p[e + "fc1.0.bias"]), ptr(p[e + "fc2.0.weight"]), ptr(d_o), ptr(hl), ptr(dhp), ptr(u),
                                   ptr(dr2), ptr(dz2), stream_ptr()), "me_stage3_tail_bwd")
        # ensemble head
        if n_img > 0:
            gemm_tn(d_o[:n_img], hl[:n_img], G(e + "fc2.0.weight"))
            colsum(d_o[:n_img], G(e + "fc2.0.bias"))
            gemm_tn(dhp[:2 * n_img], u[:2 * n_img], G(e + "fc1.0.weight"))
            colsum(dhp[:2 * n_img], G(e + "fc1.0.bias"))
        else:
            for k in ("fc2.0.weight", "fc2.0.bias", "fc1.0.weight", "fc1.0.bias"):
                G(e + k).zero_()
        # radar_net
        w3 = p[h + "radar_net.3.weight"].view(1, 10)
        gemm_tn(dr2, c["ar"], G(h + "radar_net.3.weight").view(1, 10))
        colsum(dr2, G(h + "radar_net.3.bias"))
        dar = gemm_nn(dr2, w3, RB("dar", R, 10))
        dr1 = RB("dr1", R, 10)
        bn_train_bwd(dar, c["ar"], c["xhr"], p[h + "radar_net.1.weight"], c["istdr"], c["sumsr"], G(h + "radar_net.1.weight"),
                     G(h + "radar_net.1.bias"), dr1, B("bnsr", (20,)), **sb)
        wr = p[h + "radar_net.0.weight"].view(10, 490)
        gemm_tn(dr1, c["x_rad"], G(h + "radar_net.0.weight").view(10, 490))
        colsum(dr1, G(h + "radar_net.0.bias"))
        dx_rad = gemm_nn(dr1, wr, RB("dx_rad", R, 490))
        # RoIAlign adjoint -> radar score map -> radar_cnn_layers
        ds = B("ds", (P, 10), zero=True)
        roi_align_f32(False, None, n, g, g, 10, c["rois"], R, grad_out=dx_rad, dfeat=ds)
        check(L.me_sigmoid_bwd_f32(ptr(ds), ptr(c["s"]), P * 10, stream_ptr()), "me_sigmoid_bwd_f32")
        w4 = p[q + "conv3.3.weight"].view(10, 128)
        gemm_tn(ds, c["a3"], G(q + "conv3.3.weight").view(10, 128))
        colsum(ds, G(q + "conv3.3.bias"))
        da = gemm_nn(ds, w4, B("da3", (P, 128)))
        for i, name, cin, cout in ((3, "conv3", 64, 128), (2, "conv2", 32, 64), (1, "conv1", 3, 32)):
            dz = B(f"dz{i}", (P, cout))
            bn_train_bwd(da, c[f"a{i}"], c[f"xh{i}"], p[f"{q}{name}.1.weight"], c[f"istd{i}"], c[f"sums{i}"],
                         G(f"{q}{name}.1.weight"), G(f"{q}{name}.1.bias"), dz, B(f"bns{i}", (2 * cout,)), **sb)
            w = p[f"{q}{name}.0.weight"].view(cout, cin * 9)
            gemm_tn(dz, c[f"cols{i}"], G(f"{q}{name}.0.weight").view(cout, cin * 9))
            colsum(dz, G(f"{q}{name}.0.bias"))
            if i > 1:
                dcols = gemm_nn(dz, w, B(f"dcols{i}", (P, cin * 9)))
                da = B(f"da{i - 1}", (P, cin))
                check(L.me_col2im3_f32(ptr(dcols), n, g, g, cin, ptr(da), stream_ptr()), "me_col2im3_f32")
        if not image_path:
            return grads
        # image path: net2 -> net0 -> PS-RoIAlign adjoint -> BN -> 1x1 conv
        gemm_tn(dz2, c["t"], G(h + "net2.0.weight"))
        colsum(dz2, G(h + "net2.0.bias"))
        dt = gemm_nn(dz2, p[h + "net2.0.weight"], RB("dt", R, 256))
        check(L.me_leaky_bwd_f32(ptr(dt), ptr(c["t"]), R * 256, stream_ptr()), "me_leaky_bwd_f32")
        gemm_tn(dt, c["x_img"], G(h + "net0.0.weight"))
        colsum(dt, G(h + "net0.0.bias"))
        dx_img = gemm_nn(dt, p[h + "net0.0.weight"], RB("dx_img", R, 490))
        dai = B("dai", (P, 490), zero=True)
        roi_align_f32(True, None, n, g, g, 10, c["rois"], R, grad_out=dx_img, dfeat=dai)
        dzi = B("dzi", (P, 490))
        i_ = "img_cnn_layers.net."
        bn_train_bwd(dai, c["ai"], c["xhi"], p[i_ + "batch_norm_0.weight"], c["istdi"], c["sumsi"],
                     G(i_ + "batch_norm_0.weight"), G(i_ + "batch_norm_0.bias"), dzi, B("bnsi", (980,)), **sb)
        gemm_tn(dzi, c["feat"], G(i_ + "conv_0.weight").view(490, 256))
        colsum(dzi, G(i_ + "conv_0.bias"))
        return grads


class Stage3Optimizer:
    """Adam over the trainable head parameters (train.py:158: lr 5e-4) with everything flat: parameter storage,
    gradients and the two moment buffers are single fp32 tensors, the module's nn.Parameters are re-pointed at views of
    the flat parameter buffer, so a step is ONE all-reduce of the gradient buffer over NCCL (sharded batch; the losses
    are sums over proposals, so summed shard gradients are the whole-batch gradient) and ONE Adam kernel."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, names=None):
        named = dict(model.named_parameters())
        self.names = [k for k in (names or TRAINABLE) if k in named and named[k].requires_grad]
        dev = named[self.names[0]].device
        sizes = [named[k].numel() for k in self.names]
        total = sum(sizes)
        self.flat = torch.empty((total,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.grads = {}
        off = 0
        for k, nelem in zip(self.names, sizes):
            prm = named[k]
            view = self.flat[off:off + nelem].view_as(prm)
            view.copy_(prm.data)
            prm.data = view                                  # the module now lives in the flat buffer
            self.grads[k] = self.grad[off:off + nelem].view_as(prm)
            prm.grad = self.grads[k]
            off += nelem
        self.lr, self.betas, self.eps, self.steps = lr, betas, eps, 0
        self.numel = total

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce(self, group=None):
        _all_reduce(self.grad, group)

    def step(self):
        self.steps += 1
        check(_lib.lib().me_adam_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
                                      float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), self.steps,
                                      stream_ptr()), "me_adam_step")


def half_rows_to_float(x_view_tensor, rows, cols, pitch, out):
    check(_lib.lib().me_half_rows_to_float(ptr(x_view_tensor), rows, cols, pitch, ptr(out), stream_ptr()), "me_half_rows_to_float")
    return out


def nchw_to_rows(x, out):
    n, c, h, w = x.shape
    check(_lib.lib().me_nchw_to_rows_f32(ptr(x), n, c, h * w, ptr(out), stream_ptr()), "me_nchw_to_rows_f32")
    return out


_ = ctypes  # (kept for callers that build pointer arguments by hand)
