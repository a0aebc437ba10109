"""Hand-derived backward pass of the stage-3 trainable set, op by op - the arithmetic the CUDA kernels of the next round
have to implement.  Test infrastructure only (see oracle/__init__.py); no autograd is used here, and the result is
checked against oracle/stage3_train.py (torch.autograd, itself pinned bit for bit by the reference's gradients).

Scope: every parameter that receives a gradient.  With --pretrained_module2 train.py freezes the R-CNN part
(img_cnn_layers and refinement_head.net0-2, reference module3_our_dataset/train.py:117-149) and only the radar / fusion
set below is updated; without it ("train module2 + module3 from scratch", :111-115) the image path trains too
(forward_train(..., feat=...) / backward(..., image_path=True): 1x1 conv + BatchNorm + LeakyReLU, the PS-RoIAlign
adjoint, FC 490->256, FC 256->13 + sigmoid).  Radar / fusion set: radar_cnn_layers (three 3x3 conv + BatchNorm + LeakyReLU, one
1x1 conv + sigmoid; my_models.py:130-157), refinement_head.radar_net (7x7 conv on the 7x7 crop = a 490->10 linear map,
BatchNorm over the proposals, LeakyReLU, 1x1 conv, sigmoid; :246-251, 268-276) and ensemble_head (:176-210), for the loss
of :610-635 (FocalLoss on the image proposals' masks + confidence BCE / 6).  All BatchNorms run on batch statistics.

    cache = forward_train(sd, feat..., )   # every intermediate the backward needs
    grads = backward(cache)                # {parameter name: gradient}
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5
SLOPE = 0.1


# ------------------------------------------------------------------------------------------------ building blocks
def bn_forward(z, gamma, beta, dims):
    """BatchNorm in training mode over `dims` (biased variance, eps 1e-5).  Returns y, x_hat, 1/std."""
    mu = z.mean(dims, keepdim=True)
    var = z.var(dims, unbiased=False, keepdim=True)
    inv_std = 1.0 / torch.sqrt(var + EPS)
    x_hat = (z - mu) * inv_std
    shape = [1] * z.dim()
    shape[1] = -1
    return gamma.view(shape) * x_hat + beta.view(shape), x_hat, inv_std


def bn_backward(dy, x_hat, inv_std, gamma, dims):
    """dgamma = sum(dy * x_hat), dbeta = sum(dy), dz = (gamma / std) * (dy - mean(dy) - x_hat * mean(dy * x_hat))."""
    shape = [1] * dy.dim()
    shape[1] = -1
    dgamma = (dy * x_hat).sum(dims)
    dbeta = dy.sum(dims)
    dxh = dy * gamma.view(shape)
    dz = inv_std * (dxh - dxh.mean(dims, keepdim=True) - x_hat * (dxh * x_hat).mean(dims, keepdim=True))
    return dz, dgamma, dbeta


def leaky_backward(dy, pre):
    return dy * torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, SLOPE))


def conv3x3_backward(dz, a_prev, weight, need_input_grad=True):
    """3x3 / stride 1 / pad 1: dW[o, k] = sum_{n,p} dz[n, o, p] * cols[n, k, p] (im2col), db = sum dz,
    da_prev = col2im(W^T dz)."""
    n, cout, h, w = dz.shape
    cols = F.unfold(a_prev, 3, padding=1)                         # (N, Cin*9, H*W)
    dzf = dz.flatten(2)                                           # (N, Cout, H*W)
    dw = torch.einsum("nop,nkp->ok", dzf, cols).view_as(weight)
    db = dz.sum((0, 2, 3))
    da = None
    if need_input_grad:
        da = F.fold(weight.view(cout, -1).t() @ dzf, (h, w), 3, padding=1)
    return dw, db, da


def roi_align_backward(grad_out, rois, n, channels, height, width, scale, pooled=7):
    """Adjoint of torchvision roi_align (aligned=False, sampling_ratio=-1): every sample spreads grad / count over its
    four bilinear neighbours; samples outside [-1, H] x [-1, W] contribute nothing.  grad_out (R, C, P, P) numpy."""
    g = np.zeros((n, channels, height, width), dtype=np.float64)
    f32 = np.float32
    for r in range(len(rois)):
        b = int(rois[r, 0])
        sw, sh = f32(rois[r, 1]) * f32(scale), f32(rois[r, 2]) * f32(scale)
        ew, eh = f32(rois[r, 3]) * f32(scale), f32(rois[r, 4]) * f32(scale)
        rw, rh = max(f32(ew - sw), f32(1)), max(f32(eh - sh), f32(1))
        bin_h, bin_w = f32(rh / f32(pooled)), f32(rw / f32(pooled))
        gh, gw = int(np.ceil(rh / f32(pooled))), int(np.ceil(rw / f32(pooled)))
        count = max(gh * gw, 1)
        ph = np.arange(pooled, dtype=f32)
        ys = ((ph[:, None] * bin_h + sh) + (np.arange(gh, dtype=f32)[None, :] + f32(0.5)) * bin_h / f32(gh)).reshape(-1)
        xs = ((ph[:, None] * bin_w + sw) + (np.arange(gw, dtype=f32)[None, :] + f32(0.5)) * bin_w / f32(gw)).reshape(-1)

        def axis(c, dim):
            oob = (c < -1.0) | (c > dim)
            c = np.maximum(c, 0)
            lo = c.astype(np.int64)
            edge = lo >= dim - 1
            lo = np.where(edge, dim - 1, lo)
            hi = np.where(edge, dim - 1, lo + 1)
            c = np.where(edge, lo.astype(f32), c)
            l = (c - lo).astype(np.float64)
            return lo, hi, 1.0 - l, l, oob

        ylo, yhi, hy, ly, oy = axis(ys, height)
        xlo, xhi, hx, lx, ox = axis(xs, width)
        go = grad_out[r].astype(np.float64) / count                              # (C, P, P)
        gy = np.repeat(go, gh, axis=1)                                           # (C, P*gh, P)
        gs = np.repeat(gy, gw, axis=2)                                           # (C, P*gh, P*gw): grad per sample
        gs = gs * (~oy)[None, :, None] * (~ox)[None, None, :]
        for yi, wy in ((ylo, hy), (yhi, ly)):
            for xi, wx in ((xlo, hx), (xhi, lx)):
                contrib = gs * wy[None, :, None] * wx[None, None, :]
                np.add.at(g[b], (slice(None), yi[:, None], xi[None, :]), contrib)
    return g.astype(np.float32)


def psroi_align_backward(grad_out, rois, n, height, width, scale, pooled=7):
    """Adjoint of torchvision ps_roi_align (sampling_ratio=-1): output (c, ph, pw) only reads feature channel
    (c*P + ph)*P + pw inside bin (ph, pw); start = coord*scale - 0.5, size not clamped, count = gh*gw.
    grad_out (R, C_out, P, P) numpy -> (N, C_out*P*P, H, W)."""
    c_out = grad_out.shape[1]
    g = np.zeros((n, c_out * pooled * pooled, height, width), dtype=np.float64)
    f32 = np.float32
    ch = (np.arange(c_out)[:, None, None] * pooled + np.arange(pooled)[None, :, None]) * pooled + np.arange(pooled)[None, None, :]
    for r in range(len(rois)):
        b = int(rois[r, 0])
        sw, sh = f32(rois[r, 1]) * f32(scale) - f32(0.5), f32(rois[r, 2]) * f32(scale) - f32(0.5)
        ew, eh = f32(rois[r, 3]) * f32(scale) - f32(0.5), f32(rois[r, 4]) * f32(scale) - f32(0.5)
        rw, rh = f32(ew - sw), f32(eh - sh)
        bin_h, bin_w = f32(rh / f32(pooled)), f32(rw / f32(pooled))
        gh, gw = int(np.ceil(rh / f32(pooled))), int(np.ceil(rw / f32(pooled)))
        if gh <= 0 or gw <= 0:
            continue
        count = gh * gw
        for ph in range(pooled):
            ys = (f32(ph) * bin_h + sh) + (np.arange(gh, dtype=f32) + f32(0.5)) * bin_h / f32(gh)
            for pw in range(pooled):
                xs = (f32(pw) * bin_w + sw) + (np.arange(gw, dtype=f32) + f32(0.5)) * bin_w / f32(gw)
                go = grad_out[r, :, ph, pw].astype(np.float64) / count                 # (C_out,)
                chans = ch[:, ph, pw]
                for y in ys:
                    for x in xs:
                        if y < -1.0 or y > height or x < -1.0 or x > width:
                            continue
                        yy, xx = max(y, f32(0)), max(x, f32(0))
                        yl, xl = int(yy), int(xx)
                        if yl >= height - 1:
                            yl = yh = height - 1
                            yy = f32(yl)
                        else:
                            yh = yl + 1
                        if xl >= width - 1:
                            xl = xh = width - 1
                            xx = f32(xl)
                        else:
                            xh = xl + 1
                        ly, lx = float(yy - yl), float(xx - xl)
                        hy, hx = 1.0 - ly, 1.0 - lx
                        g[b, chans, yl, xl] += go * hy * hx
                        g[b, chans, yl, xh] += go * hy * lx
                        g[b, chans, yh, xl] += go * ly * hx
                        g[b, chans, yh, xh] += go * ly * lx
    return g.astype(np.float32)


# ------------------------------------------------------------------------------------------------ forward (train mode)
def forward_train(sd, maps, box_locations, n_img, yolo_vec, cls_img_path, pos, sample_filter, alpha=0.75, lambda_conf=6.0,
                  feat=None):
    """Radar branch + heads in training mode, keeping what the backward needs.
    sd: fp32 tensors by reference key; maps (N,3,g,g); box_locations (R,5) pixels; cls_img_path (R,13): sigmoid output of
    refinement_head.net2 (column 0 enters the confidence) - taken as given (frozen image path) unless `feat`
    (N,256,g,g), the detector's feature map, is passed: then the image path is run here too and its intermediates are
    kept for backward(image_path=True); pos / sample_filter: bool (R,)."""
    from torchvision.ops import ps_roi_align, roi_align
    c = dict(sd=sd, maps=maps, rois=box_locations, n_img=n_img, yolo_vec=yolo_vec, pos=pos, sel=sample_filter, alpha=alpha,
             lam=lambda_conf)
    if feat is not None:
        zi = F.conv2d(feat, sd["img_cnn_layers.net.conv_0.weight"], sd["img_cnn_layers.net.conv_0.bias"])
        yi, xhi, istdi = bn_forward(zi, sd["img_cnn_layers.net.batch_norm_0.weight"], sd["img_cnn_layers.net.batch_norm_0.bias"],
                                    (0, 2, 3))
        ai = F.leaky_relu(yi, SLOPE)
        x_img = ps_roi_align(ai, box_locations, (7, 7), spatial_scale=1. / 16).flatten(1)
        tp = x_img @ sd["refinement_head.net0.0.weight"].t() + sd["refinement_head.net0.0.bias"]
        t = F.leaky_relu(tp, SLOPE)
        cls_img_path = torch.sigmoid(t @ sd["refinement_head.net2.0.weight"].t() + sd["refinement_head.net2.0.bias"])
        c.update(feat=feat, yi=yi, xhi=xhi, istdi=istdi, x_img=x_img, tp=tp, t=t, map_hw=tuple(ai.shape))
    c["cls"] = cls_img_path
    a = maps
    c["a0"] = a
    for i, name in enumerate(("conv1", "conv2", "conv3"), 1):
        z = F.conv2d(a, sd[f"radar_cnn_layers.{name}.0.weight"], sd[f"radar_cnn_layers.{name}.0.bias"], padding=1)
        y, xh, istd = bn_forward(z, sd[f"radar_cnn_layers.{name}.1.weight"], sd[f"radar_cnn_layers.{name}.1.bias"], (0, 2, 3))
        a = F.leaky_relu(y, SLOPE)
        c[f"y{i}"], c[f"xh{i}"], c[f"istd{i}"], c[f"a{i}"] = y, xh, istd, a
    s = torch.sigmoid(F.conv2d(a, sd["radar_cnn_layers.conv3.3.weight"], sd["radar_cnn_layers.conv3.3.bias"]))
    c["s"] = s
    x_rad = roi_align(s, box_locations, (7, 7), spatial_scale=1. / 16)           # forward only; adjoint restated above
    h = "refinement_head.radar_net."
    wr = sd[h + "0.weight"]
    r1 = x_rad.flatten(1) @ wr.view(wr.shape[0], -1).t() + sd[h + "0.bias"]       # 7x7 conv on a 7x7 crop
    yr, xhr, istdr = bn_forward(r1, sd[h + "1.weight"], sd[h + "1.bias"], (0,))
    ar = F.leaky_relu(yr, SLOPE)
    r2 = ar @ sd[h + "3.weight"].view(1, -1).t() + sd[h + "3.bias"]               # (R, 1)
    rc = torch.sigmoid(r2)
    conf = torch.sigmoid(rc + cls_img_path[:, :1])
    ref_vec = torch.cat((conf, cls_img_path[:, 1:2]), -1)
    c.update(x_rad=x_rad, yr=yr, xhr=xhr, istdr=istdr, ar=ar, rc=rc, conf=conf, ref_vec=ref_vec)
    e = "ensemble_head."
    u = torch.stack((ref_vec[:n_img], yolo_vec), -1)                              # (n, 2, 2)
    hp = u @ sd[e + "fc1.0.weight"].t() + sd[e + "fc1.0.bias"]                    # (n, 2, 32)
    hl = F.leaky_relu(hp, SLOPE)
    o = hl.flatten(1) @ sd[e + "fc2.0.weight"].t() + sd[e + "fc2.0.bias"]         # (n, 2)
    p = torch.softmax(o, dim=1)
    c.update(u=u, hp=hp, hl=hl, p=p)
    m = torch.cat((p[:, :1], ref_vec[n_img:, :1]), 0).reshape(-1)
    prob = torch.where(pos[:n_img], m[:n_img], 1 - m[:n_img])
    a_f = torch.where(pos[:n_img], torch.tensor(alpha), torch.tensor(1 - alpha))
    focal = (-a_f * (1 - prob) ** 2 * prob.log())[sample_filter[:n_img]].sum()
    y = pos.float()
    bce = -(y * conf[:, 0].log() + (1 - y) * (1 - conf[:, 0]).log())[sample_filter].sum()
    c["loss"] = float(focal + bce / lambda_conf)
    return c


# ------------------------------------------------------------------------------------------------ backward
def backward(c, image_path=False):
    sd, n_img, pos, sel = c["sd"], c["n_img"], c["pos"], c["sel"]
    grads = {}
    conf = c["conf"][:, 0]
    R = len(conf)
    # loss -> m (image rows) and conf
    p = c["p"]
    m = p[:, 0]
    is_pos = pos[:n_img]
    prob = torch.where(is_pos, m, 1 - m)
    a_f = torch.where(is_pos, torch.tensor(c["alpha"]), torch.tensor(1 - c["alpha"]))
    dprob = -a_f * (-2 * (1 - prob) * prob.log() + (1 - prob) ** 2 / prob)
    dm = torch.where(is_pos, dprob, -dprob) * sel[:n_img].float()
    y = pos.float()
    dconf = (-(y / conf - (1 - y) / (1 - conf)) / c["lam"]) * sel.float()
    # ensemble head
    e = "ensemble_head."
    dp = torch.stack((dm, torch.zeros_like(dm)), 1)
    do = p * (dp - (dp * p).sum(1, keepdim=True))                                   # softmax
    w2 = sd[e + "fc2.0.weight"]
    grads[e + "fc2.0.weight"] = do.t() @ c["hl"].flatten(1)
    grads[e + "fc2.0.bias"] = do.sum(0)
    dhl = (do @ w2).view(-1, 2, 32)
    dhp = leaky_backward(dhl, c["hp"])
    grads[e + "fc1.0.weight"] = torch.einsum("ncj,nck->jk", dhp, c["u"])
    grads[e + "fc1.0.bias"] = dhp.sum((0, 1))
    du = dhp @ sd[e + "fc1.0.weight"]                                               # (n, 2, 2)
    dconf = dconf.clone()
    dconf[:n_img] += du[:, 0, 0]                                                    # u[:, 0, 0] is the refined confidence
    # confidence -> radar_net
    drc = dconf * conf * (1 - conf)
    rc = c["rc"][:, 0]
    dr2 = (drc * rc * (1 - rc)).view(R, 1)
    h = "refinement_head.radar_net."
    grads[h + "3.weight"] = (dr2.t() @ c["ar"]).view_as(sd[h + "3.weight"])
    grads[h + "3.bias"] = dr2.sum(0)
    dar = dr2 @ sd[h + "3.weight"].view(1, -1)
    dyr = leaky_backward(dar, c["yr"])
    dr1, dg, dbeta = bn_backward(dyr, c["xhr"], c["istdr"], sd[h + "1.weight"], (0,))
    grads[h + "1.weight"], grads[h + "1.bias"] = dg, dbeta
    wr = sd[h + "0.weight"]
    grads[h + "0.weight"] = (dr1.t() @ c["x_rad"].flatten(1)).view_as(wr)
    grads[h + "0.bias"] = dr1.sum(0)
    dx_rad = (dr1 @ wr.view(wr.shape[0], -1)).view(R, 10, 7, 7)
    # RoIAlign adjoint -> radar score map -> radar_cnn_layers
    s = c["s"]
    ds = torch.from_numpy(roi_align_backward(dx_rad.numpy(), c["rois"].numpy(), s.shape[0], s.shape[1], s.shape[2], s.shape[3],
                                             1. / 16))
    dz4 = ds * s * (1 - s)
    q = "radar_cnn_layers."
    w4 = sd[q + "conv3.3.weight"]
    grads[q + "conv3.3.weight"] = torch.einsum("nohw,nchw->oc", dz4, c["a3"]).view_as(w4)
    grads[q + "conv3.3.bias"] = dz4.sum((0, 2, 3))
    da = torch.einsum("nohw,oc->nchw", dz4, w4.view(w4.shape[0], -1))
    for i, name in ((3, "conv3"), (2, "conv2"), (1, "conv1")):
        dy = leaky_backward(da, c[f"y{i}"])
        dz, dg, dbeta = bn_backward(dy, c[f"xh{i}"], c[f"istd{i}"], sd[f"{q}{name}.1.weight"], (0, 2, 3))
        grads[f"{q}{name}.1.weight"], grads[f"{q}{name}.1.bias"] = dg, dbeta
        dw, db, da = conv3x3_backward(dz, c[f"a{i - 1}"], sd[f"{q}{name}.0.weight"], need_input_grad=i > 1)
        grads[f"{q}{name}.0.weight"], grads[f"{q}{name}.0.bias"] = dw, db
    if not image_path:
        return grads
    # ---- image path: conf = sigmoid(rc + cls[:, 0]) and, for image proposals, u[:, 1, 0] = cls[:, 1]
    cls = c["cls"]
    dcls = torch.zeros_like(cls)
    dcls[:, 0] = drc                                                                # same pre-activation as rc
    dcls[:n_img, 1] = du[:, 1, 0]
    dz2 = dcls * cls * (1 - cls)
    r = "refinement_head."
    grads[r + "net2.0.weight"] = dz2.t() @ c["t"]
    grads[r + "net2.0.bias"] = dz2.sum(0)
    dtp = leaky_backward(dz2 @ sd[r + "net2.0.weight"], c["tp"])
    grads[r + "net0.0.weight"] = dtp.t() @ c["x_img"]
    grads[r + "net0.0.bias"] = dtp.sum(0)
    dx_img = (dtp @ sd[r + "net0.0.weight"]).view(R, 10, 7, 7)
    n, _, hh, ww = c["map_hw"]
    dai = torch.from_numpy(psroi_align_backward(dx_img.numpy(), c["rois"].numpy(), n, hh, ww, 1. / 16))
    dyi = leaky_backward(dai, c["yi"])
    i = "img_cnn_layers.net."
    dzi, dg, dbeta = bn_backward(dyi, c["xhi"], c["istdi"], sd[i + "batch_norm_0.weight"], (0, 2, 3))
    grads[i + "batch_norm_0.weight"], grads[i + "batch_norm_0.bias"] = dg, dbeta
    grads[i + "conv_0.weight"] = torch.einsum("nohw,nchw->oc", dzi, c["feat"]).view_as(sd[i + "conv_0.weight"])
    grads[i + "conv_0.bias"] = dzi.sum((0, 2, 3))
    return grads
