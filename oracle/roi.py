"""Oracle restatement of torchvision's ps_roi_align / roi_align CPU kernels in numpy (float32).

Test infrastructure only (see oracle/__init__.py).  Third-party arithmetic: torchvision 0.26.0
(installed here; the reference pinned 0.6.0, README.md:11), called at my_models.py:495-496 with
output_size (7,7), spatial_scale 1/16, sampling_ratio -1 (adaptive grid) and aligned=False.
Semantics (probed against the installed library):
  bilinear(y, x): 0 if y < -1 or y > H or x < -1 or x > W; clamp to >= 0; low = int(.);
                  if low >= dim-1: low = high = dim-1 and the coordinate snaps to it.
  ps_roi_align  : start = coord*scale - 0.5; size = end - start (NOT clamped); bin = size/P;
                  grid = ceil(size/P); out[k, c, ph, pw] averages channel (c*P + ph)*P + pw.
  roi_align     : start = coord*scale; size = max(end - start, 1); count = max(gh*gw, 1).
"""
import numpy as np

F32 = np.float32


def _bilinear_planes(feat, ys, xs):
    """feat: (C, H, W); ys: (ny,), xs: (nx,) sample coordinates -> (C, ny, nx) values."""
    C, H, W = feat.shape
    ys = ys.astype(F32).copy()
    xs = xs.astype(F32).copy()
    oob_y = (ys < F32(-1.0)) | (ys > F32(H))
    oob_x = (xs < F32(-1.0)) | (xs > F32(W))
    ys = np.maximum(ys, F32(0))
    xs = np.maximum(xs, F32(0))
    y_low = ys.astype(np.int64)
    x_low = xs.astype(np.int64)
    y_edge = y_low >= H - 1
    x_edge = x_low >= W - 1
    y_low = np.where(y_edge, H - 1, y_low)
    x_low = np.where(x_edge, W - 1, x_low)
    y_high = np.where(y_edge, H - 1, y_low + 1)
    x_high = np.where(x_edge, W - 1, x_low + 1)
    ys = np.where(y_edge, y_low.astype(F32), ys)
    xs = np.where(x_edge, x_low.astype(F32), xs)
    ly = (ys - y_low.astype(F32)).astype(F32)
    lx = (xs - x_low.astype(F32)).astype(F32)
    hy = F32(1) - ly
    hx = F32(1) - lx
    v1 = feat[:, y_low][:, :, x_low]
    v2 = feat[:, y_low][:, :, x_high]
    v3 = feat[:, y_high][:, :, x_low]
    v4 = feat[:, y_high][:, :, x_high]
    w1 = (hy[:, None] * hx[None, :]).astype(F32)
    w2 = (hy[:, None] * lx[None, :]).astype(F32)
    w3 = (ly[:, None] * hx[None, :]).astype(F32)
    w4 = (ly[:, None] * lx[None, :]).astype(F32)
    val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4
    val = np.where(oob_y[None, :, None] | oob_x[None, None, :], F32(0), val)
    return val.astype(F32)


def _roi_common(feat_n, roi, scale, pooled, position_sensitive, out_channels):
    off = F32(0.5) if position_sensitive else F32(0)
    sw = F32(roi[1]) * F32(scale) - off
    sh = F32(roi[2]) * F32(scale) - off
    ew = F32(roi[3]) * F32(scale) - off
    eh = F32(roi[4]) * F32(scale) - off
    rw, rh = F32(ew - sw), F32(eh - sh)
    if not position_sensitive:
        rw, rh = max(rw, F32(1)), max(rh, F32(1))
    bin_h, bin_w = F32(rh / F32(pooled)), F32(rw / F32(pooled))
    gh, gw = int(np.ceil(rh / F32(pooled))), int(np.ceil(rw / F32(pooled)))
    out = np.zeros((out_channels, pooled, pooled), dtype=F32)
    if gh <= 0 or gw <= 0:
        if position_sensitive:
            out[:] = np.nan  # torchvision divides 0 by count 0
        return out
    ph = np.arange(pooled, dtype=F32)
    iy = np.arange(gh, dtype=F32)
    ix = np.arange(gw, dtype=F32)
    # y = hstart + (iy + .5) * bin_h / gh with hstart = ph * bin_h + start
    ys = ((ph[:, None] * bin_h + sh) + (iy[None, :] + F32(0.5)) * bin_h / F32(gh)).astype(F32)  # (P, gh)
    xs = ((ph[:, None] * bin_w + sw) + (ix[None, :] + F32(0.5)) * bin_w / F32(gw)).astype(F32)  # (P, gw)
    vals = _bilinear_planes(feat_n, ys.reshape(-1), xs.reshape(-1))  # (C, P*gh, P*gw)
    C = feat_n.shape[0]
    vals = vals.reshape(C, pooled, gh, pooled, gw)
    # accumulate in the kernel's order (iy outer, ix inner) for float32 parity
    acc = np.zeros((C, pooled, pooled), dtype=F32)
    for a in range(gh):
        for b in range(gw):
            acc = (acc + vals[:, :, a, :, b]).astype(F32)
    count = F32(gh * gw) if position_sensitive else F32(max(gh * gw, 1))
    acc = (acc / count).astype(F32)
    if position_sensitive:
        acc = acc.reshape(out_channels, pooled, pooled, pooled, pooled)
        idx = np.arange(pooled)
        out = acc[:, idx[:, None], idx[None, :], idx[:, None], idx[None, :]]  # channel (c,ph,pw) at bin (ph,pw)
    else:
        out = acc
    return out.astype(F32)


def ps_roi_align(feat, rois, pooled=7, spatial_scale=1.0 / 16):
    """feat: (N, C*P*P, H, W) float32; rois: (R, 5) [batch, x1, y1, x2, y2] -> (R, C, P, P)."""
    feat = np.asarray(feat, dtype=F32)
    rois = np.asarray(rois, dtype=F32)
    c_out = feat.shape[1] // (pooled * pooled)
    out = np.zeros((rois.shape[0], c_out, pooled, pooled), dtype=F32)
    for r, roi in enumerate(rois):
        out[r] = _roi_common(feat[int(roi[0])], roi, spatial_scale, pooled, True, c_out)
    return out


def roi_align(feat, rois, pooled=7, spatial_scale=1.0 / 16):
    """feat: (N, C, H, W) float32; rois: (R, 5) -> (R, C, P, P); aligned=False, sampling_ratio=-1."""
    feat = np.asarray(feat, dtype=F32)
    rois = np.asarray(rois, dtype=F32)
    out = np.zeros((rois.shape[0], feat.shape[1], pooled, pooled), dtype=F32)
    for r, roi in enumerate(rois):
        out[r] = _roi_common(feat[int(roi[0])], roi, spatial_scale, pooled, False, feat.shape[1])
    return out
