"""TEST INFRASTRUCTURE ONLY (tests/ and nothing else may import this): float64 numpy restatement of ROIAlign as the
reference uses it -- ``batch_crop_resize`` (core/utils/zoom_utils.py:80-95) = detectron2 ``ROIAlign(output_size, 1.0, 0,
aligned)``, which is ``torchvision.ops.roi_align``.  detectron2 is NOT under /root/reference (un-vendored, unpinned:
``scripts/install_deps.sh`` installs it from source); the algorithm restated here is torchvision's published one
(``roi_align_kernel.cpp``: pre-computed bilinear taps, mean over ceil(roi/out) samples per bin and axis), and it is pinned
by golden vectors generated with torchvision 0.26's CPU op in the build container (tests/golden/make_golden_roialign.py ->
tests/golden/ref_roialign.npz), forward and backward.  Pure loops over samples, vectorised over channels: small cases only.
"""
import math

import numpy as np


def _taps(v, size):
    """one axis of bilinear_interpolate: -> (lo, hi, w_lo, w_hi) or None when the sample is outside [-1, size]"""
    if v < -1.0 or v > size:
        return None
    if v <= 0:
        v = 0.0
    lo = int(v)
    if lo >= size - 1:
        hi = lo = size - 1
        v = float(lo)
    else:
        hi = lo + 1
    l = v - lo
    return lo, hi, 1.0 - l, l


def _geom(roi, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned):
    f = np.float32                                     # the roi arithmetic is float32 in the op; samples follow in float64
    off = f(0.5) if aligned else f(0.0)
    sw, sh = f(roi[1]) * f(spatial_scale) - off, f(roi[2]) * f(spatial_scale) - off
    ew, eh = f(roi[3]) * f(spatial_scale) - off, f(roi[4]) * f(spatial_scale) - off
    rw, rh = f(ew - sw), f(eh - sh)
    if not aligned:
        rw, rh = max(rw, f(1.0)), max(rh, f(1.0))
    bh, bw = f(rh / f(pooled_h)), f(rw / f(pooled_w))
    gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(float(f(rh / f(pooled_h)))))
    gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(float(f(rw / f(pooled_w)))))
    return int(roi[0]), sw, sh, bw, bh, gw, gh


def _coord(start, p, b, i, grid):
    f = np.float32
    return float(f(f(start + f(f(p) * b)) + f(f(f(i + 0.5) * b) / f(grid))))


def roi_align(x, rois, out_h, out_w, spatial_scale=1.0, sampling_ratio=0, aligned=True, grad_out=None):
    """x [N,C,H,W], rois [R,5] -> out [R,C,out_h,out_w] (float64); with grad_out also d/dx of sum(out * grad_out)."""
    x = np.asarray(x, dtype=np.float64)
    N, C, H, W = x.shape
    R = len(rois)
    out = np.zeros((R, C, out_h, out_w), dtype=np.float64)
    gx = np.zeros_like(x) if grad_out is not None else None
    for r in range(R):
        n, sw, sh, bw, bh, gw, gh = _geom(rois[r], out_h, out_w, spatial_scale, sampling_ratio, aligned)
        count = max(gh * gw, 1)
        for ph in range(out_h):
            ty = [_taps(_coord(sh, ph, bh, iy, gh), H) for iy in range(gh)]
            for pw in range(out_w):
                for iy in range(gh):
                    if ty[iy] is None:
                        continue
                    y0, y1, hy, ly = ty[iy]
                    for ix in range(gw):
                        tx = _taps(_coord(sw, pw, bw, ix, gw), W)
                        if tx is None:
                            continue
                        x0, x1, hx, lx = tx
                        out[r, :, ph, pw] += (hy * hx * x[n, :, y0, x0] + hy * lx * x[n, :, y0, x1] +
                                              ly * hx * x[n, :, y1, x0] + ly * lx * x[n, :, y1, x1])
                        if gx is not None:
                            g = np.asarray(grad_out[r, :, ph, pw], dtype=np.float64) / count
                            gx[n, :, y0, x0] += hy * hx * g
                            gx[n, :, y0, x1] += hy * lx * g
                            gx[n, :, y1, x0] += ly * hx * g
                            gx[n, :, y1, x1] += ly * lx * g
                out[r, :, ph, pw] /= count
    return (out, gx) if grad_out is not None else out
