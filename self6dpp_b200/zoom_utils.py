"""Crop & resize of rendered images on a B200 kernel: the drop-in for ``core/utils/zoom_utils.py``.

``batch_crop_resize(x, rois, out_H, out_W, aligned=True, interpolation="bilinear")`` (zoom_utils.py:80-95) is what
``compute_self_loss_pose`` applies to the full-frame render before the photometric / normal losses
(self_engine_utils.py:528-533, 662-666, 690-692); ``deepim_boxes`` (zoom_utils.py:6-77) builds the boxes for the refiner.
The reference goes through detectron2's ``ROIAlign`` (a wrapper over ``torchvision.ops.roi_align``), which wants a
contiguous BCHW input -- the rendered image arrives as a permuted BHWC view, so it is copied first -- and scatters the
gradient with fp32 atomics.  Here ``x`` is read through its strides (no copy for any dense layout), and the backward is a
fixed-order gather that writes the dense gradient once, in ``x``'s own layout: bit-reproducible.  No CPU fallback.

``interpolation="nearest"`` is torchvision's ``RoIPool`` (maximum over quantised bins; no call site in the reference's
``core/`` uses it): its own pair of kernels, gradient by a fixed-order gather.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .rasterizer import _require_cuda_f32, _stream


def _dense_strides(x):
    """x's own strides when its elements tile a block of memory without gaps or overlap (any permutation), else None"""
    if x.numel() == 0:
        return None
    dims = sorted(range(x.dim()), key=lambda d: (x.stride(d), x.size(d)))
    expect = 1
    for d in dims:
        if x.size(d) == 1:
            continue
        if x.stride(d) != expect:
            return None
        expect *= x.size(d)
    return tuple(x.stride())


def _fill(q, x_like, rois, out_h, out_w, spatial_scale, sampling_ratio, aligned):
    n, c, h, w = x_like.shape
    q.num_rois, q.num_images, q.channels, q.height, q.width = rois.shape[0], n, c, h, w
    q.pooled_h, q.pooled_w, q.sampling_ratio, q.aligned = int(out_h), int(out_w), int(sampling_ratio), int(bool(aligned))
    q.spatial_scale = float(spatial_scale)
    q.stride_n, q.stride_c, q.stride_h, q.stride_w = (int(s) for s in x_like.stride())
    q.rois = _lib.ptr(rois)


class _RoiAlign(Function):
    @staticmethod
    def forward(ctx, x, rois, out_h, out_w, spatial_scale, sampling_ratio, aligned):
        _require_cuda_f32("x", x)
        _require_cuda_f32("rois", rois)
        if x.dim() != 4:
            raise RuntimeError("batch_crop_resize: x must be BCHW")
        if rois.dim() != 2 or rois.shape[1] != 5:
            raise RuntimeError("batch_crop_resize: rois must be Bx5 (index into x, x1, y1, x2, y2)")
        if rois.device != x.device:
            raise RuntimeError("batch_crop_resize: x and rois must be on the same device")
        xd = x.detach()
        if _dense_strides(xd) is None:
            xd = xd.contiguous()
        r_c = rois.detach().contiguous()
        out = torch.empty(r_c.shape[0], x.shape[1], int(out_h), int(out_w), dtype=torch.float32, device=x.device)
        q = _lib.DibrRoiAlign()
        _fill(q, xd, r_c, out_h, out_w, spatial_scale, sampling_ratio, aligned)
        q.input, q.output = _lib.ptr(xd), _lib.ptr(out)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().dibr_roi_align_forward(ctypes.byref(q), _stream(x.device)), "dibr_roi_align_forward")
        ctx.save_for_backward(r_c)
        ctx.meta = (tuple(x.shape), tuple(xd.stride()), int(out_h), int(out_w), float(spatial_scale), int(sampling_ratio), bool(aligned))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (r_c,) = ctx.saved_tensors
        shape, strides, out_h, out_w, spatial_scale, sampling_ratio, aligned = ctx.meta
        g = grad_out.contiguous()
        grad_x = torch.empty_strided(shape, strides, dtype=torch.float32, device=g.device)     # every element is written
        if grad_x.numel() > 0:
            q = _lib.DibrRoiAlign()
            _fill(q, grad_x, r_c, out_h, out_w, spatial_scale, sampling_ratio, aligned)
            q.grad_output, q.grad_input = _lib.ptr(g), _lib.ptr(grad_x)
            with torch.cuda.device(g.device):
                _lib.check(_lib.load().dibr_roi_align_backward(ctypes.byref(q), _stream(g.device)), "dibr_roi_align_backward")
        return grad_x, None, None, None, None, None, None


class _RoiPool(Function):
    @staticmethod
    def forward(ctx, x, rois, out_h, out_w, spatial_scale):
        _require_cuda_f32("x", x)
        _require_cuda_f32("rois", rois)
        if x.dim() != 4:
            raise RuntimeError("batch_crop_resize: x must be BCHW")
        if rois.dim() != 2 or rois.shape[1] != 5:
            raise RuntimeError("batch_crop_resize: rois must be Bx5 (index into x, x1, y1, x2, y2)")
        if rois.device != x.device:
            raise RuntimeError("batch_crop_resize: x and rois must be on the same device")
        xd = x.detach()
        if _dense_strides(xd) is None:
            xd = xd.contiguous()
        r_c = rois.detach().contiguous()
        out = torch.empty(r_c.shape[0], x.shape[1], int(out_h), int(out_w), dtype=torch.float32, device=x.device)
        arg = torch.empty(out.shape, dtype=torch.int32, device=x.device)
        q = _lib.DibrRoiPool()
        _fill_pool(q, xd, r_c, out_h, out_w, spatial_scale)
        q.input, q.output, q.argmax = _lib.ptr(xd), _lib.ptr(out), _lib.ptr(arg)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().dibr_roi_pool_forward(ctypes.byref(q), _stream(x.device)), "dibr_roi_pool_forward")
        ctx.save_for_backward(r_c, arg)
        ctx.meta = (tuple(x.shape), tuple(xd.stride()), int(out_h), int(out_w), float(spatial_scale))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        r_c, arg = ctx.saved_tensors
        shape, strides, out_h, out_w, spatial_scale = ctx.meta
        g = grad_out.contiguous()
        grad_x = torch.empty_strided(shape, strides, dtype=torch.float32, device=g.device)     # every element is written
        if grad_x.numel() > 0:
            q = _lib.DibrRoiPool()
            _fill_pool(q, grad_x, r_c, out_h, out_w, spatial_scale)
            q.argmax, q.grad_output, q.grad_input = _lib.ptr(arg), _lib.ptr(g), _lib.ptr(grad_x)
            with torch.cuda.device(g.device):
                _lib.check(_lib.load().dibr_roi_pool_backward(ctypes.byref(q), _stream(g.device)), "dibr_roi_pool_backward")
        return grad_x, None, None, None, None


def _fill_pool(q, x_like, rois, out_h, out_w, spatial_scale):
    n, c, h, w = x_like.shape
    q.num_rois, q.num_images, q.channels, q.height, q.width = rois.shape[0], n, c, h, w
    q.pooled_h, q.pooled_w, q.spatial_scale = int(out_h), int(out_w), float(spatial_scale)
    q.stride_n, q.stride_c, q.stride_h, q.stride_w = (int(s) for s in x_like.stride())
    q.rois = _lib.ptr(rois)


def roi_pool(x, rois, output_size, spatial_scale=1.0):
    """``torchvision.ops.RoIPool(output_size, spatial_scale)(x, rois)``."""
    out_h, out_w = (output_size, output_size) if isinstance(output_size, int) else output_size
    return _RoiPool.apply(x, rois, out_h, out_w, spatial_scale)


def roi_align(x, rois, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True):
    """``ROIAlign(output_size, spatial_scale, sampling_ratio, aligned)(x, rois)`` (detectron2.layers.roi_align)."""
    out_h, out_w = (output_size, output_size) if isinstance(output_size, int) else output_size
    return _RoiAlign.apply(x, rois, out_h, out_w, spatial_scale, sampling_ratio, aligned)


def batch_crop_resize(x, rois, out_H, out_W, aligned=True, interpolation="bilinear"):
    """
    Args:
        x: BCHW (any dense layout: a permuted BHWC render is read in place)
        rois: Bx5, rois[:, 0] is the idx into x
        out_H (int):
        out_W (int):
    Returns: len(rois) x C x out_H x out_W   (zoom_utils.py:80-95)
    """
    if interpolation == "bilinear":
        return roi_align(x, rois, (out_H, out_W), 1.0, 0, aligned)
    if interpolation == "nearest":
        return roi_pool(x, rois, (out_H, out_W), 1.0)              # RoIPool(output_size, 1.0), zoom_utils.py:91-92
    raise ValueError(f"Wrong interpolation type: {interpolation}")


def deepim_boxes(ren_boxes, ren_centers_2d, obs_boxes=None, lamb=1.4, imHW=(480, 640), outHW=(480, 640), clamp=False):
    """zoom_utils.py:6-77: the square-ish crop around the rendered centre that holds the rendered (and observed) box,
    enlarged by ``lamb`` with the aspect ratio of ``outHW``.  Returns (crop_boxes Nx4, resize_ratios Nx2 = (w, h))."""
    cx, cy = ren_centers_2d[:, 0], ren_centers_2d[:, 1]
    boxes = [ren_boxes] if obs_boxes is None else [obs_boxes, ren_boxes]
    xdist = torch.stack([(cx - b[:, 0]).abs() for b in boxes] + [(b[:, 2] - cx).abs() for b in boxes], dim=1).max(dim=1)[0]
    ydist = torch.stack([(cy - b[:, 1]).abs() for b in boxes] + [(b[:, 3] - cy).abs() for b in boxes], dim=1).max(dim=1)[0]
    outH, outW = outHW
    aspect_ratio = outW / outH
    crop_h = torch.max(xdist / aspect_ratio, ydist).clamp(min=1) * 2 * lamb
    crop_w = crop_h * aspect_ratio
    crop_boxes = torch.stack([cx - crop_w / 2, cy - crop_h / 2, cx + crop_w / 2, cy + crop_h / 2], dim=1)
    assert not clamp                                              # the reference asserts the same
    resize_ratios = torch.stack([outW / crop_w, outH / crop_h], dim=1)
    return crop_boxes, resize_ratios
