"""MS-SSIM on B200 kernels: the drop-in for ``core.self6dpp.losses.ssim.MS_SSIM``
(/root/reference/core/self6dpp/losses/ssim.py:188-246), which the self-supervised loop builds once
(core/self6dpp/engine/self_engine.py:352, ``MS_SSIM(data_range=1.0, normalize=True)``) and applies to the real and the
rendered crop (self_engine_utils.py:777-785).

Same constructor arguments, same ``forward(X, Y) -> (N,)``.  What is different underneath: per pyramid level ONE kernel
computes the five Gaussian moments in shared memory and reduces cs / ssim (the reference: 10 depthwise cuDNN convolutions,
TF32 by default, and ~25 elementwise launches with five full-size temporaries); the backward is one transposed-filter
kernel per level, no atomics, bit-reproducible.  Restrictions, each raised loudly: ``window_size == 11``,
``use_padding=False`` (both the reference's defaults and what the self-supervised path uses); the gradient flows to ``Y``
only (``X`` is the real image on that path).  No CPU fallback.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .rasterizer import _require_cuda_f32, _stream


def create_window(window_size, sigma):
    """ssim.py:13-30 without the channel repeat: the 1-D Gaussian, float32, normalised to sum 1."""
    coords = torch.arange(window_size, dtype=torch.float)
    coords -= window_size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g


def _fill(q, X, Y, window, weights, data_range, normalize, want_grad, use_padding=False):
    n, c, h, w = Y.shape
    q.n_img, q.channels, q.height, q.width = n, c, h, w
    q.levels, q.normalize, q.want_grad = len(weights), int(bool(normalize)), int(bool(want_grad))
    q.use_padding = int(bool(use_padding))
    q.data_range = float(data_range)
    for i, v in enumerate(window):
        q.window[i] = v
    for i, v in enumerate(weights):
        q.weights[i] = v
    q.x, q.y = _lib.ptr(X), _lib.ptr(Y)


class _MsSsim(Function):
    @staticmethod
    def forward(ctx, X, Y, window, weights, data_range, normalize, use_padding=False):
        _require_cuda_f32("X", X)
        _require_cuda_f32("Y", Y)
        if X.dim() != 4 or X.shape != Y.shape:
            raise ValueError("X and Y must both be (N, C, H, W), got {} and {}".format(tuple(X.shape), tuple(Y.shape)))
        if X.requires_grad:
            raise NotImplementedError("self6dpp_b200 MS_SSIM: no gradient w.r.t. X (the real image); pass the rendered image as Y")
        device = Y.device
        x_c, y_c = X.detach().contiguous(), Y.detach().contiguous()
        want_grad = Y.requires_grad
        lib = _lib.load()
        q = _lib.DibrMsSsim()
        _fill(q, x_c, y_c, window, weights, data_range, normalize, want_grad, use_padding)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.dibr_ms_ssim_workspace_bytes(ctypes.byref(q), ctypes.byref(nbytes)), "dibr_ms_ssim_workspace_bytes")
        ws = torch.empty((nbytes.value + 3) // 4, dtype=torch.float32, device=device)
        out = torch.empty(Y.shape[0], dtype=torch.float32, device=device)
        q.workspace, q.workspace_bytes, q.out = ws.data_ptr(), ws.numel() * 4, _lib.ptr(out)
        with torch.cuda.device(device):
            _lib.check(lib.dibr_ms_ssim_forward(ctypes.byref(q), _stream(device)), "dibr_ms_ssim_forward")
        if want_grad:
            ctx.save_for_backward(x_c, y_c, ws)
            ctx.cfg = (window, weights, data_range, normalize, use_padding)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x_c, y_c, ws = ctx.saved_tensors
        window, weights, data_range, normalize, use_padding = ctx.cfg
        device = y_c.device
        go = grad_out.detach().to(torch.float32).contiguous()
        gy = torch.empty_like(y_c)
        q = _lib.DibrMsSsim()
        _fill(q, x_c, y_c, window, weights, data_range, normalize, True, use_padding)
        q.workspace, q.workspace_bytes = ws.data_ptr(), ws.numel() * 4
        q.grad_out, q.grad_y = _lib.ptr(go), _lib.ptr(gy)
        with torch.cuda.device(device):
            _lib.check(_lib.load().dibr_ms_ssim_backward(ctypes.byref(q), _stream(device)), "dibr_ms_ssim_backward")
        return None, gy, None, None, None, None, None


class MS_SSIM(torch.nn.Module):
    """ssim.py:188-246.  ``channel`` is accepted for signature compatibility; the kernels take it from the input."""

    def __init__(self, window_size=11, window_sigma=1.5, data_range=255.0, channel=3, use_padding=False, weights=None,
                 levels=None, normalize=False):
        super().__init__()
        assert window_size % 2 == 1, "Window size must be odd."
        if window_size != 11:
            raise NotImplementedError("self6dpp_b200 MS_SSIM: window_size must be 11 (the reference's default)")
        self.data_range = data_range
        self.use_padding = use_padding
        self.normalize = normalize
        self.register_buffer("window", create_window(window_size, window_sigma))
        if weights is None:
            weights = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
        weights = torch.tensor(weights, dtype=torch.float)
        if levels is not None:
            weights = weights[:levels]
            weights = weights / weights.sum()
        self.register_buffer("weights", weights)
        self._window = tuple(float(v) for v in self.window)          # host copies: no sync per call
        self._weights = tuple(float(v) for v in weights)

    def forward(self, X, Y):
        return _MsSsim.apply(X, Y, self._window, self._weights, self.data_range, self.normalize, self.use_padding)
