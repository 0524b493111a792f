"""Losses that consume the renderer's outputs on Self6D++'s self-supervised path, on B200 kernels.

``weighted_ex_loss_probs`` is the drop-in for ``core.self6dpp.losses.mask_losses.weighted_ex_loss_probs``
(/root/reference/core/self6dpp/losses/mask_losses.py:63-108), which ``compute_self_loss_pose`` applies to the rendered
soft mask (core/self6dpp/engine/self_engine_utils.py:541-545): same arguments, same scalar.  The reference builds
boolean-indexed temporaries and synchronises with the host four times per call; here it is one reduction launch and one
elementwise launch for the backward (``dibr_mask_loss_forward`` / ``_backward``), bit-reproducible, no host sync.  The
reference's NaN diagnostics (prints) are not reproduced.  No CPU fallback.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .rasterizer import _require_cuda_f32, _stream

_SCRATCH = {}


def _scratch(n, device):
    """partial sums + the ticket word (zero between calls: the kernel re-arms it), one buffer per (device, size)"""
    floats = _lib.load().dibr_mask_loss_scratch_floats(n)
    key = (str(device), floats)
    buf = _SCRATCH.get(key)
    if buf is None:
        if len(_SCRATCH) > 16:
            _SCRATCH.clear()
        buf = torch.zeros(floats, dtype=torch.float32, device=device)
        _SCRATCH[key] = buf
    return buf


class _WeightedExLossProbs(Function):
    @staticmethod
    def forward(ctx, probs, target, weight):
        _require_cuda_f32("probs", probs)
        _require_cuda_f32("target", target)
        assert probs.size() == target.size()
        device = probs.device
        p_c, t_c = probs.detach().contiguous(), target.detach().contiguous()
        w_c = None
        if weight is not None:
            _require_cuda_f32("weight", weight)
            w_c = weight.detach().expand_as(p_c).contiguous()
        out = torch.empty(3, dtype=torch.float32, device=device)
        q = _lib.DibrMaskLoss()
        q.n = p_c.numel()
        q.probs, q.target, q.weight = _lib.ptr(p_c), _lib.ptr(t_c), _lib.ptr(w_c)
        scratch = _scratch(q.n, device)
        q.scratch, q.out = _lib.ptr(scratch), _lib.ptr(out)
        with torch.cuda.device(device):
            _lib.check(_lib.load().dibr_mask_loss_forward(ctypes.byref(q), _stream(device)), "dibr_mask_loss_forward")
        ctx.save_for_backward(p_c, t_c, out, *([w_c] if w_c is not None else []))
        ctx.has_w = w_c is not None
        ctx.shape = probs.shape
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        p_c, t_c, out = saved[:3]
        w_c = saved[3] if ctx.has_w else None
        device = p_c.device
        go = grad_out.detach().reshape(1).to(torch.float32).contiguous()
        gp = torch.empty_like(p_c)
        q = _lib.DibrMaskLoss()
        q.n = p_c.numel()
        q.probs, q.target, q.weight = _lib.ptr(p_c), _lib.ptr(t_c), _lib.ptr(w_c)
        q.out, q.grad_out, q.grad_probs = _lib.ptr(out), _lib.ptr(go), _lib.ptr(gp)
        with torch.cuda.device(device):
            _lib.check(_lib.load().dibr_mask_loss_backward(ctypes.byref(q), _stream(device)), "dibr_mask_loss_backward")
        return gp.view(ctx.shape), None, None


def weighted_ex_loss_probs(probs, target, weight=None):
    """mask_losses.py:63-108: loss = mean over target>0 of -target*log(p)*w  +  mean over target==0 of -log(1-p)*w"""
    return _WeightedExLossProbs.apply(probs, target, weight)
