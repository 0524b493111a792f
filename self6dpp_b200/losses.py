"""Losses that consume the renderer's outputs on Self6D++'s self-supervised path, on B200 kernels.

``weighted_ex_loss_probs`` is the drop-in for ``core.self6dpp.losses.mask_losses.weighted_ex_loss_probs``
(/root/reference/core/self6dpp/losses/mask_losses.py:63-108), which ``compute_self_loss_pose`` applies to the rendered
soft mask (core/self6dpp/engine/self_engine_utils.py:541-545): same arguments, same scalar.  The reference builds
boolean-indexed temporaries and synchronises with the host four times per call; here it is one reduction launch and one
elementwise launch for the backward (``dibr_mask_loss_forward`` / ``_backward``), bit-reproducible, no host sync.  The
reference's NaN diagnostics (prints) are not reproduced.  No CPU fallback.

``lab_l1_loss`` is the Lab-space colour loss of ``compute_self_loss_pose`` (self_engine_utils.py:745-773, over
lib/torch_utils/color/lab.py:16-82): ~60 elementwise torch launches there, one launch per direction here.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .rasterizer import _require_cuda_f32, _stream

_SCRATCH = {}


def _scratch(n, device, kind="mask"):
    """partial sums + the ticket word (zero between calls: the kernel re-arms it), one buffer per (device, size)"""
    lib = _lib.load()
    floats = lib.dibr_mask_loss_scratch_floats(n) if kind == "mask" else lib.dibr_lab_loss_scratch_floats(n)
    key = (str(device), floats, kind)
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


class _SoftDiceLoss(Function):
    @staticmethod
    def forward(ctx, probs, labels, smooth, eps, reduction):
        _require_cuda_f32("probs", probs)
        _require_cuda_f32("labels", labels)
        num = labels.size(0)
        p_c, l_c = probs.detach().reshape(num, -1).contiguous(), labels.detach().reshape(num, -1).contiguous()
        if p_c.shape != l_c.shape:
            raise RuntimeError("soft_dice_loss: probs and labels must hold the same number of elements per sample")
        device = probs.device
        red = {"mean": 0, "sum": 1}.get(reduction, 2)
        out = torch.empty(num if red == 2 else 1, dtype=torch.float32, device=device)
        scratch = torch.zeros(num * 3 + 1, dtype=torch.float32, device=device)          # stats + the ticket word
        q = _lib.DibrDiceLoss()
        q.num, q.reduction, q.per, q.smooth, q.eps = num, red, p_c.shape[1], float(smooth), float(eps)
        q.probs, q.labels, q.stats, q.out = _lib.ptr(p_c), _lib.ptr(l_c), _lib.ptr(scratch), _lib.ptr(out)
        q.ticket = scratch.data_ptr() + 4 * num * 3
        with torch.cuda.device(device):
            _lib.check(_lib.load().dibr_dice_loss_forward(ctypes.byref(q), _stream(device)), "dibr_dice_loss_forward")
        ctx.save_for_backward(p_c, l_c, scratch, out)
        ctx.meta = (num, red, float(smooth), float(eps), probs.shape)
        return out if red == 2 else out[0]

    @staticmethod
    def backward(ctx, grad_out):
        p_c, l_c, scratch, out = ctx.saved_tensors
        num, red, smooth, eps, shape = ctx.meta
        g = grad_out.detach().reshape(-1).contiguous().float()
        grad = torch.empty_like(p_c)
        q = _lib.DibrDiceLoss()
        q.num, q.reduction, q.per, q.smooth, q.eps = num, red, p_c.shape[1], smooth, eps
        q.probs, q.labels, q.stats, q.out = _lib.ptr(p_c), _lib.ptr(l_c), _lib.ptr(scratch), _lib.ptr(out)
        q.grad_out, q.grad_probs = _lib.ptr(g), _lib.ptr(grad)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().dibr_dice_loss_backward(ctypes.byref(q), _stream(g.device)), "dibr_dice_loss_backward")
        return grad.reshape(shape), None, None, None, None


def soft_dice_loss(probs, labels, smooth=0.0, eps=1e-7, reduction="mean"):
    """mask_losses.py:444-463 (the "dice" choice of MASK_INIT_REN_LOSS_TYPE, self_engine_utils.py:546-549; SOLOv2 uses
    eps=0.002): per sample score = 2 (sum p l + smooth) / (sum p + sum l + smooth + eps); "mean": 1 - sum(score) / num,
    "sum": sum(1 - score), anything else: 1 - score per sample.  One reduction launch + one elementwise backward;
    ``labels`` is data (no gradient)."""
    return _SoftDiceLoss.apply(probs, labels, smooth, eps, reduction)


class _NormLoss(Function):
    @staticmethod
    def forward(ctx, out_norm, gt_norm, mask, with_l1, with_cs):
        _require_cuda_f32("out_norm", out_norm)
        _require_cuda_f32("gt_norm", gt_norm)
        _require_cuda_f32("mask", mask)
        b, c, h, w = out_norm.shape
        assert out_norm.shape == gt_norm.shape, "{} != {}".format(out_norm.shape, gt_norm.shape)
        assert c == 3 and mask.shape == (b, 1, h, w), mask.shape
        device = out_norm.device
        o_c, g_c, m_c = out_norm.detach().contiguous(), gt_norm.detach().contiguous(), mask.detach().contiguous()
        lib = _lib.load()
        floats = lib.dibr_norm_loss_scratch_floats(b * h * w)
        key = (str(device), floats, "norm")
        scratch = _SCRATCH.get(key)
        if scratch is None:
            scratch = torch.zeros(floats, dtype=torch.float32, device=device)
            _SCRATCH[key] = scratch
        out = torch.empty(2, dtype=torch.float32, device=device)
        q = _lib.DibrNormLoss()
        q.n_img, q.hw, q.with_l1, q.with_cs = b, h * w, int(bool(with_l1)), int(bool(with_cs))
        q.out_norm, q.gt_norm, q.mask = _lib.ptr(o_c), _lib.ptr(g_c), _lib.ptr(m_c)
        q.scratch, q.out = _lib.ptr(scratch), _lib.ptr(out)
        with torch.cuda.device(device):
            _lib.check(lib.dibr_norm_loss_forward(ctypes.byref(q), _stream(device)), "dibr_norm_loss_forward")
        ctx.save_for_backward(o_c, g_c, m_c, out)
        ctx.flags = (int(bool(with_l1)), int(bool(with_cs)))
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        o_c, g_c, m_c, out = ctx.saved_tensors
        g = grad_out.detach().reshape(1).contiguous().float()
        grad = torch.empty_like(o_c)
        q = _lib.DibrNormLoss()
        q.n_img, q.hw = o_c.shape[0], o_c.shape[2] * o_c.shape[3]
        q.with_l1, q.with_cs = ctx.flags
        q.out_norm, q.gt_norm, q.mask = _lib.ptr(o_c), _lib.ptr(g_c), _lib.ptr(m_c)
        q.out, q.grad_out, q.grad_out_norm = _lib.ptr(out), _lib.ptr(g), _lib.ptr(grad)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().dibr_norm_loss_backward(ctypes.byref(q), _stream(g.device)), "dibr_norm_loss_backward")
        return grad, None, None, None, None


class NORMLoss(torch.nn.Module):
    """``core/self6dpp/losses/vf_norm_loss.py:56-103``: L1 + cosine loss between the network's normals and the cropped
    teacher render (self_engine_utils.py:667-680), same constructor and ``forward(out_norm, gt_norm, mask)``.
    One reduction launch and one elementwise backward instead of ~15 launches and a host sync; the gradient goes to
    ``out_norm`` (``gt_norm`` is the rendered teacher map, ``mask`` the pseudo label: data)."""

    def __init__(self, with_l1=True, with_cs=True):
        super().__init__()
        assert with_l1 or with_cs
        self._with_l1, self._with_cs = bool(with_l1), bool(with_cs)

    def forward(self, out_norm, gt_norm, mask):
        return _NormLoss.apply(out_norm, gt_norm, mask, self._with_l1, self._with_cs)


class _LabL1Loss(Function):
    @staticmethod
    def forward(ctx, gt_img, ren_img, mask, no_l, bgr):
        _require_cuda_f32("gt_img", gt_img)
        _require_cuda_f32("ren_img", ren_img)
        if gt_img.dim() != 4 or gt_img.shape[1] != 3 or gt_img.shape != ren_img.shape:
            raise ValueError("gt_img / ren_img must both be (N, 3, H, W), got {} and {}".format(
                tuple(gt_img.shape), tuple(ren_img.shape)))
        device = ren_img.device
        n, _, h, w = ren_img.shape
        g_c, r_c = gt_img.detach().contiguous(), ren_img.detach().contiguous()
        m_c = None
        if mask is not None:
            _require_cuda_f32("mask", mask)
            if mask.numel() != n * h * w:
                raise ValueError("mask must be (N, 1, H, W)")
            m_c = mask.detach().contiguous()
        out = torch.empty(3, dtype=torch.float32, device=device)
        q = _lib.DibrLabLoss()
        q.n_img, q.hw, q.bgr, q.no_l = n, h * w, int(bool(bgr)), int(bool(no_l))
        q.gt, q.ren, q.mask = _lib.ptr(g_c), _lib.ptr(r_c), _lib.ptr(m_c)
        scratch = _scratch(n * h * w, device, kind="lab")
        q.scratch, q.out = _lib.ptr(scratch), _lib.ptr(out)
        with torch.cuda.device(device):
            _lib.check(_lib.load().dibr_lab_loss_forward(ctypes.byref(q), _stream(device)), "dibr_lab_loss_forward")
        ctx.save_for_backward(g_c, r_c, out, *([m_c] if m_c is not None else []))
        ctx.has_m = m_c is not None
        ctx.flags = (int(bool(bgr)), int(bool(no_l)))
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        g_c, r_c, out = saved[:3]
        m_c = saved[3] if ctx.has_m else None
        device = r_c.device
        n, _, h, w = r_c.shape
        go = grad_out.detach().reshape(1).to(torch.float32).contiguous()
        gr = torch.empty_like(r_c)
        q = _lib.DibrLabLoss()
        q.n_img, q.hw = n, h * w
        q.bgr, q.no_l = ctx.flags
        q.gt, q.ren, q.mask = _lib.ptr(g_c), _lib.ptr(r_c), _lib.ptr(m_c)
        q.out, q.grad_out, q.grad_ren = _lib.ptr(out), _lib.ptr(go), _lib.ptr(gr)
        with torch.cuda.device(device):
            _lib.check(_lib.load().dibr_lab_loss_backward(ctypes.byref(q), _stream(device)), "dibr_lab_loss_backward")
        return None, gr, None, None, None


def lab_l1_loss(gt_img_roi, ren_img_roi, pseudo_vis_mask_roi=None, no_l=False, bgr=True):
    """self_engine_utils.py:745-773 (without the LAB_LW factor):

        lab_x = normalize_lab(rgb_to_lab(x[:, [2, 1, 0]]))
        loss  = smooth_l1_loss(lab_gt * m, lab_ren * m, beta=0, reduction="sum") / max(1, m.sum())

    over the a/b channels only when ``no_l`` (LAB_NO_L).  Images are (N, 3, H, W) with B,G,R planes like the
    reference's crops (``bgr=False`` for R,G,B planes); the mask is (N, 1, H, W) or None (ones).  The gradient flows to
    ``ren_img_roi`` only -- the real crop and the pseudo mask are data on this path."""
    return _LabL1Loss.apply(gt_img_roi, ren_img_roi, pseudo_vis_mask_roi, no_l, bgr)
