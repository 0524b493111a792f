"""Chamfer nearest-neighbour distance and the depth back-projection chamfer loss on B200.

Drop-ins for
  * ``core.csrc.torch_nndistance.torch_nndistance.nnd`` (/root/reference/core/csrc/torch_nndistance/torch_nndistance.py:13-85):
    ``nnd(xyz1 [b,n,3], xyz2 [b,m,3]) -> (dist1 [b,n], dist2 [b,m])`` squared distances to the nearest point of the other
    cloud, differentiable w.r.t. both clouds;
  * ``depth_bp_chamfer_loss`` (core/self6dpp/losses/depth_bp_chamfer_loss.py:12-62) and ``backproject_th``
    (lib/pysixd/misc.py:350-367).

The reference loops over the batch in Python (boolean-mask compaction forces a host sync per sample) and its CUDA
kernel then runs 16 blocks.  Here the whole batch is compacted on the device (cumsum + scatter, no sync), one launch
per direction covers every sample, and the backward is a deterministic gather (``dibr_nnd_forward`` /
``dibr_nnd_backward`` in include/dibr_b200.h).  No CPU fallback.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .rasterizer import _require_cuda_f32, _stream


def _launch(fn_name, p, device):
    with torch.cuda.device(device):
        _lib.check(getattr(_lib.load(), fn_name)(ctypes.byref(p), _stream(device)), fn_name)


_WS_CACHE = {}


def _workspace(p, device):
    """scratch of the grid search / inverse index, reused between calls of the same shape (stream order keeps a forward
    and the backward that follows it consistent; concurrent streams need their own call sites)"""
    n = ctypes.c_size_t(0)
    _lib.check(_lib.load().dibr_nnd_workspace_bytes(ctypes.byref(p), ctypes.byref(n)), "dibr_nnd_workspace_bytes")
    key = (str(device), n.value)
    ws = _WS_CACHE.get(key)
    if ws is None:
        if len(_WS_CACHE) > 8:
            _WS_CACHE.clear()
        ws = torch.empty(n.value + 256, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    off = (-ws.data_ptr()) % 256
    p.workspace, p.workspace_bytes = ws.data_ptr() + off, n.value
    return ws


class NNDFunction(Function):
    """Padded clouds: sample b of cloud i = rows [0, count_i[b]) of xyz_i[b] (count None: all rows).
    ``exhaustive=True`` runs the reference's O(n m) loops instead of the grid search (same answers)."""

    exhaustive = False

    @staticmethod
    def forward(ctx, xyz1, xyz2, count1=None, count2=None):
        _require_cuda_f32("xyz1", xyz1)
        _require_cuda_f32("xyz2", xyz2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        assert xyz2.shape[0] == b and xyz1.shape[2] == 3 and xyz2.shape[2] == 3
        x1, x2 = xyz1.detach().contiguous(), xyz2.detach().contiguous()
        device = x1.device
        dist1 = torch.zeros(b, n, dtype=torch.float32, device=device)
        dist2 = torch.zeros(b, m, dtype=torch.float32, device=device)
        idx1 = torch.zeros(b, n, dtype=torch.int32, device=device)
        idx2 = torch.zeros(b, m, dtype=torch.int32, device=device)
        c1 = count1.to(torch.int32).contiguous() if count1 is not None else None
        c2 = count2.to(torch.int32).contiguous() if count2 is not None else None
        p = _lib.DibrNnd()
        p.batch, p.stride1, p.stride2 = b, n, m
        p.count1, p.count2 = _lib.ptr(c1), _lib.ptr(c2)
        p.xyz1, p.xyz2 = _lib.ptr(x1), _lib.ptr(x2)
        p.dist1, p.dist2, p.idx1, p.idx2 = _lib.ptr(dist1), _lib.ptr(dist2), _lib.ptr(idx1), _lib.ptr(idx2)
        ctx.use_grid = not NNDFunction.exhaustive
        keep = _workspace(p, device) if ctx.use_grid else None
        _launch("dibr_nnd_forward", p, device)
        del keep
        ctx.save_for_backward(x1, x2, idx1, idx2, *( [c1] if c1 is not None else []), *([c2] if c2 is not None else []))
        ctx.has_counts = (c1 is not None, c2 is not None)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, _gi1, _gi2):
        saved = list(ctx.saved_tensors)
        x1, x2, idx1, idx2 = saved[:4]
        rest = saved[4:]
        c1 = rest.pop(0) if ctx.has_counts[0] else None
        c2 = rest.pop(0) if ctx.has_counts[1] else None
        device = x1.device
        b, n, _ = x1.shape
        m = x2.shape[1]
        g1 = (graddist1 if graddist1 is not None else torch.zeros(b, n, device=device)).contiguous()
        g2 = (graddist2 if graddist2 is not None else torch.zeros(b, m, device=device)).contiguous()
        gx1 = torch.empty(b, n, 3, dtype=torch.float32, device=device)
        gx2 = torch.empty(b, m, 3, dtype=torch.float32, device=device)
        p = _lib.DibrNnd()
        p.batch, p.stride1, p.stride2 = b, n, m
        p.count1, p.count2 = _lib.ptr(c1), _lib.ptr(c2)
        p.xyz1, p.xyz2, p.idx1, p.idx2 = _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(idx1), _lib.ptr(idx2)
        p.graddist1, p.graddist2, p.gradxyz1, p.gradxyz2 = _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(gx1), _lib.ptr(gx2)
        keep = _workspace(p, device) if ctx.use_grid else None
        _launch("dibr_nnd_backward", p, device)
        del keep
        return gx1, gx2, None, None


def nnd(xyz1, xyz2):
    """reference signature (torch_nndistance.py:82-85): -> (dist1, dist2)"""
    d1, d2, _, _ = NNDFunction.apply(xyz1, xyz2, None, None)
    return d1, d2


def nnd_padded(xyz1, count1, xyz2, count2):
    """ragged batch: -> (dist1, dist2, idx1, idx2); rows beyond count are zero"""
    return NNDFunction.apply(xyz1, xyz2, count1, count2)


def backproject_th(depth, K):
    """lib/pysixd/misc.py:350-367; depth [H,W] or [B,H,W], K [3,3] or [B,3,3] -> organised cloud [...,H,W,3]"""
    squeeze = depth.ndim == 2
    d = depth[None] if squeeze else depth
    Kb = K.reshape(-1, 3, 3).to(d)
    H, W = d.shape[-2:]
    ys = torch.arange(H, device=d.device, dtype=d.dtype).view(1, H, 1) - Kb[:, 1, 2].view(-1, 1, 1)
    xs = torch.arange(W, device=d.device, dtype=d.dtype).view(1, 1, W) - Kb[:, 0, 2].view(-1, 1, 1)
    out = torch.stack((xs * d / Kb[:, 0, 0].view(-1, 1, 1), ys * d / Kb[:, 1, 1].view(-1, 1, 1), d), dim=-1)
    return out[0] if squeeze else out


def compact_valid_points(cloud_bxhxwx3):
    """points with z > 0 of every sample, in row-major order (what boolean-mask indexing gives the reference), padded to
    H*W rows: -> (points [B, H*W, 3], count [B] int32).  Differentiable, no host sync."""
    B, H, W, _ = cloud_bxhxwx3.shape
    flat = cloud_bxhxwx3.reshape(B, H * W, 3)
    valid = flat[:, :, 2] > 0
    pos = torch.cumsum(valid.to(torch.int32), dim=1) - 1
    count = (pos[:, -1] + 1).to(torch.int32)
    slot = torch.where(valid, pos, torch.full_like(pos, H * W)).to(torch.int64)      # invalid -> dump row
    out = torch.zeros(B, H * W + 1, 3, dtype=flat.dtype, device=flat.device)
    out = out.scatter(1, slot.unsqueeze(-1).expand(B, H * W, 3), flat)
    return out[:, :H * W].contiguous(), count


class BackprojectCompact(Function):
    """depth [B,H,W], K [3,3] or [B,3,3] -> (points [B, H*W, 3], count [B] int32): ``backproject_th`` followed by the
    reference's per-sample ``pc[pc[:, :, 2] > 0]`` (depth_bp_chamfer_loss.py:27-36) for the whole batch in two launches
    (``dibr_backproject_compact``); the backward is one launch.  Same fp32 expressions as the torch version."""

    @staticmethod
    def forward(ctx, depth, K):
        _require_cuda_f32("depth", depth)
        B, H, W = depth.shape
        device = depth.device
        d = depth.detach().contiguous()
        Kc = K.detach().to(device=device, dtype=torch.float32).reshape(-1, 3, 3).contiguous()
        points = torch.zeros(B, H * W, 3, dtype=torch.float32, device=device)       # rows beyond count stay zero
        count = torch.empty(B, dtype=torch.int32, device=device)
        slot = torch.empty(B, H * W, dtype=torch.int32, device=device)
        chunks = torch.empty(B, (H * W + 1023) // 1024, dtype=torch.int32, device=device)
        p = _lib.DibrBackproject()
        p.batch, p.height, p.width, p.num_K = B, H, W, int(Kc.shape[0])
        p.depth, p.K, p.points = _lib.ptr(d), _lib.ptr(Kc), _lib.ptr(points)
        p.count, p.slot, p.chunk_count = _lib.ptr(count), _lib.ptr(slot), _lib.ptr(chunks)
        _launch("dibr_backproject_compact", p, device)
        ctx.save_for_backward(slot, Kc)
        ctx.dims = (B, H, W)
        ctx.mark_non_differentiable(count)
        return points, count

    @staticmethod
    def backward(ctx, grad_points, _gc):
        slot, Kc = ctx.saved_tensors
        B, H, W = ctx.dims
        gp = grad_points.contiguous()
        gd = torch.empty(B, H, W, dtype=torch.float32, device=gp.device)
        p = _lib.DibrBackproject()
        p.batch, p.height, p.width, p.num_K = B, H, W, int(Kc.shape[0])
        p.K, p.slot, p.grad_points, p.grad_depth = _lib.ptr(Kc), _lib.ptr(slot), _lib.ptr(gp), _lib.ptr(gd)
        _launch("dibr_backproject_compact_backward", p, gp.device)
        return gd, None


def backproject_compact(depth, K):
    """-> (points [B, H*W, 3], count [B]) of the pixels with depth > 0, row-major order"""
    return BackprojectCompact.apply(depth, K)


_TICKETS = {}


def _ticket(device):
    t = _TICKETS.get(str(device))
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=device)      # zero between calls: the kernel re-arms it
        _TICKETS[str(device)] = t
    return t


class ChamferReduce(Function):
    """(dist1 [B,S1], count1, dist2 [B,S2], count2, threshold) -> loss of depth_bp_chamfer_loss.py:38-62 (without the
    centre term): one reduction launch (``dibr_chamfer_reduce_forward``), one elementwise launch for the backward."""

    @staticmethod
    def forward(ctx, dist1, count1, dist2, count2, threshold):
        device = dist1.device
        d1, d2 = dist1.detach().contiguous(), dist2.detach().contiguous()
        B = d1.shape[0]
        stats = torch.empty(B, 4, dtype=torch.float32, device=device)
        out = torch.empty(2, dtype=torch.float32, device=device)
        q = _lib.DibrChamferReduce()
        q.batch, q.stride1, q.stride2, q.threshold = B, d1.shape[1], d2.shape[1], float(threshold)
        q.count1, q.count2, q.dist1, q.dist2 = _lib.ptr(count1), _lib.ptr(count2), _lib.ptr(d1), _lib.ptr(d2)
        q.stats, q.out = _lib.ptr(stats), _lib.ptr(out)
        q.ticket = ctypes.c_void_p(_ticket(device).data_ptr())
        _launch("dibr_chamfer_reduce_forward", q, device)
        ctx.save_for_backward(d1, d2, count1, count2, stats, out)
        ctx.threshold = float(threshold)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        d1, d2, count1, count2, stats, out = ctx.saved_tensors
        device = d1.device
        go = grad_out.detach().reshape(1).to(torch.float32).contiguous()
        g1, g2 = torch.empty_like(d1), torch.empty_like(d2)
        q = _lib.DibrChamferReduce()
        q.batch, q.stride1, q.stride2, q.threshold = d1.shape[0], d1.shape[1], d2.shape[1], ctx.threshold
        q.count1, q.count2, q.dist1, q.dist2 = _lib.ptr(count1), _lib.ptr(count2), _lib.ptr(d1), _lib.ptr(d2)
        q.stats, q.out, q.grad_out = _lib.ptr(stats), _lib.ptr(out), _lib.ptr(go)
        q.grad_dist1, q.grad_dist2 = _lib.ptr(g1), _lib.ptr(g2)
        _launch("dibr_chamfer_reduce_backward", q, device)
        return g1, None, g2, None, None


def depth_bp_chamfer_loss(ren_depths, real_depths, Ks, distance_threshold=0.05, center_lw=0):
    """
    Args (core/self6dpp/losses/depth_bp_chamfer_loss.py:12-19):
        ren_depths: BHW, real_depths: BHW (target points: depth(masked) => backproject (K)), Ks: [3,3] or [B,3,3]
    Returns (loss / max(num_valid,1), loss_center / max(num_valid,1)) like the reference, for the whole batch at once.
    """
    B, H, W = ren_depths.shape
    Kt = torch.as_tensor(Ks)
    real_pts, real_cnt = backproject_compact(real_depths, Kt)
    rend_pts, rend_cnt = backproject_compact(ren_depths, Kt)
    dist1, dist2, _, _ = nnd_padded(real_pts, real_cnt, rend_pts, rend_cnt)
    if not center_lw > 0:            # the usual configuration: the whole reduction is one launch
        loss = ChamferReduce.apply(dist1, real_cnt, dist2, rend_cnt, float(distance_threshold))
        return loss, torch.zeros((), dtype=ren_depths.dtype, device=ren_depths.device)
    ar = torch.arange(H * W, device=ren_depths.device).view(1, -1)
    v1, v2 = ar < real_cnt.view(-1, 1), ar < rend_cnt.view(-1, 1)
    s1, s2 = v1, v2
    if distance_threshold > 0:
        s1 = v1 & (dist1 < distance_threshold)
        s2 = v2 & (dist2 < distance_threshold)
    # A sample with an empty selection has mean = 0/0 = nan in the reference, which then skips it (:47-48).  Here the
    # denominators are made safe BEFORE dividing and the skip is decided from the counts: a masked-out nan would still
    # send 0 * inf = nan down the backward (into ren_depths, R, t and the network).
    n1, n2 = s1.sum(1), s2.sum(1)
    ok = (n1 > 0) & (n2 > 0)
    mean1 = (dist1 * s1).sum(1) / n1.clamp(min=1)
    mean2 = (dist2 * s2).sum(1) / n2.clamp(min=1)
    cur = mean1 + mean2
    num_valid = ok.sum().clamp(min=1)
    loss = torch.where(ok, cur, torch.zeros_like(cur)).sum() / num_valid
    loss_center = torch.zeros((), dtype=ren_depths.dtype, device=ren_depths.device)
    if center_lw > 0:
        c_real = (real_pts * v1.unsqueeze(-1)).sum(1) / real_cnt.view(-1, 1).clamp(min=1)
        c_rend = (rend_pts * v2.unsqueeze(-1)).sum(1) / rend_cnt.view(-1, 1).clamp(min=1)
        per = (c_real - c_rend).abs().mean(1) * center_lw       # smooth_l1(beta=0, "mean") = mean |.|
        loss_center = torch.where(ok, per, torch.zeros_like(per)).sum() / num_valid
    return loss, loss_center
