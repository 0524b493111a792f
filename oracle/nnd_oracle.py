"""ORACLE (test infrastructure only) for the chamfer nearest-neighbour op.

* ``nnd_forward`` / ``nnd_backward``: plain-C restatement (oracle/nnd_oracle.c) of the reference's CPU implementation
  /root/reference/core/csrc/torch_nndistance/src/nnd_cpu.cpp.
* ``ref_module()``: that very reference file compiled here into oracle/_ref/libnnd_ref.so (``make -C oracle ref``) and
  imported as the pybind11 module it is -- the REAL reference, used to pin the restatement and to make
  tests/golden/ref_nnd.npz.  It exists only where /root/reference existed at build time; the .so travels to the GPU box.
* ``depth_bp_chamfer_loss``: restatement of core/self6dpp/losses/depth_bp_chamfer_loss.py:12-62 on top of the oracle.
"""
import ctypes
import importlib.machinery
import importlib.util
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libnnd_oracle.so")
        src = os.path.join(_HERE, "nnd_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.run(["make", "-C", _HERE, "-s", "libnnd_oracle.so"], check=True, stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(so)
    return _LIB


def ref_module():
    """the reference's own nnd_cpu.cpp as a Python module (None if oracle/_ref was never built)"""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libnnd_ref.so")
        if not os.path.exists(path):
            return None
        loader = importlib.machinery.ExtensionFileLoader("nnd_ref", path)
        spec = importlib.util.spec_from_loader("nnd_ref", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _REF = mod
    return _REF


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def nnd_forward(xyz1, xyz2):
    """xyz1 [b,n,3], xyz2 [b,m,3] float32 CPU -> dist1 [b,n], dist2 [b,m], idx1, idx2 (int32)"""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    x1, x2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    d1, d2 = torch.zeros(b, n), torch.zeros(b, m)
    i1, i2 = torch.zeros(b, n, dtype=torch.int32), torch.zeros(b, m, dtype=torch.int32)
    lib().nnd_oracle_search(b, n, m, _p(x1), _p(x2), _p(d1), _p(i1))
    lib().nnd_oracle_search(b, m, n, _p(x2), _p(x1), _p(d2), _p(i2))
    return d1, d2, i1, i2


def nnd_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    x1, x2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    g1, g2 = torch.zeros(b, n, 3), torch.zeros(b, m, 3)
    lib().nnd_oracle_backward(b, n, m, _p(x1), _p(x2), _p(g1), _p(g2), _p(graddist1.contiguous().float()),
                              _p(graddist2.contiguous().float()), _p(idx1.contiguous()), _p(idx2.contiguous()))
    return g1, g2


class _NND(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, d2, i1, i2 = nnd_forward(xyz1.detach(), xyz2.detach())
        ctx.save_for_backward(xyz1.detach(), xyz2.detach(), i1, i2)
        return d1, d2

    @staticmethod
    def backward(ctx, g1, g2):
        x1, x2, i1, i2 = ctx.saved_tensors
        return nnd_backward(x1, x2, g1, g2, i1, i2)


def nnd(xyz1, xyz2):
    return _NND.apply(xyz1, xyz2)


def backproject_th(depth, K):
    """lib/pysixd/misc.py:350-367"""
    H, W = depth.shape[:2]
    Y, X = torch.meshgrid(torch.arange(H, dtype=depth.dtype) - K[1, 2], torch.arange(W, dtype=depth.dtype) - K[0, 2], indexing="ij")
    return torch.stack((X * depth / K[0, 0], Y * depth / K[1, 1], depth), dim=2)


def depth_bp_chamfer_loss(ren_depths, real_depths, Ks, distance_threshold=0.05, center_lw=0):
    """core/self6dpp/losses/depth_bp_chamfer_loss.py:12-62 (smooth_l1 with beta=0 is plain L1)"""
    bs = len(ren_depths)
    num_valid = 0
    loss = torch.tensor(0.0).to(ren_depths)
    loss_center = torch.tensor(0.0).to(ren_depths)
    for i in range(bs):
        K = Ks if Ks.ndim == 2 else Ks[i]
        real_pc = backproject_th(real_depths[i], K)
        real_pts = real_pc[real_pc[:, :, 2] > 0]
        rend_pc = backproject_th(ren_depths[i], K)
        rend_pts = rend_pc[rend_pc[:, :, 2] > 0]
        dist1, dist2 = nnd(real_pts[None], rend_pts[None])
        if distance_threshold > 0:
            dist1 = dist1[dist1 < distance_threshold]
            dist2 = dist2[dist2 < distance_threshold]
        cur = torch.mean(dist1) + torch.mean(dist2)
        if torch.isnan(cur):
            continue
        loss = loss + cur
        if center_lw > 0:
            loss_center = loss_center + (torch.mean(real_pts, 0) - torch.mean(rend_pts, 0)).abs().mean() * center_lw
        num_valid += 1
    return loss / max(num_valid, 1), loss_center / max(num_valid, 1)
