"""ORACLE -- test infrastructure only (never imported by self6dpp_b200/).

Python face of oracle/dibr_oracle.c: a CPU restatement of the DIB-R kernels that the reference
reaches through ``kaolin.graphics.dib_renderer.cuda.rasterizer.forward/backward``
(/root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:29,152-172,249-269).
kaolin v0.1 is an un-vendored dependency (only .gitignore:61 names it), so the kernel bodies are
restated from the published algorithm -- PARITY UNPINNED by the reference (it has no tests for
this path); pinned instead by tests/test_oracle_kat.py, tests/test_oracle_torch_xcheck.py and
finite differences.

Three things live here:
  * ``forward`` / ``backward``: the 19-positional-argument, fill-in-place extension contract
    (fp32 = kernel operation order, fp64 = tolerance oracle), usable as a drop-in stub module.
  * ``install_reference_stubs`` / ``import_reference``: register that stub (and a stub of
    core.utils.pose_utils) in ``sys.modules`` so the reference's OWN Python layers
    (VCRenderBatch, VCRenderMulti, perspective_projection, LinearRasterizer) run on CPU.
  * ``rasterize`` / ``rasterize_backward``: restatement of LinearRasterizer.forward/backward
    (rasterizer.py:73-291) that is dtype-generic (the reference hard-codes fp32 buffers).
"""
import ctypes
import os
import subprocess
import sys
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
REFERENCE_ROOT = "/root/reference"


def build(force=False):
    so = os.path.join(_HERE, "libdibr_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("dibr_oracle.c", "dibr_oracle_body.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        assert _LIB.dibr_oracle_abi_version() == 1
    return _LIB


def _suffix(t):
    if t.dtype == torch.float32:
        return "_f32"
    if t.dtype == torch.float64:
        return "_f64"
    raise TypeError(f"oracle supports float32/float64, got {t.dtype}")


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous(), "oracle works on contiguous CPU tensors"
    return ctypes.c_void_p(t.data_ptr())


def _same_dtype(*ts):
    d = ts[0].dtype
    for t in ts:
        if t.dtype != d:
            raise TypeError("oracle: mixed dtypes")


# ----------------------------------------------------------------------------------------------
# the extension contract (rasterizer.py:152-172 / :249-269)
# ----------------------------------------------------------------------------------------------
def forward(points3d_bxfx9, points2d_bxfx6, pointsdirect_bxfx1, pointsbbox_bxfx4, pointsbbox2_bxfx4,
            pointsdep_bxfx1, colors_bxfx3d, imidx_bxhxwx1, imdep_bxhxwx1, imwei_bxhxwx3,
            probface_bxhxwxk, probcase_bxhxwxk, probdis_bxhxwxk, probdep_bxhxwxk, probacc_bxhxwxk,
            im_bxhxwxd, improb_bxhxwx1, multiplier, sigmainv):
    """Fills imidx/imdep/imwei/im (K1) then prob*/improb (K2) in place; returns None."""
    _same_dtype(points3d_bxfx9, points2d_bxfx6, colors_bxfx3d, im_bxhxwxd, improb_bxhxwx1)
    sfx = _suffix(points3d_bxfx9)
    bnum, fnum = points3d_bxfx9.shape[0], points3d_bxfx9.shape[1]
    height, width, dnum = im_bxhxwxd.shape[1], im_bxhxwxd.shape[2], im_bxhxwxd.shape[3]
    knum = probface_bxhxwxk.shape[3]
    direct = pointsdirect_bxfx1.contiguous()
    L = lib()
    getattr(L, "dibr_oracle_forward_render" + sfx)(
        _p(points3d_bxfx9.contiguous()), _p(points2d_bxfx6.contiguous()), _p(direct),
        _p(pointsbbox_bxfx4.contiguous()), _p(colors_bxfx3d.contiguous()),
        _p(imidx_bxhxwx1), _p(imdep_bxhxwx1), _p(imwei_bxhxwx3), _p(im_bxhxwxd),
        bnum, height, width, fnum, dnum, int(multiplier))
    getattr(L, "dibr_oracle_forward_prob" + sfx)(
        _p(points2d_bxfx6.contiguous()), _p(pointsbbox2_bxfx4.contiguous()), _p(pointsdep_bxfx1.contiguous()),
        _p(imidx_bxhxwx1), _p(probface_bxhxwxk), _p(probcase_bxhxwxk), _p(probdis_bxhxwxk),
        _p(probdep_bxhxwxk), _p(improb_bxhxwx1),
        bnum, height, width, fnum, knum, int(multiplier), int(sigmainv))


def backward(grad_im_bxhxwxd, grad_improb_bxhxwx1, im_bxhxwxd, improb_bxhxwx1, imidx_bxhxwx1,
             imwei_bxhxwx3, probface_bxhxwxk, probcase_bxhxwxk, probdis_bxhxwxk, probdep_bxhxwxk,
             probacc_bxhxwxk, points2d_bxfx6, colors_bxfx3d, grad_points2d_bxfx6, grad_colors_bxfx3d,
             grad_points2dprob_bxfx6, debug_im_bxhxwx3, multiplier, sigmainv):
    """Accumulates into grad_points2d / grad_colors (K3) and grad_points2dprob (K4) in place."""
    sfx = _suffix(points2d_bxfx6)
    bnum, fnum = points2d_bxfx6.shape[0], points2d_bxfx6.shape[1]
    height, width, dnum = im_bxhxwxd.shape[1], im_bxhxwxd.shape[2], im_bxhxwxd.shape[3]
    knum = probface_bxhxwxk.shape[3]
    L = lib()
    gp = torch.zeros(grad_points2d_bxfx6.shape, dtype=torch.float64)
    gc = torch.zeros(grad_colors_bxfx3d.shape, dtype=torch.float64)
    gpp = torch.zeros(grad_points2dprob_bxfx6.shape, dtype=torch.float64)
    getattr(L, "dibr_oracle_backward_color" + sfx)(
        _p(grad_im_bxhxwxd.contiguous()), _p(imidx_bxhxwx1), _p(imwei_bxhxwx3),
        _p(points2d_bxfx6.contiguous()), _p(colors_bxfx3d.contiguous()), _p(gp), _p(gc),
        bnum, height, width, fnum, dnum, int(multiplier))
    getattr(L, "dibr_oracle_backward_prob" + sfx)(
        _p(grad_improb_bxhxwx1.contiguous()), _p(improb_bxhxwx1), _p(imidx_bxhxwx1),
        _p(probface_bxhxwxk), _p(probcase_bxhxwxk), _p(probdis_bxhxwxk),
        _p(points2d_bxfx6.contiguous()), _p(gpp),
        bnum, height, width, fnum, knum, int(multiplier), int(sigmainv))
    grad_points2d_bxfx6.add_(gp.to(grad_points2d_bxfx6.dtype))
    grad_colors_bxfx3d.add_(gc.to(grad_colors_bxfx3d.dtype))
    grad_points2dprob_bxfx6.add_(gpp.to(grad_points2dprob_bxfx6.dtype))


# ----------------------------------------------------------------------------------------------
# dtype-generic restatement of LinearRasterizer (rasterizer.py:36-70, 73-291)
# ----------------------------------------------------------------------------------------------
def prepare_tfpoints(points3d_bxfx9, points2d_bxfx6, multiplier, expand):
    """rasterizer.py:36-70 -- x multiplier, bbox, expanded bbox, mean depth."""
    b, f = points3d_bxfx9.shape[:2]
    p2m = float(multiplier) * points2d_bxfx6
    v = p2m.view(b, f, 3, 2)
    pmin = torch.min(v, dim=2)[0]
    pmax = torch.max(v, dim=2)[0]
    bbox = torch.cat((pmin, pmax), dim=2)
    bbox2 = torch.cat((pmin - expand * multiplier, pmax + expand * multiplier), dim=2)
    dep = ((points3d_bxfx9[..., 2] + points3d_bxfx9[..., 5] + points3d_bxfx9[..., 8]).unsqueeze(-1)) / 3.0
    return p2m.contiguous(), bbox.contiguous(), bbox2.contiguous(), dep.contiguous()


def rasterize(width, height, points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, attr_bxfx3d,
              expand=0.02, knum=30, multiplier=1000, delta=7000):
    """Forward of LinearRasterizer (rasterizer.py:73-220) in the inputs' dtype.
    Returns a dict with every buffer the reference allocates (im, improb, imidx, imwei, prob*)."""
    dt = points3d_bxfx9.dtype
    b, f = points3d_bxfx9.shape[:2]
    d = attr_bxfx3d.shape[2] // 3
    assert d * 3 == attr_bxfx3d.shape[2]
    p2m, bbox, bbox2, dep = prepare_tfpoints(points3d_bxfx9, points2d_bxfx6, multiplier, expand)
    z = lambda *s: torch.zeros(*s, dtype=dt)
    out = dict(
        imidx=z(b, height, width, 1), imdep=torch.full((b, height, width, 1), -1000.0, dtype=dt),
        imwei=z(b, height, width, 3), im=z(b, height, width, d), improb=z(b, height, width, 1),
        probface=z(b, height, width, knum), probcase=z(b, height, width, knum),
        probdis=z(b, height, width, knum), probdep=z(b, height, width, knum),
        probacc=z(b, height, width, knum), p2m=p2m, bbox=bbox, bbox2=bbox2, dep=dep,
        attr=attr_bxfx3d.contiguous(), multiplier=multiplier, delta=delta)
    forward(points3d_bxfx9.contiguous(), p2m, normalz_bxfx1.contiguous(), bbox, bbox2, dep, out["attr"],
            out["imidx"], out["imdep"], out["imwei"], out["probface"], out["probcase"], out["probdis"],
            out["probdep"], out["probacc"], out["im"], out["improb"], multiplier, delta)
    return out


def rasterize_backward(fw, dldI_bxhxwxd, dldp_bxhxwx1):
    """Backward of LinearRasterizer (rasterizer.py:222-291): returns (dL/dpoints2d_bxfx6 =
    dldp2 + dldp2_prob, dL/dattr_bxfx3d) -- the only two inputs the reference differentiates."""
    dt = fw["p2m"].dtype
    dldp2 = torch.zeros_like(fw["p2m"])
    dldp2_prob = torch.zeros_like(fw["p2m"])
    dldc = torch.zeros_like(fw["attr"])
    debug_im = torch.zeros(*fw["im"].shape[:3], 3, dtype=dt)
    backward(dldI_bxhxwxd.to(dt).contiguous(), dldp_bxhxwx1.to(dt).contiguous(), fw["im"], fw["improb"],
             fw["imidx"], fw["imwei"], fw["probface"], fw["probcase"], fw["probdis"], fw["probdep"],
             fw["probacc"], fw["p2m"], fw["attr"], dldp2, dldc, dldp2_prob, debug_im,
             fw["multiplier"], fw["delta"])
    return dldp2 + dldp2_prob, dldc


def work_counts(fw, normalz_bxfx1):
    """(N_cov, N_soft) of SURVEY.md 8(d) for a forward result."""
    sfx = _suffix(fw["p2m"])
    b, h, w, _ = fw["imidx"].shape
    f = fw["p2m"].shape[1]
    a, c = ctypes.c_longlong(0), ctypes.c_longlong(0)
    getattr(lib(), "dibr_oracle_work_counts" + sfx)(
        _p(normalz_bxfx1.contiguous()), _p(fw["bbox"]), _p(fw["bbox2"]), _p(fw["imidx"]),
        b, h, w, f, int(fw["multiplier"]), ctypes.byref(a), ctypes.byref(c))
    return a.value, c.value


def project(verts_px3, faces_fx3, cam_rot_3x3, cam_pos_3, cam_proj_4x4):
    """Fixed-operation-order vertex shader (see dibr_oracle_body.h, dibr_oracle_project).
    Returns (points3d_1xfx9, points2d_1xfx6, normalz_1xfx1, normal_1xfx3)."""
    dt = verts_px3.dtype
    sfx = _suffix(verts_px3)
    f = faces_fx3.shape[0]
    faces = faces_fx3.to(torch.int32).contiguous()
    p3 = torch.zeros(1, f, 9, dtype=dt)
    p2 = torch.zeros(1, f, 6, dtype=dt)
    nz = torch.zeros(1, f, 1, dtype=dt)
    nn = torch.zeros(1, f, 3, dtype=dt)
    getattr(lib(), "dibr_oracle_project" + sfx)(
        _p(verts_px3.contiguous()), verts_px3.shape[0], _p(faces), f,
        _p(cam_rot_3x3.to(dt).contiguous()), _p(cam_pos_3.to(dt).contiguous()),
        _p(cam_proj_4x4.to(dt).contiguous()), _p(p3), _p(p2), _p(nz), _p(nn))
    return p3, p2, nz, nn


def camera_from_pose(R_3x3, t_3, K_3x3, width, height, near=0.01, far=10.0):
    """Fixed-operation-order camera set-up (dibr_oracle_camera): -> (cam_rot 3x3, cam_pos 3, proj 4x4)."""
    dt = R_3x3.dtype
    sfx = _suffix(R_3x3)
    cr, cp, pj = torch.zeros(3, 3, dtype=dt), torch.zeros(3, dtype=dt), torch.zeros(4, 4, dtype=dt)
    getattr(lib(), "dibr_oracle_camera" + sfx)(
        _p(R_3x3.contiguous()), _p(t_3.to(dt).contiguous()), _p(K_3x3.to(dt).contiguous()), int(width), int(height),
        ctypes.c_double(near), ctypes.c_double(far), _p(cr), _p(cp), _p(pj))
    return cr, cp, pj


# ----------------------------------------------------------------------------------------------
# camera set-up restated (renderer/base.py:131-191 hard-codes .cuda(); utils/perspective.py:95-130)
# ----------------------------------------------------------------------------------------------
def projection_matrix(K, width, height, near=0.01, far=10.0, dtype=torch.float32):
    """projectiveprojection_real(K, 0, 0, w, h, nc, fc) -- utils/perspective.py:95-130."""
    K = torch.as_tensor(K, dtype=dtype)
    q = -(far + near) / float(far - near)
    qn = -2 * (far * near) / float(far - near)
    P = torch.zeros(4, 4, dtype=dtype)
    P[0, 0] = 2 * K[0, 0] / width
    P[1, 0] = -2 * K[0, 1] / width
    P[1, 1] = 2 * K[1, 1] / height
    P[2, 0] = (-2 * K[0, 2] + width) / width
    P[2, 1] = (+2 * K[1, 2] - height) / height
    P[2, 2] = q
    P[3, 2] = qn
    P[2, 3] = -1.0
    return P


def quat2mat(quat):
    """core/utils/pose_utils.py:349-400 quat2mat_torch (w,x,y,z), eps=0."""
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    qw, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    X, Y, Z = qx * 2.0, qy * 2.0, qz * 2.0
    wX, wY, wZ = qw * X, qw * Y, qw * Z
    xX, xY, xZ = qx * X, qx * Y, qx * Z
    yY, yZ, zZ = qy * Y, qy * Z, qz * Z
    return torch.stack([1.0 - (yY + zZ), xY - wZ, xZ + wY,
                        xY + wZ, 1.0 - (xX + zZ), yZ - wX,
                        xZ - wY, yZ + wX, 1.0 - (xX + yY)], dim=1).reshape(-1, 3, 3)


def camera_params_from_RT_K(Rs, ts, Ks, height, width, near=0.01, far=10.0, rot_type="mat"):
    """renderer/base.py:131-191 without the hard-coded cuda:0: returns
    [cam_view_R bx3x3 = diag(1,-1,-1) R, cam_pos bx3 = -(R^T t), proj 4x4 or bx4x4]."""
    Rs = torch.as_tensor(Rs)
    dt = Rs.dtype
    if rot_type == "quat":
        Rs = quat2mat(Rs)
    ts = torch.as_tensor(ts, dtype=dt)
    yz_flip = torch.eye(3, dtype=dt)
    yz_flip[1, 1], yz_flip[2, 2] = -1, -1
    cam_R = torch.stack([yz_flip @ Rs[i] for i in range(len(Rs))])
    cam_t = torch.stack([-(Rs[i].t() @ ts[i]) for i in range(len(Rs))])
    Ks = torch.as_tensor(Ks)
    if Ks.ndim == 2:
        proj = projection_matrix(Ks, width, height, near, far, dt)
    else:
        proj = torch.stack([projection_matrix(Ks[i], width, height, near, far, dt) for i in range(len(Ks))])
    return [cam_R, cam_t, proj]


# ----------------------------------------------------------------------------------------------
# plugging into the reference's own Python (SURVEY.md 8(c) recipe)
# ----------------------------------------------------------------------------------------------
def install_reference_stubs():
    """Register the oracle as kaolin.graphics.dib_renderer.cuda.rasterizer and a stub of
    core.utils.pose_utils (the real one needs transforms3d) so lib.dr_utils.dib_renderer_x imports."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present (it never is on the GPU box)")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    names = ["kaolin", "kaolin.graphics", "kaolin.graphics.dib_renderer", "kaolin.graphics.dib_renderer.cuda"]
    for n in names:
        sys.modules.setdefault(n, types.ModuleType(n))
    leaf = types.ModuleType("kaolin.graphics.dib_renderer.cuda.rasterizer")
    leaf.forward = forward
    leaf.backward = backward
    sys.modules[leaf.__name__] = leaf
    sys.modules["kaolin.graphics.dib_renderer.cuda"].rasterizer = leaf
    import core  # noqa: F401  (empty __init__)
    import core.utils  # noqa: F401
    pu = types.ModuleType("core.utils.pose_utils")
    pu.quat2mat_torch = lambda quat, eps=0.0: quat2mat(quat)
    sys.modules["core.utils.pose_utils"] = pu


def import_reference():
    """Returns the reference's own (VCRenderBatch, VCRenderMulti, VCRender, linear_rasterizer,
    perspective_projection) running on the oracle stub."""
    install_reference_stubs()
    from lib.dr_utils.dib_renderer_x.renderer.vcrender_batch import VCRenderBatch
    from lib.dr_utils.dib_renderer_x.renderer.vcrender_multi import VCRenderMulti
    from lib.dr_utils.dib_renderer_x.renderer.vcrender import VCRender
    from lib.dr_utils.dib_renderer_x.rasterizer import linear_rasterizer
    from lib.dr_utils.dib_renderer_x.renderer.vertex_shaders.perpsective import perspective_projection
    return dict(VCRenderBatch=VCRenderBatch, VCRenderMulti=VCRenderMulti, VCRender=VCRender,
                linear_rasterizer=linear_rasterizer, perspective_projection=perspective_projection)
