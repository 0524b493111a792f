"""``linear_rasterizer`` -- drop-in for the reference's operator of the same name
(/root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:73-294): same positional
signature, same defaults (expand 0.02, knum 30, multiplier 1000, delta 7000; :90-97), same two
outputs ``(im_bxhxwxd, improb_bxhxwx1)`` that both require grad (:217-218), and gradients for
exactly the two inputs the reference differentiates -- ``points2d_bxfx6`` and
``vertex_attr_bxfx3d`` (:278-291).

Underneath: ``dibr_setup_faces`` -> ``dibr_forward`` / ``dibr_backward_faces`` of libdibr_b200.so
(hand-written sm_100a CUDA, include/dibr_b200.h).  No CPU path: CPU tensors raise.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib

DEFAULT_EXPAND = 0.02
DEFAULT_KNUM = 30
DEFAULT_MULTIPLIER = 1000
DEFAULT_DELTA = 7000


def _require_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (self6dpp_b200 has no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device):
    """cudaStream_t of torch's current stream on ``device`` (the raw accessor is ~40x cheaper than building a
    torch.cuda.Stream object per call)"""
    if _RAW_STREAM is not None:
        index = device.index if isinstance(device, torch.device) else torch.device(device).index
        return ctypes.c_void_p(_RAW_STREAM(torch.cuda.current_device() if index is None else index))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _NoSwitch(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on_device(device):
    """``torch.cuda.device(device)`` only when that is not the current device already (the context manager costs ~6 us of
    host time per use; the training loop sits on one device)"""
    index = device.index
    if index is None or index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


def _alloc_workspace(p, device):
    nbytes = _lib.workspace_bytes(p)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
    p.workspace = ctypes.c_void_p(ws.data_ptr())
    p.workspace_bytes = nbytes
    return ws


def _base_pass(batch, height, width, num_attr, knum, multiplier, delta, expand, total_faces, faces_per_image):
    p = _lib.DibrPass()
    p.batch, p.height, p.width = int(batch), int(height), int(width)
    p.num_attr, p.knum = int(num_attr), int(knum)
    p.multiplier, p.delta, p.expand = int(multiplier), int(delta), float(expand)
    p.total_faces, p.faces_per_image = int(total_faces), int(faces_per_image)
    p.min_output = -1
    return p


def rasterize_forward_raw(width, height, points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, attr_bxfx3d,
                          expand, knum, multiplier, delta):
    """Runs set-up + forward; returns (im, improb, imidx, imcomp, workspace, pass_struct, kept tensors)."""
    for n, t in (("points3d_bxfx9", points3d_bxfx9), ("points2d_bxfx6", points2d_bxfx6),
                 ("normalz_bxfx1", normalz_bxfx1), ("vertex_attr_bxfx3d", attr_bxfx3d)):
        _require_cuda_f32(n, t)
    b, f = points3d_bxfx9.shape[0], points3d_bxfx9.shape[1]
    if points3d_bxfx9.shape[2] != 9 or tuple(points2d_bxfx6.shape) != (b, f, 6):
        raise RuntimeError("points3d must be bxfx9 and points2d bxfx6")
    num_vertex_attr = attr_bxfx3d.shape[2] / 3
    assert num_vertex_attr == int(num_vertex_attr), \
        "vertex_attr_bxfx3d has shape {} which is not a multiple of 3".format(attr_bxfx3d.shape[2])
    d = int(num_vertex_attr)
    device = points3d_bxfx9.device
    p3 = points3d_bxfx9.detach().contiguous()
    p2 = points2d_bxfx6.detach().contiguous()
    nz = normalz_bxfx1.detach().contiguous()
    at = attr_bxfx3d.detach().contiguous()
    with torch.cuda.device(device):
        p = _base_pass(b, height, width, d, knum, multiplier, delta, expand, b * f, f)
        ws = _alloc_workspace(p, device)
        im = torch.empty(b, height, width, d, dtype=torch.float32, device=device)
        improb = torch.empty(b, height, width, 1, dtype=torch.float32, device=device)
        imcomp = torch.empty(b, height, width, dtype=torch.float32, device=device)
        imidx = torch.empty(b, height, width, dtype=torch.int32, device=device)
        p.points3d, p.points2d, p.normalz = _lib.ptr(p3), _lib.ptr(p2), _lib.ptr(nz)
        p.face_attr = _lib.ptr(at)
        p.im, p.improb, p.imidx, p.imcomp = _lib.ptr(im), _lib.ptr(improb), _lib.ptr(imidx), _lib.ptr(imcomp)
        lib = _lib.load()
        st = _stream(device)
        _lib.check(lib.dibr_setup_faces(ctypes.byref(p), st), "dibr_setup_faces")
        _lib.check(lib.dibr_forward(ctypes.byref(p), st), "dibr_forward")
    return im, improb, imidx, imcomp, ws, p, (p3, p2, nz, at)


class LinearRasterizer(Function):
    @staticmethod
    def forward(ctx, width, height, tfpoints3d_bxfx9, tfpoints2d_bxfx6, tfnormalz_bxfx1, vertex_attr_bxfx3d,
                expand=None, knum=None, multiplier=None, delta=None, debug=False):
        if expand is None:
            expand = DEFAULT_EXPAND
        if knum is None:
            knum = DEFAULT_KNUM
        if multiplier is None:
            multiplier = DEFAULT_MULTIPLIER
        if delta is None:
            delta = DEFAULT_DELTA
        im, improb, imidx, imcomp, ws, p, kept = rasterize_forward_raw(
            width, height, tfpoints3d_bxfx9, tfpoints2d_bxfx6, tfnormalz_bxfx1, vertex_attr_bxfx3d,
            expand, knum, multiplier, delta)
        ctx.save_for_backward(improb, imidx, imcomp, ws, kept[3])
        ctx.cfg = (int(width), int(height), float(expand), int(knum), int(multiplier), int(delta),
                   tfpoints3d_bxfx9.shape[0], tfpoints3d_bxfx9.shape[1], im.shape[3])
        ctx.mark_non_differentiable()
        return im, improb

    @staticmethod
    def backward(ctx, dldI_bxhxwxd, dldp_bxhxwx1):
        improb, imidx, imcomp, ws, attr = ctx.saved_tensors
        width, height, expand, knum, multiplier, delta, b, f, d = ctx.cfg
        device = improb.device
        gI = dldI_bxhxwxd.contiguous() if dldI_bxhxwxd is not None else None
        gP = dldp_bxhxwx1.contiguous() if dldp_bxhxwx1 is not None else None
        with torch.cuda.device(device):
            p = _base_pass(b, height, width, d, knum, multiplier, delta, expand, b * f, f)
            p.workspace = ctypes.c_void_p(ws.data_ptr())
            p.workspace_bytes = ws.numel()
            p.face_attr = _lib.ptr(attr)
            p.improb, p.imidx, p.imcomp = _lib.ptr(improb), _lib.ptr(imidx), _lib.ptr(imcomp)
            p.grad_im, p.grad_improb = _lib.ptr(gI), _lib.ptr(gP)
            dldp2 = torch.empty(b, f, 6, dtype=torch.float32, device=device)
            dldc = torch.empty(b, f, 3 * d, dtype=torch.float32, device=device)
            p.grad_points2d, p.grad_face_attr = _lib.ptr(dldp2), _lib.ptr(dldc)
            _lib.check(_lib.load().dibr_backward_faces(ctypes.byref(p), _stream(device)), "dibr_backward_faces")
        # same 11-slot layout as the reference (rasterizer.py:278-291): only points2d and attributes
        return (None, None, None, dldp2, None, dldc, None, None, None, None, None)


def linear_rasterizer(width, height, tfpoints3d_bxfx9, tfpoints2d_bxfx6, tfnormalz_bxfx1, vertex_attr_bxfx3d,
                      expand=None, knum=None, multiplier=None, delta=None, debug=False):
    """Reference signature: rasterizer.py:73-88 / :294 (``LinearRasterizer.apply``)."""
    im, improb = LinearRasterizer.apply(width, height, tfpoints3d_bxfx9, tfpoints2d_bxfx6, tfnormalz_bxfx1,
                                        vertex_attr_bxfx3d, expand, knum, multiplier, delta, debug)
    # the reference forces requires_grad on both outputs even when no input needs grad (:217-218)
    if not im.requires_grad:
        im.requires_grad_(True)
        improb.requires_grad_(True)
    return im, improb


def linear_rasterizer_debug(width, height, points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, attr_bxfx3d,
                            expand=DEFAULT_EXPAND, knum=DEFAULT_KNUM, multiplier=DEFAULT_MULTIPLIER,
                            delta=DEFAULT_DELTA):
    """Forward only, returning the internal buffers the parity tests compare bit-exactly:
    ``imidx`` as the reference's fp32 'face+1, 0 = none' image (rasterizer.py:124) and the raw
    int32 buffer with the K-th-face encoding."""
    im, improb, imidx, imcomp, _, _, _ = rasterize_forward_raw(
        width, height, points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, attr_bxfx3d, expand, knum, multiplier, delta)
    return {"im": im, "improb": improb, "imidx": imidx.clamp(min=0).to(torch.float32).unsqueeze(-1),
            "imidx_raw": imidx, "imcomp": imcomp}
