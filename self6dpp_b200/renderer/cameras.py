"""Batched camera set-up.  Restates /root/reference/lib/dr_utils/dib_renderer_x/renderer/base.py:131-191
(``set_camera_parameters_from_RT_K``: per-sample Python loop of ~10 tiny CUDA launches) and
utils/perspective.py:95-130 (``projectiveprojection_real``) as a handful of batched, autograd-friendly
torch ops on the inputs' own device (the reference hard-codes cuda:0, base.py:69,164-166)."""
import numpy as np
import torch


def quat2mat_torch(quat, eps=0.0):
    """core/utils/pose_utils.py:349-400 (w, x, y, z)."""
    assert quat.ndim == 2 and quat.shape[1] == 4, quat.shape
    q = quat / (quat.norm(p=2, dim=1, keepdim=True) + eps)
    qw, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    X, Y, Z = qx * 2.0, qy * 2.0, qz * 2.0
    wX, wY, wZ = qw * X, qw * Y, qw * Z
    xX, xY, xZ = qx * X, qx * Y, qx * Z
    yY, yZ, zZ = qy * Y, qy * Z, qz * Z
    return torch.stack([1.0 - (yY + zZ), xY - wZ, xZ + wY,
                        xY + wZ, 1.0 - (xX + zZ), yZ - wX,
                        xZ - wY, yZ + wX, 1.0 - (xX + yY)], dim=1).reshape(-1, 3, 3)


_PROJ_CONST = {}


def _proj_constants(width, height, near, far, device):
    """proj.flatten() = K.flatten() @ M + c  reproduces perspective.py:122-129 for x0 = y0 = 0."""
    key = (int(width), int(height), float(near), float(far), str(device))
    if key not in _PROJ_CONST:
        M = np.zeros((9, 16), dtype=np.float64)
        c = np.zeros(16, dtype=np.float64)
        w, h = float(width), float(height)
        M[0, 0 * 4 + 0] = 2.0 / w            # proj[0,0] = 2 fx / w
        M[1, 1 * 4 + 0] = -2.0 / w           # proj[1,0] = -2 K01 / w
        M[4, 1 * 4 + 1] = 2.0 / h            # proj[1,1] = 2 fy / h
        M[2, 2 * 4 + 0] = -2.0 / w           # proj[2,0] = (-2 px + w) / w
        c[2 * 4 + 0] = 1.0
        M[5, 2 * 4 + 1] = 2.0 / h            # proj[2,1] = (2 py - h) / h
        c[2 * 4 + 1] = -1.0
        c[2 * 4 + 2] = -(far + near) / float(far - near)        # q
        c[3 * 4 + 2] = -2 * (far * near) / float(far - near)    # qn
        c[2 * 4 + 3] = -1.0
        _PROJ_CONST[key] = (torch.tensor(M, dtype=torch.float32, device=device),
                            torch.tensor(c, dtype=torch.float32, device=device))
    return _PROJ_CONST[key]


def projection_from_K(Ks, width, height, near, far, device):
    """Ks: [3,3] or [b,3,3] (tensor / ndarray / list of [3,3], base.py:136) -> [4,4] or [b,4,4] float32 on ``device``."""
    if isinstance(Ks, (list, tuple)) and len(Ks) > 0 and isinstance(Ks[0], torch.Tensor):
        Ks = torch.stack(list(Ks))
    Ks = torch.as_tensor(Ks)
    if Ks.device != device or Ks.dtype != torch.float32:
        Ks = Ks.to(device=device, dtype=torch.float32)
    M, c = _proj_constants(width, height, near, far, device)
    flat = torch.addmm(c, Ks.reshape(-1, 9), M)
    return flat.reshape(4, 4) if Ks.ndim == 2 else flat.reshape(-1, 4, 4)


def _as_batch(x, device, last_shape):
    """list of tensors / arrays or one tensor -> one float32 tensor on device (autograd kept)."""
    if isinstance(x, (list, tuple)):
        x = torch.stack([e if isinstance(e, torch.Tensor) else torch.tensor(np.asarray(e), dtype=torch.float32)
                         for e in x])
    elif not isinstance(x, torch.Tensor):
        x = torch.tensor(np.asarray(x), dtype=torch.float32)
    if x.device != device or x.dtype != torch.float32:
        x = x.to(device=device, dtype=torch.float32)
    return x


def camera_params_from_RT_K(Rs, ts, Ks, height, width, near=0.01, far=10.0, rot_type="mat", device=None):
    """-> [cam_view_R bx3x3 = diag(1,-1,-1) R, cam_view_pos bx3 = -(R^T t), proj 4x4 | bx4x4]."""
    assert rot_type in ["mat", "quat"], rot_type
    if device is None:
        first = Rs[0] if isinstance(Rs, (list, tuple)) else Rs
        device = first.device if isinstance(first, torch.Tensor) and first.is_cuda else torch.device("cuda", torch.cuda.current_device())
    R = _as_batch(Rs, device, None)
    t = _as_batch(ts, device, None)
    if rot_type == "quat":
        R = quat2mat_torch(R)
    flip = torch.tensor([1.0, -1.0, -1.0], dtype=torch.float32, device=device).view(1, 3, 1)
    cam_R = R * flip                                             # yz_flip @ R   (base.py:169)
    cam_t = -torch.bmm(R.transpose(1, 2), t.reshape(-1, 3, 1)).squeeze(-1)   # -(R^T t)  (base.py:170)
    proj = projection_from_K(Ks, width, height, near, far, device)
    return [cam_R, cam_t, proj]


def proj_as_4x4(proj):
    """accept the reference's three projection layouts (vcrender_batch.py:43-46, perpsective.py:33-35):
    4x4 / bx4x4 real projection, or the 3x1 (bx3x1) diagonal form  xy = p.xy * proj.xy / (p.z * proj.z)."""
    if proj.shape[-1] == 4:
        return proj.reshape(-1, 4, 4)
    d = proj.reshape(-1, 3)
    P = torch.zeros(d.shape[0], 4, 4, dtype=d.dtype, device=d.device)
    P[:, 0, 0] = d[:, 0]
    P[:, 1, 1] = d[:, 1]
    P[:, 2, 3] = d[:, 2]
    return P


def look_at_camera_params(azimuth, elevation, distance):
    """kaolin.mathutils.geometry.transformations.compute_camera_params restated (only reached through
    set_look_at_parameters, base.py:105-126, which Self6D++ never calls)."""
    theta, phi = np.deg2rad(azimuth), np.deg2rad(elevation)
    cam_y = distance * np.sin(phi)
    temp = distance * np.cos(phi)
    cam_pos = np.array([temp * np.cos(theta), cam_y, temp * np.sin(theta)], dtype=np.float32)
    axis_z = cam_pos.copy()
    axis_y = np.array([0, 1, 0], dtype=np.float32)
    axis_x = np.cross(axis_y, axis_z)
    axis_y = np.cross(axis_z, axis_x)
    mat = np.stack([axis_x, axis_y, axis_z])
    mat = mat / (np.linalg.norm(mat, axis=1, keepdims=True) + 1e-15)
    return torch.tensor(mat, dtype=torch.float32), torch.tensor(cam_pos, dtype=torch.float32)
