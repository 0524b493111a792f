"""Shared fixtures for the parity tests: seeded synthetic scenes in the reference's operator-seam
layout (points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, attr_bxfx3d), produced on CPU by the oracle's
fixed-order vertex shader so CUDA and oracle consume bit-identical fp32 inputs."""
import numpy as np
import torch

from oracle import dibr_oracle as O
from self6dpp_b200 import synth


def seam_inputs(meshes, Rs, ts, Ks, H, W, attr_key="colors", dtype=torch.float32):
    """one image per (mesh, R, t, K); all meshes must have the same face count (dense b x f layout)."""
    p3s, p2s, nzs, ats = [], [], [], []
    for m, R, t, K in zip(meshes, Rs, ts, Ks):
        cams = O.camera_params_from_RT_K(torch.tensor(R)[None], torch.tensor(t)[None], torch.tensor(K), H, W)
        v = torch.tensor(m["vertices"])
        f = torch.tensor(m["faces"]).long()
        p3, p2, nz, _ = O.project(v, f, cams[0][0], cams[1][0], cams[2])
        c = torch.tensor(m[attr_key])
        one = torch.ones(f.shape[0], 1)
        at = torch.cat([c[f[:, 0]], one, c[f[:, 1]], one, c[f[:, 2]], one], dim=1)[None]
        p3s.append(p3)
        p2s.append(p2)
        nzs.append(nz)
        ats.append(at)
    cat = lambda xs: torch.cat(xs, dim=0).to(dtype).contiguous()
    return cat(p3s), cat(p2s), cat(nzs), cat(ats)


def small_scene(batch=2, level=2, H=64, W=64, seed=0, fill=0.6):
    """icosphere(level) under `batch` random poses, cropped so it fills ~`fill` of the image."""
    mesh = synth.icosphere(level, radius=0.05, noise_sigma=0.004, seed=seed)
    Rs, _ = synth.random_rotations(batch, seed + 11)
    rng = np.random.default_rng(seed + 5)
    ts, Ks = [], []
    for i in range(batch):
        z = rng.uniform(0.5, 0.9)
        u, v = rng.uniform(200, 440), rng.uniform(160, 320)
        ts.append(np.array([(u - synth.K_LM[0, 2]) * z / synth.K_LM[0, 0],
                            (v - synth.K_LM[1, 2]) * z / synth.K_LM[1, 1], z], np.float32))
        diam = 2 * 0.055 * synth.K_LM[0, 0] / z
        Ks.append(synth.crop_K(synth.K_LM, (u, v), diam / fill, max(H, W)))
    return [mesh] * batch, Rs, np.stack(ts), np.stack(Ks)


def assert_close(name, got, ref, rtol=1e-5, atol_rel=1e-5, mask=None, outlier_frac=0.0, outlier_tol=1e-4):
    """|got - ref| <= rtol*|ref| + atol_rel*max|ref|  (fp32 result vs float64 oracle).

    ``outlier_frac``: fraction of elements allowed to exceed that bound, but never ``outlier_tol*max|ref|``.
    Needed only where the reference's own fp32 formula is ill-conditioned (weights of sliver triangles are
    k1/k3 with k3 -> 0, so ANY fp32 evaluation is 1/k3 away from float64); such elements are additionally
    pinned bit-exactly against the fp32 operation-order oracle by the callers."""
    got = torch.as_tensor(got).detach().double().cpu()
    ref = torch.as_tensor(ref).detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = float(ref.abs().max()) if ref.numel() else 0.0
    err = (got - ref).abs()
    tol = rtol * ref.abs() + atol_rel * scale + 1e-30
    if mask is not None:
        err = err[mask]
        tol = tol[mask]
    bad = err > tol
    nbad = int(bad.sum())
    assert nbad <= outlier_frac * bad.numel(), (f"{name}: {nbad} of {bad.numel()} elements out of tolerance; "
                                                f"max err {float(err.max()):.3e}, scale {scale:.3e}")
    if nbad:
        assert float(err.max()) <= outlier_tol * scale, f"{name}: outlier {float(err.max()):.3e} vs scale {scale:.3e}"
    return float(err.max()) / (scale + 1e-30) if err.numel() else 0.0


# ------------------------------------------------------------------------------------------------
# float64 end-to-end oracle of the fused path: differentiable torch vertex shader (restating
# perpsective.py:71-111) around the C oracle's rasterizer forward/backward.
# ------------------------------------------------------------------------------------------------
def load_golden(name):
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    meshes = []
    i = 0
    while f"mesh{i}_vertices" in d:
        meshes.append({k: d[f"mesh{i}_{k}"] for k in ("vertices", "faces", "colors", "normals")})
        i += 1
    return d, meshes


def torch_project(verts, faces, cam_rot, cam_pos, proj):
    """perspective_projection_real in the given dtype (differentiable)."""
    p = (verts - cam_pos.view(1, 3)) @ cam_rot.t()
    p4 = torch.cat([p, torch.ones_like(p[:, :1])], dim=1) @ proj
    xy = p4[:, :2] / p4[:, 3:4]
    f = faces.long()
    p3 = torch.cat([p[f[:, 0]], p[f[:, 1]], p[f[:, 2]]], dim=1)
    p2 = torch.cat([xy[f[:, 0]], xy[f[:, 1]], xy[f[:, 2]]], dim=1)
    n = torch.cross(p[f[:, 1]] - p[f[:, 0]], p[f[:, 2]] - p[f[:, 0]], dim=1)
    return p3, p2, n[:, 2:3], p


def oracle_render_batch64(meshes, ids, Rs, ts, Ks, H, W, attr_names, with_depth, grads, imidx_override=None, dt=torch.float64):
    """float64: every sample rendered with attributes [attr_names..., ones, (depth)]; returns images and
    dL/dRs, dL/dts for loss = sum(im * grads['im']) + sum(prob * grads['prob']).  ``dt=torch.float32`` runs the same
    pipeline in fp32 (profiles/r02_parity.md: what ANY fp32 evaluation is away from float64)."""
    Rs = torch.tensor(Rs, dtype=dt, requires_grad=True)
    ts = torch.tensor(ts, dtype=dt, requires_grad=True)
    cams = O.camera_params_from_RT_K(Rs, ts, torch.tensor(Ks, dtype=dt), H, W, near=0.01, far=100.0)
    ims, probs, idxs = [], [], []
    proxy = 0.0
    for i, mid in enumerate(ids):
        m = meshes[mid]
        v = torch.tensor(m["vertices"], dtype=dt)
        f = torch.tensor(m["faces"])
        proj = cams[2] if cams[2].ndim == 2 else cams[2][i]
        p3, p2, nz, pc = torch_project(v, f, cams[0][i], cams[1][i], proj)
        cols = [torch.tensor(m[a], dtype=dt) for a in attr_names] + [torch.ones(v.shape[0], 1, dtype=dt)]
        if with_depth:
            cols.append(-pc[:, 2:3])
        va = torch.cat(cols, dim=1)
        fl = f.long()
        at = torch.cat([va[fl[:, 0]], va[fl[:, 1]], va[fl[:, 2]]], dim=1)
        fw = O.rasterize(W, H, p3.detach()[None], p2.detach()[None], nz.detach()[None], at.detach()[None])
        ims.append(fw["im"])
        probs.append(fw["improb"])
        idxs.append(fw["imidx"])
        if grads is not None:
            dp2, dat = O.rasterize_backward(fw, grads["im"][i:i + 1].to(dt), grads["prob"][i:i + 1].to(dt))
            proxy = proxy + (p2 * dp2[0]).sum() + (at * dat[0]).sum()
    out = {"im": torch.cat(ims), "prob": torch.cat(probs), "imidx": torch.cat(idxs)}
    if grads is not None:
        proxy.backward()
        out["grad_Rs"], out["grad_ts"] = Rs.grad, ts.grad
    return out
