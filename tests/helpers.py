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


def assert_close(name, got, ref, rtol=1e-5, atol_rel=1e-5, mask=None):
    """|got - ref| <= rtol*|ref| + atol_rel*max|ref|  (fp32 result vs float64 oracle)."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = float(ref.abs().max()) if ref.numel() else 0.0
    err = (got - ref).abs()
    tol = rtol * ref.abs() + atol_rel * scale + 1e-30
    if mask is not None:
        err = err[mask]
        tol = tol[mask]
    bad = err > tol
    assert not bool(bad.any()), (f"{name}: {int(bad.sum())} of {bad.numel()} elements out of tolerance; "
                                 f"max err {float(err.max()):.3e}, scale {scale:.3e}")
    return float(err.max()) / (scale + 1e-30) if err.numel() else 0.0
