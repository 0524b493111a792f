"""-m gpu: the CUDA path at the reference's linear_rasterizer seam vs the CPU oracle on identical fp32 inputs.
Bar (BASELINE.json north_star): face-index / visibility bit-exact vs the fp32 op-order oracle; interpolated
attributes, soft mask and gradients within 1e-5 relative (plus a 1e-5*max floor) of the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import dibr_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def run_case(p3, p2, nz, at, H, W, knum=30, expand=0.02, seed=0, check_grad=True, min_same=0.999, dp2_outlier_frac=0.0):
    from self6dpp_b200 import rasterizer as Rz
    dev = torch.device("cuda:0")
    fw32 = O.rasterize(W, H, p3, p2, nz, at, expand=expand, knum=knum)
    fw64 = O.rasterize(W, H, p3.double(), p2.double(), nz.double(), at.double(), expand=expand, knum=knum)
    dbg = Rz.linear_rasterizer_debug(W, H, p3.to(dev), p2.to(dev), nz.to(dev), at.to(dev), expand=expand, knum=knum)
    # 1. face index buffer: bit-exact against the fp32 operation-order oracle
    assert torch.equal(dbg["imidx"].cpu(), fw32["imidx"]), \
        f"imidx mismatch at {int((dbg['imidx'].cpu() != fw32['imidx']).sum())} pixels"
    same = (fw32["imidx"].double() == fw64["imidx"])                  # pixels where fp32 and fp64 agree on the face
    assert same.float().mean() > min_same
    # 2. attributes and soft mask vs float64
    assert torch.equal(dbg["im"].cpu(), fw32["im"]), "im differs from the fp32 operation-order oracle"
    e_im = Hh.assert_close("im", dbg["im"], fw64["im"], mask=same.expand_as(fw64["im"]), outlier_frac=5e-4)
    e_pr = Hh.assert_close("improb", dbg["improb"], fw64["improb"], mask=same)
    out = {"e_im": e_im, "e_pr": e_pr, "covered": int((fw32["imidx"] > 0).sum())}
    if not check_grad:
        return out
    # 3. gradients through the autograd Function
    g = torch.Generator().manual_seed(seed + 100)
    gI = torch.randn(fw64["im"].shape, generator=g, dtype=torch.float64) * same
    gP = torch.randn(fw64["improb"].shape, generator=g, dtype=torch.float64) * same
    dp2_ref, dc_ref = O.rasterize_backward(fw64, gI, gP)
    P2 = p2.to(dev).requires_grad_(True)
    AT = at.to(dev).requires_grad_(True)
    im, improb = Rz.linear_rasterizer(W, H, p3.to(dev), P2, nz.to(dev), AT, expand, knum, 1000, 7000)
    assert im.requires_grad and improb.requires_grad
    (im * gI.float().to(dev)).sum().add((improb * gP.float().to(dev)).sum()).backward()
    out["e_dc"] = Hh.assert_close("dldc", AT.grad, dc_ref)
    # dp2_outlier_frac: large cases hold a few sliver faces whose fp32 1/k3 is ill-conditioned (see helpers.assert_close)
    out["e_dp2"] = Hh.assert_close("dldp2", P2.grad, dp2_ref, rtol=1e-4, atol_rel=2e-5, outlier_frac=dp2_outlier_frac, outlier_tol=3e-4)
    # determinism: a second backward gives bit-identical gradients
    P2b = p2.to(dev).requires_grad_(True)
    ATb = at.to(dev).requires_grad_(True)
    im2, improb2 = Rz.linear_rasterizer(W, H, p3.to(dev), P2b, nz.to(dev), ATb, expand, knum, 1000, 7000)
    (im2 * gI.float().to(dev)).sum().add((improb2 * gP.float().to(dev)).sum()).backward()
    assert torch.equal(P2b.grad, P2.grad) and torch.equal(ATb.grad, AT.grad)
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_small_scene_64(seed):
    meshes, Rs, ts, Ks = Hh.small_scene(batch=2, level=2, H=64, W=64, seed=seed)
    p3, p2, nz, at = Hh.seam_inputs(meshes, Rs, ts, Ks, 64, 64)
    print(run_case(p3, p2, nz, at, 64, 64, seed=seed))


def test_non_square_partial_tiles():
    meshes, Rs, ts, Ks = Hh.small_scene(batch=3, level=3, H=72, W=100, seed=4)
    p3, p2, nz, at = Hh.seam_inputs(meshes, Rs, ts, Ks, 72, 100)
    print(run_case(p3, p2, nz, at, 72, 100, seed=4))


def test_k_cap_active():
    meshes, Rs, ts, Ks = Hh.small_scene(batch=2, level=3, H=64, W=64, seed=5, fill=0.5)
    p3, p2, nz, at = Hh.seam_inputs(meshes, Rs, ts, Ks, 64, 64)
    print(run_case(p3, p2, nz, at, 64, 64, knum=4, expand=0.08, seed=5))


def test_big_triangles_and_ties():
    """two coincident full-screen triangles + a nearer small one: exercises the cooperative
    large-face path and the (max z, min index) rule."""
    H = W = 96
    p2 = torch.tensor([[[-0.9, -0.9, 0.9, -0.9, -0.9, 0.9], [-0.9, -0.9, 0.9, -0.9, -0.9, 0.9],
                        [-0.3, -0.3, 0.2, -0.3, -0.3, 0.2], [0.1, 0.1, 0.8, 0.1, 0.1, 0.8]]], dtype=torch.float32)
    p3 = torch.zeros(1, 4, 9)
    for f, z in enumerate([-1.0, -1.0, -0.5, -2.0]):
        p3[0, f, 2::3] = z
    nz = torch.tensor([[[1.0], [1.0], [1.0], [-1.0]]])
    g = torch.Generator().manual_seed(0)
    at = torch.rand(1, 4, 12, generator=g)
    # pixel centres lie exactly on the shared hypotenuse, where fp32 and fp64 may disagree on w0 >= 0
    print(run_case(p3, p2, nz, at, H, W, seed=7, min_same=0.99))


def test_empty_and_offscreen():
    H = W = 64
    p2 = torch.tensor([[[2.0, 2.0, 3.0, 2.0, 2.0, 3.0]], [[-0.5, -0.5, 0.5, -0.5, -0.5, 0.5]]], dtype=torch.float32)
    p3 = -torch.ones(2, 1, 9)
    nz = torch.ones(2, 1, 1)
    at = torch.rand(2, 1, 12, generator=torch.Generator().manual_seed(1))
    out = run_case(p3, p2, nz, at, H, W, seed=8)
    assert out["covered"] > 0


def test_cfg1_shape_reduced():
    """cfg1 geometry (icosphere level 4, 5120 faces) at 160x120 so the oracle finishes in seconds."""
    from self6dpp_b200 import synth
    mesh = synth.icosphere(4, radius=0.05, noise_sigma=0.005, seed=0)
    R, _ = synth.random_rotations(1, 0)
    K = synth.K_LM.copy()
    K[:2] *= 0.25
    p3, p2, nz, at = Hh.seam_inputs([mesh], R, np.array([[0, 0, 0.8]], np.float32), K[None], 120, 160)
    print(run_case(p3, p2, nz, at, 120, 160, seed=9))


def test_cpu_tensor_raises():
    from self6dpp_b200 import rasterizer as Rz
    with pytest.raises(RuntimeError):
        Rz.linear_rasterizer(8, 8, torch.zeros(1, 1, 9), torch.zeros(1, 1, 6), torch.zeros(1, 1, 1), torch.zeros(1, 1, 12))


def test_dense_mesh_multi_batch_lists():
    """20k faces on a 64x64 image: > 512 faces per 16x16 tile, so the tile lists are processed in several batches
    (forward re-scans for the soft pass) and the K cap is active almost everywhere near the silhouette."""
    from self6dpp_b200 import synth
    mesh = synth.icosphere(5, radius=0.05, noise_sigma=0.002, seed=3)        # 20480 faces
    R, _ = synth.random_rotations(2, 17)
    ts = np.array([[0.0, 0.0, 0.6], [0.01, -0.01, 0.7]], np.float32)
    K = synth.crop_K(synth.K_LM, (325.26, 242.05), 150.0, 64)
    p3, p2, nz, at = Hh.seam_inputs([mesh, mesh], R, ts, np.stack([K, K]), 64, 64)
    out = run_case(p3, p2, nz, at, 64, 64, seed=11)
    print(out)


def test_knum_above_hit_capacity():
    """K = 45 on the dense mesh: pixels near the silhouette collect more faces than one pass of the soft phase holds
    (30 per pixel), so the tile re-collects with a skip window; also K > 30 in the backward's 'first K' rule."""
    from self6dpp_b200 import synth
    mesh = synth.icosphere(5, radius=0.05, noise_sigma=0.002, seed=4)        # 20480 faces
    R, _ = synth.random_rotations(1, 23)
    ts = np.array([[0.0, 0.0, 0.65]], np.float32)
    K = synth.crop_K(synth.K_LM, (325.26, 242.05), 150.0, 64)
    p3, p2, nz, at = Hh.seam_inputs([mesh], R, ts, K[None], 64, 64)
    out = run_case(p3, p2, nz, at, 64, 64, knum=45, seed=13)
    print(out)


def test_large_faces_many_per_tile():
    """80 overlapping large triangles: every face exceeds the 4-lane budget and goes through the CTA-cooperative
    path, more of them than the deferred list holds (overflow rescan)."""
    H = W = 64
    g = torch.Generator().manual_seed(5)
    n = 80
    c = torch.rand(n, 1, 2, generator=g) * 1.2 - 0.6
    d = (torch.rand(n, 3, 2, generator=g) - 0.5) * 1.4
    p2 = (c + d).reshape(1, n, 6).contiguous()
    p3 = torch.zeros(1, n, 9)
    p3[0, :, 2::3] = -(torch.rand(n, 1, generator=g) + 0.5)
    e1 = p2[0, :, 2:4] - p2[0, :, 0:2]
    e2 = p2[0, :, 4:6] - p2[0, :, 0:2]
    nz = (e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]).reshape(1, n, 1).contiguous()
    at = torch.rand(1, n, 12, generator=g)
    print(run_case(p3, p2, nz, at, H, W, seed=12, min_same=0.99))


@pytest.mark.parametrize("hw", [(480, 640)])
def test_cfg1_full_size_forward(hw):
    """cfg1 at the reference's full 480x640 (renderer_base.py:7-8): forward only against the fp32 oracle
    (face ids + interpolated attributes bit-exact), soft mask vs float64."""
    from self6dpp_b200 import synth
    H, W = hw
    mesh = synth.icosphere(4, radius=0.05, noise_sigma=0.005, seed=0)
    R, _ = synth.random_rotations(1, 0)
    p3, p2, nz, at = Hh.seam_inputs([mesh], R, np.array([[0, 0, 0.8]], np.float32), synth.K_LM[None], H, W)
    print(run_case(p3, p2, nz, at, H, W, seed=13, check_grad=False))
