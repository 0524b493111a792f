"""CPU: the C oracle (recalled kaolin v0.1 kernels) against an INDEPENDENT float64 PyTorch transcription whose
backward comes from autograd -- pins that the recalled K3/K4 really are the derivative of the recalled K1/K2."""
import numpy as np
import pytest
import torch

from oracle import dibr_oracle as O, torch_oracle as TO
from self6dpp_b200 import synth


@pytest.mark.parametrize("knum", [30, 3])
def test_c_oracle_matches_torch_autograd(knum):
    m = synth.icosphere(2, radius=0.05, noise_sigma=0.003, seed=0)
    H, W = 48, 64
    R, _ = synth.random_rotations(1, 3)
    t = np.array([[0.01, -0.02, 0.5]], np.float32)
    K = synth.crop_K(synth.K_LM, (325., 242.), 160., 64)
    cams = O.camera_params_from_RT_K(torch.tensor(R), torch.tensor(t), torch.tensor(K), H, W)
    p3, p2, nz, _ = O.project(torch.tensor(m["vertices"]), torch.tensor(m["faces"]), cams[0][0], cams[1][0], cams[2])
    F = p3.shape[1]
    cols, fa = torch.tensor(m["colors"]), torch.tensor(m["faces"]).long()
    one = torch.ones(F, 1)
    attr = torch.cat([cols[fa[:, 0]], one, cols[fa[:, 1]], one, cols[fa[:, 2]], one], 1)[None]
    fw = O.rasterize(W, H, p3.double(), p2.double(), nz.double(), attr.double(), knum=knum)
    g = torch.Generator().manual_seed(0)
    gI = torch.randn(1, H, W, 4, dtype=torch.float64, generator=g)
    gP = torch.randn(1, H, W, 1, dtype=torch.float64, generator=g)
    dp2, dc = O.rasterize_backward(fw, gI, gP)
    P2 = p2[0].double().clone().requires_grad_(True)
    A = attr[0].double().clone().requires_grad_(True)
    tw = TO.rasterize(W, H, p3[0].double(), P2, nz[0, :, 0].double(), A, knum=knum)
    assert torch.equal(tw["imidx"], fw["imidx"][0, ..., 0].long())
    assert (tw["im"] - fw["im"][0]).abs().max() < 1e-12
    assert (tw["improb"] - fw["improb"][0, ..., 0]).abs().max() < 1e-12
    ((tw["im"] * gI[0]).sum() + (tw["improb"] * gP[0, ..., 0]).sum()).backward()
    assert ((P2.grad - dp2[0]).abs().max() / dp2.abs().max()) < 1e-10
    assert ((A.grad - dc[0]).abs().max() / dc.abs().max()) < 1e-12


def test_fixed_order_vertex_shader_matches_torch():
    m = synth.ellipsoid(9, 12, seed=1)
    R, _ = synth.random_rotations(1, 5)
    t = torch.tensor([0.02, 0.01, 0.6])
    cr, cp, pj = O.camera_from_pose(torch.tensor(R[0]), t, torch.tensor(synth.K_LM), 640, 480, 0.01, 100.0)
    v, f = torch.tensor(m["vertices"]), torch.tensor(m["faces"])
    p3, p2, nz, nn = O.project(v, f, cr, cp, pj)
    from tests.helpers import torch_project
    q3, q2, qz, _ = torch_project(v.double(), f, cr.double(), cp.double(), pj.double())
    assert (p3[0].double() - q3).abs().max() < 1e-6
    assert (p2[0].double() - q2).abs().max() < 1e-6
    assert (nz[0].double() - qz).abs().max() < 1e-7
