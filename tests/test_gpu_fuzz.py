"""-m gpu: randomised small scenes through the operator seam against the fp32 operation-order oracle (bit-exact face
indices, attributes, and the soft mask / gradients within tolerance): odd image sizes, ragged K / expand, faces that are
tiny, huge, degenerate, off-screen or behind each other."""
import numpy as np
import pytest
import torch

from oracle import dibr_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _random_scene(rng, B, F, H, W):
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    kind = rng.integers(0, 4)
    c = torch.rand(B, F, 1, 2, generator=g) * 2.4 - 1.2                       # centres, some off-screen
    if kind == 0:
        size = 0.02 + 0.1 * torch.rand(B, F, 1, 1, generator=g)                # small faces
    elif kind == 1:
        size = 0.02 + 1.5 * torch.rand(B, F, 1, 1, generator=g) ** 4           # a few huge ones
    elif kind == 2:
        size = torch.full((B, F, 1, 1), 0.3)
    else:
        size = 0.001 + 0.05 * torch.rand(B, F, 1, 1, generator=g)              # slivers / near-degenerate
    d = (torch.rand(B, F, 3, 2, generator=g) - 0.5) * size
    if kind == 3:
        d[:, ::3, 2] = d[:, ::3, 1]                                            # exactly degenerate faces
    p2 = (c + d).reshape(B, F, 6).contiguous()
    p3 = torch.zeros(B, F, 9)
    z = -(0.5 + torch.rand(B, F, 3, generator=g))
    if kind == 2:
        z = torch.round(z * 4) / 4                                             # many equal depths: the tie rule
    p3[:, :, 2::3] = z
    e1, e2 = p2[:, :, 2:4] - p2[:, :, 0:2], p2[:, :, 4:6] - p2[:, :, 0:2]
    nz = (e1[..., 0] * e2[..., 1] - e1[..., 1] * e2[..., 0]).unsqueeze(-1).contiguous()
    D = int(rng.integers(1, 6))
    at = torch.rand(B, F, 3 * D, generator=g)
    return p3, p2, nz, at


@pytest.mark.parametrize("seed", list(range(10)))
def test_random_scene_matches_oracle(seed):
    from self6dpp_b200 import rasterizer as Rz
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 4))
    F = int(rng.choice([1, 7, 60, 400, 1500]))
    H, W = int(rng.integers(5, 90)), int(rng.integers(5, 90))
    knum = int(rng.choice([1, 3, 30, 40]))
    expand = float(rng.choice([0.0, 0.02, 0.11]))
    p3, p2, nz, at = _random_scene(rng, B, F, H, W)
    dev = torch.device("cuda:0")
    fw32 = O.rasterize(W, H, p3, p2, nz, at, expand=expand, knum=knum)
    fw64 = O.rasterize(W, H, p3.double(), p2.double(), nz.double(), at.double(), expand=expand, knum=knum)
    dbg = Rz.linear_rasterizer_debug(W, H, p3.to(dev), p2.to(dev), nz.to(dev), at.to(dev), expand=expand, knum=knum)
    assert torch.equal(dbg["imidx"].cpu(), fw32["imidx"]), f"imidx mismatch at {int((dbg['imidx'].cpu() != fw32['imidx']).sum())} pixels"
    assert torch.equal(dbg["im"].cpu(), fw32["im"]), "im differs from the fp32 operation-order oracle"
    same = (fw32["imidx"].double() == fw64["imidx"])
    # the soft mask: fp32 vs float64 where both agree on coverage; tolerance 1e-5 relative (+1e-5 of the scale)
    Hh.assert_close("improb", dbg["improb"], fw64["improb"], mask=same, rtol=1e-4, atol_rel=1e-5, outlier_frac=2e-3, outlier_tol=1.0)
    # gradients through the autograd Function, against the float64 oracle on the pixels where coverage agrees
    g = torch.Generator().manual_seed(seed)
    gI = torch.randn(fw64["im"].shape, generator=g, dtype=torch.float64) * same
    gP = torch.randn(fw64["improb"].shape, generator=g, dtype=torch.float64) * same
    dp2_ref, dc_ref = O.rasterize_backward(fw64, gI, gP)
    P2, AT = p2.to(dev).requires_grad_(True), at.to(dev).requires_grad_(True)
    im, improb = Rz.linear_rasterizer(W, H, p3.to(dev), P2, nz.to(dev), AT, expand, knum, 1000, 7000)
    (im * gI.float().to(dev)).sum().add((improb * gP.float().to(dev)).sum()).backward()
    Hh.assert_close("dldc", AT.grad, dc_ref, rtol=1e-4, atol_rel=1e-5, outlier_frac=2e-3, outlier_tol=1.0)
    Hh.assert_close("dldp2", P2.grad, dp2_ref, rtol=1e-3, atol_rel=1e-4, outlier_frac=5e-3, outlier_tol=1.0)
