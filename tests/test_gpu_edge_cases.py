"""-m gpu: edge cases of the operator seam -- a batch without faces, the maximum channel count, K = 0, one channel too
many (the reference's tests have none of these; the oracle defines the expected values)."""
import pytest
import torch

from oracle import dibr_oracle as O

pytestmark = pytest.mark.gpu


def test_empty_batch_max_channels_and_k_zero():
    from self6dpp_b200 import rasterizer as Rz
    dev = "cuda:0"
    # 1. zero faces
    im, pr = Rz.linear_rasterizer(40, 24, torch.zeros(2, 0, 9, device=dev), torch.zeros(2, 0, 6, device=dev), torch.zeros(2, 0, 1, device=dev), torch.zeros(2, 0, 9, device=dev))
    assert im.shape == (2, 24, 40, 3) and float(im.abs().max()) == 0 and float(pr.abs().max()) == 0, "empty"
    # 2. D = 12, knum = 0, and a requires-grad backward with knum = 0
    g = torch.Generator().manual_seed(0)
    F = 50
    c = torch.rand(1, F, 1, 2, generator=g) * 1.6 - 0.8
    p2 = (c + (torch.rand(1, F, 3, 2, generator=g) - 0.5) * 0.4).reshape(1, F, 6)
    p3 = torch.zeros(1, F, 9); p3[:, :, 2::3] = -(0.5 + torch.rand(1, F, 3, generator=g))
    e1, e2 = p2[:, :, 2:4] - p2[:, :, 0:2], p2[:, :, 4:6] - p2[:, :, 0:2]
    nz = (e1[..., 0] * e2[..., 1] - e1[..., 1] * e2[..., 0]).unsqueeze(-1)
    at = torch.rand(1, F, 36, generator=g)
    for knum in (0, 30):
        fw = O.rasterize(33, 21, p3, p2, nz, at, knum=knum)
        P2 = p2.to(dev).requires_grad_(True); AT = at.to(dev).requires_grad_(True)
        im, pr = Rz.linear_rasterizer(33, 21, p3.to(dev), P2, nz.to(dev), AT, 0.02, knum, 1000, 7000)
        assert torch.equal(im.detach().cpu(), fw["im"]), ("D=12 im", knum)
        assert float((pr.detach().cpu() - fw["improb"]).abs().max()) < 1e-5, ("prob", knum)
        (im.sum() + pr.sum()).backward()
        assert torch.isfinite(P2.grad).all() and torch.isfinite(AT.grad).all()
    # 3. D = 13 must be refused
    with pytest.raises(RuntimeError):
        Rz.linear_rasterizer(33, 21, p3.to(dev), p2.to(dev), nz.to(dev), torch.rand(1, F, 39).to(dev))
