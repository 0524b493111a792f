"""CPU: the chamfer nearest-neighbour oracle (oracle/nnd_oracle.c) against (1) golden vectors produced by the
reference's own torch_nndistance.py + depth_bp_chamfer_loss.py running on the reference's own nnd_cpu.cpp
(tests/golden/make_golden.py --nnd) and (2), where oracle/_ref was built, that compiled reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import nnd_oracle as N

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nnd.npz")


def test_oracle_matches_reference_golden():
    d = np.load(GOLD)
    x1, x2 = torch.tensor(d["x1"]), torch.tensor(d["x2"])
    d1, d2, i1, i2 = N.nnd_forward(x1, x2)
    assert np.array_equal(d1.numpy(), d["d1"]) and np.array_equal(d2.numpy(), d["d2"])
    assert int(i1[0][torch.tensor(d["x1"][0]).sub(torch.tensor(d["x2"][0, 3])).pow(2).sum(1).argmin()]) != 10   # duplicate: first index wins
    g1, g2 = N.nnd_backward(x1, x2, torch.tensor(d["g1"]), torch.tensor(d["g2"]), i1, i2)
    assert np.array_equal(g1.numpy(), d["gx1"]) and np.array_equal(g2.numpy(), d["gx2"])


def test_depth_loss_restatement_matches_reference_golden():
    d = np.load(GOLD)
    ren = torch.tensor(d["ren"], requires_grad=True)
    loss, loss_c = N.depth_bp_chamfer_loss(ren, torch.tensor(d["real"]), torch.tensor(d["K"]), 0.05, 0.5)
    (loss + loss_c).backward()
    assert np.allclose(loss.detach().numpy(), d["loss"], rtol=1e-6) and np.allclose(loss_c.detach().numpy(), d["loss_center"], rtol=1e-6)
    assert np.allclose(ren.grad.numpy(), d["g_ren"], rtol=1e-5, atol=1e-9)


@pytest.mark.skipif(N.ref_module() is None, reason="oracle/_ref/libnnd_ref.so not built (needs /root/reference)")
def test_oracle_is_bit_identical_to_the_compiled_reference():
    ref = N.ref_module()
    g = torch.Generator().manual_seed(3)
    for b, n, m in ((1, 1, 1), (2, 257, 1031), (3, 64, 5)):
        x1, x2 = torch.randn(b, n, 3, generator=g), torch.randn(b, m, 3, generator=g)
        d1, d2, i1, i2 = N.nnd_forward(x1, x2)
        rd1, rd2 = torch.zeros(b, n), torch.zeros(b, m)
        ri1, ri2 = torch.zeros(b, n, dtype=torch.int32), torch.zeros(b, m, dtype=torch.int32)
        ref.nnd_forward(x1, x2, rd1, rd2, ri1, ri2)
        assert torch.equal(d1, rd1) and torch.equal(d2, rd2) and torch.equal(i1, ri1) and torch.equal(i2, ri2)
        g1, g2 = torch.randn(b, n, generator=g), torch.randn(b, m, generator=g)
        o1, o2 = N.nnd_backward(x1, x2, g1, g2, i1, i2)
        r1, r2 = torch.zeros(b, n, 3), torch.zeros(b, m, 3)
        ref.nnd_backward(x1, x2, r1, r2, g1, g2, ri1, ri2)
        assert torch.equal(o1, r1) and torch.equal(o2, r2)
