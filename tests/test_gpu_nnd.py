"""-m gpu: chamfer nearest-neighbour kernels and the batched depth back-projection chamfer loss against the golden
vectors of the reference's compiled CPU implementation and against the oracle on ragged random clouds.
Bar: distances and indices bit-exact; gradients / loss within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import nnd_oracle as N

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nnd.npz")


def test_nnd_matches_reference_golden():
    from self6dpp_b200.nndistance import nnd
    d = np.load(GOLD)
    x1 = torch.tensor(d["x1"], device=DEV, requires_grad=True)
    x2 = torch.tensor(d["x2"], device=DEV, requires_grad=True)
    d1, d2 = nnd(x1, x2)
    assert np.array_equal(d1.detach().cpu().numpy(), d["d1"]) and np.array_equal(d2.detach().cpu().numpy(), d["d2"])
    ((d1 * torch.tensor(d["g1"], device=DEV)).sum() + (d2 * torch.tensor(d["g2"], device=DEV)).sum()).backward()
    for got, ref in ((x1.grad, d["gx1"]), (x2.grad, d["gx2"])):
        err = np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-5, err


def test_nnd_ragged_vs_oracle():
    from self6dpp_b200.nndistance import nnd_padded
    g = torch.Generator().manual_seed(5)
    B, S1, S2 = 4, 2500, 3000
    c1 = torch.tensor([2500, 1, 777, 0], dtype=torch.int32)
    c2 = torch.tensor([3000, 1234, 2, 50], dtype=torch.int32)
    x1, x2 = torch.randn(B, S1, 3, generator=g) * 0.1, torch.randn(B, S2, 3, generator=g) * 0.1
    X1 = x1.to(DEV).requires_grad_(True)
    X2 = x2.to(DEV).requires_grad_(True)
    d1, d2, i1, i2 = nnd_padded(X1, c1.to(DEV), X2, c2.to(DEV))
    g1, g2 = torch.randn(B, S1, generator=g), torch.randn(B, S2, generator=g)
    ((d1 * g1.to(DEV)).sum() + (d2 * g2.to(DEV)).sum()).backward()
    for b in range(B):
        n, m = int(c1[b]), int(c2[b])
        if n == 0 or m == 0:
            assert float(X1.grad[b].abs().max()) == 0.0 or n > 0
            continue
        od1, od2, oi1, oi2 = N.nnd_forward(x1[b:b + 1, :n], x2[b:b + 1, :m])
        assert torch.equal(d1[b, :n].cpu(), od1[0]) and torch.equal(d2[b, :m].cpu(), od2[0])
        assert torch.equal(i1[b, :n].cpu(), oi1[0]) and torch.equal(i2[b, :m].cpu(), oi2[0])
        o1, o2 = N.nnd_backward(x1[b:b + 1, :n], x2[b:b + 1, :m], g1[b:b + 1, :n], g2[b:b + 1, :m], oi1, oi2)
        assert float((X1.grad[b, :n].cpu() - o1[0]).abs().max()) <= 1e-5 * float(o1.abs().max()) + 1e-12
        assert float((X2.grad[b, :m].cpu() - o2[0]).abs().max()) <= 1e-5 * float(o2.abs().max()) + 1e-12
        assert float(X1.grad[b, n:].abs().max() if n < S1 else 0.0) == 0.0


def test_depth_bp_chamfer_loss_matches_reference_golden():
    from self6dpp_b200.nndistance import depth_bp_chamfer_loss
    d = np.load(GOLD)
    ren = torch.tensor(d["ren"], device=DEV, requires_grad=True)
    loss, loss_c = depth_bp_chamfer_loss(ren, torch.tensor(d["real"], device=DEV), torch.tensor(d["K"], device=DEV), 0.05, 0.5)
    (loss + loss_c).backward()
    assert abs(float(loss) - float(d["loss"])) <= 1e-5 * abs(float(d["loss"]))
    assert abs(float(loss_c) - float(d["loss_center"])) <= 1e-5 * abs(float(d["loss_center"]))
    err = np.abs(ren.grad.cpu().numpy() - d["g_ren"]).max() / np.abs(d["g_ren"]).max()
    assert err < 1e-5, err


def test_chamfer_loss_with_an_empty_sample_has_finite_gradients():
    """A sample whose rendered depth is all zero has no points: the reference skips it (depth_bp_chamfer_loss.py:47-48) and so
    it must contribute NO gradient -- in particular no 0 * inf = NaN from a masked-out mean (centre term on)."""
    from self6dpp_b200.nndistance import depth_bp_chamfer_loss
    d = np.load(GOLD)
    ren0 = torch.tensor(d["ren"], device=DEV)
    ren0[0] = 0.0                                  # sample 0 renders nothing
    ren = ren0.clone().requires_grad_(True)
    loss, loss_c = depth_bp_chamfer_loss(ren, torch.tensor(d["real"], device=DEV), torch.tensor(d["K"], device=DEV), 0.05, 0.5)
    (loss + loss_c).backward()
    assert torch.isfinite(loss) and torch.isfinite(loss_c)
    assert torch.isfinite(ren.grad).all()
    assert float(ren.grad[0].abs().max()) == 0.0 and float(ren.grad[1:].abs().max()) > 0.0


def test_rendered_depth_feeds_the_chamfer_loss():
    """end of the chain the reference builds in compute_self_loss_pose: rendered depth -> chamfer loss -> dL/dR, dL/dt"""
    from self6dpp_b200 import Renderer_dibr, synth
    from self6dpp_b200.nndistance import depth_bp_chamfer_loss
    mesh = synth.icosphere(3, radius=0.05, noise_sigma=0.003, seed=1)
    models = [{"vertices": torch.tensor(mesh["vertices"], device=DEV), "colors": torch.tensor(mesh["colors"], device=DEV),
               "normals": torch.tensor(mesh["normals"], device=DEV), "faces": torch.tensor(mesh["faces"], device=DEV, dtype=torch.int32)}]
    H = W = 64
    batch = synth.roi_batch([mesh, mesh], 2, res=W, seed=4, fill=(0.5, 0.7))
    K = torch.tensor(batch["Ks"], device=DEV)
    ren = Renderer_dibr(H, W, "VertexColorBatch")
    with torch.no_grad():
        tgt = ren.render_batch(torch.tensor(batch["Rs"], device=DEV), torch.tensor(batch["ts"], device=DEV) + 0.004, models * 2,
                               Ks=K, width=W, height=H, mode=["depth"])["depth"]
    Rs = torch.tensor(batch["Rs"], device=DEV, requires_grad=True)
    ts = torch.tensor(batch["ts"], device=DEV, requires_grad=True)
    out = ren.render_batch(Rs, ts, models * 2, Ks=K, width=W, height=H, mode=["depth"])
    loss, _ = depth_bp_chamfer_loss(out["depth"], tgt, K, distance_threshold=0.05)
    loss.backward()
    assert float(loss) > 0 and torch.isfinite(Rs.grad).all() and float(ts.grad.abs().max()) > 0
    # moving t towards the target (+0.004) must reduce the loss: the gradient points the other way
    assert float((ts.grad * 0.004).sum()) < 0


def test_nnd_cpu_tensor_raises():
    from self6dpp_b200.nndistance import nnd
    with pytest.raises(RuntimeError):
        nnd(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))


def _both(x1, c1, x2, c2, g1, g2):
    """grid search + inverse-index backward vs the exhaustive kernels on the same inputs"""
    from self6dpp_b200 import nndistance as ND
    outs = []
    for exhaustive in (False, True):
        ND.NNDFunction.exhaustive = exhaustive
        try:
            X1, X2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
            d1, d2, i1, i2 = ND.nnd_padded(X1, c1, X2, c2)
            ((d1 * g1).sum() + (d2 * g2).sum()).backward()
            outs.append((d1.detach(), d2.detach(), i1, i2, X1.grad, X2.grad))
        finally:
            ND.NNDFunction.exhaustive = False
    return outs


@pytest.mark.parametrize("case", ["surfaces", "far_apart", "duplicates_and_lines", "tiny_and_empty", "crowd", "non_finite"])
def test_grid_search_is_bit_identical_to_the_exhaustive_search(case):
    g = torch.Generator().manual_seed(hash(case) % 1000)
    B, S1, S2 = 3, 4000, 5000
    x1, x2 = torch.zeros(B, S1, 3), torch.zeros(B, S2, 3)
    c1, c2 = torch.full((B,), S1, dtype=torch.int32), torch.full((B,), S2, dtype=torch.int32)

    def surface(n, shift):
        uv = torch.rand(n, 2, generator=g) * 0.12 - 0.06
        z = 0.8 + 0.3 * uv[:, 0] ** 2 - 0.2 * uv[:, 1] + 0.001 * torch.randn(n, generator=g)
        return torch.stack((uv[:, 0], uv[:, 1], z), 1) + torch.tensor(shift)
    for b in range(B):
        x1[b], x2[b] = surface(S1, [0.0, 0.0, 0.0]), surface(S2, [0.002 * b, -0.001, 0.003])
    if case == "far_apart":                       # every query is far outside the target's box: the exhaustive fallback
        x2 += torch.tensor([0.5, -0.3, 0.4])
    elif case == "duplicates_and_lines":
        x2[0, 100:200] = x2[0, 5]                 # equal minima: the lowest index must win
        x1[1, :, 1:] = 0.0                        # a cloud on a line: two grid axes collapse
        x2[1, :, 1:] = 0.0
        x1[2] = torch.round(x1[2] * 500) / 500    # points on a lattice: many exact ties
        x2[2] = torch.round(x2[2] * 500) / 500
    elif case == "tiny_and_empty":
        c1 = torch.tensor([1, 0, 3], dtype=torch.int32)
        c2 = torch.tensor([2, 7, 0], dtype=torch.int32)
    elif case == "non_finite":                    # nnd_cpu.cpp:17 starts from target 0 and only ever replaces it on 'd < best'
        x1[0, 7] = float("nan")                   # a NaN query: every distance is NaN -> (NaN, 0)
        x1[0, 9, 2] = float("inf")                # an infinite query: every distance is inf -> (inf, 0)
        x2[1, 0, 0] = float("nan")                # target 0 is NaN: d(q, 0) = NaN sticks for every query of the sample
        x2[2, 17] = float("nan")                  # a NaN target elsewhere is simply never the nearest
    elif case == "crowd":                         # thousands of queries share one nearest neighbour (backward fallback)
        x2[0, :, :] = x2[0, :, :] * 0.001 + torch.tensor([0.3, 0.3, 1.5])
        x2[1, 10:] = x2[1, 3]
    g1, g2 = torch.randn(B, S1, generator=g).to(DEV), torch.randn(B, S2, generator=g).to(DEV)
    a, e = _both(x1.to(DEV), c1.to(DEV), x2.to(DEV), c2.to(DEV), g1, g2)
    for k, name in enumerate(("dist1", "dist2", "idx1", "idx2", "grad1", "grad2")):
        same = (a[k] == e[k]) | ((a[k] != a[k]) & (e[k] != e[k]))                 # NaN == NaN for this comparison
        assert bool(same.all()), f"{case}: {name} differs at {int((~same).sum())} entries"
    assert int(a[2].min()) >= 0 and int(a[2].max()) < S2 and int(a[3].min()) >= 0 and int(a[3].max()) < S1, "indices stay inside the clouds"


def test_backproject_compact_equals_the_torch_expressions():
    """dibr_backproject_compact vs backproject_th + compact_valid_points (the restated reference expressions):
    points and counts bit-exact, d loss / d depth to 1e-6"""
    from self6dpp_b200.nndistance import backproject_compact, backproject_th, compact_valid_points
    g = torch.Generator().manual_seed(9)
    B, H, W = 3, 37, 53                                  # not a multiple of the 1024-pixel chunk
    depth = (torch.rand(B, H, W, generator=g) + 0.5) * (torch.rand(B, H, W, generator=g) > 0.4)
    depth[1] = 0                                          # an empty map
    K = torch.tensor([[[60.0, 0, 25.5], [0, 62.0, 18.0], [0, 0, 1]]]).repeat(B, 1, 1)
    K[2, 0, 2] += 3.0
    for Kin in (K, K[0]):
        d1 = depth.to(DEV).requires_grad_(True)
        d2 = depth.to(DEV).requires_grad_(True)
        p1, c1 = backproject_compact(d1, Kin.to(DEV))
        p2, c2 = compact_valid_points(backproject_th(d2, Kin.to(DEV)))
        assert torch.equal(c1, c2) and torch.equal(p1, p2)
        w = torch.randn(p1.shape, generator=g).to(DEV)
        (p1 * w).sum().backward()
        (p2 * w).sum().backward()
        assert float((d1.grad - d2.grad).abs().max()) <= 1e-6 * float(d2.grad.abs().max())


@pytest.mark.parametrize("thr", [0.05, 0.0])
def test_fused_loss_reduction_equals_the_torch_expressions(thr):
    """center_lw = 0 takes the one-launch reduction (dibr_chamfer_reduce_*); it must agree with the torch restatement of
    depth_bp_chamfer_loss.py:38-62 (same path with the centre term switched on at a negligible weight), incl. an empty sample"""
    from self6dpp_b200.nndistance import depth_bp_chamfer_loss
    d = np.load(GOLD)
    real, K = torch.tensor(d["real"], device=DEV), torch.tensor(d["K"], device=DEV)
    r1 = torch.tensor(d["ren"], device=DEV, requires_grad=True)
    r2 = torch.tensor(d["ren"], device=DEV, requires_grad=True)
    l1, _ = depth_bp_chamfer_loss(r1, real, K, thr, 0)
    l2, _ = depth_bp_chamfer_loss(r2, real, K, thr, 1e-30)          # torch path (centre term contributes ~0)
    l1.backward()
    l2.backward()
    assert abs(float(l1) - float(l2)) <= 1e-5 * abs(float(l2))
    assert float((r1.grad - r2.grad).abs().max()) <= 1e-5 * float(r2.grad.abs().max())
