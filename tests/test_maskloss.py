"""Re-weighted BCE on probabilities (self6dpp_b200.losses.weighted_ex_loss_probs) against golden vectors produced by the
reference's own function (tests/golden/make_golden.py --maskloss)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_maskloss.npz")


def test_oracle_restatement_matches_reference_golden():
    from oracle import maskloss_oracle as M
    d = np.load(GOLD)
    for tag in "abc":
        w = d[f"{tag}_weight"] if f"{tag}_weight" in d else None
        loss, grad = M.weighted_ex_loss_probs(d[f"{tag}_probs"], d[f"{tag}_target"], w)
        assert abs(loss - float(d[f"{tag}_loss"])) <= 1e-5 * abs(float(d[f"{tag}_loss"]))
        ref = d[f"{tag}_grad"] / 1.7
        assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
def test_gpu_matches_reference_golden_and_is_reproducible():
    from self6dpp_b200.losses import weighted_ex_loss_probs
    dev = "cuda:0"
    d = np.load(GOLD)
    for tag in "abc":
        w = torch.tensor(d[f"{tag}_weight"], device=dev) if f"{tag}_weight" in d else None
        outs = []
        for _ in range(2):
            p = torch.tensor(d[f"{tag}_probs"], device=dev, requires_grad=True)
            loss = weighted_ex_loss_probs(p, torch.tensor(d[f"{tag}_target"], device=dev), weight=w)
            (loss * 1.7).backward()
            outs.append((loss.detach().clone(), p.grad.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])      # bit-reproducible
        assert abs(float(outs[0][0]) - float(d[f"{tag}_loss"])) <= 1e-5 * abs(float(d[f"{tag}_loss"]))   # tolerance: 1e-5 relative
        ref = d[f"{tag}_grad"]
        assert np.abs(outs[0][1].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
def test_gpu_large_input_against_oracle():
    """cfg2-sized soft mask (32 x 256 x 256) straight from random data: many CTAs, the last-CTA reduction"""
    from oracle import maskloss_oracle as M
    from self6dpp_b200.losses import weighted_ex_loss_probs
    g = torch.Generator().manual_seed(2)
    probs = torch.rand(32, 1, 256, 256, generator=g)
    target = (torch.rand(32, 1, 256, 256, generator=g) > 0.7).float()
    weight = torch.rand(32, 1, 256, 256, generator=g) + 0.5
    p = probs.to("cuda:0").requires_grad_(True)
    loss = weighted_ex_loss_probs(p, target.to("cuda:0"), weight=weight.to("cuda:0"))
    loss.backward()
    ref_loss, ref_grad = M.weighted_ex_loss_probs(probs.numpy(), target.numpy(), weight.numpy())
    assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
    assert np.abs(p.grad.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()


def test_cpu_tensor_raises():
    from self6dpp_b200.losses import weighted_ex_loss_probs
    with pytest.raises(Exception):
        weighted_ex_loss_probs(torch.rand(2, 1, 4, 4), torch.zeros(2, 1, 4, 4))


# ------------------------------------------------------------------------------------------------------------------
# soft dice loss (mask_losses.py:444-463), golden vectors from the reference's own function (make_golden.py --diceloss).
# Tolerance: 1e-5 relative on the value, 1e-5 of the largest entry on the gradient.
# ------------------------------------------------------------------------------------------------------------------
DICE_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_diceloss.npz")
DICE_TAGS = "abcd"


def test_dice_oracle_matches_reference_golden():
    from oracle import maskloss_oracle as M
    d = np.load(DICE_GOLD)
    for tag in DICE_TAGS:
        smooth, eps = (float(v) for v in d[f"{tag}_cfg"])
        loss, grad = M.soft_dice_loss(d[f"{tag}_probs"], d[f"{tag}_labels"], smooth, eps, str(d[f"{tag}_red"]), grad_out=d[f"{tag}_go"])
        np.testing.assert_allclose(loss, d[f"{tag}_loss"], rtol=1e-5)
        ref = d[f"{tag}_grad"]
        assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
def test_dice_gpu_matches_reference_golden_and_is_reproducible():
    from self6dpp_b200.losses import soft_dice_loss
    dev = "cuda:0"
    d = np.load(DICE_GOLD)
    for tag in DICE_TAGS:
        smooth, eps = (float(v) for v in d[f"{tag}_cfg"])
        red = str(d[f"{tag}_red"])
        go = torch.tensor(d[f"{tag}_go"], device=dev)
        outs = []
        for _ in range(2):
            p = torch.tensor(d[f"{tag}_probs"], device=dev, requires_grad=True)
            loss = soft_dice_loss(p, torch.tensor(d[f"{tag}_labels"], device=dev), smooth=smooth, eps=eps, reduction=red)
            (loss * (go if red == "none" else go[0])).sum().backward()
            outs.append((loss.detach().clone(), p.grad.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])      # bit-reproducible
        np.testing.assert_allclose(outs[0][0].cpu().numpy(), d[f"{tag}_loss"], rtol=1e-5)
        ref = d[f"{tag}_grad"]
        assert outs[0][1].shape == ref.shape
        assert np.abs(outs[0][1].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
def test_dice_gpu_crop_size_against_oracle():
    """the loop's size (32 x 1 x 256 x 256, eps = 0.002 as self_engine_utils.py:548 passes it)"""
    from oracle import maskloss_oracle as M
    from self6dpp_b200.losses import soft_dice_loss
    g = torch.Generator().manual_seed(6)
    probs = torch.rand(32, 1, 256, 256, generator=g)
    labels = (torch.rand(32, 1, 256, 256, generator=g) > 0.7).float()
    p = probs.to("cuda:0").requires_grad_(True)
    loss = soft_dice_loss(p, labels.to("cuda:0"), eps=0.002)
    loss.backward()
    ref_loss, ref_grad = M.soft_dice_loss(probs.numpy(), labels.numpy(), 0.0, 0.002, "mean")
    assert abs(float(loss.detach()) - ref_loss) <= 1e-5 * abs(ref_loss)
    assert np.abs(p.grad.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()


def test_dice_cpu_tensor_raises():
    from self6dpp_b200.losses import soft_dice_loss
    with pytest.raises(Exception):
        soft_dice_loss(torch.rand(2, 1, 4, 4), torch.ones(2, 1, 4, 4))


# ------------------------------------------------------------------------------------------------------------------
# normal-map loss NORMLoss (vf_norm_loss.py:56-103), golden vectors from the reference's own module
# (make_golden.py --normloss).  Tolerance: 1e-5 relative on the value, 1e-5 of the largest entry on the gradient.
# ------------------------------------------------------------------------------------------------------------------
NORM_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_normloss.npz")


def test_normloss_oracle_matches_reference_golden():
    from oracle import maskloss_oracle as M
    d = np.load(NORM_GOLD)
    for tag in "abc":
        l1, cs = (bool(v) for v in d[f"{tag}_flags"])
        loss, grad = M.norm_loss(d[f"{tag}_out"], d[f"{tag}_gt"], d[f"{tag}_mask"], l1, cs)
        assert abs(loss - float(d[f"{tag}_loss"])) <= 1e-5 * abs(float(d[f"{tag}_loss"]))
        ref = d[f"{tag}_grad"] / 1.3
        assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
def test_normloss_gpu_matches_reference_golden_and_is_reproducible():
    from self6dpp_b200.losses import NORMLoss
    dev = "cuda:0"
    d = np.load(NORM_GOLD)
    for tag in "abc":
        l1, cs = (bool(v) for v in d[f"{tag}_flags"])
        mod = NORMLoss(with_l1=l1, with_cs=cs)
        outs = []
        for _ in range(2):
            o = torch.tensor(d[f"{tag}_out"], device=dev, requires_grad=True)
            loss = mod(o, torch.tensor(d[f"{tag}_gt"], device=dev), torch.tensor(d[f"{tag}_mask"], device=dev))
            (loss * 1.3).backward()
            outs.append((loss.detach().clone(), o.grad.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])      # bit-reproducible
        assert abs(float(outs[0][0]) - float(d[f"{tag}_loss"])) <= 1e-5 * abs(float(d[f"{tag}_loss"]))
        ref = d[f"{tag}_grad"]
        assert np.abs(outs[0][1].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
def test_normloss_gpu_crop_size_against_oracle():
    """the loop's size: 32 x 3 x 64 x 64 normals against the cropped teacher render"""
    from oracle import maskloss_oracle as M
    from self6dpp_b200.losses import NORMLoss
    g = torch.Generator().manual_seed(9)
    o = torch.randn(32, 3, 64, 64, generator=g)
    gt = torch.nn.functional.normalize(torch.randn(32, 3, 64, 64, generator=g), dim=1)
    m = (torch.rand(32, 1, 64, 64, generator=g) > 0.5).float()
    x = o.to("cuda:0").requires_grad_(True)
    loss = NORMLoss()(x, gt.to("cuda:0"), m.to("cuda:0"))
    loss.backward()
    ref_loss, ref_grad = M.norm_loss(o.numpy(), gt.numpy(), m.numpy())
    assert abs(float(loss.detach()) - ref_loss) <= 1e-5 * abs(ref_loss)
    assert np.abs(x.grad.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()


def test_normloss_cpu_tensor_raises():
    from self6dpp_b200.losses import NORMLoss
    with pytest.raises(Exception):
        NORMLoss()(torch.rand(1, 3, 4, 4), torch.rand(1, 3, 4, 4), torch.ones(1, 1, 4, 4))
    with pytest.raises(AssertionError):
        NORMLoss(with_l1=False, with_cs=False)
