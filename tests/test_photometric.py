"""Photometric losses on the rendered colour crop against golden vectors the reference's own
lab.py produced through autograd (tests/golden/make_golden_photometric.py).  Tolerance: 1e-5 relative on the loss,
1e-5 of the largest gradient entry on the gradient (fp32 pow / cbrt differ in the last ulp between libraries); NaNs --
the reference's autograd yields NaN at exactly-black rendered pixels -- must sit at the same pixels."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_photometric.npz")
LAB_TAGS = "abc"


def _close_with_nans(got, ref, rtol):
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    # right at the sRGB knee / the Lab knee one ulp of the input flips the branch: allow a handful of such pixels
    bad = np.abs(got[ok] - ref[ok]) > rtol * np.abs(ref[ok]).max()
    assert bad.sum() <= 2, (int(bad.sum()), float(np.abs(got[ok] - ref[ok]).max()), float(np.abs(ref[ok]).max()))


def test_lab_oracle_matches_reference_golden():
    from oracle import photometric_oracle as P
    d = np.load(GOLD)
    for tag in LAB_TAGS:
        loss, grad = P.lab_l1_loss(d[f"lab_{tag}_gt"], d[f"lab_{tag}_ren"], d[f"lab_{tag}_mask"],
                                   no_l=bool(d[f"lab_{tag}_no_l"]), bgr=True)
        ref = float(d[f"lab_{tag}_loss"])
        assert abs(loss - ref) <= 1e-5 * abs(ref)
        _close_with_nans(grad * 0.2, d[f"lab_{tag}_grad"], 1e-5)


@pytest.mark.gpu
def test_lab_gpu_matches_reference_golden_and_is_reproducible():
    from self6dpp_b200.losses import lab_l1_loss
    dev = "cuda:0"
    d = np.load(GOLD)
    for tag in LAB_TAGS:
        outs = []
        for _ in range(2):
            ren = torch.tensor(d[f"lab_{tag}_ren"], device=dev, requires_grad=True)
            loss = lab_l1_loss(torch.tensor(d[f"lab_{tag}_gt"], device=dev), ren, torch.tensor(d[f"lab_{tag}_mask"], device=dev),
                               no_l=bool(d[f"lab_{tag}_no_l"]))
            (loss * 0.2).backward()
            outs.append((loss.detach().clone(), ren.grad.clone()))
        assert torch.equal(outs[0][0], outs[1][0])
        assert torch.equal(torch.nan_to_num(outs[0][1], nan=7.0), torch.nan_to_num(outs[1][1], nan=7.0))   # bit-reproducible
        ref = float(d[f"lab_{tag}_loss"])
        assert abs(float(outs[0][0]) - ref) <= 1e-5 * abs(ref)
        _close_with_nans(outs[0][1].cpu().numpy(), d[f"lab_{tag}_grad"], 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("bgr,no_l,with_mask", [(True, True, True), (False, False, True), (True, False, False)])
def test_lab_gpu_crop_size_against_oracle(bgr, no_l, with_mask):
    """cfg2-sized crops (32 x 3 x 256 x 256): many CTAs, the last-CTA reduction, both plane orders, no mask"""
    from oracle import photometric_oracle as P
    from self6dpp_b200.losses import lab_l1_loss
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(32, 3, 256, 256, generator=g)
    ren = (gt + 0.1 * torch.randn(32, 3, 256, 256, generator=g)).clamp(0.002, 1)
    mask = (torch.rand(32, 1, 256, 256, generator=g) > 0.4).float() if with_mask else None
    r = ren.to("cuda:0").requires_grad_(True)
    loss = lab_l1_loss(gt.to("cuda:0"), r, mask.to("cuda:0") if with_mask else None, no_l=no_l, bgr=bgr)
    loss.backward()
    ref_loss, ref_grad = P.lab_l1_loss(gt.numpy(), ren.numpy(), mask.numpy() if with_mask else None, no_l=no_l, bgr=bgr)
    assert abs(float(loss.detach()) - ref_loss) <= 1e-5 * abs(ref_loss)
    got = r.grad.cpu().numpy()
    bad = np.abs(got - ref_grad) > 1e-5 * np.abs(ref_grad).max()
    # the oracle is float64: where |lab_gt - lab_ren| or a knee test is within one fp32 ulp the sign / branch may flip
    assert bad.sum() <= 1e-5 * got.size, int(bad.sum())


def test_lab_cpu_tensor_raises():
    from self6dpp_b200.losses import lab_l1_loss
    with pytest.raises(Exception):
        lab_l1_loss(torch.rand(1, 3, 4, 4), torch.rand(1, 3, 4, 4), torch.ones(1, 1, 4, 4))


# ------------------------------------------------------------------------------------------------------------------
# MS-SSIM (core/self6dpp/losses/ssim.py).  Tolerances: 1e-5 relative on the value; on the gradient 1e-4 of its largest
# entry against the reference's fp32 autograd (whose own conv / cancellation noise is of that order: the float64 oracle
# agrees with it only to ~1e-5..1e-4) and 2e-5 against the float64 oracle.
# ------------------------------------------------------------------------------------------------------------------
SSIM_TAGS = "abcde"         # d, e: use_padding=True (zero-padded windows)


def test_ms_ssim_oracle_matches_reference_golden():
    from oracle import photometric_oracle as P
    d = np.load(GOLD)
    for tag in SSIM_TAGS:
        y = d[f"ssim_{tag}_ren"] * d[f"ssim_{tag}_mask"]
        np.testing.assert_allclose(P.create_window(), d[f"ssim_{tag}_window"], rtol=1e-6)
        ms, g_y = P.ms_ssim(d[f"ssim_{tag}_x"], y, data_range=1.0, weights=d[f"ssim_{tag}_weights"],
                            normalize=bool(d[f"ssim_{tag}_normalize"]), window=d[f"ssim_{tag}_window"], grad_out=d[f"ssim_{tag}_go"],
                            use_padding=bool(d[f"ssim_{tag}_pad"]))
        np.testing.assert_allclose(ms, d[f"ssim_{tag}_val"], rtol=1e-5)
        ref = d[f"ssim_{tag}_grad"]
        got = g_y * d[f"ssim_{tag}_mask"]                 # chain through ren * mask
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), (np.abs(got - ref).max(), np.abs(ref).max())


def test_ms_ssim_module_rejects_unbuilt_options():
    from self6dpp_b200.ssim import MS_SSIM, create_window
    assert MS_SSIM(use_padding=True).use_padding
    with pytest.raises(NotImplementedError):
        MS_SSIM(window_size=7)
    m = MS_SSIM(data_range=1.0, normalize=True, levels=3)
    assert len(m.weights) == 3 and abs(float(m.weights.sum()) - 1.0) < 1e-6
    d = np.load(GOLD)
    np.testing.assert_array_equal(create_window(11, 1.5).numpy(), d["ssim_a_window"])
    with pytest.raises(Exception):
        m(torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64))      # CPU tensors: no fallback


@pytest.mark.gpu
def test_ms_ssim_gpu_matches_reference_golden_and_is_reproducible():
    from self6dpp_b200.ssim import MS_SSIM
    dev = "cuda:0"
    d = np.load(GOLD)
    for tag in SSIM_TAGS:
        levels = int(d[f"ssim_{tag}_levels"])
        m = MS_SSIM(data_range=1.0, normalize=bool(d[f"ssim_{tag}_normalize"]), levels=levels if levels != 5 else None,
                    use_padding=bool(d[f"ssim_{tag}_pad"])).to(dev)
        np.testing.assert_allclose(m.weights.cpu().numpy(), d[f"ssim_{tag}_weights"], rtol=1e-6)
        mask = torch.tensor(d[f"ssim_{tag}_mask"], device=dev)
        outs = []
        for _ in range(2):
            ren = torch.tensor(d[f"ssim_{tag}_ren"], device=dev, requires_grad=True)
            val = m(torch.tensor(d[f"ssim_{tag}_x"], device=dev), ren * mask)
            (val * torch.tensor(d[f"ssim_{tag}_go"], device=dev)).sum().backward()
            outs.append((val.detach().clone(), ren.grad.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])       # bit-reproducible
        np.testing.assert_allclose(outs[0][0].cpu().numpy(), d[f"ssim_{tag}_val"], rtol=1e-5)
        ref = d[f"ssim_{tag}_grad"]
        got = outs[0][1].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), (np.abs(got - ref).max(), np.abs(ref).max())
        # forward without a gradient takes the path that stores no maps: same values
        with torch.no_grad():
            v2 = m(torch.tensor(d[f"ssim_{tag}_x"], device=dev), torch.tensor(d[f"ssim_{tag}_ren"], device=dev) * mask)
        assert torch.equal(v2, outs[0][0])


@pytest.mark.gpu
def test_ms_ssim_gpu_crop_size_against_oracle():
    """the loop's size: 32 x 3 x 256 x 256, 5 levels, normalize -- against the float64 oracle"""
    from oracle import photometric_oracle as P
    from self6dpp_b200.ssim import MS_SSIM
    g = torch.Generator().manual_seed(8)
    n = 8                                           # the oracle is numpy: 8 images keep it at a few seconds
    gt = torch.rand(n, 3, 256, 256, generator=g)
    ren = (gt + 0.1 * torch.randn(n, 3, 256, 256, generator=g)).clamp(0, 1)
    mask = (torch.rand(n, 1, 256, 256, generator=g) > 0.3).float()
    go = torch.rand(n, generator=g) + 0.5
    m = MS_SSIM(data_range=1.0, normalize=True).to("cuda:0")
    y = (ren * mask).to("cuda:0").requires_grad_(True)
    val = m((gt * mask).to("cuda:0"), y)
    (val * go.to("cuda:0")).sum().backward()
    ref_val, ref_g = P.ms_ssim((gt * mask).numpy(), (ren * mask).numpy(), data_range=1.0, normalize=True,
                               window=np.array(m._window), grad_out=go.numpy())
    np.testing.assert_allclose(val.detach().cpu().numpy(), ref_val, rtol=1e-5)
    got = y.grad.cpu().numpy()
    assert np.abs(got - ref_g).max() <= 2e-5 * np.abs(ref_g).max(), (np.abs(got - ref_g).max(), np.abs(ref_g).max())
