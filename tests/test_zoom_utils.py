"""Crop & resize of rendered images (core/utils/zoom_utils.py: batch_crop_resize = ROIAlign(out, 1.0, 0, aligned=True)).
Golden vectors come from torchvision's CPU roi_align, the op behind detectron2's ROIAlign (tests/golden/
make_golden_roialign.py).  Tolerances: 1e-5 relative to the largest entry, forward and backward (fp32 sums of <= a few
dozen taps; the reference's own backward is an order-dependent atomicAdd scatter)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_roialign.npz")
TAGS = "abc"


def _close(name, got, want, tol=1e-5):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = np.abs(got - want).max() if got.size else 0.0
    assert err <= tol * max(np.abs(want).max() if want.size else 0.0, 1e-30), (name, err)


def test_roialign_oracle_matches_torchvision_golden():
    from oracle import roialign_oracle as O
    d = np.load(GOLD)
    for tag in TAGS:
        oh, ow, aligned, sr = (int(v) for v in d[f"{tag}_cfg"])
        y, gx = O.roi_align(d[f"{tag}_x"], d[f"{tag}_rois"], oh, ow, 1.0, sr, bool(aligned), grad_out=d[f"{tag}_go"])
        _close(tag + " y", y, d[f"{tag}_y"], 2e-6)
        _close(tag + " gx", gx, d[f"{tag}_gx"], 2e-6)
    # a roi with x2 < x1 yields zeros and no gradient (ceil of a negative size: no samples)
    assert np.all(d["a_y"][4] == 0)


def test_deepim_boxes_restated():
    """zoom_utils.py:6-77 by hand: centre (10, 20), rendered box 6..16 x 14..24, observed 8..18 x 10..22, lamb 1.4, 4:3 output"""
    from self6dpp_b200.zoom_utils import deepim_boxes
    ren = torch.tensor([[6.0, 14.0, 16.0, 24.0]])
    obs = torch.tensor([[8.0, 10.0, 18.0, 22.0]])
    c = torch.tensor([[10.0, 20.0]])
    boxes, ratios = deepim_boxes(ren, c, obs, lamb=1.4, outHW=(480, 640))
    xdist, ydist = 8.0, 10.0                                     # max |cx - x|, max |cy - y| over both boxes
    crop_h = max(xdist / (640 / 480), ydist) * 2 * 1.4
    crop_w = crop_h * 640 / 480
    np.testing.assert_allclose(boxes.numpy(), [[10 - crop_w / 2, 20 - crop_h / 2, 10 + crop_w / 2, 20 + crop_h / 2]], rtol=1e-6)
    np.testing.assert_allclose(ratios.numpy(), [[640 / crop_w, 480 / crop_h]], rtol=1e-6)
    b2, _ = deepim_boxes(ren, c, None, lamb=1.0, outHW=(64, 64))
    np.testing.assert_allclose(b2.numpy(), [[4.0, 14.0, 16.0, 26.0]], rtol=1e-6)


def test_batch_crop_resize_refuses_cpu_and_unknown_modes():
    from self6dpp_b200.zoom_utils import batch_crop_resize
    x, r = torch.rand(1, 1, 8, 8), torch.tensor([[0.0, 1, 1, 5, 5]])
    with pytest.raises(Exception):
        batch_crop_resize(x, r, 4, 4)
    with pytest.raises(ValueError):
        batch_crop_resize(x, r, 4, 4, interpolation="cubic")
    with pytest.raises(Exception):
        batch_crop_resize(x, r, 4, 4, interpolation="nearest")      # CPU tensors: no fallback for RoIPool either


@pytest.mark.gpu
def test_roialign_gpu_matches_golden_and_is_reproducible():
    from self6dpp_b200.zoom_utils import roi_align
    dev = "cuda:0"
    d = np.load(GOLD)
    for tag in TAGS:
        oh, ow, aligned, sr = (int(v) for v in d[f"{tag}_cfg"])
        outs = []
        for layout in ("bchw", "bhwc", "bchw"):
            x = torch.tensor(d[f"{tag}_x"], device=dev)
            if layout == "bhwc":                                  # the renderer's layout, viewed as BCHW: read in place
                x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
            x.requires_grad_(True)
            y = roi_align(x, torch.tensor(d[f"{tag}_rois"], device=dev), (oh, ow), 1.0, sr, bool(aligned))
            (y * torch.tensor(d[f"{tag}_go"], device=dev)).sum().backward()
            assert x.grad.stride() == x.stride()
            outs.append((y.detach().clone(), x.grad.clone()))
            _close(f"{tag} {layout} y", y.detach().cpu().numpy(), d[f"{tag}_y"])
            _close(f"{tag} {layout} gx", x.grad.cpu().numpy(), d[f"{tag}_gx"])
        for a, b in zip(outs[0], outs[2]):
            assert torch.equal(a, b)                              # bit-reproducible
        for a, b in zip(outs[0], outs[1]):
            assert torch.equal(a, b)                              # and independent of the memory layout


@pytest.mark.gpu
def test_roialign_gpu_edge_cases():
    from oracle import roialign_oracle as O
    from self6dpp_b200.zoom_utils import batch_crop_resize
    dev = "cuda:0"
    g = torch.Generator().manual_seed(3)
    # no rois: empty output, zero gradient
    x = torch.randn(2, 3, 9, 11, generator=g).to(dev).requires_grad_(True)
    y = batch_crop_resize(x, torch.zeros(0, 5, device=dev), 4, 4)
    assert y.shape == (0, 3, 4, 4)
    (y.sum() + 0 * x.sum()).backward()
    assert float(x.grad.abs().sum()) == 0
    # more rois on one tile than the backward's shared-memory list holds (512: several flushes), many on the same pixels,
    # channel count not a multiple of 4
    R = 1300
    xs = torch.randn(2, 5, 10, 12, generator=g)
    c = torch.rand(R, 2, generator=g) * torch.tensor([12.0, 10.0])
    s = torch.rand(R, 2, generator=g) * 6 + 0.2
    rois = torch.cat([torch.randint(0, 2, (R, 1), generator=g).float(), c - s, c + s], dim=1)
    go = torch.randn(R, 5, 3, 3, generator=g)
    x = xs.to(dev).requires_grad_(True)
    y = batch_crop_resize(x, rois.to(dev), 3, 3)
    (y * go.to(dev)).sum().backward()
    ry, rgx = O.roi_align(xs.numpy(), rois.numpy(), 3, 3, grad_out=go.numpy())
    _close("many rois y", y.detach().cpu().numpy(), ry)
    _close("many rois gx", x.grad.cpu().numpy(), rgx)
    # a non-dense view (a crop of a larger tensor) is accepted (copied)
    big = torch.randn(1, 2, 16, 16, generator=g).to(dev)
    v = big[:, :, 2:12, 3:13]
    r = torch.tensor([[0.0, 1.0, 1.5, 8.0, 7.5]], device=dev)
    assert torch.equal(batch_crop_resize(v, r, 4, 4), batch_crop_resize(v.contiguous(), r, 4, 4))


@pytest.mark.gpu
def test_batch_crop_resize_full_size_against_torchvision_on_the_gpu():
    """the loop's size: 32 full-frame renders (480x640, channels-last as the renderer writes them) -> 256x256 colour crops
    and 64x64 normal crops, forward and backward, against torchvision's CUDA roi_align (atomicAdd backward)"""
    tv = pytest.importorskip("torchvision")
    from self6dpp_b200.zoom_utils import batch_crop_resize
    dev = "cuda:0"
    g = torch.Generator().manual_seed(21)
    B, H, W = 32, 480, 640
    ren = torch.rand(B, H, W, 3, generator=g).to(dev)
    c = torch.stack([torch.rand(B, generator=g) * W, torch.rand(B, generator=g) * H], dim=1)
    half = torch.rand(B, 1, generator=g) * 120 + 30
    rois = torch.cat([torch.arange(B).float().view(-1, 1), c - half, c + half], dim=1).to(dev)
    for out_res in (256, 64):
        go = torch.randn(B, 3, out_res, out_res, generator=g).to(dev)
        x = ren.clone().requires_grad_(True)
        y = batch_crop_resize(x.permute(0, 3, 1, 2), rois, out_res, out_res)
        (y * go).sum().backward()
        x2 = ren.clone().requires_grad_(True)
        y2 = tv.ops.roi_align(x2.permute(0, 3, 1, 2).contiguous(), rois, (out_res, out_res), 1.0, 0, True)
        (y2 * go).sum().backward()
        # torchvision's CUDA build contracts the sample coordinate into FMAs, its CPU build (the golden vectors) and this
        # kernel do not: coordinates near 640 differ by one fp32 ulp (6e-5), and a random image has unit slopes
        _close(f"{out_res} y", y.detach().cpu().numpy(), y2.detach().cpu().numpy(), 2.5e-4)
        _close(f"{out_res} gx", x.grad.cpu().numpy(), x2.grad.cpu().numpy(), 2.5e-4)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_roialign_oracle_matches_torchvision_cpu_on_random_rois(seed):
    """beyond the committed vectors: random rois (inside, straddling the border, tiny, larger than the image), both `aligned`
    settings, fixed and adaptive sampling -- the restatement against torchvision's CPU op wherever that package is present"""
    tv = pytest.importorskip("torchvision")
    from oracle import roialign_oracle as O
    g = torch.Generator().manual_seed(100 + seed)
    N, C, H, W = 2, 2, 11 + seed, 14
    x = torch.randn(N, C, H, W, generator=g, requires_grad=True)
    R = 7
    c = torch.rand(R, 2, generator=g) * torch.tensor([W + 4.0, H + 4.0]) - 2.0
    s = torch.cat([torch.rand(R - 2, 2, generator=g) * 8 + 0.05, torch.tensor([[0.02, 0.03], [20.0, 17.0]])])
    rois = torch.cat([torch.randint(0, N, (R, 1), generator=g).float(), c - s, c + s], dim=1)
    for aligned, sr, (oh, ow) in ((True, 0, (4, 5)), (False, 0, (3, 3)), (True, 3, (2, 4))):
        go = torch.randn(R, C, oh, ow, generator=g)
        x.grad = None
        y = tv.ops.roi_align(x, rois, (oh, ow), 1.0, sr, aligned)
        (y * go).sum().backward()
        ry, rgx = O.roi_align(x.detach().numpy(), rois.numpy(), oh, ow, 1.0, sr, aligned, grad_out=go.numpy())
        _close(f"y aligned={aligned} sr={sr}", ry, y.detach().numpy(), 5e-6)
        _close(f"gx aligned={aligned} sr={sr}", rgx, x.grad.numpy(), 5e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_roialign_gpu_random_rois_against_oracle(seed):
    """the gather backward's sample-range estimate under stress: rois far outside, straddling every border, a 0.01-pixel
    roi (pitch below the estimate's threshold: all samples examined), one much larger than the image; odd sizes"""
    from oracle import roialign_oracle as O
    from self6dpp_b200.zoom_utils import roi_align
    g = torch.Generator().manual_seed(200 + seed)
    N, C, H, W = 2, 3, 37 + seed, 70 + 3 * seed                    # wider than one 64-pixel tile of the backward
    xs = torch.randn(N, C, H, W, generator=g)
    R = 9
    c = torch.rand(R, 2, generator=g) * torch.tensor([W + 10.0, H + 10.0]) - 5.0
    s = torch.cat([torch.rand(R - 3, 2, generator=g) * 20 + 0.3, torch.tensor([[0.005, 0.004], [0.3, 25.0], [90.0, 60.0]])])
    rois = torch.cat([torch.randint(0, N, (R, 1), generator=g).float(), c - s, c + s], dim=1)
    for aligned, sr, (oh, ow) in ((True, 0, (6, 5)), (False, 0, (3, 4)), (True, 2, (4, 4))):
        go = torch.randn(R, C, oh, ow, generator=g)
        x = xs.to("cuda:0").requires_grad_(True)
        y = roi_align(x, rois.to("cuda:0"), (oh, ow), 1.0, sr, aligned)
        (y * go.to("cuda:0")).sum().backward()
        ry, rgx = O.roi_align(xs.numpy(), rois.numpy(), oh, ow, 1.0, sr, aligned, grad_out=go.numpy())
        _close(f"y aligned={aligned} sr={sr}", y.detach().cpu().numpy(), ry)
        _close(f"gx aligned={aligned} sr={sr}", x.grad.cpu().numpy(), rgx)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_roipool_nearest_against_torchvision(seed):
    """batch_crop_resize(interpolation="nearest") = torchvision.ops.RoIPool(output_size, 1.0) (zoom_utils.py:91-92), against
    torchvision's own CPU op on random rois (inside, straddling the border, outside, malformed, smaller than the output grid):
    the forward bit for bit (a maximum is a selection), the gradient to 1e-6 (torchvision adds with atomics in arbitrary order,
    ours is a fixed-order gather) and bit-reproducible, in both memory layouts."""
    from torchvision.ops import roi_pool as tv_roi_pool
    from self6dpp_b200.zoom_utils import batch_crop_resize
    dev = "cuda:0"
    g = torch.Generator().manual_seed(40 + seed)
    B, C, H, W, R = 3, 4, 37, 45, 60
    x = torch.randn(B, C, H, W, generator=g)
    c = torch.rand(R, 2, generator=g) * torch.tensor([W + 10.0, H + 10.0]) - 5.0
    s = torch.rand(R, 2, generator=g) * 14 + 0.3
    rois = torch.cat([torch.randint(0, B, (R, 1), generator=g).float(), c - s, c + s], dim=1)
    rois[3, 1:] = torch.tensor([20.0, 20.0, 10.0, 5.0])             # x2 < x1, y2 < y1: becomes 1 x 1
    rois[4, 1:] = torch.tensor([-30.0, -30.0, -20.0, -25.0])        # outside: empty bins, zeros, no gradient
    rois[5, 1:] = torch.tensor([10.2, 11.7, 12.4, 13.1])            # smaller than the 6 x 5 output grid
    oh, ow = 6, 5
    go = torch.randn(R, C, oh, ow, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = tv_roi_pool(xr, rois, (oh, ow), 1.0)
    (yr * go).sum().backward()
    outs = []
    for layout in ("bchw", "bhwc", "bchw"):
        xd = x.to(dev)
        if layout == "bhwc":
            xd = xd.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        xd.requires_grad_(True)
        y = batch_crop_resize(xd, rois.to(dev), oh, ow, interpolation="nearest")
        (y * go.to(dev)).sum().backward()
        assert xd.grad.stride() == xd.stride()
        assert torch.equal(y.detach().cpu(), yr.detach()), layout
        _close("roipool gx " + layout, xd.grad.cpu().numpy(), xr.grad.numpy(), 1e-6)
        outs.append((y.detach().clone(), xd.grad.clone()))
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    y0 = batch_crop_resize(x.to(dev), torch.zeros(0, 5, device=dev), oh, ow, interpolation="nearest")
    assert y0.shape == (0, C, oh, ow)
