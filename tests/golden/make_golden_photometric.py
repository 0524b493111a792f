"""Generates tests/golden/ref_photometric.npz by running the REFERENCE's own photometric code on CPU through autograd:
  lib/torch_utils/color/lab.py (rgb_to_lab, normalize_lab) inside the expression of
  core/self6dpp/engine/self_engine_utils.py:745-773 (fvcore's smooth_l1_loss(beta=0, reduction="sum") = sum |a - b|).

Run here (the reference tree does not exist on the GPU box):  python tests/golden/make_golden_photometric.py
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def crops(g, n, h, w, black=True):
    """a 'real' crop, a 'rendered' crop with an exactly-black background (as the rasteriser leaves it) and a 0/1 mask"""
    gt = torch.rand(n, 3, h, w, generator=g)
    ren = (gt + 0.15 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    inside = (((yy - h / 2) / (0.4 * h)) ** 2 + ((xx - w / 2) / (0.35 * w)) ** 2 < 1).float()[None, None]
    mask = inside * (torch.rand(n, 1, h, w, generator=g) > 0.1).float()
    if black:
        ren = ren * inside            # background pixels of the rendered crop are exactly 0
    ren.view(-1)[:4] = torch.tensor([0.04045, 0.0404, 0.041, 1.0])     # both sides of the sRGB knee
    return gt, ren, mask


def golden_lab():
    from lib.torch_utils.color.lab import rgb_to_lab, normalize_lab
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (n, h, w), no_l, black in (("a", (2, 24, 32), True, False), ("b", (3, 17, 23), False, False),
                                       ("c", (2, 16, 16), True, True)):
        gt, ren, mask = crops(g, n, h, w, black=black)
        ren.requires_grad_(True)
        lab_gt = normalize_lab(rgb_to_lab(gt[:, [2, 1, 0]]))                      # self_engine_utils.py:746-750
        lab_ren = normalize_lab(rgb_to_lab(ren[:, [2, 1, 0]].contiguous()))
        if no_l:
            loss = torch.abs(lab_gt[:, 1:] * mask - lab_ren[:, 1:] * mask).sum() / max(1, mask.sum())
        else:
            loss = torch.abs(lab_gt * mask - lab_ren * mask).sum() / max(1, mask.sum())
        (loss * 0.2).backward()                                                     # LAB_LW = 0.2 in the configs
        out.update({f"lab_{tag}_gt": gt.numpy(), f"lab_{tag}_ren": ren.detach().numpy(), f"lab_{tag}_mask": mask.numpy(),
                    f"lab_{tag}_no_l": np.array(int(no_l)), f"lab_{tag}_loss": loss.detach().numpy(),
                    f"lab_{tag}_grad": ren.grad.numpy()})
        print("lab", tag, float(loss), "NaN grads:", int(torch.isnan(ren.grad).sum()))
    return out


if __name__ == "__main__":
    out = golden_lab()
    np.savez_compressed(os.path.join(OUT, "ref_photometric.npz"), **out)


def golden_ms_ssim():
    """core/self6dpp/losses/ssim.py: the MS_SSIM module itself, as self_engine.py:352 builds it, and two variations"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ssim", "/root/reference/core/self6dpp/losses/ssim.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(91)
    out = {}
    cases = (("a", (2, 3, 192, 208), dict(data_range=1.0, normalize=True), None),         # the loop's configuration, 5 levels
             ("b", (1, 3, 97, 113), dict(data_range=1.0, normalize=True, levels=3), None),  # odd sizes: padded pooling
             ("c", (2, 1, 64, 48), dict(data_range=1.0, normalize=False, levels=2, channel=1), None),
             ("d", (2, 3, 96, 80), dict(data_range=1.0, normalize=True, use_padding=True), None),     # zero-padded windows, 5 levels
             ("e", (1, 3, 45, 37), dict(data_range=1.0, normalize=False, use_padding=True, levels=3), None))   # odd sizes, levels below 11 px
    for tag, shape, kw, _ in cases:
        n, c, h, w = shape
        gt, ren, mask = crops(g, n, h, w, black=True)
        gt, ren = gt[:, :c], ren[:, :c].detach()
        X = gt * mask
        ren.requires_grad_(True)
        m = mod.MS_SSIM(**kw)
        val = m(X, ren * mask)                                                        # self_engine_utils.py:779-784
        go = torch.rand(n, generator=g) + 0.5
        (val * go).sum().backward()
        out.update({f"ssim_{tag}_x": X.numpy(), f"ssim_{tag}_ren": ren.detach().numpy(), f"ssim_{tag}_mask": mask.numpy(),
                    f"ssim_{tag}_levels": np.array(len(m.weights)), f"ssim_{tag}_weights": m.weights.numpy(),
                    f"ssim_{tag}_normalize": np.array(int(kw["normalize"])), f"ssim_{tag}_pad": np.array(int(kw.get("use_padding", False))), f"ssim_{tag}_window": m.window[0, 0, 0].numpy(),
                    f"ssim_{tag}_val": val.detach().numpy(), f"ssim_{tag}_go": go.numpy(), f"ssim_{tag}_grad": ren.grad.numpy()})
        print("ms_ssim", tag, val.detach().numpy())
    return out


if __name__ == "__main__":
    out.update(golden_ms_ssim())
    np.savez_compressed(os.path.join(OUT, "ref_photometric.npz"), **out)
