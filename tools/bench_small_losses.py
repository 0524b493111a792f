#!/usr/bin/env python
"""Soft dice loss on [32,1,256,256] and NORMLoss on [32,3,64,64], fwd+bwd: ours vs the reference's expressions
(mask_losses.py:444-463, vf_norm_loss.py:56-103) restated in torch eager on the same GPU.  CUDA events, median of 20."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from self6dpp_b200.losses import soft_dice_loss, NORMLoss

DEV = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


probs = torch.rand(32, 1, 256, 256, generator=g).to(DEV)
labels = (torch.rand(32, 1, 256, 256, generator=g) > 0.6).float().to(DEV)
def dice_ours():
    p = probs.clone().requires_grad_(True); soft_dice_loss(p, labels, eps=0.002).backward()
def dice_torch():
    p = probs.clone().requires_grad_(True)
    m1, m2 = p.view(32, -1), labels.view(32, -1)
    score = 2.0 * (m1 * m2).sum(1) / (m1.sum(1) + m2.sum(1) + 0.002)
    (1 - score.sum() / 32).backward()
print(json.dumps({"config": "soft dice loss fwd+bwd on [32,1,256,256]", "ms": timed(dice_ours), "torch_expression_ms": timed(dice_torch)}))

o = torch.randn(32, 3, 64, 64, generator=g).to(DEV)
gt = F.normalize(torch.randn(32, 3, 64, 64, generator=g), dim=1).to(DEV)
m = (torch.rand(32, 1, 64, 64, generator=g) > 0.5).float().to(DEV)
mod = NORMLoss()
def norm_ours():
    x = o.clone().requires_grad_(True); mod(x, gt, m).backward()
def norm_torch():
    x = o.clone().requires_grad_(True)
    a, b = m * x, m * gt
    n = (m != 0).sum().item()
    (F.l1_loss(a, b) + (m.squeeze(1) * (1 - F.cosine_similarity(a, b, dim=1))).sum() / n).backward()
print(json.dumps({"config": "NORMLoss (L1 + masked cosine) fwd+bwd on [32,3,64,64]", "ms": timed(norm_ours), "torch_expression_ms": timed(norm_torch)}))
