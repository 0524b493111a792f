"""ORACLE (test infrastructure only): float64 numpy restatement of the reference's re-weighted BCE on probabilities,
core/self6dpp/losses/mask_losses.py:63-108 (weighted_ex_loss_probs), with its gradient w.r.t. probs.
Pinned by tests/golden/ref_maskloss.npz, which the reference's OWN function produced (make_golden.py --maskloss)."""
import numpy as np


def weighted_ex_loss_probs(probs, target, weight=None):
    p64 = np.asarray(probs, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    w = np.ones_like(p64) if weight is None else np.broadcast_to(np.asarray(weight, dtype=np.float64), p64.shape)
    lo, hi = np.float64(np.float32(1e-7)), np.float64(np.float32(1.0) - np.float32(1e-7))       # torch clamps in fp32
    p = np.clip(p64, lo, hi)
    pos, neg = t > 0, t == 0                                                                        # :70-71
    npos, nneg = int(pos.sum()), int(neg.sum())
    loss = 0.0
    grad = np.zeros_like(p64)
    inside = (p64 >= lo) & (p64 <= hi)
    if npos > 0:                                                                                    # :102-103
        loss += (-t[pos] * np.log(p[pos]) * w[pos]).sum() / npos
        grad[pos] = np.where(inside[pos], -t[pos] * w[pos] / p[pos] / npos, 0.0)
    if nneg > 0:                                                                                    # :104-105
        loss += (-np.log(1 - p[neg]) * w[neg]).sum() / nneg
        grad[neg] = np.where(inside[neg], w[neg] / (1 - p[neg]) / nneg, 0.0)
    return loss, grad
