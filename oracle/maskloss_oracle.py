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


def soft_dice_loss(probs, labels, smooth=0.0, eps=1e-7, reduction="mean", grad_out=None):
    """mask_losses.py:444-463 in float64, with d (sum(loss * grad_out)) / d probs.  Pinned by tests/golden/ref_diceloss.npz,
    which the reference's OWN function produced (make_golden.py --diceloss)."""
    p = np.asarray(probs, dtype=np.float64)
    l = np.asarray(labels, dtype=np.float64)
    num = l.shape[0]
    m1, m2 = p.reshape(num, -1), l.reshape(num, -1)
    inter = (m1 * m2).sum(1) + smooth
    den = m1.sum(1) + m2.sum(1) + smooth + eps
    score = 2.0 * inter / den                                                                      # :456
    if reduction == "mean":
        loss, c = 1 - score.sum() / num, np.full(num, 1.0 / num)
    elif reduction == "sum":
        loss, c = (1 - score).sum(), np.ones(num)
    else:
        loss, c = 1 - score, np.ones(num)
    go = np.ones_like(c) if grad_out is None else np.broadcast_to(np.asarray(grad_out, dtype=np.float64).reshape(-1), c.shape)
    grad = -(c * go)[:, None] * 2.0 * (m2 * den[:, None] - inter[:, None]) / (den[:, None] ** 2)
    return loss, grad.reshape(p.shape)


def norm_loss(out_norm, gt_norm, mask, with_l1=True, with_cs=True):
    """NORMLoss (core/self6dpp/losses/vf_norm_loss.py:56-103) in float64 with d loss / d out_norm.  Pinned by
    tests/golden/ref_normloss.npz, which the reference's OWN module produced (make_golden.py --normloss)."""
    o, g, m = (np.asarray(v, dtype=np.float64) for v in (out_norm, gt_norm, mask))
    a, b = m * o, m * g                                                                            # :80-81
    loss, grad_a = 0.0, np.zeros_like(a)
    if with_l1:                                                                                    # :84-85, mean over all elements
        loss += np.abs(a - b).mean()
        grad_a += np.sign(a - b) / a.size
    if with_cs:                                                                                    # :90-103
        nfg = float((m != 0).sum())
        na = np.maximum(np.sqrt((a * a).sum(1, keepdims=True)), 1e-8)
        nb = np.maximum(np.sqrt((b * b).sum(1, keepdims=True)), 1e-8)
        cs = ((a / na) * (b / nb)).sum(1, keepdims=True)
        loss += (m * (1 - cs)).sum() / nfg
        grad_a += -(m / nfg) * ((b / nb) / na - cs * a / (na * na))
    return loss, grad_a * m
