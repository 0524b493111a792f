"""ORACLE (test infrastructure only): float64 numpy restatement of the photometric losses on the rendered colour crop.

lab_l1_loss:  core/self6dpp/engine/self_engine_utils.py:745-773 over lib/torch_utils/color/lab.py:16-82 (rgb_to_lab,
normalize_lab) and lib/torch_utils/color/xyz.py:28-30, with the gradient w.r.t. the rendered image derived by hand in the
order autograd applies it -- including the 0 * inf = NaN that pow(x, 1/3)'s backward produces at x == 0.
(fvcore.nn.smooth_l1_loss with beta=0 is |input - target|; fvcore is a pip dependency of the reference, not in its tree.)

Pinned by tests/golden/ref_photometric.npz, produced by the reference's OWN lab.py / ssim.py through torch autograd
(tests/golden/make_golden_photometric.py)."""
import numpy as np

_M = np.array([[0.412453, 0.357580, 0.180423],
               [0.212671, 0.715160, 0.072169],
               [0.019334, 0.119193, 0.950227]], dtype=np.float64)      # xyz.py:28-30
_WHITE = np.array([0.95047, 1.0, 1.08883], dtype=np.float64)          # lab.py:50
_MIN = np.array([0.0, -110.0, -110.0]).reshape(1, 3, 1, 1)            # lab.py:78-79
_MAX = np.array([100.0, 110.0, 110.0]).reshape(1, 3, 1, 1)


def _lab_norm(rgb):
    """rgb: (N,3,H,W) float64 in R,G,B order -> (normalised lab, intermediates)"""
    with np.errstate(invalid="ignore"):
        lin = np.where(rgb > 0.04045, np.power((rgb + 0.055) / 1.055, 2.4), rgb / 12.92)          # lab.py:43-45
    xyz = np.einsum("ij,njhw->nihw", _M, lin)
    n = xyz / _WHITE.reshape(1, 3, 1, 1)
    with np.errstate(invalid="ignore"):
        f = np.where(n > 0.008856, np.power(n, 1.0 / 3.0), 7.787 * n + 4.0 / 29.0)                # lab.py:55-57
    L = 116.0 * f[:, 1] - 16.0
    a = 500.0 * (f[:, 0] - f[:, 1])
    b = 200.0 * (f[:, 1] - f[:, 2])
    lab = np.stack([L, a, b], axis=1)
    return (lab - _MIN) / (_MAX - _MIN), (lin, n)


def lab_l1_loss(gt, ren, mask=None, no_l=False, bgr=True):
    """returns (loss, d loss / d ren); gt, ren (N,3,H,W), mask (N,1,H,W) or None"""
    gt = np.asarray(gt, dtype=np.float64)
    ren = np.asarray(ren, dtype=np.float64)
    if bgr:
        gt, ren = gt[:, ::-1], ren[:, ::-1]                                                      # self_engine_utils.py:746,749
    m = np.ones((gt.shape[0], 1) + gt.shape[2:]) if mask is None else np.asarray(mask, dtype=np.float64).reshape(
        gt.shape[0], 1, gt.shape[2], gt.shape[3])
    lab_g, _ = _lab_norm(gt)
    lab_r, (lin, n) = _lab_norm(ren)
    diff = lab_g * m - lab_r * m
    if no_l:
        diff = diff.copy()
        diff[:, 0] = 0.0                                                                          # :751-757 uses [:, 1:]
    den = max(1.0, float(m.sum()))
    loss = np.abs(diff).sum() / den
    # backward
    g_lab = -np.sign(diff) * m / den / (_MAX - _MIN)
    gL, ga, gb = g_lab[:, 0], g_lab[:, 1], g_lab[:, 2]
    g_f = np.stack([500.0 * ga, 116.0 * gL - 500.0 * ga + 200.0 * gb, -200.0 * gb], axis=1)
    hi = n > 0.008856
    with np.errstate(divide="ignore", invalid="ignore"):
        g_n = np.where(hi, g_f, 0.0) * ((1.0 / 3.0) * np.power(n, 1.0 / 3.0 - 1.0)) + np.where(hi, 0.0, g_f) * 7.787
    g_xyz = g_n / _WHITE.reshape(1, 3, 1, 1)
    g_lin = np.einsum("ij,nihw->njhw", _M, g_xyz)
    hi = ren > 0.04045
    with np.errstate(invalid="ignore"):
        g_rgb = np.where(hi, g_lin, 0.0) * (2.4 * np.power((ren + 0.055) / 1.055, 1.4)) / 1.055 + np.where(hi, 0.0, g_lin) / 12.92
    if bgr:
        g_rgb = g_rgb[:, ::-1]
    return loss, np.ascontiguousarray(g_rgb)
