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


# ------------------------------------------------------------------------------------------------------------------
# MS-SSIM: core/self6dpp/losses/ssim.py:13-160 (create_window, _gaussian_filter, ssim, ms_ssim; use_padding False or True)
# ------------------------------------------------------------------------------------------------------------------
def create_window(window_size=11, sigma=1.5):
    """ssim.py:13-30 (float32 arithmetic like torch's, then promoted)"""
    coords = np.arange(window_size, dtype=np.float32) - np.float32(window_size // 2)
    g = np.exp(-(coords ** 2) / np.float32(2 * sigma ** 2)).astype(np.float32)
    g = g / g.sum(dtype=np.float32)
    return g.astype(np.float64)


def _filt(img, win, pad=0):
    """separable correlation over the last two axes, valid or zero padded by `pad` on every side (ssim.py:33-55)"""
    k = len(win)
    if pad:
        img = np.pad(img, [(0, 0)] * (img.ndim - 2) + [(pad, pad), (pad, pad)])
    H, W = img.shape[-2:]
    out = np.zeros(img.shape[:-1] + (W - k + 1,))
    for i in range(k):
        out += win[i] * img[..., :, i:i + W - k + 1]
    out2 = np.zeros(img.shape[:-2] + (H - k + 1, W - k + 1))
    for i in range(k):
        out2 += win[i] * out[..., i:i + H - k + 1, :]
    return out2


def _filt_T(m, win, pad=0):
    """adjoint of _filt: (.., H-k+1+2 pad, W-k+1+2 pad) -> (.., H, W)"""
    if pad:
        full = _filt_T(m, win)
        return full[..., pad:full.shape[-2] - pad, pad:full.shape[-1] - pad]
    k = len(win)
    Ho, Wo = m.shape[-2:]
    mid = np.zeros(m.shape[:-2] + (Ho + k - 1, Wo))
    for i in range(k):
        mid[..., i:i + Ho, :] += win[i] * m
    out = np.zeros(m.shape[:-2] + (Ho + k - 1, Wo + k - 1))
    for i in range(k):
        out[..., :, i:i + Wo] += win[i] * mid
    return out


def _pool(img):
    """F.avg_pool2d(kernel 2, stride 2, padding (H % 2, W % 2)), zeros counted (ssim.py:143-145)"""
    H, W = img.shape[-2:]
    ph, pw = H % 2, W % 2
    p = np.pad(img, [(0, 0)] * (img.ndim - 2) + [(ph, ph), (pw, pw)])
    Ho, Wo = (H + 2 * ph - 2) // 2 + 1, (W + 2 * pw - 2) // 2 + 1
    p = p[..., :2 * Ho, :2 * Wo]
    return 0.25 * (p[..., 0::2, 0::2] + p[..., 0::2, 1::2] + p[..., 1::2, 0::2] + p[..., 1::2, 1::2])


def _pool_T(g, H, W):
    ph, pw = H % 2, W % 2
    Ho, Wo = g.shape[-2:]
    up = np.zeros(g.shape[:-2] + (2 * Ho, 2 * Wo))
    for a in (0, 1):
        for b in (0, 1):
            up[..., a::2, b::2] = 0.25 * g
    full = np.zeros(g.shape[:-2] + (H + 2 * ph, W + 2 * pw))
    hh, ww = min(full.shape[-2], 2 * Ho), min(full.shape[-1], 2 * Wo)
    full[..., :hh, :ww] = up[..., :hh, :ww]
    return full[..., ph:ph + H, pw:pw + W]


def ms_ssim(X, Y, data_range=1.0, weights=(0.0448, 0.2856, 0.3001, 0.2363, 0.1333), normalize=False, window=None,
            grad_out=None, use_padding=False):
    """returns (ms [N], d sum(grad_out * ms) / d Y or None)"""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    win = create_window() if window is None else np.asarray(window, dtype=np.float64)
    pad = len(win) // 2 if use_padding else 0                                                     # ssim.py:42-45
    w = np.asarray(np.asarray(weights, dtype=np.float32), dtype=np.float64)
    L = len(w)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    xs, ys, keep = [X], [Y], []
    cs_vals, ssim_vals = [], []
    for l in range(L):
        x, y = xs[-1], ys[-1]
        mu1, mu2 = _filt(x, win, pad), _filt(y, win, pad)
        e11, e22, e12 = _filt(x * x, win, pad), _filt(y * y, win, pad), _filt(x * y, win, pad)
        s11, s22, s12 = e11 - mu1 ** 2, e22 - mu2 ** 2, e12 - mu1 * mu2
        dcs, dl = s11 + s22 + C2, mu1 ** 2 + mu2 ** 2 + C1
        cs = (2 * s12 + C2) / dcs
        lum = (2 * mu1 * mu2 + C1) / dl
        ssim_vals.append((lum * cs).mean(axis=(1, 2, 3)))
        cs_vals.append(cs.mean(axis=(1, 2, 3)))
        keep.append((mu1, mu2, cs, lum, dcs, dl))
        xs.append(_pool(x))
        ys.append(_pool(y))
    half = 1.0
    if normalize:                                                                                 # ssim.py:150-152
        ssim_vals = [(v + 1) / 2 for v in ssim_vals]
        cs_vals = [(v + 1) / 2 for v in cs_vals]
        half = 0.5
    last = ssim_vals[-1] ** w[-1]
    ms = np.ones_like(last)
    for l in range(L - 1):                                                                        # ssim.py:153-156
        ms = ms * (cs_vals[l] ** w[l]) * last
    if grad_out is None:
        return ms, None
    go = np.asarray(grad_out, dtype=np.float64)
    g_next = None
    for l in range(L - 1, -1, -1):
        x, y = xs[l], ys[l]
        mu1, mu2, cs, lum, dcs, dl = keep[l]
        npx = cs[0].size
        g_mu = (2 * mu2 * cs - 2 * mu1) / dcs
        g_22 = -cs / dcs
        g_12 = 2.0 / dcs
        if l == L - 1:
            g_mu = lum * g_mu + cs * (2 * (mu1 - lum * mu2) / dl)
            g_22 = lum * g_22
            g_12 = lum * g_12
            s = ms * (L - 1) * w[l] / ssim_vals[l] * half / npx
        else:
            s = ms * w[l] / cs_vals[l] * half / npx
        s = (s * go).reshape(-1, 1, 1, 1)
        g = s * (_filt_T(g_mu, win, pad) + 2 * y * _filt_T(g_22, win, pad) + x * _filt_T(g_12, win, pad))
        if g_next is not None:
            g = g + _pool_T(g_next, x.shape[-2], x.shape[-1])
        g_next = g
    return ms, g_next
