"""ORACLE #2 -- test infrastructure only.  An INDEPENDENT float64 PyTorch transcription of the DIB-R
forward (vectorised over pixels, all faces at once -- small cases only) whose backward comes from
torch autograd rather than from the hand-derived K3/K4 formulas.  Purpose: pin oracle/dibr_oracle.c
(the recalled kaolin v0.1 kernels) from a second direction:
  * forward values (idx, weights, im, improb) must agree with the C oracle;
  * autograd(d loss / d points2d, d loss / d attr) must agree with the C oracle's analytic K3+K4,
    i.e. the recalled backward really is the derivative of the recalled forward with the
    winning face / selected case held fixed.
Reference seams: rasterizer.py:36-70 (prepare_tfpoints), :73-220 (buffers), :222-291 (which inputs
get gradients).  Not fast, not used by the product.
"""
import torch


def pixel_centres(width, height, multiplier, dtype=torch.float64):
    w = torch.arange(width, dtype=torch.float64)
    h = torch.arange(height, dtype=torch.float64)
    x0 = 1.0 * multiplier / width * (2 * w + 1 - width)
    y0 = 1.0 * multiplier / height * (height - 2 * h - 1)
    return x0.to(dtype), y0.to(dtype)


def rasterize(width, height, points3d_fx9, points2d_fx6, normalz_f, attr_fx3d,
              expand=0.02, knum=30, multiplier=1000, delta=7000):
    """Single image.  Differentiable w.r.t. points2d_fx6 and attr_fx3d.
    Returns dict(im HxWxD, improb HxW, imidx HxW (long, 0 = none), imwei HxWx3)."""
    dt = torch.float64
    eps = 1e-15
    F = points2d_fx6.shape[0]
    D = attr_fx3d.shape[1] // 3
    p3 = points3d_fx9.to(dt)
    p2m = float(multiplier) * points2d_fx6.to(dt)
    attr = attr_fx3d.to(dt)
    x0, y0 = pixel_centres(width, height, multiplier, dt)
    X = x0.view(1, 1, width).expand(1, height, width)
    Y = y0.view(1, height, 1).expand(1, height, width)

    v = p2m.view(F, 3, 2)
    pmin, pmax = v.min(dim=1)[0].detach(), v.max(dim=1)[0].detach()
    e = expand * multiplier

    def inside(lo, hi):
        return ((X >= lo[:, 0].view(F, 1, 1)) & (X < hi[:, 0].view(F, 1, 1)) &
                (Y >= lo[:, 1].view(F, 1, 1)) & (Y < hi[:, 1].view(F, 1, 1)))

    ax, ay, bx, by, cx, cy = [p2m[:, i].view(F, 1, 1) for i in range(6)]
    m, p, n, q = bx - ax, by - ay, cx - ax, cy - ay
    s, t = X - ax, Y - ay
    k1, k2, k3 = s * q - n * t, m * t - s * p, m * q - n * p
    w1, w2 = k1 / (k3 + eps), k2 / (k3 + eps)
    w0 = 1 - w1 - w2
    z0 = w0 * p3[:, 2].view(F, 1, 1) + w1 * p3[:, 5].view(F, 1, 1) + w2 * p3[:, 8].view(F, 1, 1)
    ok = inside(pmin, pmax) & (normalz_f.to(dt).view(F, 1, 1) >= 0) & (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
    zz = torch.where(ok, z0.detach(), torch.full_like(z0, -float("inf")))
    zbest, fbest = zz.max(dim=0)                      # first maximal index (torch.max doc) ...
    # ... but make the tie rule explicit: lowest index among equal z
    tie = (zz == zbest.unsqueeze(0)) & ok
    fbest = torch.where(tie.any(0), tie.to(torch.int64).argmax(dim=0), fbest)
    covered = zbest > -1000.0
    imidx = torch.where(covered, fbest + 1, torch.zeros_like(fbest))

    g = fbest.unsqueeze(0)
    W0 = torch.gather(w0, 0, g)[0]
    W1 = torch.gather(w1, 0, g)[0]
    W2 = torch.gather(w2, 0, g)[0]
    imwei = torch.stack([W0, W1, W2], dim=-1) * covered.unsqueeze(-1)
    a = attr[fbest]                                  # H x W x 3D
    im = (W0.unsqueeze(-1) * a[..., 0:D] + W1.unsqueeze(-1) * a[..., D:2 * D] + W2.unsqueeze(-1) * a[..., 2 * D:3 * D])
    im = im * covered.unsqueeze(-1)

    # ---- soft silhouette over uncovered pixels: first knum faces (ascending) whose expanded bbox holds the pixel
    hit = inside(pmin - e, pmax + e) & (~covered).unsqueeze(0)
    rank = torch.cumsum(hit.to(torch.int64), dim=0)
    keep = hit & (rank <= knum)
    big = float(4 * multiplier * multiplier)
    dists = []
    for i in range(3):
        x1, y1 = v[:, i, 0].view(F, 1, 1), v[:, i, 1].view(F, 1, 1)
        x2, y2 = v[:, (i + 1) % 3, 0].view(F, 1, 1), v[:, (i + 1) % 3, 1].view(F, 1, 1)
        A, B, C = y2 - y1, x1 - x2, x2 * y1 - x1 * y2
        up, down = A * X + B * Y + C, A * A + B * B
        x3 = (B * B * X - A * B * Y - A * C) / (down + eps)
        y3 = (A * A * Y - A * B * X - B * C) / (down + eps)
        direct = (x3 - x1) * (x3 - x2) + (y3 - y1) * (y3 - y2)
        dists.append(torch.where(direct.detach() > 0, torch.full_like(up, big), up * up / (down + eps)))
    for i in range(3):
        x1, y1 = v[:, i, 0].view(F, 1, 1), v[:, i, 1].view(F, 1, 1)
        dists.append((X - x1) ** 2 + (Y - y1) ** 2)
    dstack = torch.stack(dists, dim=0)               # 6 x F x H x W
    case = dstack.detach().argmin(dim=0, keepdim=True)   # first minimal index
    dmin = torch.gather(dstack, 0, case)[0]
    prob = torch.exp(-(float(delta) * dmin / multiplier / multiplier))
    one_minus = torch.where(keep, 1.0 - prob, torch.ones_like(prob))
    improb = 1.0 - torch.prod(one_minus, dim=0)
    improb = torch.where(covered, torch.ones_like(improb), improb)
    return dict(im=im, improb=improb, imidx=imidx, imwei=imwei, case=case[0], keep=keep)
