"""Known-answer tests that pin the CPU oracle.  The reference has
no tests for this path ("parity unpinned"), so these hand-derived answers are the pins."""
import math

import numpy as np
import pytest
import torch

from oracle import dibr_oracle as O

M = 1000.0


def px(w, W):
    return M / W * (2 * w + 1 - W)


def py(h, H):
    return M / H * (H - 2 * h - 1)


def tri(p2d, z=(-1.0, -1.0, -1.0), nz=1.0, dt=torch.float64):
    """one face from NDC (unscaled) 2D points; returns (p3d 1x1x9, p2d 1x1x6, normalz 1x1x1)"""
    p2 = torch.tensor(p2d, dtype=dt).view(1, 1, 6)
    p3 = torch.zeros(1, 1, 9, dtype=dt)
    p3[0, 0, 2], p3[0, 0, 5], p3[0, 0, 8] = z
    return p3, p2, torch.full((1, 1, 1), nz, dtype=dt)


def cat_faces(*fs):
    return tuple(torch.cat([f[i] for f in fs], dim=1) for i in range(3))


def attrs(F, D=4, dt=torch.float64, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(1, F, 3 * D, generator=g, dtype=dt)
    a[..., D - 1::D] = 1.0
    return a


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_kat1_single_triangle_8x8(dt):
    H = W = 8
    # a triangle with vertices at exact fractions of NDC, counter-clockwise (k3 > 0)
    p3, p2, nz = tri([-0.5, -0.5, 0.8, -0.5, -0.5, 0.7], dt=dt)
    a = attrs(1, dt=dt)
    fw = O.rasterize(W, H, p3, p2, nz, a)
    ax, ay, bx, by, cx, cy = [v * M for v in (-0.5, -0.5, 0.8, -0.5, -0.5, 0.7)]
    k3 = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay)
    for h in range(H):
        for w in range(W):
            x0, y0 = px(w, W), py(h, H)
            k1 = (x0 - ax) * (cy - ay) - (cx - ax) * (y0 - ay)
            k2 = (bx - ax) * (y0 - ay) - (x0 - ax) * (by - ay)
            w1, w2 = k1 / k3, k2 / k3
            w0 = 1 - w1 - w2
            inbb = (min(ax, bx, cx) <= x0 < max(ax, bx, cx)) and (min(ay, by, cy) <= y0 < max(ay, by, cy))
            inside = inbb and w0 >= 0 and w1 >= 0 and w2 >= 0
            assert (fw["imidx"][0, h, w, 0].item() == 1.0) == inside, (h, w)
            if inside:
                got = fw["imwei"][0, h, w].double()
                assert torch.allclose(got, torch.tensor([w0, w1, w2], dtype=torch.float64), atol=1e-6)
                exp_im = w0 * a[0, 0, 0:4] + w1 * a[0, 0, 4:8] + w2 * a[0, 0, 8:12]
                assert torch.allclose(fw["im"][0, h, w].double(), exp_im.double(), atol=1e-6)
                assert abs(fw["im"][0, h, w, 3].item() - 1.0) < 1e-6          # hardmask ~ 1
                assert fw["improb"][0, h, w, 0].item() == 1.0
    assert fw["imidx"].sum() > 0


def test_kat1b_soft_prob_value():
    """uncovered pixel next to a vertical edge: prob = exp(-delta * d^2 / m^2) with d the
    perpendicular distance in multiplier units."""
    H = W = 8
    # right edge x = 0.01 (NDC); pixel column 4 has centre x0 = 125 > 10 -> outside, distance 115
    p3, p2, nz = tri([-0.9, -0.9, 0.01, -0.9, 0.01, 0.9])
    fw = O.rasterize(W, H, p3, p2, nz, attrs(1), expand=0.2)
    h, w = 5, 4                                  # y0 = -375 within the edge's y-range [-900, 900]
    x0 = px(w, W)
    assert fw["imidx"][0, h, w, 0] == 0
    d2 = (x0 - 10.0) ** 2
    assert math.isclose(fw["improb"][0, h, w, 0].item(), 1 - (1 - math.exp(-7000 * d2 / M / M)), rel_tol=1e-12)
    assert fw["probcase"][0, h, w, 0].item() == 2.0          # edge 1 (b->c), stored +1
    assert fw["probface"][0, h, w, 0].item() == 1.0
    # outside the expanded bbox (expand 0.2 -> 200 units): column 6 has x0 = 625 > 10 + 200
    assert fw["improb"][0, h, 6, 0].item() == 0.0


def test_kat1c_vertex_case():
    H = W = 64
    p3, p2, nz = tri([-0.9, -0.9, 0.0, -0.9, -0.9, 0.0])
    fw = O.rasterize(W, H, p3, p2, nz, attrs(1), expand=0.5)
    h, w = 60, 32        # x0 = 15.625, y0 = -890.625 ; nearest feature is vertex b = (0, -900)
    x0, y0 = px(w, W), py(h, H)
    assert (x0, y0) == (15.625, -890.625)
    d2 = (x0 - 0.0) ** 2 + (y0 + 900.0) ** 2
    # feet on edges a->b (y = -900) and b->c (x + y = -900) both fall beyond b -> sentinel 4 m^2,
    # so vertex b wins (case index 4, stored +1)
    assert fw["imidx"][0, h, w, 0] == 0
    assert fw["probcase"][0, h, w, 0].item() == 5.0
    assert math.isclose(fw["improb"][0, h, w, 0].item(), math.exp(-7000 * d2 / M / M), rel_tol=1e-9)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_kat2_equal_z_lower_index_wins_and_nearer_wins(dt):
    H = W = 8
    big = [-0.9, -0.9, 0.9, -0.9, -0.9, 0.9]
    f0, f1 = tri(big, z=(-1, -1, -1), dt=dt), tri(big, z=(-1, -1, -1), dt=dt)
    p3, p2, nz = cat_faces(f0, f1)
    fw = O.rasterize(W, H, p3, p2, nz, attrs(2, dt=dt))
    cov = fw["imidx"] > 0
    assert cov.any() and (fw["imidx"][cov] == 1.0).all()          # tie -> first face
    f1n = tri(big, z=(-0.5, -0.5, -0.5), dt=dt)                   # larger z = nearer (imdep starts at -1000)
    p3, p2, nz = cat_faces(f0, f1n)
    fw = O.rasterize(W, H, p3, p2, nz, attrs(2, dt=dt))
    assert (fw["imidx"][cov] == 2.0).all()
    # a face behind z = -1000 never beats the initial depth (-1000 exactly is rounding-dependent)
    far = tri(big, z=(-1001.0, -1001.0, -1001.0), dt=dt)
    fw = O.rasterize(W, H, *far, attrs(1, dt=dt))
    assert (fw["imidx"] == 0).all()


def test_kat3_back_face_invisible_but_soft():
    H = W = 8
    p3, p2, nz = tri([-0.5, -0.5, 0.75, -0.5, -0.5, 0.75], nz=-1.0)
    fw = O.rasterize(W, H, p3, p2, nz, attrs(1))
    assert (fw["imidx"] == 0).all() and (fw["im"] == 0).all()
    assert fw["improb"].max() > 0.5                               # K2 does not cull (frozen choice)
    p3, p2, nz = tri([-0.5, -0.5, 0.75, -0.5, -0.5, 0.75], nz=0.0)
    assert (O.rasterize(W, H, p3, p2, nz, attrs(1))["imidx"] > 0).any()   # normalz == 0 is front


def test_kat4_first_k_faces_only():
    H = W = 8
    K = 5
    # 8 far-away slivers whose expanded bboxes all contain pixel (4,4); only the first K count
    faces = []
    for i in range(8):
        off = 0.3 + 0.01 * i
        faces.append(tri([off, off, off + 0.05, off, off, off + 0.05]))
    p3, p2, nz = cat_faces(*faces)
    fw = O.rasterize(W, H, p3, p2, nz, attrs(8), expand=0.5, knum=K)
    h, w = 4, 4
    assert fw["imidx"][0, h, w, 0] == 0
    assert fw["probface"][0, h, w].tolist() == [1.0, 2.0, 3.0, 4.0, 5.0]
    sub = O.rasterize(W, H, p3[:, :K], p2[:, :K], nz[:, :K], attrs(8)[:, :K], expand=0.5, knum=30)
    assert fw["improb"][0, h, w, 0] == sub["improb"][0, h, w, 0]


def test_kat5_shared_edge_pixel_goes_to_lower_index():
    H = W = 4
    # two triangles sharing the diagonal through pixel centres (x0 == y0 on the diagonal h = 3 - w)
    t0 = tri([-1.0, -1.0, 1.0, -1.0, 1.0, 1.0])      # below the diagonal y = x
    t1 = tri([-1.0, -1.0, 1.0, 1.0, -1.0, 1.0])      # above
    p3, p2, nz = cat_faces(t0, t1)
    fw = O.rasterize(W, H, p3, p2, nz, attrs(2))
    for w in range(W):
        h = H - 1 - w                                # y0 == x0
        assert fw["imidx"][0, h, w, 0] == 1.0
    p3, p2, nz = cat_faces(t1, t0)
    fw = O.rasterize(W, H, p3, p2, nz, attrs(2))
    for w in range(W):
        assert fw["imidx"][0, H - 1 - w, w, 0] == 1.0


def test_kat6_attribute_gradient_is_weight_sum():
    H = W = 16
    g = torch.Generator().manual_seed(1)
    p3, p2, nz = cat_faces(tri([-0.8, -0.7, 0.6, -0.5, -0.1, 0.9]), tri([-0.2, -0.9, 0.9, 0.1, 0.3, 0.8], z=(-2, -2, -2)))
    a = attrs(2)
    fw = O.rasterize(W, H, p3, p2, nz, a)
    gI = torch.randn(1, H, W, 4, generator=g, dtype=torch.float64)
    _, dc = O.rasterize_backward(fw, gI, torch.zeros(1, H, W, 1, dtype=torch.float64))
    exp = torch.zeros_like(dc)
    for h in range(H):
        for w in range(W):
            f = int(fw["imidx"][0, h, w, 0]) - 1
            if f >= 0:
                for i in range(3):
                    exp[0, f, i * 4:(i + 1) * 4] += gI[0, h, w] * fw["imwei"][0, h, w, i]
    assert torch.allclose(dc, exp, rtol=1e-12, atol=1e-14)


def test_kat7_finite_differences_with_fixed_visibility():
    H = W = 16
    g = torch.Generator().manual_seed(2)
    p3, p2, nz = cat_faces(tri([-0.8, -0.7, 0.6, -0.5, -0.1, 0.9]), tri([-0.2, -0.9, 0.9, 0.1, 0.3, 0.8], z=(-2, -2, -2)))
    a = attrs(2)
    gI = torch.randn(1, H, W, 4, generator=g, dtype=torch.float64)
    gP = torch.randn(1, H, W, 1, generator=g, dtype=torch.float64)
    fw = O.rasterize(W, H, p3, p2, nz, a, expand=0.1)
    dp2, _ = O.rasterize_backward(fw, gI, gP)

    def loss(p2x):
        f = O.rasterize(W, H, p3, p2x, nz, a, expand=0.1)
        same = (f["imidx"] == fw["imidx"])
        return ((f["im"] * gI) * same).sum() + ((f["improb"] * gP) * same).sum(), same.all()

    h = 1e-7
    for j in range(6):
        for f in range(2):
            d = torch.zeros_like(p2)
            d[0, f, j] = h
            lp, okp = loss(p2 + d)
            lm, okm = loss(p2 - d)
            if not (okp and okm):
                continue
            fd = (lp - lm) / (2 * h)
            assert math.isclose(fd.item(), dp2[0, f, j].item(), rel_tol=2e-4, abs_tol=1e-4), (f, j, fd.item(), dp2[0, f, j].item())


def test_kat8_projection_lands_on_pixel_grid():
    """A camera-space point (X,Y,Z) must land at u = fx X/Z + px, v = fy Y/Z + py in the convention
    'pixel (w,h) has centre (w+0.5, h+0.5)' once mapped through x0,y0 (perspective.py:122-129)."""
    H, W = 480, 640
    K = torch.tensor([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], dtype=torch.float64)
    R = torch.eye(3, dtype=torch.float64)[None]
    t = torch.tensor([[0.0, 0.0, 0.0]], dtype=torch.float64)
    cams = O.camera_params_from_RT_K(R, t, K, H, W)
    P = torch.tensor([[0.03, -0.02, 0.7], [0.1, 0.05, 0.9], [-0.08, 0.06, 0.5]], dtype=torch.float64)
    faces = torch.tensor([[0, 1, 2]])
    _, p2, _, _ = O.project(P, faces, cams[0][0], cams[1][0], cams[2])
    for i in range(3):
        u = K[0, 0] * P[i, 0] / P[i, 2] + K[0, 2]
        v = K[1, 1] * P[i, 1] / P[i, 2] + K[1, 2]
        xn, yn = p2[0, 0, 2 * i].item(), p2[0, 0, 2 * i + 1].item()
        # x0 = (2w+1-W)/W at pixel centre w+0.5  ->  u = (xn*W + W)/2 ;  y0 = (H-2h-1)/H -> v = (H - yn*H)/2
        assert math.isclose((xn * W + W) / 2, u.item(), rel_tol=1e-12)
        assert math.isclose((H - yn * H) / 2, v.item(), rel_tol=1e-12)


def test_pixel_centre_is_rounded_once_from_double():
    fw = O.rasterize(7, 3, *tri([-2.0, -2.0, 2.0, -2.0, -2.0, 2.0], dt=torch.float32), attrs(1, dt=torch.float32))
    assert (fw["imidx"] > 0).any()
    # weights reproduce fp32(double expr): w1 = (x0 - ax)/4000 ... check one pixel
    x0 = np.float32(1.0 * 1000 / 7 * (2 * 2 + 1 - 7))
    w1 = np.float32(np.float64(np.float32(x0 + np.float32(2000.0)) * np.float32(4000.0)) / (np.float64(np.float32(16000000.0)) + 1e-15))
    assert fw["imwei"][0, 1, 2, 1].item() == float(w1)
