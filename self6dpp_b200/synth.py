"""Seeded synthetic meshes, poses and intrinsics shaped like the reference's workloads.

The reference renders BOP meshes (``models_scaled_f5k`` ~5k faces, /root/reference/ref/lm_full.py:26-31)
with LINEMOD intrinsics (ref/lm_full.py:106) or YCB-V intrinsics (ref/ycbv.py:89).  There is no
network for datasets, so tests and bench.py use these generators.
Everything is numpy + explicit seeds, float32 out.
"""
import numpy as np

K_LM = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], dtype=np.float32)
K_YCBV = np.array([[1066.778, 0.0, 312.9869], [0.0, 1067.487, 241.3109], [0.0, 0.0, 1.0]], dtype=np.float32)


def icosphere(level, radius=0.05, noise_sigma=0.0, seed=0):
    """Subdivided icosahedron: level 4 -> V=2562, F=5120 (cfg1); level 6 -> F=81920."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(level):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        edges_sorted = np.sort(edges, axis=1)
        uniq, inv = np.unique(edges_sorted, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        nf = len(f)
        m01, m12, m20 = base + inv[:nf], base + inv[nf:2 * nf], base + inv[2 * nf:]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    rng = np.random.default_rng(seed)
    r = radius + (noise_sigma * rng.standard_normal(len(v)) if noise_sigma > 0 else 0.0)
    verts = (v * np.reshape(r, (-1, 1))).astype(np.float32)
    return _finish(verts, f, seed)


def ellipsoid(n_lat, n_lon, radii=(0.05, 0.04, 0.03), noise_sigma=0.003, seed=0):
    """Closed lat-long grid ellipsoid with radial noise: F = 2*n_lon*(n_lat-1) triangles
    (e.g. n_lat=49, n_lon=52 -> 4992 faces; SURVEY.md cfg2: F in [4k,6k])."""
    rng = np.random.default_rng(seed)
    lat = np.linspace(0.0, np.pi, n_lat + 1)[1:-1]              # interior rings
    lon = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False)
    ring = np.stack([np.outer(np.sin(lat), np.cos(lon)), np.outer(np.sin(lat), np.sin(lon)),
                     np.outer(np.cos(lat), np.ones_like(lon))], axis=-1).reshape(-1, 3)
    v = np.concatenate([[[0, 0, 1.0]], ring, [[0, 0, -1.0]]], axis=0)
    nr = n_lat - 1
    idx = lambda r, c: 1 + r * n_lon + (c % n_lon)
    faces = []
    for c in range(n_lon):
        faces.append([0, idx(0, c), idx(0, c + 1)])
        faces.append([len(v) - 1, idx(nr - 1, c + 1), idx(nr - 1, c)])
    for r in range(nr - 1):
        for c in range(n_lon):
            a, b, cc, d = idx(r, c), idx(r, c + 1), idx(r + 1, c), idx(r + 1, c + 1)
            faces.append([a, cc, b])
            faces.append([b, cc, d])
    f = np.asarray(faces, dtype=np.int64)
    scale = 1.0 + noise_sigma / max(radii) * rng.standard_normal(len(v))
    verts = (v * np.asarray(radii)[None] * scale[:, None]).astype(np.float32)
    return _finish(verts, f, seed)


def _finish(verts, faces, seed):
    rng = np.random.default_rng(seed + 7919)
    # outward orientation (front faces must have normal_z >= 0 in view space, vcrender_batch.py:75)
    c = verts[faces].mean(axis=1)
    n = np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]])
    flip = (n * c).sum(1) < 0
    faces = faces.copy()
    faces[flip] = faces[flip][:, [0, 2, 1]]
    # shuffle face order so index order is not spatially sorted (exercises first-K / tie rules)
    faces = faces[rng.permutation(len(faces))]
    normals = verts / (np.linalg.norm(verts, axis=1, keepdims=True) + 1e-12)
    colors = rng.uniform(0.0, 1.0, size=verts.shape)
    return {"vertices": verts.astype(np.float32), "faces": faces.astype(np.int32),
            "colors": colors.astype(np.float32), "normals": normals.astype(np.float32)}


def lm13_meshes():
    """13 LINEMOD-shaped meshes, F in [4k,6k] (cfg2): lat-long ellipsoids, seeds 0..12."""
    out = []
    for i in range(13):
        n_lat = 46 + (i * 5) % 9        # 46..54
        n_lon = 46 + (i * 7) % 10       # 46..55
        radii = (0.035 + 0.004 * (i % 5), 0.03 + 0.005 * (i % 4), 0.025 + 0.006 * (i % 3))
        out.append(ellipsoid(n_lat, n_lon, radii=radii, noise_sigma=0.002, seed=i))
    return out


def random_rotations(n, seed):
    """Rotations from normalised N(0,1) quaternions."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                  2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                  2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(n, 3, 3)
    return R.astype(np.float32), q.astype(np.float32)


def crop_K(K, center_uv, scale, out_res):
    """Intrinsics of a square crop of side ``scale`` around ``center_uv`` resized to ``out_res``
    (the get_K_crop_resize semantics the reference uses for ROI rendering)."""
    r = out_res / float(scale)
    Kc = np.array(K, dtype=np.float64).copy()
    Kc[0, 2] -= center_uv[0] - scale / 2.0
    Kc[1, 2] -= center_uv[1] - scale / 2.0
    Kc[:2] *= r
    return Kc.astype(np.float32)


def roi_batch(meshes, batch, res=256, seed=0, fill=(0.4, 0.7), K=K_LM):
    """cfg2/cfg5 sample set: per sample a mesh id (i mod len), a random pose with the object
    somewhere in a 640x480 frame, and the crop-K that makes it fill 40-70% of a res x res crop."""
    rng = np.random.default_rng(seed)
    Rs, _ = random_rotations(batch, seed + 1)
    ts = np.zeros((batch, 3), dtype=np.float32)
    Ks = np.zeros((batch, 3, 3), dtype=np.float32)
    ids = np.arange(batch) % len(meshes)
    for i in range(batch):
        m = meshes[ids[i]]
        radius = float(np.linalg.norm(m["vertices"], axis=1).max())
        z = rng.uniform(0.5, 1.1)
        u = rng.uniform(120, 520)
        v = rng.uniform(100, 380)
        x = (u - K[0, 2]) * z / K[0, 0]
        y = (v - K[1, 2]) * z / K[1, 1]
        ts[i] = (x, y, z)
        diam_px = 2 * radius * K[0, 0] / z
        scale = diam_px / rng.uniform(*fill)
        Ks[i] = crop_K(K, (u, v), scale, res)
    return {"ids": ids.astype(np.int64), "Rs": Rs, "ts": ts, "Ks": Ks}


def upstream_grads(shape_im, shape_prob, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape_im).astype(np.float32),
            rng.standard_normal(shape_prob).astype(np.float32))
