#!/usr/bin/env python
"""CPU-side statistics of the cfg2 workload's per-tile face lists (how large LCAP / sub-lists must be)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from oracle import dibr_oracle as O
meshes, student, teacher = bench.workload(0)
H = W = bench.RES
m_ = 1000.0
ex = 0.02 * m_
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cnts, fronts, covfrac, softpix, pairs = [], [], [], [], []
for i in range(0, 32, 2):
    m = meshes[int(student["ids"][i])]
    v, f = torch.tensor(m["vertices"]), torch.tensor(m["faces"])
    cams = O.camera_params_from_RT_K(torch.tensor(student["Rs"][i:i + 1]), torch.tensor(student["ts"][i:i + 1]),
                                     torch.tensor(student["Ks"][i]), H, W, near=0.01, far=100.0)
    p3, p2, nz, _ = O.project(v, f, cams[0][0], cams[1][0], cams[2])
    p2 = p2[0].numpy().reshape(-1, 3, 2) * m_
    nzv = nz[0].numpy().reshape(-1)
    xmin, xmax = p2[:, :, 0].min(1) - ex, p2[:, :, 0].max(1) + ex
    ymin, ymax = p2[:, :, 1].min(1) - ex, p2[:, :, 1].max(1) + ex
    xs = m_ / W * (2 * np.arange(W) + 1 - W)
    ys = m_ / H * (H - 2 * np.arange(H) - 1)
    nt = W // tile
    for ty in range(nt):
        for tx in range(nt):
            x_lo, x_hi = xs[tx * tile], xs[tx * tile + tile - 1]
            y_hi, y_lo = ys[ty * tile], ys[ty * tile + tile - 1]
            hit = (xmin <= x_hi) & (xmax > x_lo) & (ymin <= y_hi) & (ymax > y_lo)
            if hit.any():
                cnts.append(int(hit.sum())); fronts.append(int((hit & (nzv >= 0)).sum()))
    fw = O.rasterize(W, H, p3, p2=torch.tensor(p2.reshape(1, -1, 6) / m_, dtype=torch.float32), nz=nz, attr=torch.ones(1, f.shape[0], 3)) if False else None
cnts = np.array(cnts); fronts = np.array(fronts)
print(f"tile {tile}: touched tiles/img {len(cnts)/16:.1f}  list mean {cnts.mean():.0f} p50 {np.median(cnts):.0f} p90 {np.percentile(cnts,90):.0f} p99 {np.percentile(cnts,99):.0f} max {cnts.max()}  front mean {fronts.mean():.0f} max {fronts.max()}")
