#!/usr/bin/env python
"""Debug: the cfg4 step (B=16, 480x640, 100k-face meshes) under `ncu --metrics gpu__time_duration.sum`: which kernel takes what."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from self6dpp_b200 import Renderer_dibr, synth
DEV = torch.device("cuda:0")
H, W = 480, 640
big = synth.ellipsoid(225, 224, radii=(0.06, 0.05, 0.045), noise_sigma=0.001, seed=5)       # the mesh of tools/bench_configs.py cfg4
mb = [{"vertices": torch.tensor(big["vertices"], device=DEV), "colors": torch.tensor(big["colors"], device=DEV),
       "normals": torch.tensor(big["normals"], device=DEV), "faces": torch.tensor(big["faces"], device=DEV, dtype=torch.int32)}]
print("faces", big["faces"].shape[0])
B = 16; Rb, _ = synth.random_rotations(B, 6); tb = np.tile(np.array([[0.0, 0.0, 0.9]], np.float32), (B, 1))
bren = Renderer_dibr(H, W, "VertexColorBatch")
g = torch.Generator().manual_seed(1)
gbc, gbp, gbd = torch.randn(B, H, W, 3, generator=g).to(DEV), torch.randn(B, H, W, generator=g).to(DEV), torch.randn(B, H, W, generator=g).to(DEV)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    Rs = torch.tensor(Rb, device=DEV, requires_grad=True); ts = torch.tensor(tb, device=DEV, requires_grad=True)
    ret = bren.render_batch(Rs, ts, mb * B, Ks=torch.tensor(synth.K_YCBV, device=DEV), width=W, height=H, mode=["color", "depth", "mask", "prob"])
    torch.autograd.backward([ret["color"], ret["prob"], ret["depth"]], [gbc, gbp, gbd])
torch.cuda.synchronize()
