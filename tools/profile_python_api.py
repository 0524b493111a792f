#!/usr/bin/env python
"""Debug: where the host time of the drop-in Python API goes (cProfile over 200 steps of render_batch x2 + backward)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200 import Renderer_dibr
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
B, RES = bench.BATCH, bench.RES
ren = Renderer_dibr(RES, RES, "VertexColorBatch")
g = torch.Generator(device=dev).manual_seed(0)
gc, gp, gd = torch.randn(B, RES, RES, 3, device=dev, generator=g), torch.randn(B, RES, RES, device=dev, generator=g), torch.randn(B, RES, RES, device=dev, generator=g)
Ks = torch.tensor(student["Ks"], device=dev)
tR, tt = torch.tensor(teacher["Rs"], device=dev), torch.tensor(teacher["ts"], device=dev)
def step():
    Rs = torch.tensor(student["Rs"], device=dev, requires_grad=True) if False else sR.detach().requires_grad_(True)
    ts = sT.detach().requires_grad_(True)
    ret = ren.render_batch(Rs, ts, cur, Ks=Ks, width=RES, height=RES, mode=["color", "depth", "mask", "norm", "prob"])
    with torch.no_grad():
        ren.render_batch(tR, tt, cur, Ks=Ks, width=RES, height=RES, mode=["norm"])
    torch.autograd.backward([ret["color"], ret["prob"], ret["depth"]], [gc, gp, gd])
    return Rs.grad
sR, sT = torch.tensor(student["Rs"], device=dev), torch.tensor(student["ts"], device=dev)
for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per step %.1f us (enqueue only), %.1f us with the final sync" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
