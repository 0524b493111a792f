import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench
from self6dpp_b200 import Renderer_dibr
from self6dpp_b200.nndistance import depth_bp_chamfer_loss
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
ren = Renderer_dibr(256, 256, "VertexColorBatch")
K = torch.tensor(student["Ks"], device=dev)
with torch.no_grad():
    d_s = ren.render_batch(torch.tensor(student["Rs"], device=dev), torch.tensor(student["ts"], device=dev), cur, Ks=K, width=256, height=256, mode=["depth"])["depth"]
    d_t = ren.render_batch(torch.tensor(teacher["Rs"], device=dev), torch.tensor(teacher["ts"], device=dev), cur, Ks=K, width=256, height=256, mode=["depth"])["depth"]
ds = d_s.clone().requires_grad_(True)
for it in range(4):
    ds.grad = None
    l, _ = depth_bp_chamfer_loss(ds, d_t, K, 0.05); l.backward()
    torch.cuda.synchronize()
print("ok")
