"""Host-side cost of RenderSession.step (enqueue only) on the cfg2 batch: the step must be enqueued faster than the GPU
executes it (0.38 ms), also with 8 ranks sharing the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200.session import RenderSession
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
B, RES = bench.BATCH, bench.RES
sess = RenderSession(models, B, RES, RES, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
gc = torch.randn(B, RES, RES, 3, device=dev, generator=g); gp = torch.randn(B, RES, RES, device=dev, generator=g); gd = torch.randn(B, RES, RES, device=dev, generator=g)
def step(up):
    sess.step(student["Rs"], student["ts"], student["Ks"], cur, teacher["Rs"], teacher["ts"], grad_color=gc, grad_prob=gp, grad_depth=gd, upload=up, download=up)
for up in (False, True):
    for _ in range(20): step(up)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): step(up)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("upload/download=%s: host enqueue %.3f ms per step, with drain %.3f ms per step" % (up, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
