"""cProfile of the bench step on the GPU box: where does HOST time go?"""
import cProfile, pstats, sys, os, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200 import Renderer_dibr
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
ren = Renderer_dibr(256, 256, "VertexColorBatch")
B = 32
g_color = torch.randn(B, 256, 256, 3, device=dev); g_prob = torch.randn(B, 256, 256, device=dev); g_depth = torch.randn(B, 256, 256, device=dev)
dev_in = {k: torch.tensor(student[k], device=dev) for k in ("Rs", "ts", "Ks")}
dev_te = {k: torch.tensor(teacher[k], device=dev) for k in ("Rs", "ts")}
def step():
    Rs = dev_in["Rs"].detach().clone().requires_grad_(True); ts = dev_in["ts"].detach().clone().requires_grad_(True)
    ret = ren.render_batch(Rs, ts, cur, Ks=dev_in["Ks"], width=256, height=256, mode=["color", "depth", "mask", "norm", "prob"])
    with torch.no_grad():
        ren.render_batch(dev_te["Rs"], dev_te["ts"], cur, Ks=dev_in["Ks"], width=256, height=256, mode=["norm"])
    torch.autograd.backward([ret["color"], ret["prob"], ret["depth"]], [g_color, g_prob, g_depth])
    return Rs.grad
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host ms/step (enqueue only): %.3f   incl. drain: %.3f" % ((t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
