"""Throughput of the chamfer nearest-neighbour op on cfg2-shaped clouds (32 samples, rendered 256x256 depth vs a
perturbed target), and of the reference's own CPU implementation (oracle/_ref) on one sample."""
import sys, os, json, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from self6dpp_b200 import Renderer_dibr
from self6dpp_b200.nndistance import depth_bp_chamfer_loss, backproject_th, compact_valid_points, nnd_padded
from oracle import nnd_oracle as N
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
p1, c1 = compact_valid_points(backproject_th(d_t, K)); p2, c2 = compact_valid_points(backproject_th(d_s, K))
pairs = float((c1.double() * c2.double()).sum()) * 2
print("points per sample: real %.0f rendered %.0f; pair tests per step %.3g" % (c1.float().mean(), c2.float().mean(), pairs))
def run_fwd():
    return nnd_padded(p1, c1, p2, c2)
for _ in range(3): run_fwd()
torch.cuda.synchronize(); ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run_fwd(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = statistics.median(ts)
ds = d_s.clone().requires_grad_(True)
def run_loss():
    ds.grad = None
    l, _ = depth_bp_chamfer_loss(ds, d_t, K, 0.05); l.backward()
for _ in range(3): run_loss()
torch.cuda.synchronize(); tl = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run_loss(); b.record(); torch.cuda.synchronize(); tl.append(a.elapsed_time(b))
# CPU reference: the compiled nnd_cpu.cpp on sample 0
ref = N.ref_module()
n, m = int(c1[0]), int(c2[0])
x1, x2 = p1[0:1, :n].cpu().contiguous(), p2[0:1, :m].cpu().contiguous()
cpu = None
if ref is not None:
    rd1, rd2 = torch.zeros(1, n), torch.zeros(1, m); ri1, ri2 = torch.zeros(1, n, dtype=torch.int32), torch.zeros(1, m, dtype=torch.int32)
    t0 = time.perf_counter(); ref.nnd_forward(x1, x2, rd1, rd2, ri1, ri2); cpu = time.perf_counter() - t0
print(json.dumps({"op": "chamfer nnd forward (both directions), 32 samples", "ms": ms, "pair_tests_per_s": pairs / (ms * 1e-3),
                  "fp32_flop_per_s": pairs * 8 / (ms * 1e-3), "loss_fwd_bwd_ms": statistics.median(tl),
                  "reference_cpu_one_sample_s": cpu, "reference_cpu_pairs_per_s": (2.0 * n * m / cpu) if cpu else None}))
