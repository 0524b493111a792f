"""How long does dibr_forward take with the soft-silhouette phase disabled (knum = 0)?  Upper bound of what
restructuring phase D can win."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200 import Renderer_dibr, fused, _lib
from self6dpp_b200.bench_util import time_forward_kernel
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
ren = Renderer_dibr(256, 256, "VertexColorBatch")
dev_in = {k: torch.tensor(student[k], device=dev) for k in ("Rs", "ts", "Ks")}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
mode = ["color", "depth", "mask", "norm", "prob"]
print("full      : %.1f us" % (1e3 * time_forward_kernel(ren, dev_in, cur, mode, 256, flush)))
orig = fused.DEFAULT_KNUM
import self6dpp_b200.fused as F
# knum=0 through build_meta default
old_build = F.build_meta
def bm(*a, **k):
    k["knum"] = 0
    return old_build(*a, **k)
F.build_meta = bm
print("knum = 0  : %.1f us" % (1e3 * time_forward_kernel(ren, dev_in, cur, mode, 256, flush)))
