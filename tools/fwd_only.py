#!/usr/bin/env python
"""Debug/profiling driver: set up the cfg2 student pass once, then launch dibr_forward N times (L2 flushed between
launches) and print the mean kernel time.  Used under ncu (-k regex:dibr_forward):
    python tools/fwd_only.py [reps] [teacher]       # 'teacher' = the norm-only pass (D = 4)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200 import Renderer_dibr
from self6dpp_b200.bench_util import time_forward_kernel
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = ["norm"] if "teacher" in sys.argv else ["color", "depth", "mask", "norm", "prob"]
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
ren = Renderer_dibr(256, 256, "VertexColorBatch")
dev_in = {k: torch.tensor(student[k], device=dev) for k in ("Rs", "ts", "Ks")}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ms = time_forward_kernel(ren, dev_in, cur, mode, 256, flush, reps=reps)
print("forward kernel (%s) %.1f us" % ("+".join(mode), ms * 1e3), flush=True)
