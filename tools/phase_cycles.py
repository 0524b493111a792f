"""Debug: per-phase cycles (thread 0 of every touched tile CTA) of dibr_forward_kernel, from a -DDIBR_PHASE_TIMING build:
  cd self6dpp_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DDIBR_PHASE_TIMING -shared -cudart static -o ../lib/libdibr_b200_timing.so dibr_abi.cu dibr_setup.cu dibr_forward.cu dibr_backward.cu"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self6dpp_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libdibr_b200_timing.so")
import torch, bench
from self6dpp_b200 import Renderer_dibr
from self6dpp_b200.bench_util import time_forward_kernel
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
ren = Renderer_dibr(256, 256, "VertexColorBatch")
dev_in = {k: torch.tensor(student[k], device=dev) for k in ("Rs", "ts", "Ks")}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 8)()
lib.dibr_debug_phase_cycles(buf, 1)
reps = 10
ms = time_forward_kernel(ren, dev_in, cur, ["color", "depth", "mask", "norm", "prob"], 256, flush, reps=reps)
lib.dibr_debug_phase_cycles(buf, 1)
n = buf[7]
names = ["setup", "scan(fill_list)", "raster", "resolve", "soft", "-", "-", "CTAs"]
print("kernel %.1f us; touched CTAs per launch %.0f" % (ms * 1e3, n / (reps + 4)))
tot = sum(buf[:5])
for k in range(5):
    print("%-16s %8.0f cycles/CTA  %5.1f%%" % (names[k], buf[k] / max(n, 1), 100.0 * buf[k] / max(tot, 1)))
