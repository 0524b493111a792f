import sys, os, ctypes
sys.path.insert(0, "/root/repo")
os.environ["DIBR_B200_LIB"] = "/root/repo/self6dpp_b200/lib/libdibr_b200_stats.so"
import subprocess
csrc = "/root/repo/self6dpp_b200/csrc"
subprocess.run(["nvcc","-O3","-std=c++17","-gencode","arch=compute_100a,code=sm_100a","-Xcompiler","-fPIC","--expt-relaxed-constexpr","-shared","-cudart","static","-DDIBR_NND_STATS","-o",os.environ["DIBR_B200_LIB"]]+[os.path.join(csrc,f) for f in ("dibr_abi.cu","dibr_setup.cu","dibr_forward.cu","dibr_backward.cu","dibr_nnd.cu","dibr_nnd_grid.cu")], check=True)
import torch, bench
from self6dpp_b200 import Renderer_dibr, _lib
from self6dpp_b200.nndistance import backproject_th, compact_valid_points, nnd_padded
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
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
lib.dibr_debug_nnd_stats(buf, 1)
d1, d2, i1, i2 = nnd_padded(p1, c1, p2, c2)
lib.dibr_debug_nnd_stats(buf, 1)
print("queries by last ring r (0..7+):", list(buf[:8]), "candidate points visited per query: %.1f" % (buf[8] / max(1, sum(buf[:8]))))
print("sample 0 extents:", (p2[0,:int(c2[0])].max(0)[0]-p2[0,:int(c2[0])].min(0)[0]).tolist(), "median nn dist (mm): %.3f" % (d1[0,:int(c1[0])].sqrt().median().item()*1e3))
os.remove(os.environ["DIBR_B200_LIB"])
