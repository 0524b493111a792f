#!/usr/bin/env python
"""Debug: time dibr_backward_faces of the cfg2 student pass alone (L2 flushed before every call)."""
import ctypes, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200 import _lib
from self6dpp_b200.session import RenderSession
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
B, RES = bench.BATCH, bench.RES
sess = RenderSession(models, B, RES, RES, device=dev, cuda_graphs=False)
g = torch.Generator(device=dev).manual_seed(0)
gc, gp, gd = torch.randn(B, RES, RES, 3, device=dev, generator=g), torch.randn(B, RES, RES, device=dev, generator=g), torch.randn(B, RES, RES, device=dev, generator=g)
sess.step(student["Rs"], student["ts"], student["Ks"], cur, teacher["Rs"], teacher["ts"], grad_color=gc, grad_prob=gp, grad_depth=gd)
sess.synchronize()
lib = _lib.load()
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timed(label):
    ts = []
    for it in range(5 if "short" in sys.argv else 23):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        assert lib.dibr_backward_faces(ctypes.byref(sess.st.student), stream) == 0
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b) * 1e3)
    print("backward_faces call (%s) %.1f us (median of %d)" % (label, statistics.median(ts), len(ts)), flush=True)
al = lambda x: (x + 255) // 256 * 256
_p = sess.student.p
_nt = B * ((RES + 15) // 16) ** 2
_F = _p.total_faces
_off = al(64 * _F) + al(16 * _F) + al(4 * 32 * _nt) + al(16 * 32 * _nt) + al(4 * 4 * 32) + al(4 * _nt) + al(4 * B) + al(4 * (_nt // B) * ((_F + 31) // 32 + B + 1))
print("work lists (colour, soft):", sess.student.ws[_off:_off + 8].view(torch.int32).cpu().tolist(), "of", int(sess._last_total), "faces in use")
timed("colour + soft")
if "split" in sys.argv:
    sess._set_grads(gc, None, gd)
    timed("colour part only")
    sess._set_grads(None, gp, None)
    timed("soft part only")
    sess._set_grads(gc, gp, gd)
