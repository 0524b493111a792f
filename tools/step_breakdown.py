#!/usr/bin/env python
"""Per-entry-point timing of one cfg2 step under the step's real cache conditions: L2 flushed once at the start of the
step, then every C-ABI call of dibr_render_step issued separately with CUDA events in between (median of N steps)."""
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
sess = RenderSession(models, B, RES, RES, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
g_color = torch.randn(B, RES, RES, 3, device=dev, generator=g)
g_prob = torch.randn(B, RES, RES, device=dev, generator=g)
g_depth = torch.randn(B, RES, RES, device=dev, generator=g)
sess.step(student["Rs"], student["ts"], student["Ks"], cur, teacher["Rs"], teacher["ts"], grad_color=g_color, grad_prob=g_prob, grad_depth=g_depth)
sess.synchronize()
lib = _lib.load()
st = sess.st
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
npix = B * RES * RES
sp, tp = st.student, st.teacher
calls = [
    ("student setup_meshes", lambda: lib.dibr_setup_meshes(ctypes.byref(sp), stream)),
    ("student forward", lambda: lib.dibr_forward(ctypes.byref(sp), stream)),
    ("student normal_map", lambda: lib.dibr_normal_map_pass(ctypes.byref(sp), ctypes.c_void_p(st.student_normal_in), ctypes.c_void_p(st.student_mask_in), ctypes.c_void_p(st.student_normal_out), stream)),
    ("teacher setup_meshes", lambda: lib.dibr_setup_meshes(ctypes.byref(tp), stream)),
    ("teacher forward", lambda: lib.dibr_forward(ctypes.byref(tp), stream)),
    ("teacher normal_map", lambda: lib.dibr_normal_map_pass(ctypes.byref(tp), ctypes.c_void_p(st.teacher_normal_in), ctypes.c_void_p(st.teacher_mask_in), ctypes.c_void_p(st.teacher_normal_out), stream)),
    ("backward_faces", lambda: lib.dibr_backward_faces(ctypes.byref(sp), stream)),
    ("backward_meshes", lambda: lib.dibr_backward_meshes(ctypes.byref(sp), stream)),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
N = 20
acc = {n: [] for n, _ in calls}
tot = []
for it in range(N + 3):
    flush.zero_()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
    evs[0].record()
    for k, (n, f) in enumerate(calls):
        rc = f()
        assert rc == 0, (n, lib.dibr_last_error())
        evs[k + 1].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k, (n, _) in enumerate(calls):
            acc[n].append(evs[k].elapsed_time(evs[k + 1]) * 1e3)
        tot.append(evs[0].elapsed_time(evs[-1]) * 1e3)
for n, _ in calls:
    print("%-24s %7.1f us" % (n, statistics.median(acc[n])))
print("%-24s %7.1f us  (events add ~1-2 us per call)" % ("sum", statistics.median(tot)))
