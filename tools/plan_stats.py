#!/usr/bin/env python
"""Debug: tile statistics of the cfg2 student pass -- how many 16x16 tiles some face reaches, the histogram of their
list lengths (the plan's 32-face buckets), the work-list lengths the forward leaves for the backward.  Reads the plan the
set-up call left in the pass workspace (layout of carve() in csrc/dibr_abi.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from self6dpp_b200.session import RenderSession
dev = torch.device("cuda:0")
meshes, student, teacher = bench.workload(0)
models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
           "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)} for m in meshes]
cur = [models[int(i)] for i in student["ids"]]
B, RES = bench.BATCH, bench.RES
sess = RenderSession(models, B, RES, RES, device=dev, cuda_graphs=False)
sess.forward(student["Rs"], student["ts"], student["Ks"], cur, teacher["Rs"], teacher["ts"])
sess.synchronize()
al = lambda x: (x + 255) // 256 * 256
for name, pb in (("student", sess.student), ("teacher", sess.teacher)):
    F = pb.p.total_faces
    ntiles = B * ((RES + 15) // 16) ** 2
    off = al(64 * F) + al(16 * F) + al(4 * 32 * ntiles) + al(16 * 32 * ntiles)
    base = 0                                  # carve() starts at p.workspace (torch allocations are 512 B aligned)
    cnt = pb.ws[base + off: base + off + 512].view(torch.int32).cpu().tolist()
    hist = cnt[:32]
    touched = cnt[64]
    print(name, "tiles", ntiles, "touched", touched, "untouched", hist[0], "work CTAs", cnt[65])
    print("  faces-per-tile buckets (k: lists of 32(k-1)+1 .. 32k faces):", {k: v for k, v in enumerate(hist) if v and k})
    tc_off = off + al(4 * 4 * 32)
    tc = pb.ws[base + tc_off: base + tc_off + 4 * ntiles].view(torch.int32).cpu().float()
    t = tc[tc > 0]
    print("  listed faces per touched tile: mean %.1f median %.0f p90 %.0f max %.0f; sum %.0f (faces in use %d -> %.2f tiles per face)"
          % (t.mean(), t.median(), t.quantile(0.9), t.max(), t.sum(), int(sess.student.p.total_faces), t.sum() / 156682))
    idx = pb.imidx
    cov = (idx > 0).sum().item()
    print("  covered pixels per sample %.0f; pixels closed by the K-th face %d" % (cov / B, (idx < 0).sum().item()))
    # silhouette tiles: touched tiles with at least one uncovered pixel
    unc = (idx <= 0).view(B, RES // 16, 16, RES // 16, 16).any(4).any(2)
    tcv = tc.view(B, RES // 16, RES // 16) > 0
    tcv = tcv.to(unc.device)
    print("  touched tiles with an uncovered pixel: %d; fully covered: %d" % ((unc & tcv).sum().item(), (tcv & ~unc).sum().item()))
