#!/usr/bin/env python
"""bench.py -- DIB-R forward+backward throughput on B200 (BASELINE.json metric) and its CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload = "cfg2"): per GPU a batch of 32 (image, object) samples,
256x256 ROI crops, 13 LINEMOD-shaped synthetic meshes (4.1k-5.9k faces), per-sample crop intrinsics.
One STEP = what Self6D++'s compute_self_loss_pose asks of the renderer for that batch
(/root/reference/core/self6dpp/engine/self_engine_utils.py:426-447):
    student  render_batch(mode=[color, depth, mask, norm, prob])   (grads flow to pred_rot / pred_trans)
    teacher  render_batch(mode=[norm])                             (no grad)
    backward of the mask (prob), depth and RGB terms -> dL/dR, dL/dt
The reference does this as 4 forward + 2 backward rasterisations per sample, one sample at a time; here it is
2 fused rasterisations + 1 backward for the whole batch.  The step is issued the way a training loop has to issue it:
``RenderSession.forward`` (dibr_render_forward), then -- with upstream gradients that exist only after the forward --
``RenderSession.backward`` (dibr_render_backward).  The upstream gradients dL/dcolor, dL/dprob, dL/ddepth are fixed
N(0,1) tensors (evaluating the losses is not part of the metric).

Timing: CUDA events on the launching stream around every step, an L2 flush (256 MiB memset) between steps,
barrier + synchronize on both sides of the timed region, max over ranks.  ``value`` has poses/intrinsics
resident on the GPU; ``e2e`` goes through the same public API from pinned HOST buffers (H2D of R, t, K for
student and teacher every step) and reads the pose gradients back to the host (D2H) inside the timed region.
After the timed runs the rendered batch is compared with the CPU oracle on a few samples (``parity_check``); a
mismatch makes the run fail.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
RES = 256
STRONG_BATCH = 512          # cfg5: the global batch of the strong-scaling sweep (BASELINE.json configs[4])
METRIC = "dibr_fwd_bwd_samples_per_sec"
UNIT = "samples/s"


def workload(rank, batch=BATCH):
    from self6dpp_b200 import synth
    meshes = synth.lm13_meshes()
    student = synth.roi_batch(meshes, batch, res=RES, seed=2000 + rank)
    # teacher pose = student pose perturbed a little (pseudo label vs prediction)
    rng = np.random.default_rng(99 + rank)
    dR, _ = synth.random_rotations(batch, 500 + rank)
    eye = np.eye(3, dtype=np.float32)[None]
    small = eye + 0.03 * (dR - eye)
    u, _, vt = np.linalg.svd(small)
    small = (u @ vt).astype(np.float32)
    teacher = {"Rs": small @ student["Rs"], "ts": student["ts"] + rng.normal(0, 0.003, student["ts"].shape).astype(np.float32),
               "Ks": student["Ks"], "ids": student["ids"]}
    return meshes, student, teacher


def algorithmic_bytes_forward(D, H, W, faces_total, batch):
    """bytes_fwd = 4*[(D+5)*H*W + (16+3D)*F] per image."""
    return 4.0 * ((D + 5) * H * W * batch + (16 + 3 * D) * faces_total)


def csrc_hash():
    """hash of the kernel sources: profiles/traffic.json records it at capture time, so a stale capture is recognised"""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "self6dpp_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's kernels (the reference's own CUDA extension lives in the
# un-vendored kaolin v0.1 wheel and cannot be built; DESIGN.md "Oracle")
# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_with_threads():
    """Load the oracle with an EXPLICIT OpenMP thread count: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would silently make the CPU arm single-threaded.  Returns (module, threads actually in use)."""
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import torch
    torch.set_num_threads(cores)
    from oracle import dibr_oracle as O
    L = O.lib()
    L.dibr_oracle_set_threads(int(cores))
    return O, int(L.dibr_oracle_get_threads())


def sample_seam(O, mesh, pose, i, H=RES, W=RES):
    """fp32 operator-seam inputs of sample i (camera + vertex shader in the oracle's fixed order)"""
    import torch
    v, f = torch.tensor(mesh["vertices"]), torch.tensor(mesh["faces"])
    cams = O.camera_params_from_RT_K(torch.tensor(pose["Rs"][i:i + 1]), torch.tensor(pose["ts"][i:i + 1]),
                                     torch.tensor(pose["Ks"][i]), H, W, near=0.01, far=100.0)
    p3, p2, nz, _ = O.project(v, f, cams[0][0], cams[1][0], cams[2])
    return v, f, p3, p2, nz


def face_attr(a, fl):
    import torch
    one = torch.ones(fl.shape[0], 1)
    return torch.cat([a[fl[:, 0]], one, a[fl[:, 1]], one, a[fl[:, 2]], one], 1)[None]


def cpu_step(O, meshes, student, teacher, sample_ids, H=RES, W=RES, counts=None):
    """Reference pass structure per sample (renderer_dibr.py:273-301): colour fwd+bwd, normals fwd, depth(xyz)
    fwd+bwd for the student, normals fwd for the teacher -- 4 forward + 2 backward rasterisations, fp32 oracle."""
    import torch
    g = torch.Generator().manual_seed(0)
    for i in sample_ids:
        m = meshes[int(student["ids"][i])]
        for which, pose in (("student", student), ("teacher", teacher)):
            v, f, p3, p2, nz = sample_seam(O, m, pose, i, H, W)
            fl = f.long()
            fw_n = O.rasterize(W, H, p3, p2, nz, face_attr(torch.tensor(m["normals"]), fl))          # normals, fwd only
            if which == "teacher":
                continue
            fw_c = O.rasterize(W, H, p3, p2, nz, face_attr(torch.tensor(m["colors"]), fl))
            if counts is not None:
                a, c = O.work_counts(fw_c, nz)
                counts["n_cov"] += a
                counts["n_soft"] += c
                counts["covered"] += int((fw_c["imidx"] > 0).sum())
                counts["samples"] += 1
            O.rasterize_backward(fw_c, torch.randn(fw_c["im"].shape, generator=g), torch.randn(fw_c["improb"].shape, generator=g))
            xyz = (torch.tensor(pose["Rs"][i]) @ v.t()).t() + torch.tensor(pose["ts"][i])
            fw_d = O.rasterize(W, H, p3, p2, nz, face_attr(xyz, fl))
            O.rasterize_backward(fw_d, torch.randn(fw_d["im"].shape, generator=g), torch.zeros(fw_d["improb"].shape))
            del fw_n


def run_cpu(meshes, student, teacher, n_samples):
    O, threads = oracle_with_threads()
    cpu_step(O, meshes, student, teacher, [0])               # warm-up (builds/loads the oracle, pages memory)
    counts = {"n_cov": 0, "n_soft": 0, "covered": 0, "samples": 0}
    t0 = time.perf_counter()
    cpu_step(O, meshes, student, teacher, list(range(n_samples)), counts=counts)
    dt = time.perf_counter() - t0
    return n_samples / dt, threads, dt, counts


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, threads = oracle_with_threads()
    meshes, student, teacher = workload(0)
    per_step = 2
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cpu_step(O, meshes, student, teacher, [(2 * s) % BATCH, (2 * s + 1) % BATCH])
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = per_step * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 256x256 ROI crops, 13 LINEMOD-shaped meshes (4.1k-5.9k faces); 4 fwd + 2 bwd rasterisations per sample",
                       "samples_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d samples per step of the cfg2 batch through oracle/dibr_oracle.c (OpenMP over pixels, %d threads set "
                                       "explicitly; OMP_NUM_THREADS of the launcher is overridden); the reference's own kernels are in the "
                                       "un-vendored kaolin v0.1 wheel and cannot be built" % (per_step, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def parity_check(sess, chk, meshes, student, n_samples):
    """The rendered cfg2 batch against the CPU oracle on the first ``n_samples`` samples: face ids and interpolated attributes
    bit for bit (fp32 operation-order oracle), soft mask against float64.  ``sess`` is the timed session, which turns the
    interpolated normals into the normal map IN PLACE; ``chk`` rendered the same inputs with raw_normals=True (same kernels,
    the map goes to its own tensor), so the raw normals can be compared with the oracle -- and every output of the timed
    session, its normal map included, must equal the checked session's bit for bit."""
    import torch
    from oracle import dibr_oracle as O
    s = chk.student
    im_all = torch.cat([s.out[k] for k in s.keys], -1).cpu()
    idx_all, prob_all = s.imidx.cpu(), s.improb.cpu()
    same_out = bool(torch.equal(sess.student.imidx, s.imidx) and torch.equal(sess.student.improb, s.improb)
                    and torch.equal(sess.student.normal_map, s.normal_map) and torch.equal(sess.teacher.normal_map, chk.teacher.normal_map)
                    and all(torch.equal(sess.student.out[k], s.out[k]) for k in s.keys if k != "norm"))
    res = {"samples": n_samples, "timed_session_equals_checked_session": same_out, "imidx_mismatch_pixels": 0, "im_mismatch_values": 0, "prob_err_over_tol": 0.0,
           "tolerance_prob": "1e-5 * |ref| + 1e-5 vs the float64 oracle on identical fp32 corners"}
    for i in range(n_samples):
        m = meshes[int(student["ids"][i])]
        v, f, p3, p2, nz = sample_seam(O, m, student, i)
        fl = f.long()
        cols = torch.cat([torch.tensor(m["colors"]), torch.tensor(m["normals"]), torch.ones(len(v), 1)], 1)
        at = torch.cat([torch.cat([cols[fl[:, c]], -p3[0, :, 3 * c + 2:3 * c + 3]], 1) for c in range(3)], 1)[None].contiguous()
        fw32 = O.rasterize(RES, RES, p3, p2, nz, at)
        res["imidx_mismatch_pixels"] += int((idx_all[i].clamp(min=0).float() != fw32["imidx"][0, ..., 0]).sum())
        res["im_mismatch_values"] += int((im_all[i:i + 1] != fw32["im"]).sum())
        fw64 = O.rasterize(RES, RES, p3.double(), p2.double(), nz.double(), at.double())
        same = fw32["imidx"].double() == fw64["imidx"]
        err = ((prob_all[i:i + 1].double() - fw64["improb"]).abs() / (1e-5 * fw64["improb"].abs() + 1e-5))[same]
        res["prob_err_over_tol"] = max(res["prob_err_over_tol"], float(err.max()) if err.numel() else 0.0)
    res["ok"] = same_out and res["imidx_mismatch_pixels"] == 0 and res["im_mismatch_values"] == 0 and res["prob_err_over_tol"] <= 1.0
    return res


def kaolin_structure_gpu(meshes, student, teacher, dev, n_samples, flush):
    """The reference's pass structure (4 forward + 2 backward LinearRasterizer calls per sample, one sample at a time,
    renderer_dibr.py:273-301) on the stand-in kernels of oracle/kaolin_structure.cu: per-pixel loops over all faces, fp32
    atomics.  Projection happens before the timed region (the reference's torch vertex shader is not the subject)."""
    import torch
    from oracle import dibr_oracle as O
    from oracle import kaolin_structure as KS
    passes = []
    g = torch.Generator().manual_seed(0)
    for i in range(n_samples):
        m = meshes[int(student["ids"][i])]
        per = {}
        for which, pose in (("student", student), ("teacher", teacher)):
            v, f, p3, p2, nz = sample_seam(O, m, pose, i)
            fl = f.long()
            dv = lambda t: t.to(dev)
            per[which + "_norm"] = KS.Pass(dv(p3), dv(p2), dv(nz), dv(face_attr(torch.tensor(m["normals"]), fl)), RES, RES)
            if which == "student":
                per["color"] = KS.Pass(dv(p3), dv(p2), dv(nz), dv(face_attr(torch.tensor(m["colors"]), fl)), RES, RES)
                xyz = (torch.tensor(pose["Rs"][i]) @ v.t()).t() + torch.tensor(pose["ts"][i])
                per["xyz"] = KS.Pass(dv(p3), dv(p2), dv(nz), dv(face_attr(xyz, fl)), RES, RES)
        per["g_im"] = torch.randn(1, RES, RES, 4, generator=g).to(dev)
        per["g_prob"] = torch.randn(1, RES, RES, 1, generator=g).to(dev)
        per["g_zero"] = torch.zeros(1, RES, RES, 1, device=dev)
        passes.append(per)

    def run():
        for per in passes:
            per["student_norm"].forward()
            per["teacher_norm"].forward()
            per["color"].forward()
            per["color"].backward(per["g_im"], per["g_prob"])
            per["xyz"].forward()
            per["xyz"].backward(per["g_im"], per["g_zero"])
    run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = min(ms)
    return {"value": n_samples / (best * 1e-3), "unit": UNIT, "ms_per_sample": best / n_samples, "samples": n_samples,
            "note": "oracle/kaolin_structure.cu: a transcription of the reference kernels' STRUCTURE (one thread per pixel over all faces, "
                    "fp32 atomics) compiled for sm_100a, 4 fwd + 2 bwd rasterisations per sample; kaolin's own sources are not available"}


def main_b200(args):
    import torch
    import torch.distributed as dist
    from self6dpp_b200 import Renderer_dibr, _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- self6dpp_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    meshes, student, teacher = workload(rank)
    models = [{"vertices": torch.tensor(m["vertices"], device=dev), "colors": torch.tensor(m["colors"], device=dev),
               "normals": torch.tensor(m["normals"], device=dev), "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)}
              for m in meshes]
    cur_models = [models[int(i)] for i in student["ids"]]
    faces_total = int(sum(m["faces"].shape[0] for m in cur_models))
    ren = Renderer_dibr(RES, RES, "VertexColorBatch")
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    g_color = torch.randn(BATCH, RES, RES, 3, generator=g).to(dev)
    g_prob = torch.randn(BATCH, RES, RES, generator=g).to(dev)
    g_depth = torch.randn(BATCH, RES, RES, generator=g).to(dev)
    dev_in = {k: torch.tensor(student[k], device=dev) for k in ("Rs", "ts", "Ks")}
    dev_te = {k: torch.tensor(teacher[k], device=dev) for k in ("Rs", "ts")}
    host_in = {k: torch.tensor(student[k]).pin_memory() for k in ("Rs", "ts", "Ks")}
    host_te = {k: torch.tensor(teacher[k]).pin_memory() for k in ("Rs", "ts")}
    host_out = torch.empty(BATCH, 12).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stu_mode = ["color", "depth", "mask", "norm", "prob"]

    def step(Rs, ts, Ks, tRs, tts):
        Rs.requires_grad_(True)
        ts.requires_grad_(True)
        ret = ren.render_batch(Rs, ts, cur_models, Ks=Ks, width=RES, height=RES, mode=stu_mode)
        with torch.no_grad():
            ret_t = ren.render_batch(tRs, tts, cur_models, Ks=Ks, width=RES, height=RES, mode=["norm"])
        torch.autograd.backward([ret["color"], ret["prob"], ret["depth"]], [g_color, g_prob, g_depth])
        return Rs.grad, ts.grad, ret, ret_t

    def step_resident():
        Rs = dev_in["Rs"].detach().clone()
        ts = dev_in["ts"].detach().clone()
        return step(Rs, ts, dev_in["Ks"], dev_te["Rs"], dev_te["ts"])

    def step_e2e():
        Rs = host_in["Rs"].to(dev, non_blocking=True)
        ts = host_in["ts"].to(dev, non_blocking=True)
        Ks = host_in["Ks"].to(dev, non_blocking=True)
        tRs = host_te["Rs"].to(dev, non_blocking=True)
        tts = host_te["ts"].to(dev, non_blocking=True)
        gR, gt, _, _ = step(Rs, ts, Ks, tRs, tts)
        host_out.copy_(torch.cat([gR.reshape(BATCH, 9), gt], dim=1), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller consumes the pose gradients on the host
        return host_out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from self6dpp_b200 import dist_utils as du

    pending = []

    def allreduce_check(gR, gt):
        """cfg5: 12-float pose-gradient check-sum, one NCCL all-reduce over NVLink per step, issued asynchronously
        (NCCL's own stream) so it overlaps the following steps' kernels; up to four are in flight, the oldest is waited
        for first (stream-side wait), and all of them are drained inside the timed region."""
        if world > 1:
            if len(pending) >= 4:
                pending.pop(0).wait()
            # session paths hand over the check-sum row the backward kernel wrote (no extra launch); the Python API path adds it up
            vec = gR if gt is None else du.pose_grad_checksum(gR, gt)
            pending.append(dist.all_reduce(vec, async_op=True))
            return vec
        return None

    def timed(fn, steps, warmup, with_collective):
        for _ in range(warmup):
            out = fn()
            if with_collective:
                allreduce_check(out[0], out[1])
        sync_all()
        lib.dibr_launch_count(1)
        evs = []
        for _ in range(steps):
            flush.zero_()                                   # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            if with_collective:
                allreduce_check(out[0], out[1])
            e1.record()
            evs.append((e0, e1))
        while pending:
            pending.pop().wait()
        sync_all()
        launches = lib.dibr_launch_count(1)
        ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(ms)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, launches, ms

    # ---- the C-ABI step through RenderSession: forward call, then backward call, HOST buffers in ---------------------
    from self6dpp_b200.session import RenderSession
    sess = RenderSession(models, BATCH, RES, RES, student_mode=tuple(stu_mode), teacher_mode=("norm",), device=dev)
    g_prob3 = g_prob.contiguous()

    def make_steps(s, grads, st, te, cur):
        gc, gp, gd = grads

        def resident():     # poses / intrinsics already on the device, no host round trip
            s.forward(st["Rs"], st["ts"], st["Ks"], cur, te["Rs"], te["ts"], upload=False)
            s.backward(gc, gp, gd, download=False)
            return s.g_pose_sum, None

        def e2e():          # pinned host inputs -> H2D, kernels, D2H of the pose gradients, host waits for them
            s.forward(st["Rs"], st["ts"], st["Ks"], cur, te["Rs"], te["ts"], upload=True)
            s.backward(gc, gp, gd, download=True)
            s.synchronize()
            return s.g_pose_sum, None

        def one_call():     # dibr_render_step: the whole step in ONE C-ABI call (gradients known beforehand)
            s.step(st["Rs"], st["ts"], st["Ks"], cur, te["Rs"], te["ts"], grad_color=gc, grad_prob=gp, grad_depth=gd,
                   upload=False, download=False)
            return s.g_pose_sum, None
        return resident, e2e, one_call

    sess_resident, sess_e2e, sess_one_call = make_steps(sess, (g_color, g_prob3, g_depth), student, teacher, cur_models)
    sess_e2e()                                              # first upload
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches, per_step_ms = timed(sess_resident, args.steps, args.warmup, True)
    one_ms, _, _ = timed(sess_one_call, args.steps, args.warmup, False)
    # the same forward + backward replayed as ONE CUDA graph (launch overhead of the ~12 kernels, 4 memsets and the copies removed)
    graph_ms = None
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            sess_resident()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sess_resident()

        def sess_graph():
            graph.replay()
            return sess.g_pose_sum, None
        graph_ms, _, _ = timed(sess_graph, args.steps, args.warmup, False)
    except Exception as exc:                                # capture is an extra, never the headline
        graph_ms = None
        graph_err = str(exc)[:200]
        torch.cuda.synchronize()
    # extra (not the headline): the same step without the teacher's soft-silhouette phase, which the reference computes
    # and throws away when mode has no "color" -- an API option (RenderSession(teacher_soft_mask=False)), off by default
    sess_lean = RenderSession(models, BATCH, RES, RES, student_mode=tuple(stu_mode), teacher_mode=("norm",), device=dev,
                              teacher_soft_mask=False)
    lean_resident, lean_e2e, _ = make_steps(sess_lean, (g_color, g_prob3, g_depth), student, teacher, cur_models)
    lean_e2e()
    lean_ms, _, _ = timed(lean_resident, args.steps, args.warmup, False)
    del sess_lean
    e2e_ms, _, _ = timed(sess_e2e, args.steps, max(3, args.warmup // 2), False)
    py_ms, _, _ = timed(step_resident, args.steps, args.warmup, False)
    py_e2e_ms, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2), False)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # ---- cfg5 as BASELINE states it: a global batch of 512 crops sharded 512 / N per rank (strong scaling) ----------
    strong = None
    if STRONG_BATCH % world == 0:
        sb = STRONG_BATCH // world
        s_meshes, s_student, s_teacher = workload(rank, batch=sb)
        s_cur = [models[int(i)] for i in s_student["ids"]]
        gs = torch.Generator(device="cpu").manual_seed(99 + rank)
        s_grads = (torch.randn(sb, RES, RES, 3, generator=gs).to(dev), torch.randn(sb, RES, RES, generator=gs).to(dev),
                   torch.randn(sb, RES, RES, generator=gs).to(dev))
        s_sess = RenderSession(models, sb, RES, RES, student_mode=tuple(stu_mode), teacher_mode=("norm",), device=dev)
        s_resident, s_e2e, _ = make_steps(s_sess, s_grads, s_student, s_teacher, s_cur)
        s_e2e()
        s_steps = max(3, args.steps // 4)
        s_ms, _, _ = timed(s_resident, s_steps, 3, True)
        strong = {"global_batch": STRONG_BATCH, "batch_per_gpu": sb, "value": STRONG_BATCH * s_steps / (s_ms * 1e-3), "unit": UNIT,
                  "ms_per_step": s_ms / s_steps, "steps": s_steps, "scaling": "strong",
                  "note": "cfg5: 512 crops sharded 512 / N per rank, 12-float NCCL all-reduce of the pose-gradient check-sum per step"}
        del s_sess, s_grads

    # ---- dominant kernel: the student forward rasterisation, timed alone with CUDA events --------------------
    roof = None
    if rank == 0:
        from self6dpp_b200.bench_util import time_forward_kernel
        D = 8     # colour 3 + normal 3 + ones + depth
        k_ms = time_forward_kernel(ren, dev_in, cur_models, stu_mode, RES, flush, reps=20)
        alg = algorithmic_bytes_forward(D, RES, RES, faces_total, BATCH)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg / (k_ms * 1e-3) / 1e9
        traffic, ncu = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            cap = json.load(open(tpath))
            fresh = cap.get("csrc_hash") == csrc_hash()
            ncu = {"committed_capture": True, "file": "profiles/traffic.json", "captured_at_csrc_hash": cap.get("csrc_hash"),
                   "matches_current_sources": fresh}
            if fresh:                                        # a capture of other kernel sources says nothing about this build
                traffic = cap.get("dibr_forward_kernel_bytes_per_launch")
                ncu.update({k: v for k, v in cap.items() if k not in ("dibr_forward_kernel_bytes_per_launch", "csrc_hash")})
        # pure-write bandwidth of this GPU (a device fill): 109 of the kernel's 134 MB are writes, and a write stream tops out
        # well below the copy figure (read + write bytes) of MEASURED_PEAKS.json
        wbuf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
        for _ in range(2):
            wbuf.zero_()
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(5):
            wbuf.zero_()
        w1.record()
        torch.cuda.synchronize()
        write_peak = 5 * wbuf.numel() / (w0.elapsed_time(w1) * 1e-3) / 1e9
        del wbuf
        bytes_written = 4.0 * (D + 5) * RES * RES * BATCH
        roof = {"bound": "hbm", "kernel": "dibr_forward_kernel<FUSED> (student pass, D=8; with the list-reset memset of dibr_forward)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "kernel_ms": k_ms,
                "algorithmic_bytes": alg, "peak_source": peak_src,
                "limiter": "latency of short dependent phases between CTA barriers at half occupancy, not HBM (DESIGN.md 6): the HBM "
                           "roofline is the binding one only because the FP32 work is tiny -- see fp32",
                "write_floor": {"bytes_written": bytes_written, "fill_peak_gbs_measured": write_peak,
                                "floor_ms": bytes_written / (write_peak * 1e9) * 1e3,
                                "frac_of_floor": bytes_written / (write_peak * 1e9) * 1e3 / k_ms,
                                "note": "81 % of the algorithmic bytes are stores; a store-only stream (torch fill, 512 MiB x 5) reaches "
                                        "this rate on the same GPU, so the kernel cannot be faster than floor_ms whatever it computes"},
                "ncu": ncu}
        try:
            from oracle import kaolin_structure as KS
            roof["fp32"] = {"peak_tflops_measured": KS.fma_peak_tflops(5), "how": "oracle/kaolin_structure.cu fma_peak_kernel, 8 FMA chains per thread, best of 5"}
        except Exception as exc:
            roof["fp32"] = {"peak_tflops_measured": None, "error": str(exc)[:120]}

    cpu, par, ks = None, None, None
    if rank == 0 and not args.no_cpu:
        v, threads, dt, counts = run_cpu(meshes, student, teacher, args.cpu_samples)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of the %d cfg2 samples (4 fwd + 2 bwd rasterisations each) through oracle/dibr_oracle.c, "
                         "OpenMP over pixels with %d threads (set explicitly), %.1f s" % (args.cpu_samples, BATCH, threads, dt)}
        if roof is not None and counts["samples"]:
            n = counts["samples"]
            n_cov, n_soft, covered = counts["n_cov"] / n, counts["n_soft"] / n, counts["covered"] / n
            # per sample and forward rasterisation: ~40 flops per bbox test, ~60 per soft pair, ~6 D per covered pixel
            flops = BATCH * (40.0 * n_cov + 60.0 * n_soft + 6.0 * 8 * covered)
            roof["fp32"].update({"N_cov_per_sample": n_cov, "N_soft_per_sample": n_soft, "covered_pixels_per_sample": covered,
                                 "flops_fwd_per_launch": flops, "achieved_tflops": flops / (roof["kernel_ms"] * 1e-3) / 1e12,
                                 "counted_on": "%d cfg2 samples, oracle.work_counts" % n})
            if roof["fp32"].get("peak_tflops_measured"):
                roof["fp32"]["frac"] = roof["fp32"]["achieved_tflops"] / roof["fp32"]["peak_tflops_measured"]
        sess_e2e()
        sess.synchronize()
        chk = RenderSession(models, BATCH, RES, RES, student_mode=tuple(stu_mode), teacher_mode=("norm",), device=dev, raw_normals=True)
        chk.forward(student["Rs"], student["ts"], student["Ks"], cur_models, teacher["Rs"], teacher["ts"])
        chk.synchronize()
        par = parity_check(sess, chk, meshes, student, min(4, args.cpu_samples))
        del chk
        try:
            ks = kaolin_structure_gpu(meshes, student, teacher, dev, 4, flush)
        except Exception as exc:
            ks = {"value": None, "error": str(exc)[:200]}

    rc = 0
    if rank == 0:
        n = world
        value = BATCH * n * args.steps / (total_ms * 1e-3)
        e2e_value = BATCH * n * args.steps / (e2e_ms * 1e-3)
        h2d = 4 * sess.stage_words
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg2: batch 32 x 256x256 ROI crops per GPU, 13 LINEMOD-shaped meshes (4.1k-5.9k faces); "
                                       "student colour+depth+mask+norm+prob fwd+bwd, teacher norm fwd",
                           "batch_per_gpu": BATCH, "faces_in_batch": faces_total, "l2": "flushed between steps (256 MiB memset)",
                           "passes": "2 fused rasterisations + 1 backward per step (reference: 4 fwd + 2 bwd per sample); inside the forward call the teacher rasterisation runs on the session's side stream next to the student chain",
                           "api": "value/e2e: RenderSession.forward -> dibr_render_forward, then RenderSession.backward -> dibr_render_backward "
                                  "(two C-ABI calls per step, gradients handed over after the forward; the session captures the launches of "
                                  "each call into a CUDA graph the second time it is made and replays it afterwards); one_call_step: dibr_render_step; "
                                  "python_api: Renderer_dibr.render_batch x2 + torch.autograd.backward (drop-in reference API)"},
                "one_call_step": {"value": BATCH * n * args.steps / (one_ms * 1e-3), "unit": UNIT, "ms_per_step": one_ms / args.steps,
                                  "note": "dibr_render_step: forward + backward in one call, student chain on the side stream throughout"},
                "cuda_graph": ({"value": BATCH * n * args.steps / (graph_ms * 1e-3), "unit": UNIT, "ms_per_step": graph_ms / args.steps,
                                "note": "the forward AND backward calls captured together and replayed as ONE CUDA graph (`value` replays one graph per call: RenderSession captures each call the second time it is made)"}
                               if graph_ms else {"value": None, "error": locals().get("graph_err")}),
                "option_teacher_without_soft_mask": {"value": BATCH * n * args.steps / (lean_ms * 1e-3), "unit": UNIT,
                                                     "ms_per_step": lean_ms / args.steps,
                                                     "note": "RenderSession(teacher_soft_mask=False): skips work whose result the reference discards; NOT the headline"},
                "python_api": {"value": BATCH * n * args.steps / (py_ms * 1e-3), "e2e": BATCH * n * args.steps / (py_e2e_ms * 1e-3),
                               "unit": UNIT, "ms_per_step": py_ms / args.steps},
                "strong_scaling": strong,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": BATCH * 12 * 4,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
                "parity_check": par, "kaolin_structure_gpu": ks,
                "step_ms_min_med_max": [min(per_step_ms), statistics.median(per_step_ms), max(per_step_ms)]}
        print(json.dumps(line))
        if par is not None and not par["ok"]:
            print("bench.py: PARITY CHECK FAILED: %s" % json.dumps(par), file=sys.stderr)
            rc = 3
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-samples", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
