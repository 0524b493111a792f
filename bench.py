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
2 fused rasterisations + 1 backward for the whole batch.  Upstream gradients dL/dcolor, dL/dprob, dL/ddepth are
fixed N(0,1) tensors (loss evaluation itself is not part of the metric).

Timing: CUDA events on the launching stream around every step, an L2 flush (256 MiB memset) between steps,
barrier + synchronize on both sides of the timed region, max over ranks.  ``value`` has poses/intrinsics
resident on the GPU; ``e2e`` goes through the same public API from pinned HOST buffers (H2D of R, t, K for
student and teacher every step) and reads the pose gradients back to the host (D2H) inside the timed region.
"""
import argparse
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
def cpu_step(meshes, student, teacher, sample_ids, H=RES, W=RES):
    """Reference pass structure per sample (renderer_dibr.py:273-301): colour fwd+bwd, normals fwd, depth(xyz)
    fwd+bwd for the student, normals fwd for the teacher -- 4 forward + 2 backward rasterisations, fp32 oracle."""
    import torch
    from oracle import dibr_oracle as O
    g = torch.Generator().manual_seed(0)
    for i in sample_ids:
        m = meshes[int(student["ids"][i])]
        v, f = torch.tensor(m["vertices"]), torch.tensor(m["faces"])
        fl = f.long()
        one = torch.ones(f.shape[0], 1)

        def face_attr(a):
            return torch.cat([a[fl[:, 0]], one, a[fl[:, 1]], one, a[fl[:, 2]], one], 1)[None]
        for which, pose in (("student", student), ("teacher", teacher)):
            cams = O.camera_params_from_RT_K(torch.tensor(pose["Rs"][i:i + 1]), torch.tensor(pose["ts"][i:i + 1]),
                                             torch.tensor(pose["Ks"][i]), H, W, near=0.01, far=100.0)
            p3, p2, nz, _ = O.project(v, f, cams[0][0], cams[1][0], cams[2])
            fw_n = O.rasterize(W, H, p3, p2, nz, face_attr(torch.tensor(m["normals"])))          # normals, fwd only
            if which == "teacher":
                continue
            fw_c = O.rasterize(W, H, p3, p2, nz, face_attr(torch.tensor(m["colors"])))
            O.rasterize_backward(fw_c, torch.randn(fw_c["im"].shape, generator=g), torch.randn(fw_c["improb"].shape, generator=g))
            xyz = (torch.tensor(pose["Rs"][i]) @ v.t()).t() + torch.tensor(pose["ts"][i])
            fw_d = O.rasterize(W, H, p3, p2, nz, face_attr(xyz))
            O.rasterize_backward(fw_d, torch.randn(fw_d["im"].shape, generator=g), torch.zeros(fw_d["improb"].shape))
            del fw_n


def run_cpu(meshes, student, teacher, n_samples):
    import torch
    from oracle import dibr_oracle as O
    O.lib()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    cpu_step(meshes, student, teacher, [0])               # warm-up (builds/loads the oracle, pages memory)
    t0 = time.perf_counter()
    cpu_step(meshes, student, teacher, list(range(n_samples)))
    dt = time.perf_counter() - t0
    return n_samples / dt, cores, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    meshes, student, teacher = workload(0)
    per_step = 2
    times = []
    from oracle import dibr_oracle as O
    O.lib()
    cores = os.cpu_count() or 1
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cpu_step(meshes, student, teacher, [(2 * s) % BATCH, (2 * s + 1) % BATCH])
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = per_step * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 256x256 ROI crops, 13 LINEMOD-shaped meshes (4.1k-5.9k faces); 4 fwd + 2 bwd rasterisations per sample",
                       "samples_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d samples per step of the cfg2 batch through oracle/dibr_oracle.c (OpenMP over pixels); "
                                       "the reference's own kernels are in the un-vendored kaolin v0.1 wheel and cannot be built" % per_step},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
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
            vec = gR.sum(0) if gt is None else du.pose_grad_checksum(gR, gt)
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

    # ---- the C-ABI step (dibr_render_step through RenderSession): one call per step, HOST buffers in ------------
    from self6dpp_b200.session import RenderSession
    sess = RenderSession(models, BATCH, RES, RES, student_mode=tuple(stu_mode), teacher_mode=("norm",), device=dev)
    g_prob3 = g_prob.contiguous()

    def sess_resident():      # poses / intrinsics already on the device, no host round trip
        sess.step(student["Rs"], student["ts"], student["Ks"], cur_models, teacher["Rs"], teacher["ts"],
                  grad_color=g_color, grad_prob=g_prob3, grad_depth=g_depth, upload=False, download=False)
        return sess.g_pose_dev, None

    def sess_e2e():           # pinned host inputs -> H2D, kernels, D2H of the pose gradients, host waits for them
        sess.step(student["Rs"], student["ts"], student["Ks"], cur_models, teacher["Rs"], teacher["ts"],
                  grad_color=g_color, grad_prob=g_prob3, grad_depth=g_depth, upload=True, download=True)
        sess.synchronize()
        return sess.g_pose_dev, None

    sess.step(student["Rs"], student["ts"], student["Ks"], cur_models, teacher["Rs"], teacher["ts"],
              grad_color=g_color, grad_prob=g_prob3, grad_depth=g_depth)      # first upload
    sess.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches, per_step_ms = timed(sess_resident, args.steps, args.warmup, True)
    # extra (not the headline): the same step without the teacher's soft-silhouette phase, which the reference computes
    # and throws away when mode has no "color" -- an API option (RenderSession(teacher_soft_mask=False)), off by default
    sess_lean = RenderSession(models, BATCH, RES, RES, student_mode=tuple(stu_mode), teacher_mode=("norm",), device=dev,
                              teacher_soft_mask=False)

    def sess_lean_resident():
        sess_lean.step(student["Rs"], student["ts"], student["Ks"], cur_models, teacher["Rs"], teacher["ts"],
                       grad_color=g_color, grad_prob=g_prob3, grad_depth=g_depth, upload=False, download=False)
        return sess_lean.g_pose_dev, None
    sess_lean.step(student["Rs"], student["ts"], student["Ks"], cur_models, teacher["Rs"], teacher["ts"],
                   grad_color=g_color, grad_prob=g_prob3, grad_depth=g_depth)
    sess_lean.synchronize()
    lean_ms, _, _ = timed(sess_lean_resident, args.steps, args.warmup, False)
    del sess_lean
    e2e_ms, _, _ = timed(sess_e2e, args.steps, max(3, args.warmup // 2), False)
    py_ms, _, _ = timed(step_resident, args.steps, args.warmup, False)
    py_e2e_ms, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2), False)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

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
            ncu = json.load(open(tpath))
            traffic = ncu.pop("dibr_forward_kernel_bytes_per_launch", None)
        roof = {"bound": "hbm", "kernel": "dibr_forward_kernel (student pass, D=8)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "kernel_ms": k_ms,
                "algorithmic_bytes": alg, "peak_source": peak_src,
                # what ncu says actually limits the kernel (committed capture, profiles/): instruction issue and
                # barrier latency, not HBM -- the DRAM traffic is half the algorithmic bytes
                "ncu": ncu}

    cpu = None
    if rank == 0 and not args.no_cpu:
        v, cores, dt = run_cpu(meshes, student, teacher, args.cpu_samples)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the %d cfg2 samples (4 fwd + 2 bwd rasterisations each) through oracle/dibr_oracle.c, "
                         "OpenMP over pixels, %.1f s" % (args.cpu_samples, BATCH, dt)}

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
                           "passes": "2 fused rasterisations + 1 backward per step (reference: 4 fwd + 2 bwd per sample); inside the call the teacher rasterisation runs on a side stream next to the student chain",
                           "api": "value/e2e: RenderSession.step -> dibr_render_step (one C-ABI call per step); "
                                  "python_api: Renderer_dibr.render_batch x2 + torch.autograd.backward (drop-in reference API)"},
                "option_teacher_without_soft_mask": {"value": BATCH * n * args.steps / (lean_ms * 1e-3), "unit": UNIT,
                                                     "ms_per_step": lean_ms / args.steps,
                                                     "note": "RenderSession(teacher_soft_mask=False): skips work whose result the reference discards; NOT the headline"},
                "python_api": {"value": BATCH * n * args.steps / (py_ms * 1e-3), "e2e": BATCH * n * args.steps / (py_e2e_ms * 1e-3),
                               "unit": UNIT, "ms_per_step": py_ms / args.steps},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": BATCH * 12 * 4,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
                "step_ms_min_med_max": [min(per_step_ms), statistics.median(per_step_ms), max(per_step_ms)]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
