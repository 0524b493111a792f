#!/usr/bin/env python
"""Debug: build variant libraries with -D switches (on the GPU box) and time dibr_forward_kernel with each.
usage: variants.py NAME=-DFLAG1,-DFLAG2 ...   (NAME 'base' = no flags; a NAME starting with v4 runs the barrier-free forward design,
DIBR_FWD_IMPL=4); add :phase to a name to print phase cycles"""
import sys, os, subprocess, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "self6dpp_b200", "csrc")
script = None
argv = sys.argv[1:]
if argv and argv[0] == "--run":          # run another tool against every variant instead of the built-in forward timing
    script, argv = argv[1], argv[2:]
specs = [a.split("=", 1) if "=" in a else [a, ""] for a in argv] or [["base", ""]]
child = os.environ.get("DIBR_VARIANT_LIB")
if child:
    from self6dpp_b200 import _lib
    _lib.LIB_PATH = child
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
    phase = os.environ.get("DIBR_VARIANT_PHASE") == "1"
    buf = (ctypes.c_ulonglong * 24)()
    if phase:
        lib.dibr_debug_item_cycles(buf, 1)
    reps = 20
    ms = time_forward_kernel(ren, dev_in, cur, ["color", "depth", "mask", "norm", "prob"], 256, flush, reps=reps)
    print("%-28s forward kernel %.1f us" % (os.environ["DIBR_VARIANT_NAME"], ms * 1e3), flush=True)
    if phase:
        lib.dibr_debug_item_cycles(buf, 1)
        print("   longest item: %d cycles, flushes %d, last flush ids %d, ring %d; items with 2+ flushes per launch %.0f, mean %.0f cycles" % (buf[0] >> 32, (buf[0] >> 24) & 255, (buf[0] >> 12) & 0xfff, buf[0] & 0xfff, buf[14] / (reps + 3), buf[13] / max(buf[14], 1)))
        buf[0] >>= 32
        print("   block items: longest %d cycles, mean %.0f, count/launch %.0f, >20k %.0f, >50k %.0f, >100k %.0f per launch; items > 6k cycles: %.0f per launch, mean %.0f cycles" % (
            buf[0], buf[1] / max(buf[2], 1), buf[2] / (reps + 3), buf[3] / (reps + 3), buf[4] / (reps + 3), buf[5] / (reps + 3), buf[7] / (reps + 3), buf[6] / max(buf[7], 1)))
        print("   mean cycles per item by segment: zbuf %.0f, resolve %.0f, write %.0f, flags+soft %.0f, final stores %.0f" % tuple(buf[8 + k] / max(buf[2], 1) for k in range(5)))
        print("   soft, cycles per launch / 1000: filter %.0f, corner wait %.0f, rounds %.0f, words+expansion %.0f; all items %.0f" % tuple([buf[8 + k] / (reps + 3) / 1e3 for k in (8, 9, 10, 11)] + [buf[1] / (reps + 3) / 1e3]))
        sys.exit(0)
        n = buf[7]
        names = ["list+gather", "prep", "coverage", "resolve", "soft", "write", "fill"]
        print("   touched tiles per launch %.0f" % (n / (reps + 3)))
        tot = sum(buf[:7])
        for k in range(7):
            print("   %-12s %8.0f cycles/tile  %5.1f%%" % (names[k], buf[k] / max(n, 1), 100.0 * buf[k] / max(tot, 1)))
    sys.exit(0)
for name, flags in specs:
    phase = name.endswith(":phase")
    name = name.replace(":phase", "")
    fl = [f for f in flags.split(",") if f] + (["-DDIBR_ITEM_TIMING"] if phase else [])
    out = os.path.join(ROOT, "self6dpp_b200", "lib", f"libdibr_b200_{name}.so")
    cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-shared", "-cudart", "static", "-o", out] + fl + \
          [os.path.join(csrc, f) for f in ("dibr_abi.cu", "dibr_setup.cu", "dibr_forward.cu", "dibr_forward_v2.cu", "dibr_backward.cu", "dibr_nnd.cu", "dibr_nnd_grid.cu", "dibr_backproject.cu", "dibr_maskloss.cu", "dibr_photometric.cu", "dibr_roialign.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(name, "BUILD FAILED", r.stderr[-2000:]); continue
    if name.startswith("v4"):
        os.environ["DIBR_FWD_IMPL"] = "4"
    else:
        os.environ.pop("DIBR_FWD_IMPL", None)
    env = dict(os.environ, DIBR_VARIANT_LIB=out, DIBR_VARIANT_NAME=name + " " + " ".join(fl), DIBR_VARIANT_PHASE="1" if phase else "0")
    if script:
        print("==== variant", name, " ".join(fl), flush=True)
        env.pop("DIBR_VARIANT_LIB")
        env["DIBR_B200_LIB"] = out
        subprocess.run([sys.executable, os.path.join(ROOT, script)], env=env)
    else:
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env)
    os.remove(out)
