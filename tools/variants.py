#!/usr/bin/env python
"""Debug: build variant libraries with -D switches (on the GPU box) and time dibr_forward_kernel with each.
usage: variants.py [--run TOOL.py] NAME=-DFLAG1,-DFLAG2 ...   (NAME 'base' = no flags)"""
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
    reps = 20
    ms = time_forward_kernel(ren, dev_in, cur, ["color", "depth", "mask", "norm", "prob"], 256, flush, reps=reps)
    print("%-28s forward kernel %.1f us" % (os.environ["DIBR_VARIANT_NAME"], ms * 1e3), flush=True)
    sys.exit(0)
for name, flags in specs:
    fl = [f for f in flags.split(",") if f]
    out = os.path.join(ROOT, "self6dpp_b200", "lib", f"libdibr_b200_{name}.so")
    cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-shared", "-cudart", "static", "-o", out] + fl + \
          [os.path.join(csrc, f) for f in ("dibr_abi.cu", "dibr_setup.cu", "dibr_forward.cu", "dibr_normalmap.cu", "dibr_backward.cu", "dibr_nnd.cu", "dibr_nnd_grid.cu", "dibr_backproject.cu", "dibr_maskloss.cu", "dibr_photometric.cu", "dibr_roialign.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(name, "BUILD FAILED", r.stderr[-2000:]); continue
    env = dict(os.environ, DIBR_VARIANT_LIB=out, DIBR_VARIANT_NAME=name + " " + " ".join(fl))
    if script:
        print("==== variant", name, " ".join(fl), flush=True)
        env.pop("DIBR_VARIANT_LIB")
        env["DIBR_B200_LIB"] = out
        subprocess.run([sys.executable, os.path.join(ROOT, script)], env=env)
    else:
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env)
    os.remove(out)
