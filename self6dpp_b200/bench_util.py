"""Measurement helpers used by bench.py (not part of the reference-facing API)."""
import ctypes

import torch

from . import _lib, fused
from .renderer import vc
from .renderer_dibr import _faces_int32, _model_attrs


def time_forward_kernel(ren, dev_in, cur_models, mode, res, flush, reps=20):
    """Average duration (ms) of ONE ``dibr_forward_kernel`` launch for the student pass, CUDA events on the
    launching stream, L2 flushed before every launch (B200_PROFILING.md timing hygiene)."""
    dev = dev_in["Rs"].device
    names = [a for k, a in (("color", "colors"), ("norm", "normals"), ("xyz", "vertices")) if k in mode]
    split = [3] * len(names) + [1] + ([1] if "depth" in mode else [])
    flags = fused.FLAG_ONES | (fused.FLAG_DEPTH if "depth" in mode else 0)
    ren.dib_ren.set_camera_parameters_from_RT_K(dev_in["Rs"], dev_in["ts"], dev_in["Ks"], res, res, near=0.01, far=100)
    points = [[m["vertices"], _faces_int32(m["faces"])] for m in cur_models]
    attrs = [_model_attrs(m, names) for m in cur_models]
    orig = fused.build_meta

    def build_meta_keep(*a, **k):
        meta = orig(*a, **k)
        meta["keep_pass"] = True
        return meta
    fused.build_meta = build_meta_keep
    try:
        with torch.no_grad():
            _, _, _, meta = vc.render_instances(points, attrs, ren.dib_ren.camera_params, res, res, multi=False,
                                                want_normals=False, attr_flags=flags, out_split=split)
    finally:
        fused.build_meta = orig
    p, keep = meta["_last_pass"]
    lib = _lib.load()
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(3):
        _lib.check(lib.dibr_forward(ctypes.byref(p), st), "dibr_forward")
    torch.cuda.synchronize()
    evs = []
    import os
    for _ in range(reps):
        if os.environ.get("DIBR_NO_FLUSH") != "1":
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dibr_forward(ctypes.byref(p), st), "dibr_forward")
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    del keep
    return sum(a.elapsed_time(b) for a, b in evs) / reps
