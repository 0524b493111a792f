"""BENCH / TEST INFRASTRUCTURE ONLY (never imported by self6dpp_b200/).

Python face of oracle/kaolin_structure.cu: a GPU stand-in with the STRUCTURE of the kernels the reference reaches through
``kaolin.graphics.dib_renderer.cuda.rasterizer`` (one thread per pixel over all faces, fp32 atomics in the backward;
/root/reference/lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172, 249-269), plus an FP32 FMA peak probe.
bench.py reports it as ``kaolin_structure_gpu``: the reference's own algorithm on the same box."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libkaolin_structure.so")
    src = os.path.join(_HERE, "kaolin_structure.cu")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.run(["make", "-C", _HERE, "-s", "gpu"], check=True, stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ks_fma_peak_tflops.restype = ctypes.c_double
        _LIB.ks_fma_peak_tflops.argtypes = [ctypes.c_int]
    return _LIB


def _p(t):
    assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    return ctypes.c_void_p(t.data_ptr())


def fma_peak_tflops(reps=5):
    return float(lib().ks_fma_peak_tflops(int(reps)))


class Pass(object):
    """Buffers of one LinearRasterizer call in the reference's layout (rasterizer.py:120-150: the five B x H x W x K
    scratch tensors included)."""

    def __init__(self, p3, p2, nz, attr, height, width, expand=0.02, knum=30, multiplier=1000, delta=7000):
        dev = p3.device
        self.b, self.f = p3.shape[:2]
        self.h, self.w, self.d, self.k, self.m, self.delta = height, width, attr.shape[2] // 3, knum, multiplier, delta
        self.p3 = p3.contiguous()
        self.p2m = (multiplier * p2).contiguous()
        v = self.p2m.view(self.b, self.f, 3, 2)
        pmin, pmax = v.min(2)[0], v.max(2)[0]
        self.bbox = torch.cat((pmin, pmax), 2).contiguous()
        self.bbox2 = torch.cat((pmin - expand * multiplier, pmax + expand * multiplier), 2).contiguous()
        self.nz, self.attr = nz.contiguous(), attr.contiguous()
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.imidx, self.imdep, self.imwei = z(self.b, height, width, 1), z(self.b, height, width, 1), z(self.b, height, width, 3)
        self.im, self.improb = z(self.b, height, width, self.d), z(self.b, height, width, 1)
        self.probface, self.probcase, self.probdis = (z(self.b, height, width, knum) for _ in range(3))
        self.g_p2, self.g_p2prob, self.g_attr = z(self.b, self.f, 6), z(self.b, self.f, 6), z(self.b, self.f, 3 * self.d)

    def forward(self):
        for t in (self.imidx, self.imwei, self.im, self.improb, self.probface, self.probcase, self.probdis):
            t.zero_()
        self.imdep.fill_(-1000.0)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.p3.device).cuda_stream)
        rc = lib().ks_forward(_p(self.p3), _p(self.p2m), _p(self.nz), _p(self.bbox), _p(self.bbox2), _p(self.attr), _p(self.imidx),
                              _p(self.imdep), _p(self.imwei), _p(self.probface), _p(self.probcase), _p(self.probdis), _p(self.im),
                              _p(self.improb), self.b, self.h, self.w, self.f, self.d, self.k, self.m, self.delta, st)
        assert rc == 0, rc

    def backward(self, g_im, g_prob):
        for t in (self.g_p2, self.g_p2prob, self.g_attr):
            t.zero_()
        st = ctypes.c_void_p(torch.cuda.current_stream(self.p3.device).cuda_stream)
        rc = lib().ks_backward(_p(g_im), _p(g_prob), _p(self.improb), _p(self.imidx), _p(self.imwei), _p(self.probface), _p(self.probcase),
                               _p(self.probdis), _p(self.p2m), _p(self.attr), _p(self.g_p2), _p(self.g_attr), _p(self.g_p2prob),
                               self.b, self.h, self.w, self.f, self.d, self.k, self.m, self.delta, st)
        assert rc == 0, rc
        return self.g_p2 + self.g_p2prob, self.g_attr
