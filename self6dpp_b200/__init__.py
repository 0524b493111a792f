"""self6dpp_b200 -- B200-native DIB-R differentiable rasterizer behind the Python API of
EricZhengYX/self6dpp's ``lib/dr_utils/dib_renderer_x``.

    from self6dpp_b200 import DIBRenderer, Renderer_dibr, linear_rasterizer

mirror ``lib.dr_utils.dib_renderer_x.DIBRenderer`` (__init__.py:2),
``lib.dr_utils.dib_renderer_x.renderer_dibr.Renderer_dibr`` (renderer_dibr.py:95) and
``lib.dr_utils.dib_renderer_x.rasterizer.linear_rasterizer`` (rasterizer.py:294).
Compute happens in hand-written sm_100a CUDA (libdibr_b200.so); nothing here falls back to
PyTorch or the CPU.
"""
from .rasterizer import LinearRasterizer, linear_rasterizer  # noqa: F401
from .renderer.base import Renderer as DIBRenderer  # noqa: F401
from .renderer_dibr import Renderer_dibr  # noqa: F401

__all__ = ["DIBRenderer", "Renderer_dibr", "linear_rasterizer", "LinearRasterizer"]
