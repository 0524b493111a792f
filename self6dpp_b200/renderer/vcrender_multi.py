"""import-path parity with the reference's renderer/vcrender_multi.py"""
from .vc import VCRenderMulti  # noqa: F401
