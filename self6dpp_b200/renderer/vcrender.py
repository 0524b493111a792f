"""import-path parity with the reference's renderer/vcrender.py"""
from .vc import VCRender  # noqa: F401
