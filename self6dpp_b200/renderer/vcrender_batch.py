"""import-path parity with the reference's renderer/vcrender_batch.py"""
from .vc import VCRenderBatch  # noqa: F401
