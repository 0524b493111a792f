from .base import Renderer, K_to_fov, renderers
from .vc import VCRender, VCRenderBatch, VCRenderMulti
from .cameras import camera_params_from_RT_K, projection_from_K, quat2mat_torch
