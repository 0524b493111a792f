"""``Renderer`` (exported as ``DIBRenderer``): same constructor, attributes and methods as the
reference's /root/reference/lib/dr_utils/dib_renderer_x/renderer/base.py:53-191 -- ``mode``,
``camera_params``, ``renderer``, ``forward(points, *args, **kwargs)``, ``set_camera_parameters``,
``set_camera_parameters_from_RT_K``, ``set_look_at_parameters`` -- with the vertex-colour modes
running on the fused B200 kernels.  Differences, all deliberate:
  * tensors live on the inputs' device (the reference hard-codes cuda:0, base.py:69,153,164-166);
  * camera set-up is a few batched torch ops instead of a Python loop over the batch;
  * the texture / SH / Phong modes (renderer/tex.py) keep the reference's structure -- torch vertex shader, the B200
    rasterizer operator, torch fragment shader -- with the per-object rasterizer loop replaced by one padded call.
"""
import numpy as np
import torch
import torch.nn as nn

from .cameras import camera_params_from_RT_K, look_at_camera_params
from .tex import PhongRender, SHRender, TexRender, TexRenderBatch, TexRenderMulti
from .vc import VCRender, VCRenderBatch, VCRenderMulti

renderers = {
    "VertexColor": VCRender,
    "VertexColorMulti": VCRenderMulti,
    "VertexColorBatch": VCRenderBatch,
    "Lambertian": TexRender,
    "Texture": TexRender,  # alias
    "TextureMulti": TexRenderMulti,
    "TextureBatch": TexRenderBatch,
    "SphericalHarmonics": SHRender,
    "Phong": PhongRender,
}


def perspectiveprojectionnp(fovy, ratio=1.0, near=0.01, far=10.0):
    """utils/perspective.py:75-93: the 3x1 diagonal projection used by the look-at path."""
    tanfov = np.tan(fovy / 2.0)
    return np.array([[1.0 / (ratio * tanfov)], [1.0 / tanfov], [-1]], dtype=np.float32)


class Renderer(nn.Module):
    def __init__(self, height, width, mode="VertexColor", camera_center=None, camera_up=None, camera_fov_y=None):
        super(Renderer, self).__init__()
        assert mode in renderers, "Passed mode {0} must in in list of accepted modes: {1}".format(mode, renderers)
        self.mode = mode
        self.height = height
        self.width = width
        self.renderer = renderers[mode](height, width)
        if camera_center is None:
            self.camera_center = np.array([0, 0, 0], dtype=np.float32)
        if camera_up is None:
            self.camera_up = np.array([0, 1, 0], dtype=np.float32)
        if camera_fov_y is None:
            self.camera_fov_y = 49.13434207744484 * np.pi / 180.0
        self._camera_params = None
        self._camera_pending = None

    @property
    def camera_params(self):
        if self._camera_params is None and self._camera_pending is not None:
            Rs, ts, Ks, height, width, near, far, rot_type = self._camera_pending
            self._camera_params = camera_params_from_RT_K(Rs, ts, Ks, height, width, near=near, far=far, rot_type=rot_type)
            self._camera_pending = None
        return self._camera_params

    @camera_params.setter
    def camera_params(self, value):
        self._camera_params = value
        self._camera_pending = None

    def set_camera_parameters_lazy(self, Rs, ts, Ks, height, width, near, far, rot_type):
        """same result as set_camera_parameters_from_RT_K, computed only if somebody reads .camera_params"""
        d = self.__dict__             # plain attributes: nn.Module.__setattr__ costs ~6 us per store on this per-call path
        d["_camera_params"] = None
        d["_camera_pending"] = (Rs, ts, Ks, height, width, near, far, rot_type)

    def forward(self, points, *args, **kwargs):
        if self.camera_params is None:
            print("Camera parameters have not been set, default perspective parameters of distance = 1, "
                  "elevation = 30, azimuth = 0 are being used")
            self.set_look_at_parameters([0], [30], [1])
        if self.mode in ["VertexColorMulti", "VertexColorBatch", "TextureMulti", "TextureBatch"]:
            assert self.camera_params[0].shape[0] == len(points), \
                "multi mode need the same length of camera parameters and points"
        else:
            assert self.camera_params[0].shape[0] == points[0].shape[0], \
                "Set camera parameters batch size must equal batch size of passed points"
        return self.renderer(points, self.camera_params, *args, **kwargs)

    def set_look_at_parameters(self, azimuth, elevation, distance):
        device = torch.device("cuda", torch.cuda.current_device())
        proj = torch.tensor(perspectiveprojectionnp(self.camera_fov_y, 1.0), dtype=torch.float32, device=device)
        mats, shifts = [], []
        for a, e, d in zip(azimuth, elevation, distance):
            mat, pos = look_at_camera_params(a, e, d)
            mats.append(mat)
            shifts.append(pos)
        self.camera_params = [torch.stack(mats).to(device), torch.stack(shifts).to(device), proj]

    def set_camera_parameters(self, parameters):
        self.camera_params = parameters

    def set_camera_parameters_from_RT_K(self, Rs, ts, Ks, height, width, near=0.01, far=10.0, rot_type="mat"):
        """
        Rs: a list of rotations tensor (or a [b,3,3] / [b,4] tensor)
        ts: a list of translations tensor (or [b,3])
        Ks: a list of camera intrinsic matrices, a [b,3,3] tensor, or a single [3,3] matrix
        ----
        [cam_view_R, cam_view_pos, cam_proj]
        """
        self.camera_params = camera_params_from_RT_K(Rs, ts, Ks, height, width, near=near, far=far, rot_type=rot_type)


def K_to_fov(K, height, width):
    fx = K[0, 0]
    fy = K[1, 1]
    fov_x = 2 * np.arctan2(width, 2 * fx)  # radian
    fov_y = 2 * np.arctan2(height, 2 * fy)
    return fov_x, fov_y
