"""Texture / spherical-harmonics / Phong render modes.

Same constructors, ``forward`` signatures and return tuples ``(imrender, improb, normal1, hardmask)`` as the
reference modules
  renderer/texrender.py:11-93         TexRender ("Lambertian" / "Texture")
  renderer/texrender_batch.py:13-128  TexRenderBatch     b x [verts_1xpx3, faces_fx3], b x uv, b x texture
  renderer/texrender_multi.py:12-138  TexRenderMulti     painter's-order composite of the per-object renders
  renderer/shrender.py:10-137         SHRender           (set_smooth(pfmtx) supported)
  renderer/phongrender.py:12-151      PhongRender
and the fragment shaders of renderer/fragment_shaders/{interpolation,frag_tex,frag_shtex,frag_phongtex}.py.

The structure is the reference's: vertex shader -> per-face features -> rasterizer -> fragment shader in plain torch.
The rasterizer is the B200 operator (``linear_rasterizer``: dibr_setup_faces / dibr_forward / dibr_backward_faces);
where the reference loops over the samples and calls the rasterizer once per object, the ragged batch is padded with
culled off-screen faces and rasterised in ONE call.  No CPU fallback: the operator raises on CPU tensors.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..rasterizer import linear_rasterizer

_EPS = 1e-15


def datanormalize(data, axis):
    """utils/utils.py:30-33"""
    return data / (torch.sqrt(torch.sum(data ** 2, dim=axis, keepdim=True)) + _EPS)


def project_faces(points_bxpx3, faces_fx3, cameras):
    """vertex_shaders/perpsective.py:26-111 (3x1 diagonal and 4x4 real projection): camera-space corner positions
    [b,f,9], projected corners [b,f,6], un-normalised face normals [b,f,3]."""
    cam_rot, cam_pos, cam_proj = cameras
    p = torch.matmul(points_bxpx3 - cam_pos.view(-1, 1, 3), cam_rot.permute(0, 2, 1))
    if cam_proj.shape[-1] == 4:
        p4 = torch.cat((p, torch.ones_like(p[:, :, :1])), dim=2)
        q = torch.matmul(p4, cam_proj.view(-1, 4, 4))
        xy = q[:, :, :2] / q[:, :, 3:4]
    else:
        q = p * cam_proj.view(-1, 1, 3)
        xy = q[:, :, :2] / q[:, :, 2:3]
    f = faces_fx3.long()
    c0, c1, c2 = p[:, f[:, 0]], p[:, f[:, 1]], p[:, f[:, 2]]
    points3d = torch.cat((c0, c1, c2), dim=2)
    points2d = torch.cat((xy[:, f[:, 0]], xy[:, f[:, 1]], xy[:, f[:, 2]]), dim=2)
    normal = torch.cross(c1 - c0, c2 - c0, dim=2)
    return points3d, points2d, normal


def face_uv_features(uv_bxpx2, ft_fx3):
    """[b,f,3,3]: (u, v, 1) per corner -- the trailing one rasterises to the hard mask (texrender.py:52-58)"""
    t = ft_fx3.long()
    uv = torch.stack((uv_bxpx2[:, t[:, 0]], uv_bxpx2[:, t[:, 1]], uv_bxpx2[:, t[:, 2]]), dim=2)       # b,f,3,2
    return torch.cat((uv, torch.ones_like(uv[..., :1])), dim=3)


def texinterpolation(imtexcoord_bxhxwx2, texture_bx3xthxtw, filtering="nearest"):
    """fragment_shaders/interpolation.py:27-48: OpenGL texture coordinates (0..1, y up, wrapping) looked up with
    grid_sample (-1..1, y down)."""
    uv = torch.remainder(imtexcoord_bxhxwx2, 1.0) * 2 - 1
    grid = torch.stack((uv[..., 0], -uv[..., 1]), dim=-1)
    return F.grid_sample(texture_bx3xthxtw, grid, mode=filtering).permute(0, 2, 3, 1)


def shade_tex(imtexcoord, texture, hardmask, filtering="nearest"):
    """frag_tex.py:28-39"""
    return torch.clamp(texinterpolation(imtexcoord, texture, filtering=filtering) * hardmask, 0, 1)


_SH = (0.2820948, 0.3257350, 0.2731371, 0.1365686, 0.0788479, 0.1931371)


def shade_sh(imnormal1, lightparam_bx9, imtexcoord, texture, hardmask):
    """frag_shtex.py:28-76: 9 real spherical-harmonics bands of the unit normal dotted with the light parameters"""
    x, y, z = imnormal1[..., 0:1], imnormal1[..., 1:2], imnormal1[..., 2:3]
    c0, c1, c2, c3, c4, c5 = _SH
    bands = torch.cat((c0 * torch.ones_like(x), -c1 * y, c1 * z, -c1 * x, c2 * (x * y), -c2 * (y * z),
                       c3 * (z * z) - c4, -c5 * (x * z), c3 * (x * x - y * y)), dim=3)
    coef = torch.sum(bands * lightparam_bx9.view(-1, 1, 1, 9), dim=3, keepdim=True)
    return torch.clamp(coef * texinterpolation(imtexcoord, texture) * hardmask, 0, 1)


def shade_phong(imnormal1, lightdirect1_bx3, eyedirect1, material_bx3x3, shininess_bx1, imtexcoord, texture, hardmask):
    """frag_phongtex.py:28-67: ambient + diffuse on the texture colour, specular added on top"""
    light = lightdirect1_bx3.view(-1, 1, 1, 3)
    cos_t = torch.clamp(torch.sum(imnormal1 * light, dim=3, keepdim=True), 0, 1)
    reflect = -light + 2 * cos_t * imnormal1
    cos_a = torch.clamp(torch.sum(reflect * eyedirect1, dim=3, keepdim=True), 1e-5, 1)
    cos_a = torch.pow(cos_a, shininess_bx1.view(-1, 1, 1, 1))
    amb = material_bx3x3[:, 0:1, :].view(-1, 1, 1, 3)
    dif = material_bx3x3[:, 1:2, :].view(-1, 1, 1, 3) * cos_t
    spe = material_bx3x3[:, 2:3, :].view(-1, 1, 1, 3) * cos_a
    color = (amb + dif) * texinterpolation(imtexcoord, texture) + spe
    return torch.clamp(color * hardmask, 0, 1)


# ------------------------------------------------------------------------------------------------
def _camera_of(cameras, i, single_intrinsic):
    return [cameras[0][i:i + 1], cameras[1][i:i + 1], cameras[2] if single_intrinsic else cameras[2][i]]


def _project_objects(points, cameras, uv_list, ft_list):
    """per object: vertex shader + uv features; then pad the ragged face lists to one [b,fmax,*] batch.  Padding faces
    are back-facing and sit at NDC (2,2): outside the image even after the soft-silhouette expansion."""
    b = len(points)
    single_intrinsic = True
    if cameras[2].ndim == 3:                       # texrender_batch.py:38-41
        assert cameras[2].shape[0] == b
        single_intrinsic = False
    p3s, p2s, nzs, n1s, feats = [], [], [], [], []
    for i in range(b):
        verts_1xpx3, faces = points[i]
        ft = faces if ft_list is None else ft_list[i]
        p3, p2, nrm = project_faces(verts_1xpx3, faces, _camera_of(cameras, i, single_intrinsic))
        p3s.append(p3)
        p2s.append(p2)
        nzs.append(nrm[:, :, 2:3])
        n1s.append(datanormalize(nrm, axis=2))
        feats.append(face_uv_features(uv_list[i], ft).reshape(1, faces.shape[0], 9))
    fmax = max(x.shape[1] for x in p3s)

    def pad(xs, value):
        out = []
        for x in xs:
            if x.shape[1] < fmax:
                x = torch.cat((x, x.new_full((1, fmax - x.shape[1], x.shape[2]), value)), dim=1)
            out.append(x)
        return torch.cat(out, dim=0).contiguous()
    return pad(p3s, 0.0), pad(p2s, 2.0), pad(nzs, -1.0), n1s, pad(feats, 0.0)


class TexRender(nn.Module):
    """one topology, a batch of vertex sets (texrender.py:11-93)"""

    def __init__(self, height, width, filtering="nearest"):
        super(TexRender, self).__init__()
        self.height, self.width, self.filtering = height, width, filtering

    def forward(self, points, cameras, uv_bxpx2, texture_bx3xthxtw, ft_fx3=None):
        points_bxpx3, faces_fx3 = points
        ft = faces_fx3 if ft_fx3 is None else ft_fx3
        p3, p2, nrm = project_faces(points_bxpx3, faces_fx3, cameras)
        feat = face_uv_features(uv_bxpx2, ft).reshape(p3.shape[0], p3.shape[1], 9)
        imfeat, improb = linear_rasterizer(self.width, self.height, p3, p2, nrm[:, :, 2:3], feat)
        hardmask = imfeat[:, :, :, 2:3]
        imrender = shade_tex(imfeat[:, :, :, :2], texture_bx3xthxtw, hardmask, filtering=self.filtering)
        return imrender, improb, datanormalize(nrm, axis=2), hardmask


class TexRenderBatch(nn.Module):
    """one textured object per image, different objects allowed (texrender_batch.py:13-128)"""

    def __init__(self, height, width, filtering="nearest"):
        super(TexRenderBatch, self).__init__()
        self.height, self.width, self.filtering = height, width, filtering

    def _render(self, points, cameras, uv_bxpx2, texture_bx3xthxtw, ft_fx3):
        assert len(points) > 0, len(points)
        p3, p2, nz, normal1_list, feat = _project_objects(points, cameras, uv_bxpx2, ft_fx3)
        imfeat, improb = linear_rasterizer(self.width, self.height, p3, p2, nz, feat)
        hardmask = imfeat[:, :, :, 2:3]
        # textures may differ in size from object to object: the lookup stays per object (frag_tex.py), as in the reference
        ims = [shade_tex(imfeat[i:i + 1, :, :, :2], texture_bx3xthxtw[i], hardmask[i:i + 1]) for i in range(len(points))]
        return torch.cat(ims, dim=0), improb, normal1_list, hardmask

    def forward(self, points, cameras, uv_bxpx2, texture_bx3xthxtw, ft_fx3=None):
        return self._render(points, cameras, uv_bxpx2, texture_bx3xthxtw, ft_fx3)


class TexRenderMulti(TexRenderBatch):
    """a scene of textured objects: every object rendered on its own, then composited far-to-near by the z of its
    translation with the hard masks (texrender_multi.py:32-33,109-136 -- the reference's painter's order, 'not True but
    very close', kept as is)"""

    def forward(self, points, cameras, uv_bxpx2, texture_bx3xthxtw, ts, ft_fx3=None):
        ims, probs, normal1_list, masks = self._render(points, cameras, uv_bxpx2, texture_bx3xthxtw, ft_fx3)
        order = np.argsort(np.array([float(t[2]) for t in ts]))[::-1]
        first = int(order[0])
        imrender, improb, fg = ims[first:first + 1], probs[first:first + 1], masks[first:first + 1]
        for i in order[1:]:
            i = int(i)
            on = masks[i:i + 1] > 0.5
            imrender = torch.where(on, ims[i:i + 1], imrender)
            improb = torch.where(on, probs[i:i + 1], improb)
            fg = torch.where(on, masks[i:i + 1], fg)
        return imrender, improb, normal1_list, fg


class _LitRender(nn.Module):
    def __init__(self, height, width):
        super(_LitRender, self).__init__()
        self.height, self.width = height, width
        self.smooth = False
        self.pfmtx = None

    def _corner_normals(self, normal_bxfx3, faces_fx3):
        """flat: the face normal at every corner; smooth: vertex normals = pfmtx @ face normals (shrender.py:69-77)"""
        if not self.smooth:
            return normal_bxfx3.unsqueeze(2).expand(-1, -1, 3, -1)
        pf = self.pfmtx if torch.is_tensor(self.pfmtx) else torch.as_tensor(self.pfmtx)
        pf = pf.to(normal_bxfx3).reshape(-1, pf.shape[-2], pf.shape[-1])
        vn = torch.matmul(pf.expand(normal_bxfx3.shape[0], -1, -1) if pf.shape[0] == 1 else pf, normal_bxfx3)
        f = faces_fx3.long()
        return torch.stack((vn[:, f[:, 0]], vn[:, f[:, 1]], vn[:, f[:, 2]]), dim=2)


class SHRender(_LitRender):
    """spherical-harmonics lighting (shrender.py:10-137)"""

    def set_smooth(self, pfmtx):
        self.smooth = True
        self.pfmtx = pfmtx

    def forward(self, points, cameras, uv_bxpx2, texture_bx3xthxtw, lightparam, ft_fx3=None):
        assert lightparam is not None, "When using the Spherical Harmonics model, light parameters must be passed"
        points_bxpx3, faces_fx3 = points
        ft = faces_fx3 if ft_fx3 is None else ft_fx3
        p3, p2, nrm = project_faces(points_bxpx3, faces_fx3, cameras)
        b, f = p3.shape[0], p3.shape[1]
        feat = torch.cat((self._corner_normals(nrm, faces_fx3), face_uv_features(uv_bxpx2, ft)), dim=3).reshape(b, f, 18)
        imfeat, improb = linear_rasterizer(self.width, self.height, p3, p2, nrm[:, :, 2:3], feat)
        hardmask = imfeat[:, :, :, 5:]
        imrender = shade_sh(datanormalize(imfeat[:, :, :, :3], axis=3), lightparam, imfeat[:, :, :, 3:5],
                            texture_bx3xthxtw, hardmask)
        return imrender, improb, datanormalize(nrm, axis=2), hardmask


class PhongRender(_LitRender):
    """Phong lighting (phongrender.py:12-151)"""

    def set_smooth(self, pfmtx):
        self.smooth = True
        self.pfmtx = torch.as_tensor(pfmtx).reshape(1, pfmtx.shape[0], pfmtx.shape[1])

    def forward(self, points, cameras, uv_bxpx2, texture_bx3xthxtw, lightdirect_bx3, material_bx3x3, shininess_bx1,
                ft_fx3=None):
        assert lightdirect_bx3 is not None, "When using the Phong model, light parameters must be passed"
        assert material_bx3x3 is not None, "When using the Phong model, material parameters must be passed"
        assert shininess_bx1 is not None, "When using the Phong model, shininess parameters must be passed"
        points_bxpx3, faces_fx3 = points
        ft = faces_fx3 if ft_fx3 is None else ft_fx3
        p3, p2, nrm = project_faces(points_bxpx3, faces_fx3, cameras)
        b, f = p3.shape[0], p3.shape[1]
        eye = -p3.reshape(b, f, 3, 3)                                   # towards the camera, per corner
        feat = torch.cat((self._corner_normals(nrm, faces_fx3), eye, face_uv_features(uv_bxpx2, ft)), dim=3).reshape(b, f, 27)
        imfeat, improb = linear_rasterizer(self.width, self.height, p3, p2, nrm[:, :, 2:3], feat)
        immask = imfeat[:, :, :, 8:9]
        imrender = shade_phong(datanormalize(imfeat[:, :, :, :3], axis=3), datanormalize(lightdirect_bx3, axis=1),
                               datanormalize(imfeat[:, :, :, 3:6], axis=3), material_bx3x3, shininess_bx1,
                               imfeat[:, :, :, 6:8], texture_bx3xthxtw, immask)
        return imrender, improb, datanormalize(nrm, axis=2), immask
