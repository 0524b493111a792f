"""Helper entry points of ``lib/dr_utils/dr_utils.py`` on the B200 renderer: same names, arguments and return
tuples, so scripts written against ``lib.dr_utils`` switch by changing the import.

  load_objs               dr_utils.py:17-72     OBJ (+ texture) files -> list of model dicts with a leading 1-axis
  render_dib_vc_batch     dr_utils.py:75-122    b objects, one image each, vertex colours
  render_dib_tex_batch    dr_utils.py:125-182   b objects, one image each, textured
  render_dib_vc_multi     dr_utils.py:185-207   all objects z-buffered into one image, vertex colours
  render_dib_tex_multi    dr_utils.py:210-245   all objects composited into one image, textured

``with_depth=True`` costs the reference a second full rasterisation of the same geometry (the camera-space
vertices rendered as a colour, dr_utils.py:104-118).  Here the vertex-colour helper asks the fused rasterisation
for the view depth as one more channel of the SAME launch (the depth group of ``dibr_forward``); the values are
the same interpolation of z(R v + t).
"""
import os.path as osp

import numpy as np
import torch

from . import fused
from .models import load_obj
from .renderer.base import Renderer as DIBRenderer
from .renderer.cameras import quat2mat_torch
from .renderer.vc import render_instances


def load_objs(obj_paths, texture_paths=None, height=480, width=640, centring=True, tex_resize=True, tex_fmt="CHW",
              tex_vflip=False, device="cuda"):
    """dr_utils.py:17-72.  Each model: 'vertices' [1,n,3] (minus the middle of the global min/max coordinate when
    ``centring``), 'colors' [1,n,3], 'faces' int32 [1,m,3]; with textures also 'face_uvs' [1,t,2], 'face_uv_ids'
    [1,m,3] and 'texture' ([1,3,h,w] for "CHW", [1,h,w,3] otherwise; RGB in [0,1]).  ``height``/``width`` only
    matter when ``tex_resize``.  ``device`` is ours (the reference hard-codes ``.cuda()``)."""
    assert all(".obj" in p for p in obj_paths)
    if texture_paths is not None:
        assert len(obj_paths) == len(texture_paths)
    models = []
    for i, obj_path in enumerate(obj_paths):
        mesh = load_obj(obj_path)
        v = mesh["vertices"]
        vertices, colors = v[:, :3], v[:, 3:6]
        if centring:
            vertices = vertices - (vertices.max() + vertices.min()) / 2.0
        model = {"vertices": vertices[None].contiguous().to(device), "colors": colors[None].contiguous().to(device),
                 "faces": mesh["faces"].int()[None].to(device)}
        if texture_paths is not None:
            import cv2  # only the textured path needs it

            assert osp.exists(texture_paths[i]), texture_paths[i]
            img = cv2.imread(texture_paths[i], cv2.IMREAD_COLOR)
            img = img[::-1, :, ::-1] if tex_vflip else img[:, :, ::-1]
            texture = img.astype(np.float32) / 255.0
            if tex_resize:
                texture = cv2.resize(texture, (width, height), interpolation=cv2.INTER_AREA)
            texture = np.ascontiguousarray(texture.transpose(2, 0, 1) if tex_fmt == "CHW" else texture)
            model["face_uvs"] = mesh["uvs"][None].to(device)
            model["face_uv_ids"] = mesh["face_textures"][None].to(device)
            model["texture"] = torch.from_numpy(texture)[None].to(device)
        models.append(model)
    return models


def _per_sample_Ks(Ks, bs):
    if len(Ks) == 1:                                             # dr_utils.py:95-96
        return [Ks[0] for _ in range(bs)]
    return Ks


def _points(models, obj_ids):
    # the reference converts faces with .long() per call; the kernels take int32 and the conversion is cached per model
    return [[models[i]["vertices"], models[i]["faces"][0]] for i in obj_ids]


def _view_depth_xyz(Rs, ts, rot_type, models, obj_ids):
    """camera-space vertices per object (dr_utils.py:104-116 / lib/pysixd/misc.py:985-1004)"""
    if not isinstance(Rs, torch.Tensor):
        Rs = torch.stack(list(Rs))
    R_mats = quat2mat_torch(Rs) if rot_type == "quat" else Rs
    return [(models[o]["vertices"][0] @ R_mats[i].t() + ts[i].view(1, 3))[None] for i, o in enumerate(obj_ids)]


def render_dib_vc_batch(ren, Rs, ts, Ks, obj_ids, models, rot_type="quat", H=480, W=640, near=0.01, far=100.0,
                        with_depth=False):
    """-> (color b,h,w,3; prob b,h,w,1; mask b,h,w,1; depth b,h,w or None)"""
    assert ren.mode in ["VertexColorBatch"], ren.mode
    bs = len(Rs)
    Ks = _per_sample_Ks(Ks, bs)
    ren.set_camera_parameters_from_RT_K(Rs, ts, Ks, height=H, width=W, near=near, far=far, rot_type=rot_type)
    colors = [models[i]["colors"] for i in obj_ids]
    points = _points(models, obj_ids)
    if not with_depth:
        predictions, im_probs, _, im_masks = ren.forward(points=points, colors=colors)
        return predictions, im_probs, im_masks, None
    # colour, ones and the view depth out of one rasterisation
    (predictions, im_masks, depth), im_probs, _, _ = render_instances(
        points, colors, ren.camera_params, H, W, multi=False, want_normals=False,
        attr_flags=fused.FLAG_ONES | fused.FLAG_DEPTH, out_split=[3, 1, 1])
    return predictions, im_probs, im_masks, depth.squeeze(-1)


def render_dib_tex_batch(ren, Rs, ts, Ks, obj_ids, models, rot_type="quat", H=480, W=640, near=0.01, far=100.0,
                         with_depth=False):
    """-> (rgb b,h,w,3; prob b,h,w,1; mask b,h,w,1; depth b,h,w or None)"""
    assert ren.mode in ["TextureBatch"], ren.mode
    bs = len(Rs)
    Ks = _per_sample_Ks(Ks, bs)
    ren.set_camera_parameters_from_RT_K(Rs, ts, Ks, height=H, width=W, near=near, far=far, rot_type=rot_type)
    points = _points(models, obj_ids)
    im, prob, _, mask = ren.forward(points=points, uv_bxpx2=[models[i]["face_uvs"] for i in obj_ids],
                                    texture_bx3xthxtw=[models[i]["texture"] for i in obj_ids],
                                    ft_fx3=[models[i]["face_uv_ids"][0] for i in obj_ids])
    depth = None
    if with_depth:
        xyzs = _view_depth_xyz(Rs, ts, rot_type, models, obj_ids)
        vc = DIBRenderer(height=H, width=W, mode="VertexColorBatch")
        vc.set_camera_parameters(ren.camera_params)
        ren_xyzs, _, _, _ = vc.forward(points=points, colors=xyzs)
        depth = ren_xyzs[:, :, :, 2]
    return im, prob, mask, depth


def render_dib_vc_multi(ren, Rs, ts, K, obj_ids, models, rot_type="quat", H=480, W=640, near=0.01, far=100.0):
    """-> (rgb 1,h,w,3; prob 1,h,w,1; mask 1,h,w,1)"""
    assert ren.mode in ["VertexColorMulti"], ren.mode
    ren.set_camera_parameters_from_RT_K(Rs, ts, K, height=H, width=W, near=near, far=far, rot_type=rot_type)
    predictions, im_prob, _, im_mask = ren.forward(points=_points(models, obj_ids),
                                                   colors=[models[i]["colors"] for i in obj_ids])
    return predictions, im_prob, im_mask


def render_dib_tex_multi(ren, Rs, ts, K, obj_ids, models, rot_type="quat", H=480, W=640, near=0.01, far=100.0):
    """-> (rgb 1,h,w,3; prob 1,h,w,1; mask 1,h,w,1)"""
    assert ren.mode in ["TextureMulti"], ren.mode
    ren.set_camera_parameters_from_RT_K(Rs, ts, K, height=H, width=W, near=near, far=far, rot_type=rot_type)
    im, prob, _, mask = ren.forward(points=_points(models, obj_ids), uv_bxpx2=[models[i]["face_uvs"] for i in obj_ids],
                                    texture_bx3xthxtw=[models[i]["texture"] for i in obj_ids], ts=ts,
                                    ft_fx3=[models[i]["face_uv_ids"][0] for i in obj_ids])
    return im, prob, mask
