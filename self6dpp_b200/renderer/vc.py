"""Vertex-colour render modes on the fused B200 op.  Same constructor / forward signatures and
return tuples as the reference modules
  renderer/vcrender.py:41-87        VCRender        points=[verts_bxpx3, faces_fx3], colors_bxpx3
  renderer/vcrender_batch.py:30-139 VCRenderBatch   points = b x [verts_1xpx3, faces_fx3], colors = b x [1xpx3]
  renderer/vcrender_multi.py:20-130 VCRenderMulti   same inputs, all objects rasterised into ONE image
all returning ``(imrender, improb, normal1, hardmask)``.  The reference loops over samples in
Python and issues b rasterizer calls; here every mode is one ``dibr_setup_meshes`` + one
``dibr_forward`` launch for the whole batch.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import fused
from .cameras import proj_as_4x4


def _flat_verts(v):
    return v.reshape(-1, 3)


def _cat_cached(tensors, cache):
    """torch.cat of static tensors, memoised by storage identity + version (models are resident and
    unchanged between iterations); anything that requires grad is concatenated afresh so autograd sees it."""
    if len(tensors) == 1:
        return tensors[0]
    if any(t.requires_grad for t in tensors):
        return torch.cat(tensors, dim=0)
    key = tuple((t.data_ptr(), tuple(t.shape), t._version) for t in tensors)
    hit = cache.get(key)
    if hit is None:
        if len(cache) > 32:
            cache.clear()
        hit = (torch.cat(tensors, dim=0), list(tensors))
        cache[key] = hit
    return hit[0]


_CAT_CACHE = {}


def render_instances(points, colors, cameras, height, width, multi, want_normals=True, attr_flags=fused.FLAG_ONES,
                     expand=None, knum=None, multiplier=None, delta=None, out_split=None):
    """points: n x [verts (..,p,3), faces (f,3)]; colors: n x [(..,p,c)] or None; cameras [rot nx3x3, pos nx3, proj].
    Returns (outs: list of [B,H,W,c] per out_split group, improb [B,H,W,1], face_normal [TF,3], meta)."""
    n = len(points)
    cam_rot, cam_pos, cam_proj = cameras[0], cameras[1], proj_as_4x4(cameras[2])
    assert cam_rot.shape[0] == n and cam_pos.shape[0] == n, "multi mode need the same length of camera parameters and points"
    single_proj = cam_proj.shape[0] == 1
    if not single_proj:
        assert cam_proj.shape[0] == n
    verts_in = [_flat_verts(p[0]) for p in points]
    faces_in = [p[1] for p in points]
    for v in verts_in:
        fused._require_cuda_f32("points", v)
    # distinct meshes (topology + vertices) by storage identity
    pairs, mesh_ids, seen = [], [], {}
    for v, f in zip(verts_in, faces_in):
        k = (v.data_ptr(), f.data_ptr(), tuple(v.shape), tuple(f.shape))
        if k not in seen:
            seen[k] = len(pairs)
            pairs.append((v, f))
        mesh_ids.append(seen[k])
    pack = fused.get_mesh_pack([p[0] for p in pairs], [p[1] for p in pairs])
    verts_packed = _cat_cached([p[0] for p in pairs], _CAT_CACHE)
    if colors is not None:
        cols_in = [c.reshape(-1, c.shape[-1]) for c in colors]
        attr_dim = cols_in[0].shape[-1]
        cdist, attr_ids = fused.dedup(cols_in)
        for i, c in enumerate(cols_in):
            if c.shape[0] != verts_in[i].shape[0]:
                raise RuntimeError("colors and points must have the same number of vertices")
        attr_base = np.concatenate([[0], np.cumsum([c.shape[0] for c in cdist])]).astype(np.int64)
        vattr_packed = _cat_cached(cdist, _CAT_CACHE)
    else:
        attr_dim, attr_ids, attr_base = 0, [0] * n, np.zeros(2, np.int64)
        vattr_packed = torch.zeros(0, 1, dtype=torch.float32, device=verts_packed.device)
    need_rows = verts_packed.requires_grad or vattr_packed.requires_grad
    image_ids = [0] * n if multi else list(range(n))
    kw = {}
    if expand is not None:
        kw["expand"] = expand
    if knum is not None:
        kw["knum"] = knum
    if multiplier is not None:
        kw["multiplier"] = multiplier
    if delta is not None:
        kw["delta"] = delta
    meta = fused.build_meta(pack, mesh_ids, attr_ids, attr_base, image_ids, height, width, attr_dim, attr_flags,
                            proj_ids=None if single_proj else list(range(n)), want_normals=want_normals,
                            need_rows=need_rows, num_attr_rows=int(attr_base[-1]), out_split=out_split, **kw)
    res = fused.render_meshes(verts_packed, vattr_packed, cam_rot, cam_pos, cam_proj, meta)
    return list(res[:-2]), res[-2], res[-1], meta


class VCRenderBatch(nn.Module):
    """Vertex-Color Renderer Batch: one object per image, different objects allowed (vcrender_batch.py:19-28)."""

    def __init__(self, height, width):
        super(VCRenderBatch, self).__init__()
        self.height = height
        self.width = width

    def forward(self, points, cameras, colors):
        d = colors[0].shape[-1]
        (imrender, hardmask), improb, fnormal, meta = render_instances(points, colors, cameras, self.height, self.width,
                                                                       multi=False, out_split=[d, 1])
        off = meta["face_offsets_host"]
        normal1_list = [fnormal[int(off[i]):int(off[i + 1])].unsqueeze(0) for i in range(len(points))]
        return imrender, improb, normal1_list, hardmask


class VCRenderMulti(nn.Module):
    """Vertex-Color Renderer for a scene: all objects z-buffered into one image (vcrender_multi.py:92-106)."""

    def __init__(self, height, width):
        super(VCRenderMulti, self).__init__()
        self.height = height
        self.width = width

    def forward(self, points, cameras, colors):
        d = colors[0].shape[-1]
        (imrender, hardmask), improb, fnormal, meta = render_instances(points, colors, cameras, self.height, self.width,
                                                                       multi=True, out_split=[d, 1])
        return imrender, improb, fnormal.unsqueeze(0), hardmask


class VCRender(nn.Module):
    """Vertex-Color Renderer: one topology, a batch of vertex sets (vcrender.py:41-87)."""

    def __init__(self, height, width):
        super(VCRender, self).__init__()
        self.height = height
        self.width = width

    def forward(self, points, cameras, colors_bxpx3):
        points_bxpx3, faces_fx3 = points
        b, p = points_bxpx3.shape[0], points_bxpx3.shape[1]
        fused._require_cuda_f32("points", points_bxpx3)
        pack = fused.get_mesh_pack([points_bxpx3[0]], [faces_fx3])
        cam_rot, cam_pos, cam_proj = cameras[0], cameras[1], proj_as_4x4(cameras[2])
        single_proj = cam_proj.shape[0] == 1
        verts_packed = points_bxpx3.reshape(-1, 3)
        vattr_packed = colors_bxpx3.reshape(-1, colors_bxpx3.shape[-1])
        rows = [i * p for i in range(b)]
        need_rows = verts_packed.requires_grad or vattr_packed.requires_grad
        meta = fused.build_meta(pack, [0] * b, list(range(b)), np.asarray(rows + [b * p]), list(range(b)),
                                self.height, self.width, colors_bxpx3.shape[-1], fused.FLAG_ONES,
                                proj_ids=None if single_proj else list(range(b)), vert_rows_base=rows,
                                want_normals=True, need_rows=need_rows, num_attr_rows=b * p,
                                out_split=[colors_bxpx3.shape[-1], 1])
        imrender, hardmask, improb, fnormal = fused.render_meshes(verts_packed, vattr_packed, cam_rot, cam_pos, cam_proj, meta)
        return imrender, improb, fnormal.reshape(b, -1, 3), hardmask
