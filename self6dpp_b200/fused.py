"""Fused vertex-shader + rasterizer op: object-space meshes + per-instance cameras in, images out.

One autograd ``Function`` replaces, for a whole ragged batch, what the reference does per sample in
Python: ``perspective_projection`` (renderer/vertex_shaders/perpsective.py:29-111), the per-face
attribute gather with the ones channel (renderer/vcrender_batch.py:84-88), ``prepare_tfpoints`` and
``LinearRasterizer`` (rasterizer/rasterizer.py:36-294) -- and, in backward, the torch autograd tail
down to ``cam_view_R`` / ``cam_view_pos`` (renderer/base.py:169-170).

Terminology: an *instance* is one mesh under one camera; an *image* is rendered from one or more
instances (VertexColorBatch: one instance per image; VertexColorMulti: all instances in one image,
vcrender_multi.py:92-96).
"""
import ctypes

import numpy as np
import torch
from torch.autograd import Function

from . import _lib
from .rasterizer import (DEFAULT_DELTA, DEFAULT_EXPAND, DEFAULT_KNUM, DEFAULT_MULTIPLIER, _alloc_workspace, _on_device,
                         _base_pass, _require_cuda_f32, _stream)

INST_STRIDE = 12
FLAG_ONES = 1
FLAG_DEPTH = 2
FLAG_ATTR_GRAD_SCRATCH = 4      # include/dibr_b200.h attr_flags bit 2: grad_face_attr is scratch (depth column only, [F, 3])


class MeshPack(object):
    """Distinct meshes packed into flat device arrays + the vertex -> (face, corner) CSR adjacency
    the deterministic vertex gather in ``dibr_backward_meshes`` walks.  Built with device-side torch
    ops only (no host sync) and cached per set of face tensors (the reference keeps its models
    resident on the GPU too, self_engine_utils.py:1361-1373)."""

    def __init__(self, verts_list, faces_list, device):
        self.device = device
        self.n_meshes = len(verts_list)
        nv = [int(v.shape[-2]) for v in verts_list]
        nf = [int(f.shape[0]) for f in faces_list]
        self.num_verts = nv
        self.num_faces = nf
        self.vert_base = np.concatenate([[0], np.cumsum(nv)]).astype(np.int64)
        self.face_base = np.concatenate([[0], np.cumsum(nf)]).astype(np.int64)
        total_v, total_f = int(self.vert_base[-1]), int(self.face_base[-1])
        faces = [f.detach().to(device=device, dtype=torch.int32).reshape(-1, 3) for f in faces_list]
        self.faces = (faces[0] if len(faces) == 1 else torch.cat(faces, dim=0)).contiguous()
        # CSR over packed vertices: ascending list of LOCAL entries (face*3 + corner) per vertex
        if total_f == 0:
            self.vert_face_ptr = torch.zeros(total_v + 1, dtype=torch.int32, device=device)
            self.vert_face_idx = torch.zeros(0, dtype=torch.int32, device=device)
            return
        vbase_per_face = torch.from_numpy(np.repeat(self.vert_base[:-1], nf).astype(np.int64)).to(device, non_blocking=True)
        ebase_per_face = torch.from_numpy(np.repeat(3 * self.face_base[:-1], nf).astype(np.int64)).to(device, non_blocking=True)
        glob_vert = (self.faces.to(torch.int64) + vbase_per_face[:, None]).reshape(-1)       # entry -> packed vertex
        entry = torch.arange(3 * total_f, device=device, dtype=torch.int64)
        local_entry = entry - ebase_per_face.repeat_interleave(3)
        sorted_vert, order = torch.sort(glob_vert, stable=True)
        self.vert_face_idx = local_entry[order].to(torch.int32).contiguous()
        self.vert_face_ptr = torch.searchsorted(
            sorted_vert, torch.arange(total_v + 1, device=device, dtype=torch.int64)).to(torch.int32).contiguous()


_PACK_CACHE = {}


def get_mesh_pack(verts_list, faces_list):
    """verts_list[i]: [..., p, 3] tensor, faces_list[i]: [f, 3] integer tensor (distinct meshes only)."""
    key = tuple((f.data_ptr(), tuple(f.shape), f._version, int(v.shape[-2])) for v, f in zip(verts_list, faces_list))
    hit = _PACK_CACHE.get(key)
    if hit is not None:
        return hit[0]
    pack = MeshPack(verts_list, faces_list, verts_list[0].device)
    if len(_PACK_CACHE) > 64:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = (pack, list(faces_list))      # keep the face tensors alive so data_ptr stays unique
    return pack


def dedup(tensors):
    """-> (distinct tensors in first-seen order, index of each input in that list); identity by storage."""
    seen, distinct, ids = {}, [], []
    for t in tensors:
        k = (t.data_ptr(), tuple(t.shape), t.dtype)
        if k not in seen:
            seen[k] = len(distinct)
            distinct.append(t)
        ids.append(seen[k])
    return distinct, ids


_PASS_TEMPLATES = {}     # signature of a fused pass -> (bytes of the filled DibrPass, workspace bytes)


def _pass_from_template(key, build):
    """A DibrPass whose size / option fields are already filled: ~60 ctypes field stores and the workspace-size call are paid
    once per configuration, every later call copies the struct (one memmove) and sets its pointers."""
    hit = _PASS_TEMPLATES.get(key)
    if hit is None:
        p = build()
        nbytes = _lib.workspace_bytes(p)
        if len(_PASS_TEMPLATES) > 256:
            _PASS_TEMPLATES.clear()
        hit = _PASS_TEMPLATES[key] = (bytes(p), nbytes)
    return _lib.DibrPass.from_buffer_copy(hit[0]), hit[1]


class RenderMeshes(Function):
    """forward(verts_packed [SV,3], vattr_packed [SA,A], cam_rot [I,3,3], cam_pos [I,3], cam_proj [P,4,4], meta)
    -> (out_0 [B,H,W,c0], ..., out_n [B,H,W,cn], improb [B,H,W,1], face_normal [TF,3] or empty).
    The D = A (+ones)(+depth) interpolated channels are split over separately allocated tensors as
    meta['out_split'] says (sum = D), written directly by the kernel: no views to slice, and in backward an
    output nobody differentiated costs nothing (its grad pointer is NULL)."""

    @staticmethod
    def forward(ctx, verts, vattr, cam_rot, cam_pos, cam_proj, meta):
        for n, t in (("vertices", verts), ("camera rotation", cam_rot), ("camera position", cam_pos),
                     ("camera projection", cam_proj)):
            _require_cuda_f32(n, t)
        device = verts.device
        B, H, W = meta["batch"], meta["height"], meta["width"]
        A, flags = meta["attr_dim"], meta["attr_flags"]
        D = A + (1 if flags & FLAG_ONES else 0) + (1 if flags & FLAG_DEPTH else 0)
        TF = meta["total_faces"]
        # (inside Function.forward autograd is off: a contiguous input is used as it is, no detach() alias per call)
        verts_c = verts if verts.is_contiguous() else verts.contiguous()
        vattr_c = (vattr if vattr.is_contiguous() else vattr.contiguous()) if A > 0 else None
        if A > 0:
            _require_cuda_f32("vertex attributes", vattr)
        rot_c = cam_rot if cam_rot.is_contiguous() else cam_rot.contiguous()
        pos_c = cam_pos if cam_pos.is_contiguous() else cam_pos.contiguous()
        proj_c = cam_proj if cam_proj.is_contiguous() else cam_proj.contiguous()
        pose_mode = bool(meta.get("pose_mode"))
        split = meta.get("out_split") or [D]
        assert sum(split) == D and len(split) <= 6, (split, D)
        min_output = meta.get("min_output")
        nmap_spec = meta.get("normal_map")          # (normals group, mask group): the normal map comes out of this call too
        num_K = int(proj_c.shape[0]) if pose_mode else 0
        vs, vas = int(meta.get("verts_stride", 0)), int(meta.get("vert_attr_stride", 0))
        key = (B, H, W, D, meta["knum"], meta["multiplier"], meta["delta"], meta["expand"], TF, meta["num_instances"], A, flags,
               vs, vas, pose_mode, num_K, float(meta.get("znear", 0.0)), float(meta.get("zfar", 0.0)), tuple(split), min_output)

        def build():
            q = _base_pass(B, H, W, D, meta["knum"], meta["multiplier"], meta["delta"], meta["expand"], TF, 0)
            q.num_instances = meta["num_instances"]
            q.vert_attr_dim, q.attr_flags = A, flags
            q.verts_stride, q.vert_attr_stride = vs, vas
            if pose_mode:
                q.num_K = num_K                          # the workspace holds the derived cameras
                q.znear, q.zfar = float(meta["znear"]), float(meta["zfar"])
            q.num_outputs = len(split)
            for g, c in enumerate(split):
                q.out_channels[g] = c
            if min_output is not None:
                q.min_output = int(min_output)
            return q
        with _on_device(device):
            p, nbytes = _pass_from_template(key, build)
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
            p.face_offsets = meta["face_offsets"].data_ptr()
            # (no [F, 3, D] corner-attribute array: the kernels gather from the vertex table through the faces' row ids)
            face_normal = torch.empty(TF, 3, dtype=torch.float32, device=device) if meta["want_normals"] else None
            outs = [torch.empty(B, H, W, c, dtype=torch.float32, device=device) for c in split]
            improb = torch.empty(B, H, W, 1, dtype=torch.float32, device=device)
            imcomp = torch.empty(B, H, W, dtype=torch.float32, device=device)
            imidx = torch.empty(B, H, W, dtype=torch.int32, device=device)
            pack = meta["pack"]
            p.inst_desc = meta["inst_desc"].data_ptr()
            p.verts, p.mesh_faces = verts_c.data_ptr(), pack.faces.data_ptr()
            p.vert_attr = vattr_c.data_ptr() if vattr_c is not None else None
            if pose_mode:      # cam_rot / cam_pos / cam_proj carry R [I,3,3], t [I,3], K [nK,3,3]
                p.pose_R, p.pose_t, p.pose_K = rot_c.data_ptr(), pos_c.data_ptr(), proj_c.data_ptr()
            else:
                p.cam_rot, p.cam_pos, p.cam_proj = rot_c.data_ptr(), pos_c.data_ptr(), proj_c.data_ptr()
            p.face_normal = _lib.ptr(face_normal)
            p.improb, p.imidx, p.imcomp = improb.data_ptr(), imidx.data_ptr(), imcomp.data_ptr()
            for g, o in enumerate(outs):
                p.out[g] = o.data_ptr()
            out_min = None
            if min_output is not None:      # batch-global minimum of one output group (normal maps)
                out_min = torch.empty(1, dtype=torch.int32, device=device)
                p.out_min_ordered = out_min.data_ptr()
                meta["out_min"] = out_min
            lib = _lib.load()
            st = _stream(device)
            _lib.check(lib.dibr_setup_meshes(ctypes.byref(p), st), "dibr_setup_meshes")
            _lib.check(lib.dibr_forward(ctypes.byref(p), st), "dibr_forward")
            nmap = None
            if nmap_spec is not None and out_min is not None:
                # renderer_dibr.py:281-286 over the tiles of the pass some face reaches, the others zero-filled: same values as
                # dibr_normal_map over every pixel (tests/test_gpu_fused_parity.py), a fraction of the traffic
                nmap = torch.empty_like(outs[nmap_spec[0]])
                _lib.check(lib.dibr_normal_map_pass(ctypes.byref(p), ctypes.c_void_p(outs[nmap_spec[0]].data_ptr()),
                                                    ctypes.c_void_p(outs[nmap_spec[1]].data_ptr()), ctypes.c_void_p(nmap.data_ptr()), st),
                           "dibr_normal_map_pass")
        if meta.get("keep_pass"):      # bench_util.time_forward_kernel re-launches dibr_forward on these buffers
            meta["_last_pass"] = (p, [verts_c, vattr_c, rot_c, pos_c, proj_c, face_normal, outs, improb, imcomp, imidx, ws])
        # (outputs go through save_for_backward, never onto ctx directly: an output held by its own node is a reference cycle,
        # and 200 MB of images per call would wait for the cyclic collector)
        if nmap is not None:
            ctx.save_for_backward(verts_c, rot_c, pos_c, proj_c, vattr_c if A > 0 else None, improb, imcomp, imidx, ws,
                                  outs[nmap_spec[0]], outs[nmap_spec[1]])
        else:
            ctx.save_for_backward(verts_c, rot_c, pos_c, proj_c, vattr_c if A > 0 else None, improb, imcomp, imidx, ws)
        ctx.pass_struct = p          # the backward completes this struct (its pointers stay valid: the tensors are saved above / in meta)
        ctx.nmap_spec = nmap_spec if nmap is not None else None
        ctx.meta = meta
        ctx.dims = (D, A, flags)
        ctx.split = list(split)
        ctx.needs = (verts.requires_grad, vattr.requires_grad if A > 0 else False)
        meta["last_imidx"] = imidx
        if face_normal is None:
            face_normal = torch.empty(0, 3, dtype=torch.float32, device=device)
        ctx.mark_non_differentiable(face_normal)
        ctx.set_materialize_grads(False)
        if nmap is not None:
            return (*outs, improb, face_normal, nmap)
        return (*outs, improb, face_normal)

    @staticmethod
    def backward(ctx, *grads):
        g_nmap = None
        saved = ctx.saved_tensors
        if ctx.nmap_spec is not None:
            grads, g_nmap = grads[:-1], grads[-1]
        g_outs, g_prob = grads[:-2], grads[-2]
        verts_c, rot_c, pos_c, proj_c, vattr_c, improb, imcomp, imidx, ws = saved[:9]
        meta = ctx.meta
        D, A, flags = ctx.dims
        need_verts, need_vattr = ctx.needs
        device = verts_c.device
        B, H, W, TF, I = meta["batch"], meta["height"], meta["width"], meta["total_faces"], meta["num_instances"]
        g_outs = [g.contiguous() if g is not None else None for g in g_outs]
        if g_nmap is not None:      # rare (nobody in Self6D++ differentiates the normal map): the torch expression under autograd
            n_src, m_src, (ni, mi) = saved[9], saved[10], ctx.nmap_spec
            with torch.enable_grad():
                n = n_src.detach().requires_grad_(True)
                m = m_src.detach().requires_grad_(True)
                shift = n - n.min()
                y = shift / (torch.norm(shift, dim=-1, keepdim=True) + 1e-5) * m
                gn, gm = torch.autograd.grad(y, [n, m], g_nmap)
            g_outs[ni] = gn if g_outs[ni] is None else g_outs[ni] + gn
            g_outs[mi] = gm if g_outs[mi] is None else g_outs[mi] + gm
        gP = g_prob.contiguous() if g_prob is not None else None
        pack = meta["pack"]
        pose_mode = bool(meta.get("pose_mode"))
        with _on_device(device):
            p = ctx.pass_struct
            p.grad_improb = _lib.ptr(gP)
            for g, go in enumerate(g_outs):
                p.grad_out[g] = go.data_ptr() if go is not None else None
            g_p2d = torch.empty(max(TF, 1), 6, dtype=torch.float32, device=device)
            # nobody reads dL/d(corner attributes) unless the vertex attributes take a gradient: attr_flags bit 2 keeps only the
            # depth column the vertex stage needs ([F, 3] instead of [F, 3, D])
            p.attr_flags = flags if need_vattr else (flags | FLAG_ATTR_GRAD_SCRATCH)
            g_fattr = torch.empty(max(TF, 1), 3, D if need_vattr else 1, dtype=torch.float32, device=device)
            g_rot = torch.empty(I, 3, 3, dtype=torch.float32, device=device)
            g_pos = torch.empty(I, 3, dtype=torch.float32, device=device)
            n_rows = meta["num_inst_verts"]
            g_verts = torch.empty(n_rows, 3, dtype=torch.float32, device=device) if need_verts else None
            g_vattr = torch.empty(n_rows, max(A, 1), dtype=torch.float32, device=device) if need_vattr else None
            p.grad_points2d, p.grad_face_attr = _lib.ptr(g_p2d), _lib.ptr(g_fattr)
            if pose_mode:
                p.grad_pose_R, p.grad_pose_t = _lib.ptr(g_rot), _lib.ptr(g_pos)
            else:
                p.grad_cam_rot, p.grad_cam_pos = _lib.ptr(g_rot), _lib.ptr(g_pos)
            p.grad_verts, p.grad_vert_attr = _lib.ptr(g_verts), _lib.ptr(g_vattr)
            p.vert_face_ptr, p.vert_face_idx = _lib.ptr(pack.vert_face_ptr), _lib.ptr(pack.vert_face_idx)
            lib = _lib.load()
            st = _stream(device)
            _lib.check(lib.dibr_backward_faces(ctypes.byref(p), st), "dibr_backward_faces")
            _lib.check(lib.dibr_backward_meshes(ctypes.byref(p), st), "dibr_backward_meshes")
        gv = ga = None
        if need_verts:      # instances that share a mesh sum into the same packed rows
            gv = torch.zeros_like(verts_c).index_add_(0, meta["inst_vert_rows"], g_verts)
        if need_vattr:
            ga = torch.zeros(meta["num_attr_rows"], A, dtype=torch.float32, device=device)
            ga.index_add_(0, meta["inst_attr_rows"], g_vattr)
        # proj gradients (dL/dK) are not produced: intrinsics are data in Self6D++
        return gv, ga, g_rot, g_pos, None, None


def build_meta(pack, mesh_ids, attr_ids, attr_rows_base, image_ids, height, width, attr_dim, attr_flags,
               proj_ids=None, vert_rows_base=None, out_split=None, knum=DEFAULT_KNUM, multiplier=DEFAULT_MULTIPLIER, delta=DEFAULT_DELTA,
               expand=DEFAULT_EXPAND, want_normals=False, need_rows=False, num_attr_rows=0):
    """Instance table for one render call.

    mesh_ids[i]   : which packed mesh (topology) instance i renders
    vert_rows_base[i]: first row of instance i's vertices in verts_packed (default: the pack's own base)
    attr_ids/attr_rows_base: attribute tensor of instance i starts at row attr_rows_base[attr_ids[i]]
    image_ids[i]  : image the instance lands in (non-decreasing)
    proj_ids[i]   : row of cam_proj (default: 0 for all)
    """
    I = len(mesh_ids)
    device = pack.device
    desc = np.zeros((I, INST_STRIDE), dtype=np.int32)
    out_base = 0
    gv_base = 0
    nimg = int(image_ids[-1]) + 1 if I else 0
    face_off = np.zeros(nimg + 1, dtype=np.int32)
    vert_rows, attr_rows = [], []
    for i in range(I):
        m = mesh_ids[i]
        nv, nf = pack.num_verts[m], pack.num_faces[m]
        vb = pack.vert_base[m] if vert_rows_base is None else vert_rows_base[i]
        desc[i] = [vb, nv, pack.face_base[m], nf, out_base, i,
                   0 if proj_ids is None else proj_ids[i], attr_rows_base[attr_ids[i]], gv_base, image_ids[i],
                   pack.vert_base[m], 0]
        face_off[image_ids[i] + 1] += nf
        if need_rows:
            vert_rows.append(np.arange(vb, vb + nv))
            attr_rows.append(np.arange(attr_rows_base[attr_ids[i]], attr_rows_base[attr_ids[i]] + nv))
        out_base += nf
        gv_base += nv
    face_off = np.cumsum(face_off).astype(np.int32)
    # image-local face numbering: out_face_base is global; forward subtracts face_offsets[image]
    meta = dict(batch=nimg, height=int(height), width=int(width), attr_dim=int(attr_dim), attr_flags=int(attr_flags),
                total_faces=int(out_base), num_instances=I, num_inst_verts=int(gv_base), pack=pack,
                knum=int(knum), multiplier=int(multiplier), delta=int(delta), expand=float(expand),
                want_normals=bool(want_normals), num_attr_rows=int(num_attr_rows), out_split=out_split,
                inst_desc=torch.from_numpy(desc).to(device, non_blocking=True),
                face_offsets=torch.from_numpy(face_off).to(device, non_blocking=True),
                face_offsets_host=face_off)
    if need_rows:
        meta["inst_vert_rows"] = torch.from_numpy(np.concatenate(vert_rows)).to(device)
        meta["inst_attr_rows"] = torch.from_numpy(np.concatenate(attr_rows)).to(device)
    return meta


class NormalMap(Function):
    """(n - min) / (||n - min|| + 1e-5) * mask with the batch-global min the forward kernel accumulated
    (renderer_dibr.py:284-285).  Forward: one fused kernel (dibr_normal_map).  Backward (rare: nobody in
    Self6D++ differentiates the rendered normal map) re-evaluates the torch expression under autograd."""

    @staticmethod
    def forward(ctx, normals, mask, out_min):
        n_c, m_c = normals.contiguous(), mask.contiguous()
        out = torch.empty_like(n_c)
        with torch.cuda.device(n_c.device):
            _lib.check(_lib.load().dibr_normal_map(n_c.data_ptr(), m_c.data_ptr(), out_min.data_ptr(), out.data_ptr(),
                                                   n_c.numel() // 3, _stream(n_c.device)), "dibr_normal_map")
        ctx.save_for_backward(n_c, m_c)
        return out

    @staticmethod
    def backward(ctx, g):
        n_c, m_c = ctx.saved_tensors
        with torch.enable_grad():
            n = n_c.detach().requires_grad_(True)
            m = m_c.detach().requires_grad_(True)
            shift = n - n.min()
            y = shift / (torch.norm(shift, dim=-1, keepdim=True) + 1e-5) * m
            gn, gm = torch.autograd.grad(y, [n, m], g)
        return gn, gm, None


class _NoGradCtx(object):
    """stands in for the autograd context when nothing can ask for a gradient (torch.no_grad(), or no input requires one):
    the forward runs as a plain call, without an autograd node (~20 us of host time per render)"""

    def save_for_backward(self, *tensors):
        pass

    def mark_non_differentiable(self, *tensors):
        pass

    def set_materialize_grads(self, value):
        pass


def render_meshes(verts_packed, vattr_packed, cam_rot, cam_pos, cam_proj, meta):
    if not torch.is_grad_enabled() or not (verts_packed.requires_grad or vattr_packed.requires_grad or cam_rot.requires_grad
                                           or cam_pos.requires_grad or cam_proj.requires_grad):
        return RenderMeshes.forward(_NoGradCtx(), verts_packed, vattr_packed, cam_rot, cam_pos, cam_proj, meta)
    return RenderMeshes.apply(verts_packed, vattr_packed, cam_rot, cam_pos, cam_proj, meta)
