"""``RenderSession`` -- the render-and-compare step of Self6D++'s ``compute_self_loss_pose``
(/root/reference/core/self6dpp/engine/self_engine_utils.py:426-447) behind C-ABI calls that take HOST buffers
(include/dibr_b200.h): ``forward()`` = ``dibr_render_forward`` (both rasterisations), then -- once the caller has
turned the rendered images into upstream gradients (self_engine_utils.py:541-558, 736-813) -- ``backward()`` =
``dibr_render_backward``; ``step()`` = both in one call (``dibr_render_step``) for gradients known beforehand.

``Renderer_dibr.render_batch`` is the drop-in for the reference's Python API; it pays ~0.4 ms of
Python / torch.autograd bookkeeping per call, which is more than the kernels need.  A session
pre-allocates every device buffer once (resident meshes, workspaces, outputs, gradient buffers, a
pinned staging block) and per step only (1) writes poses / intrinsics / the instance table into
pinned memory with numpy, (2) makes one ctypes call.  The call enqueues: one H2D copy, student
rasterisation (colour + normal + mask + depth + soft mask in ONE pass), teacher rasterisation
(normal map), the deterministic backward to dL/dR, dL/dt, one D2H copy.  Inside the call the student chain
runs on a side stream next to the teacher rasterisation; the caller's stream waits for both before the call
returns, so for the caller everything is ordered in its own stream.

Outputs are persistent device tensors (overwritten by the next step); the pose gradients arrive in
pinned host memory after ``session.synchronize()``.
"""
import ctypes

import numpy as np
import torch

from . import _lib, fused
from .rasterizer import _on_device, _stream
from .renderer_dibr import _ModelRegistry

_MODE_ATTR = (("color", "colors"), ("norm", "normals"), ("xyz", "vertices"))


class _PassBuffers(object):
    def __init__(self, reg, mode, batch, height, width, max_faces, device, znear, zfar, knum=None, raw_normals=False, face_attr_grad=True):
        self.names = [a for k, a in _MODE_ATTR if k in mode]
        self.keys = [k for k, a in _MODE_ATTR if k in mode] + ["ones"] + (["depth"] if "depth" in mode else [])
        self.split = [3] * len(self.names) + [1] + ([1] if "depth" in mode else [])
        self.flags = fused.FLAG_ONES | (fused.FLAG_DEPTH if "depth" in mode else 0) | (0 if face_attr_grad else fused.FLAG_ATTR_GRAD_SCRATCH)
        self.A = 3 * len(self.names)
        self.D = sum(self.split)
        self.vattr = reg.attr_matrix(self.names) if self.names else torch.zeros(0, 4, dtype=torch.float32, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.outs = [torch.empty(batch, height, width, c, **f32) for c in self.split]
        self.out = dict(zip(self.keys, self.outs))
        self.improb = torch.empty(batch, height, width, 1, **f32)
        self.imcomp = torch.empty(batch, height, width, **f32)
        self.imidx = torch.empty(batch, height, width, dtype=torch.int32, device=device)
        self.out_min = torch.empty(1, dtype=torch.int32, device=device)
        # the normal map is computed in place over the pass's "norm" output group (dibr_normal_map_pass touches only the
        # tiles some face reaches; the forward's zero fill of the others already is the map)
        # (raw_normals=True keeps the interpolated normals in out["norm"] and writes the map to its own tensor.)
        self.normal_map = None
        if "norm" in mode:
            self.normal_map = torch.empty(batch, height, width, 3, **f32) if raw_normals else self.out["norm"]
        p = _lib.DibrPass()
        p.batch, p.height, p.width = batch, height, width
        p.num_attr, p.knum = self.D, fused.DEFAULT_KNUM if knum is None else int(knum)
        p.multiplier, p.delta, p.expand = fused.DEFAULT_MULTIPLIER, fused.DEFAULT_DELTA, fused.DEFAULT_EXPAND
        p.total_faces, p.faces_per_image = max_faces, 0
        p.num_instances, p.num_K = batch, batch
        p.znear, p.zfar = znear, zfar
        nbytes = _lib.workspace_bytes(p)                      # worst case: every sample renders the largest mesh
        self.ws = torch.empty(nbytes + 4096, dtype=torch.uint8, device=device)
        p.workspace, p.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        p.verts, p.mesh_faces = reg.verts4.data_ptr(), reg.pack.faces.data_ptr()          # rows padded to 16 B
        p.vert_attr = self.vattr.data_ptr() if self.A else None
        p.vert_attr_dim, p.attr_flags = self.A, self.flags
        p.verts_stride, p.vert_attr_stride = 4, int(self.vattr.shape[1]) if self.A else 0
        p.improb, p.imidx, p.imcomp = self.improb.data_ptr(), self.imidx.data_ptr(), self.imcomp.data_ptr()
        p.num_outputs = len(self.split)
        for g, (c, o) in enumerate(zip(self.split, self.outs)):
            p.out_channels[g] = c
            p.out[g] = o.data_ptr()
        if "norm" in mode:
            p.min_output = self.keys.index("norm")
            p.out_min_ordered = self.out_min.data_ptr()
        else:
            p.min_output = -1
        p.vert_face_ptr, p.vert_face_idx = reg.pack.vert_face_ptr.data_ptr(), reg.pack.vert_face_idx.data_ptr()
        self.p = p


class RenderSession(object):
    def __init__(self, models, batch, height, width, student_mode=("color", "depth", "mask", "norm", "prob"),
                 teacher_mode=("norm",), device="cuda:0", znear=0.01, zfar=100.0, teacher_soft_mask=True, raw_normals=False,
                 cuda_graphs=True, face_attr_grad=False):
        """``teacher_soft_mask=False`` skips the soft-silhouette phase of the teacher rasterisation (K = 0).  The reference
        always computes it (kaolin's forward does) and then drops it when ``mode`` has no "color"
        (renderer_dibr.py:273-286), so nothing a caller can observe changes; the default keeps the reference's work."""
        self.device = torch.device(device)
        self.B, self.H, self.W = int(batch), int(height), int(width)
        self.lib = _lib.load()
        reg = _ModelRegistry()
        slots = reg.slots(models)
        if slots is None:
            raise RuntimeError("RenderSession needs resident models that do not require grad")
        self.reg = reg
        self.model_slot = {id(m): int(s) for m, s in zip(models, slots)}
        B = self.B
        max_faces = B * int(reg.table[:, 3].max())
        with torch.cuda.device(self.device):
            self.student = _PassBuffers(reg, student_mode, B, self.H, self.W, max_faces, self.device, znear, zfar, raw_normals=raw_normals, face_attr_grad=face_attr_grad)
            self.teacher = _PassBuffers(reg, teacher_mode, B, self.H, self.W, max_faces, self.device, znear, zfar, raw_normals=raw_normals,
                                        knum=None if teacher_soft_mask else 0) if teacher_mode else None
            f32 = dict(dtype=torch.float32, device=self.device)
            self.g_p2d = torch.empty(max_faces, 6, **f32)
            # dL/d(corner attributes): between the two backward kernels only.  The resident models take no gradient, so the
            # vertex stage reads its depth column alone -- kept as [F, 3] unless face_attr_grad=True asks for the full array
            self.g_fattr = torch.empty(max_faces, 3, self.student.D, **f32) if face_attr_grad else torch.empty(max_faces, 3, **f32)
            self.g_pose_R = torch.empty(B, 9, **f32)
            self.g_pose_t = torch.empty(B, 3, **f32)
            self._g_pose_all = torch.zeros(B + 1, 12, **f32)       # rows 0..B-1: dL/dR | dL/dt per sample; row B: their column sums
            self.g_pose_dev = self._g_pose_all[:B]
            self.g_pose_sum = self._g_pose_all[B]                  # written by the backward kernel: what a data-parallel step all-reduces
            self.g_pose_host = torch.empty(B, 12, dtype=torch.float32, pin_memory=True)
            # staging block, 4-byte words: sR[9B] st[3B] K[9B] tR[9B] tt[3B] desc[12B] face_off[B+1]
            self.off = {}
            o = 0
            for name, n in (("sR", 9 * B), ("st", 3 * B), ("K", 9 * B), ("tR", 9 * B), ("tt", 3 * B),
                            ("desc", fused.INST_STRIDE * B), ("foff", B + 1)):
                o = (o + 3) // 4 * 4                             # 16-byte aligned sections
                self.off[name] = (o, n)
                o += n
            self.stage_words = o
            self.stage_host = torch.empty(o, dtype=torch.int32, pin_memory=True)
            self.stage_dev = torch.empty(o, dtype=torch.int32, device=self.device)
        self._h_i32 = self.stage_host.numpy()
        self._h_f32 = self._h_i32.view(np.float32)
        st = _lib.DibrStep()
        base = self.stage_dev.data_ptr()
        for name, pb in (("student", self.student), ("teacher", self.teacher)):
            if pb is None:
                continue
            p = pb.p
            p.pose_R = base + 4 * self.off["sR" if name == "student" else "tR"][0]
            p.pose_t = base + 4 * self.off["st" if name == "student" else "tt"][0]
            p.pose_K = base + 4 * self.off["K"][0]
            p.inst_desc = base + 4 * self.off["desc"][0]
            p.face_offsets = base + 4 * self.off["foff"][0]
        sp = self.student.p
        sp.grad_points2d, sp.grad_face_attr = self.g_p2d.data_ptr(), self.g_fattr.data_ptr()
        sp.grad_pose_R, sp.grad_pose_t = self.g_pose_R.data_ptr(), self.g_pose_t.data_ptr()
        st.student = sp
        if self.teacher is not None:
            st.teacher = self.teacher.p
        st.staging_host, st.staging_device = self.stage_host.data_ptr(), base
        st.staging_bytes = 4 * self.stage_words
        if self.student.normal_map is not None:
            st.student_normal_in = self.student.out["norm"].data_ptr()
            st.student_mask_in = self.student.out["ones"].data_ptr()
            st.student_normal_out = self.student.normal_map.data_ptr()
        if self.teacher is not None and self.teacher.normal_map is not None:
            st.teacher_normal_in = self.teacher.out["norm"].data_ptr()
            st.teacher_mask_in = self.teacher.out["ones"].data_ptr()
            st.teacher_normal_out = self.teacher.normal_map.data_ptr()
        st.host_grad_pose, st.device_grad_pose = self.g_pose_host.data_ptr(), self._g_pose_all.data_ptr()
        st.grad_pose_sum = 1
        # side stream + events for the student / teacher overlap: owned by THIS session (the library keeps none)
        self._overlap = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dibr_overlap_create(ctypes.byref(self._overlap)), "dibr_overlap_create")
        st.overlap = self._overlap
        self.st = st
        self._ar = np.arange(B)
        self._keep = None
        # The launches of a forward / backward call (2 + 1 memsets, kernels, copies on two streams) are captured into a CUDA
        # graph the second time the call is made with the same arguments and replayed from then on: every pointer in them
        # is owned by the session (or, for the backward, is the caller's gradient tensor: another tensor -> another graph).
        self.cuda_graphs = bool(cuda_graphs)
        self._graphs = {}          # key -> [calls seen, CUDAGraph or None, kernels launched per replay]
        self._bwd_prepared = False # the last forward() prepared the backward and no backward has consumed that yet
        self._last_slots, self._last_total, self._last_ids = None, 0, None
        self._views = {name: self._h_f32[o:o + n].reshape(B, -1) for name, (o, n) in self.off.items() if name in ("sR", "st", "K", "tR", "tt")}

    # ------------------------------------------------------------------------------------------
    def _fill(self, name, arr):
        o, n = self.off[name]
        self._h_f32[o:o + n] = np.asarray(arr, dtype=np.float32).reshape(-1)

    def __del__(self):
        try:
            if getattr(self, "_overlap", None):
                self.lib.dibr_overlap_destroy(self._overlap)
                self._overlap = None
        except Exception:
            pass

    def _stage_inputs(self, Rs, ts, Ks, models, teacher_Rs, teacher_ts, upload):
        """poses / intrinsics / instance table -> pinned staging block; returns whether it must be uploaded"""
        B = self.B
        assert len(models) == B
        v = self._views                                        # float32 views of the staging block's sections, shaped like the inputs
        np.copyto(v["sR"], np.asarray(Rs, dtype=np.float32).reshape(B, 9))
        np.copyto(v["st"], np.asarray(ts, dtype=np.float32).reshape(B, 3))
        np.copyto(v["K"], np.asarray(Ks, dtype=np.float32).reshape(-1, 9))          # one K broadcasts over the batch
        if self.teacher is not None:
            np.copyto(v["tR"], np.asarray(teacher_Rs, dtype=np.float32).reshape(B, 9))
            np.copyto(v["tt"], np.asarray(teacher_ts, dtype=np.float32).reshape(B, 3))
        ids = tuple(map(id, models))                           # the composition: which resident model sits in which sample
        if ids != self._last_ids:
            self._last_ids = ids
            slots = np.fromiter((self.model_slot[i] for i in ids), dtype=np.int64, count=B)
            # the instance table depends on WHICH models sit in the batch only: rebuilt when the composition changes
            tab = self.reg.table[slots]
            nf, nv = tab[:, 3], tab[:, 1]
            o, n = self.off["desc"]
            desc = self._h_i32[o:o + n].reshape(B, fused.INST_STRIDE)
            cf = np.cumsum(nf)
            desc[:, 0], desc[:, 1], desc[:, 2], desc[:, 3], desc[:, 4] = tab[:, 0], nv, tab[:, 2], nf, cf - nf
            desc[:, 5] = self._ar
            desc[:, 6] = self._ar
            desc[:, 7] = tab[:, 0]
            desc[:, 8] = np.cumsum(nv) - nv
            desc[:, 9] = self._ar
            desc[:, 10] = tab[:, 0]
            desc[:, 11] = 0
            o, n = self.off["foff"]
            self._h_i32[o] = 0
            self._h_i32[o + 1:o + n] = cf
            self._last_total = int(cf[-1])
            self._last_slots = slots
            upload = True                 # a new instance table must reach the device even if the caller says the inputs are resident
        # (total_faces of both passes stays at the session's capacity: the kernels read the faces in use from
        # face_offsets[B] on the device, so the workspace layout, the grids -- and a captured graph -- survive a new composition)
        return upload

    def _set_grads(self, grad_color, grad_prob, grad_depth):
        """point the student pass at the caller's gradient tensors; returns their addresses (the identity of a captured graph)"""
        sp = self.st.student
        keep, key = [], []
        grads = {"color": grad_color, "depth": grad_depth}
        npix = self.B * self.H * self.W
        for g, key_name in enumerate(self.student.keys):
            t = grads.get(key_name)
            if t is not None:
                if not t.is_cuda or t.dtype != torch.float32 or t.numel() != npix * self.student.split[g]:
                    raise RuntimeError("RenderSession.backward: grad_%s must be a float32 CUDA tensor with %d elements" % (key_name, npix * self.student.split[g]))
                t = t.contiguous()
                keep.append(t)
                sp.grad_out[g] = t.data_ptr()
                key.append(t.data_ptr())
            else:
                sp.grad_out[g] = None
                key.append(0)
        if grad_prob is not None:
            if not grad_prob.is_cuda or grad_prob.dtype != torch.float32 or grad_prob.numel() != npix:
                raise RuntimeError("RenderSession.backward: grad_prob must be a float32 CUDA tensor with %d elements" % npix)
            gp = grad_prob.contiguous()
            keep.append(gp)
            sp.grad_improb = gp.data_ptr()
            key.append(gp.data_ptr())
        else:
            sp.grad_improb = None
            key.append(0)
        self._keep = keep
        return tuple(key)

    def forward(self, Rs, ts, Ks, models, teacher_Rs=None, teacher_ts=None, upload=True):
        """Rs [B,3,3], ts [B,3], Ks [B,3,3] (+ teacher pose): HOST arrays.  Renders the student and the teacher pass and
        returns the dict of persistent output tensors; everything ``backward`` needs stays resident."""
        st = self.st
        upload = self._stage_inputs(Rs, ts, Ks, models, teacher_Rs, teacher_ts, upload)
        st.staging_bytes = 4 * self.stage_words if upload else 0       # upload=False: inputs already resident
        # run_backward bit 1: the forward call also prepares the backward (zeroed gradient slots, work lists from the flags the
        # rasterisation sets) in the shadow of the teacher pass; the first backward after it starts with the face kernel
        st.run_backward = 2
        self._run(("forward", bool(upload)), self.lib.dibr_render_forward, "dibr_render_forward")
        self._bwd_prepared = True
        return self.outputs()

    def _run(self, key, fn, what):
        """call ``fn(step, stream)`` eagerly the first time ``key`` is seen, capture it into a CUDA graph the second time,
        replay the graph afterwards"""
        st = self.st
        with _on_device(self.device):
            if torch.cuda.is_current_stream_capturing():                # the caller is capturing a graph of its own: plain launches
                _lib.check(fn(ctypes.byref(st), _stream(self.device)), what)
                return
            ent = self._graphs.get(key) if self.cuda_graphs else None
            if ent is not None and ent[1] is not None:
                ent[1].replay()
                self.lib.dibr_launch_count_add(ent[2])
                return
            if self.cuda_graphs:
                if ent is None:
                    if len(self._graphs) > 8:                           # callers that hand over fresh gradient tensors every step
                        self._graphs = {k: v for k, v in self._graphs.items() if v[1] is not None and k[0] == "forward"}
                    ent = self._graphs[key] = [0, None, 0]
                ent[0] += 1
                if ent[0] == 2:
                    try:
                        torch.cuda.current_stream(self.device).synchronize()
                        g = torch.cuda.CUDAGraph()
                        before = self.lib.dibr_launch_count(0)
                        with torch.cuda.graph(g):
                            _lib.check(fn(ctypes.byref(st), _stream(self.device)), what)
                        ent[1], ent[2] = g, int(self.lib.dibr_launch_count(0) - before)
                        self.lib.dibr_launch_count_add(-ent[2])         # the capture launched nothing
                        g.replay()
                        self.lib.dibr_launch_count_add(ent[2])
                        return
                    except Exception:
                        ent[1] = None
                        ent[0] = 3                                      # never try again for this key
                        torch.cuda.synchronize(self.device)
            _lib.check(fn(ctypes.byref(st), _stream(self.device)), what)

    def backward(self, grad_color=None, grad_prob=None, grad_depth=None, download=True):
        """grad_*: DEVICE tensors (dL/dcolor [B,H,W,3], dL/dprob [B,H,W], dL/ddepth [B,H,W]) or None, usually computed from
        ``forward``'s images.  Runs the backward of the student pass; ``grad_pose`` (pinned [B,12]: dL/dR then dL/dt) is
        valid after ``synchronize()``, ``g_pose_dev`` holds the same on the device.  May be called again with other gradients."""
        st = self.st
        ptrs = self._set_grads(grad_color, grad_prob, grad_depth)
        st.host_grad_pose = self.g_pose_host.data_ptr() if download else None
        prepared, self._bwd_prepared = self._bwd_prepared, False        # a second backward over the same forward prepares itself
        st.run_backward = 2 if prepared else 0
        self._run(("backward", bool(download), prepared) + ptrs, self.lib.dibr_render_backward, "dibr_render_backward")
        return self.g_pose_dev

    def step(self, Rs, ts, Ks, models, teacher_Rs=None, teacher_ts=None, grad_color=None, grad_prob=None,
             grad_depth=None, backward=True, upload=True, download=True):
        """``forward`` and ``backward`` in ONE call (dibr_render_step), for upstream gradients that do not depend on this
        step's images.  Returns the dict of persistent output tensors; call ``synchronize()`` before reading ``grad_pose``."""
        st = self.st
        upload = self._stage_inputs(Rs, ts, Ks, models, teacher_Rs, teacher_ts, upload)
        self._set_grads(grad_color, grad_prob, grad_depth)
        st.run_backward = 3 if backward else 0                         # bit 0: backward in this call; bit 1: prepared inside the forward half
        self._bwd_prepared = False
        st.staging_bytes = 4 * self.stage_words if upload else 0       # upload=False: inputs already resident
        st.host_grad_pose = self.g_pose_host.data_ptr() if download else None
        with _on_device(self.device):
            stream = _stream(self.device)
            _lib.check(self.lib.dibr_render_step(ctypes.byref(st), stream), "dibr_render_step")
        return self.outputs()

    def outputs(self):
        s = self.student
        ret = {"prob": s.improb.squeeze(-1), "mask": s.out["ones"].squeeze(-1)}
        if "color" in s.out:
            ret["color"] = s.out["color"]
        if "depth" in s.out:
            ret["depth"] = s.out["depth"].squeeze(-1)
        if s.normal_map is not None:
            ret["norm"] = s.normal_map
        if self.teacher is not None and self.teacher.normal_map is not None:
            ret["teacher_norm"] = self.teacher.normal_map
        return ret

    @property
    def grad_pose(self):
        return self.g_pose_host

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()
