"""Helper entry points of lib/dr_utils/dr_utils.py (load_objs, render_dib_*): host logic on CPU, renders on the GPU
against Renderer_dibr.render_batch / render_scene, which the parity tests pin to the oracle."""
import os

import numpy as np
import pytest
import torch

from self6dpp_b200 import dr_utils as U
from self6dpp_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_models.npz")


def _write_obj(path, verts, colors, faces, uvs=None):
    with open(path, "w") as f:
        for v, c in zip(verts, colors):
            f.write("v %r %r %r %r %r %r\n" % (*map(float, v), *map(float, c)))
        if uvs is not None:
            for u in uvs:
                f.write("vt %r %r\n" % (float(u[0]), float(u[1])))
        for t in faces:
            if uvs is not None:
                f.write("f %d/%d %d/%d %d/%d\n" % (t[0] + 1, t[0] + 1, t[1] + 1, t[1] + 1, t[2] + 1, t[2] + 1))
            else:
                f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def test_load_objs_layout_and_centring(tmp_path):
    """dr_utils.py:17-72: leading 1-axis on every tensor, one global min/max middle subtracted, int32 faces"""
    m = synth.icosphere(1, radius=0.05, noise_sigma=0.002, seed=4)
    verts = m["vertices"] + np.array([0.3, -0.1, 0.2], np.float32)
    p = str(tmp_path / "a.obj")
    _write_obj(p, verts, m["colors"], m["faces"])
    for centring in (True, False):
        (got,) = U.load_objs([p], centring=centring, device="cpu")
        want = verts - (verts.max() + verts.min()) / 2.0 if centring else verts
        assert got["vertices"].shape == (1, len(verts), 3) and got["colors"].shape == (1, len(verts), 3)
        assert got["faces"].shape == (1, len(m["faces"]), 3) and got["faces"].dtype == torch.int32
        np.testing.assert_allclose(got["vertices"][0].numpy(), want, rtol=0, atol=1e-7)
        np.testing.assert_array_equal(got["faces"][0].numpy(), m["faces"])
        np.testing.assert_allclose(got["colors"][0].numpy(), m["colors"], rtol=0, atol=1e-7)
    with pytest.raises(AssertionError):
        U.load_objs([str(tmp_path / "a.ply")], device="cpu")
    with pytest.raises(AssertionError):
        U.load_objs([p], texture_paths=[], device="cpu")


def test_load_objs_texture_formats(tmp_path):
    cv2 = pytest.importorskip("cv2")
    m = synth.icosphere(1, radius=0.05, seed=1)
    uv = np.random.default_rng(0).random((len(m["vertices"]), 2)).astype(np.float32)
    p = str(tmp_path / "t.obj")
    _write_obj(p, m["vertices"], m["colors"], m["faces"], uvs=uv)
    img = np.random.default_rng(1).integers(0, 256, size=(6, 8, 3), dtype=np.uint8)
    tp = str(tmp_path / "t.png")
    cv2.imwrite(tp, img)
    (a,) = U.load_objs([p], [tp], tex_resize=False, tex_fmt="CHW", device="cpu")
    assert a["texture"].shape == (1, 3, 6, 8) and a["face_uvs"].shape == (1, len(uv), 2)
    assert a["face_uv_ids"].shape == (1, len(m["faces"]), 3)
    np.testing.assert_allclose(a["texture"][0].numpy(), img[:, :, ::-1].transpose(2, 0, 1) / 255.0, atol=1e-7)
    (b,) = U.load_objs([p], [tp], tex_resize=False, tex_fmt="HWC", tex_vflip=True, device="cpu")
    np.testing.assert_allclose(b["texture"][0].numpy(), img[::-1, :, ::-1] / 255.0, atol=1e-7)
    (c,) = U.load_objs([p], [tp], height=12, width=16, tex_resize=True, device="cpu")
    assert c["texture"].shape == (1, 3, 12, 16)


def test_camera_parameters_accept_lists_like_the_reference():
    """base.py:131-140: Rs / ts / Ks may each be a list of per-sample tensors (dr_utils.py:95-97 builds such a list of Ks)"""
    from self6dpp_b200.renderer.cameras import camera_params_from_RT_K
    cpu = torch.device("cpu")
    g = torch.Generator().manual_seed(0)
    Rs = torch.linalg.qr(torch.randn(3, 3, 3, generator=g))[0]
    ts = torch.randn(3, 3, generator=g) + torch.tensor([0.0, 0.0, 1.0])
    Ks = torch.tensor([[572.4, 0.0, 325.3], [0.0, 573.6, 242.0], [0.0, 0.0, 1.0]]).repeat(3, 1, 1) + torch.rand(3, 3, 3, generator=g)
    a = camera_params_from_RT_K(Rs, ts, Ks, 48, 64, device=cpu)
    b = camera_params_from_RT_K(list(Rs), list(ts), list(Ks), 48, 64, device=cpu)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert a[2].shape == (3, 4, 4)
    one = camera_params_from_RT_K(Rs, ts, Ks[0], 48, 64, device=cpu)
    assert one[2].shape == (4, 4) and torch.equal(one[2], a[2][0])


def _scene(dev, n_obj=3, res=64, seed=2):
    meshes = [synth.icosphere(2, radius=0.05, noise_sigma=0.003, seed=0), synth.ellipsoid(9, 12, seed=1)]
    ids = [0, 1, 0][:n_obj]
    batch = synth.roi_batch([meshes[i] for i in ids], len(ids), res=res, seed=seed, fill=(0.5, 0.7))
    models = [{"vertices": torch.tensor(m["vertices"], device=dev)[None], "colors": torch.tensor(m["colors"], device=dev)[None],
               "faces": torch.tensor(m["faces"], device=dev, dtype=torch.int32)[None]} for m in meshes]
    flat = [{"vertices": m["vertices"][0], "colors": m["colors"][0], "faces": m["faces"][0]} for m in models]
    return models, flat, ids, batch


@pytest.mark.gpu
def test_render_dib_vc_batch_equals_render_batch_and_the_two_pass_depth():
    from self6dpp_b200 import DIBRenderer, Renderer_dibr
    dev = torch.device("cuda:0")
    H = W = 64
    models, flat, ids, batch = _scene(dev)
    Rs = torch.tensor(batch["Rs"], device=dev)
    ts = torch.tensor(batch["ts"], device=dev)
    Ks = torch.tensor(batch["Ks"], device=dev)
    ren = DIBRenderer(H, W, mode="VertexColorBatch")
    color, prob, mask, depth = U.render_dib_vc_batch(ren, Rs, ts, Ks, ids, models, rot_type="mat", H=H, W=W, with_depth=True)
    assert color.shape == (3, H, W, 3) and prob.shape == (3, H, W, 1) and mask.shape == (3, H, W, 1) and depth.shape == (3, H, W)
    want = Renderer_dibr(H, W, "VertexColorBatch").render_batch(Rs, ts, [flat[i] for i in ids], Ks=Ks, width=W, height=H,
                                                               znear=0.01, zfar=100.0, mode=["color", "depth", "prob", "mask"])
    assert torch.equal(mask.squeeze(-1) > 0.5, want["mask"] > 0.5)
    # render_batch derives the camera inside the set-up kernel, the helper goes through set_camera_parameters_from_RT_K like
    # the reference: projected corners differ in the last fp32 bits, thin triangles amplify that in their weights
    tol = dict(rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(color, want["color"], **tol)
    torch.testing.assert_close(prob.squeeze(-1), want["prob"], **tol)
    torch.testing.assert_close(depth, want["depth"], **tol)
    # the reference's way: a second rasterisation with the camera-space vertices as colours (dr_utils.py:104-118)
    xyzs = U._view_depth_xyz(Rs, ts, "mat", models, ids)
    pts = [[models[i]["vertices"], models[i]["faces"][0].long()] for i in ids]
    ren_xyz, _, _, _ = ren.forward(points=pts, colors=xyzs)
    torch.testing.assert_close(depth, ren_xyz[..., 2], rtol=1e-5, atol=1e-6)      # same camera route: tight
    # without depth, a list of per-sample Ks of length 1 is broadcast (dr_utils.py:95-96) and quaternions are accepted
    c2, p2, m2, d2 = U.render_dib_vc_batch(ren, Rs, ts, [Ks[0]], ids, models, rot_type="mat", H=H, W=W)
    assert d2 is None and c2.shape == color.shape
    with pytest.raises(AssertionError):
        U.render_dib_vc_batch(DIBRenderer(H, W, mode="VertexColorMulti"), Rs, ts, Ks, ids, models, rot_type="mat", H=H, W=W)


@pytest.mark.gpu
def test_render_dib_vc_multi_equals_render_scene_and_backpropagates():
    from self6dpp_b200 import DIBRenderer, Renderer_dibr
    dev = torch.device("cuda:0")
    H = W = 64
    models, flat, ids, batch = _scene(dev)
    Rs = torch.tensor(batch["Rs"], device=dev, requires_grad=True)
    ts = torch.tensor(batch["ts"], device=dev, requires_grad=True)
    K = torch.tensor(batch["Ks"][0], device=dev)
    ren = DIBRenderer(H, W, mode="VertexColorMulti")
    im, prob, mask = U.render_dib_vc_multi(ren, Rs, ts, K, ids, models, rot_type="mat", H=H, W=W)
    assert im.shape == (1, H, W, 3) and prob.shape == (1, H, W, 1) and mask.shape == (1, H, W, 1)
    want = Renderer_dibr(H, W, "VertexColorMulti").render_scene(Rs.detach(), ts.detach(), [flat[i] for i in ids], K=K, width=W,
                                                               height=H, znear=0.01, zfar=100.0)
    torch.testing.assert_close(im[0], want["color"], rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(prob[0, ..., 0], want["prob"], rtol=1e-3, atol=2e-4)
    (im.sum() + prob.sum()).backward()
    assert Rs.grad is not None and torch.isfinite(Rs.grad).all() and float(ts.grad.abs().sum()) > 0
