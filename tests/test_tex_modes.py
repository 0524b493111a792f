"""Texture / SH / Phong render modes (self6dpp_b200/renderer/tex.py) against golden vectors produced by the REFERENCE's
own TexRender / TexRenderBatch / TexRenderMulti / SHRender / PhongRender (tests/golden/make_golden.py --tex).

  * not gpu: the host logic (vertex shader, per-face features, padding of the ragged batch, fragment shaders, painter's
    order composite) with the CPU oracle standing in for the rasterizer operator -- must match to fp32 round-off;
  * gpu: the same through the B200 operator.
"""
import numpy as np
import pytest
import torch

from tests import helpers as Hh


class _OracleRasterizer(torch.autograd.Function):
    """oracle-backed stand-in for self6dpp_b200.rasterizer.linear_rasterizer (CPU tests only)"""

    @staticmethod
    def forward(ctx, width, height, p3, p2, nz, attr):
        from oracle import dibr_oracle as O
        fw = O.rasterize(width, height, p3.detach(), p2.detach(), nz.detach(), attr.detach())
        ctx.fw = fw
        return fw["im"], fw["improb"]

    @staticmethod
    def backward(ctx, g_im, g_prob):
        from oracle import dibr_oracle as O
        dp2, dc = O.rasterize_backward(ctx.fw, g_im.contiguous(), g_prob.contiguous())
        return None, None, None, dp2, None, dc


def _oracle_linear_rasterizer(width, height, p3, p2, nz, attr, *a, **k):
    return _OracleRasterizer.apply(width, height, p3, p2, nz, attr)


def _load():
    return Hh.load_golden("ref_tex.npz")[0]


def _cams(d, n, dev):
    from self6dpp_b200.renderer.cameras import camera_params_from_RT_K
    return camera_params_from_RT_K(torch.tensor(d["Rs"][:n], device=dev), torch.tensor(d["ts"][:n], device=dev),
                                   torch.tensor(d["K"], device=dev), int(d["H"]), int(d["W"]), near=0.01, far=100.0, device=dev)


def _check_modes(dev, tol_im, outliers):
    from self6dpp_b200.renderer import tex as T
    d = _load()
    H, W = int(d["H"]), int(d["W"])
    t = lambda k, **kw: torch.tensor(d[k], device=dev, **kw)
    # ---- ragged batch of three objects (object 1 with per-face uv rows), then the painter's-order scene
    cams = _cams(d, 3, dev)
    verts = [t(f"verts{k}").requires_grad_(True) for k in range(3)]
    faces = [t(f"faces{k}") for k in range(3)]
    uvs = [t(f"uv{k}") for k in range(3)]
    fts = [t(f"ft{k}") for k in range(3)]
    texs = [t(f"tex{k}").requires_grad_(True) for k in range(3)]
    points = [[v[None], f] for v, f in zip(verts, faces)]
    im, prob, normal1, mask = T.TexRenderBatch(H, W)(points, cams, [u[None] for u in uvs], texs, ft_fx3=fts)
    Hh.assert_close("TexRenderBatch im", im, torch.as_tensor(d["TexRenderBatch_im"]), rtol=tol_im, atol_rel=tol_im, outlier_frac=outliers, outlier_tol=2.0)
    Hh.assert_close("TexRenderBatch prob", prob, torch.as_tensor(d["TexRenderBatch_prob"]), rtol=1e-4, atol_rel=1e-5, outlier_frac=outliers, outlier_tol=2.0)
    Hh.assert_close("TexRenderBatch mask", mask, torch.as_tensor(d["TexRenderBatch_mask"]), rtol=1e-5, atol_rel=1e-5, outlier_frac=outliers, outlier_tol=2.0)
    for k in range(3):
        Hh.assert_close(f"normal1[{k}]", normal1[k], torch.as_tensor(d[f"TexRenderBatch_normal1_{k}"]), rtol=1e-4, atol_rel=1e-5)
    ((im * t("TexRenderBatch_gi")).sum() + (prob * t("TexRenderBatch_gp")).sum()).backward()
    for k in range(3):
        Hh.assert_close(f"dL/dtexture[{k}]", texs[k].grad, torch.as_tensor(d[f"TexRenderBatch_gtex{k}"]), rtol=1e-3, atol_rel=1e-3)
        Hh.assert_close(f"dL/dverts[{k}]", verts[k].grad, torch.as_tensor(d[f"TexRenderBatch_gv{k}"]), rtol=2e-2, atol_rel=2e-2)
    with torch.no_grad():
        im, prob, _, mask = T.TexRenderMulti(H, W)(points, cams, [u[None] for u in uvs], texs, ts=t("ts"), ft_fx3=fts)
    Hh.assert_close("TexRenderMulti im", im, torch.as_tensor(d["TexRenderMulti_im"]), rtol=tol_im, atol_rel=tol_im, outlier_frac=outliers, outlier_tol=2.0)
    Hh.assert_close("TexRenderMulti prob", prob, torch.as_tensor(d["TexRenderMulti_prob"]), rtol=1e-4, atol_rel=1e-5, outlier_frac=outliers, outlier_tol=2.0)
    Hh.assert_close("TexRenderMulti mask", mask, torch.as_tensor(d["TexRenderMulti_mask"]), rtol=1e-5, atol_rel=1e-5, outlier_frac=outliers, outlier_tol=2.0)
    # ---- one topology, two vertex sets: bilinear texture, SH (flat / smooth normals), Phong
    cams2 = _cams(d, 2, dev)
    f1, uvb = t("f1"), t("uvb")
    runs = [("TexRender", T.TexRender(H, W, filtering="bilinear"), ()),
            ("SHflat", T.SHRender(H, W), (t("light9"),)),
            ("SHsmooth", T.SHRender(H, W), (t("light9"),)),
            ("Phong", T.PhongRender(H, W), (t("lightdir"), t("material"), t("shin")))]
    runs[2][1].set_smooth(t("pf")[None])
    for tag, ren, extra in runs:
        vb = t("vb").requires_grad_(True)
        texb = t("texb").requires_grad_(True)
        im, prob, normal1, mask = ren([vb, f1], cams2, uvb, texb, *extra)
        Hh.assert_close(tag + " im", im, d[tag + "_im"], rtol=tol_im, atol_rel=tol_im, outlier_frac=outliers, outlier_tol=2.0)
        Hh.assert_close(tag + " prob", prob, d[tag + "_prob"], rtol=1e-4, atol_rel=1e-5, outlier_frac=outliers, outlier_tol=2.0)
        Hh.assert_close(tag + " normal1", normal1, d[tag + "_normal1"], rtol=1e-4, atol_rel=1e-5)
        ((im * t(tag + "_gi")).sum() + (prob * t(tag + "_gp")).sum()).backward()
        Hh.assert_close(tag + " dL/dtexture", texb.grad, d[tag + "_gtex"], rtol=1e-3, atol_rel=1e-3)
        Hh.assert_close(tag + " dL/dverts", vb.grad, d[tag + "_gv"], rtol=2e-2, atol_rel=2e-2)


def test_tex_host_logic_on_cpu_with_oracle_rasterizer(monkeypatch):
    from self6dpp_b200.renderer import tex as T
    monkeypatch.setattr(T, "linear_rasterizer", _oracle_linear_rasterizer)
    _check_modes(torch.device("cpu"), tol_im=1e-4, outliers=1e-3)   # bilinear lookups amplify uv round-off by the texture gradient


def test_renderer_accepts_texture_modes():
    from self6dpp_b200 import DIBRenderer
    for mode in ("Lambertian", "Texture", "TextureBatch", "TextureMulti", "SphericalHarmonics", "Phong"):
        assert DIBRenderer(32, 32, mode=mode).mode == mode


@pytest.mark.gpu
def test_tex_modes_on_gpu_match_reference_golden():
    _check_modes(torch.device("cuda:0"), tol_im=3e-4, outliers=3e-3)


@pytest.mark.gpu
def test_render_batch_tex_entry_point():
    """Renderer_dibr.render_batch_tex / render_scene_tex (renderer_dibr.py:159-235,309-412): keys, shapes and that the
    colour equals TexRenderBatch's on the same inputs."""
    from self6dpp_b200 import Renderer_dibr
    from self6dpp_b200.renderer import tex as T
    dev = torch.device("cuda:0")
    d = _load()
    H, W = int(d["H"]), int(d["W"])
    t = lambda k: torch.tensor(d[k], device=dev)
    models = []
    for k in range(3):
        m = {"vertices": t(f"verts{k}"), "faces": t(f"faces{k}"), "texture": t(f"tex{k}")[0]}
        if k == 1:
            m["face_uvs"], m["face_uv_ids"] = t(f"uv{k}"), t(f"ft{k}")
        models.append(m)
    face_models = [dict(m) for m in models]
    for k in (0, 2):                                                   # face-uv form of the vertex-uv objects
        face_models[k]["face_uvs"], face_models[k]["face_uv_ids"] = t(f"uv{k}"), t(f"faces{k}")
    ren = Renderer_dibr(H, W, mode="TextureBatch")
    out = ren.render_batch_tex(t("Rs"), t("ts"), face_models, Ks=t("K"), width=W, height=H, uv_type="face", mode=["color", "depth", "xyz"])
    assert out["color"].shape == (3, H, W, 3) and out["prob"].shape == (3, H, W) and out["mask"].shape == (3, H, W)
    assert out["depth"].shape == (3, H, W) and out["xyz"].shape == (3, H, W, 3)
    Hh.assert_close("render_batch_tex color", out["color"], torch.as_tensor(d["TexRenderBatch_im"]), rtol=1e-4, atol_rel=1e-4, outlier_frac=3e-3, outlier_tol=2.0)
    covered = out["mask"] > 0.5
    assert float((out["depth"][covered] - t("ts")[:, 2].view(3, 1, 1).expand(3, H, W)[covered]).abs().max()) < 0.1
    sc = ren.render_scene_tex(t("Rs"), t("ts"), face_models, K=t("K"), width=W, height=H, uv_type="face")
    assert sc["color"].shape == (H, W, 3) and sc["depth"].shape == (H, W)
    Hh.assert_close("render_scene_tex color", sc["color"], torch.as_tensor(d["TexRenderMulti_im"])[0], rtol=1e-4, atol_rel=1e-4, outlier_frac=3e-3, outlier_tol=2.0)


@pytest.mark.gpu
def test_render_batch_tex_vertex_uv_fast_path_equals_the_module():
    """per-vertex uvs take the fused rasterisation (uv as a 2-channel vertex attribute, depth in the same pass); the
    result must be what TexRenderBatch gives through the operator seam, gradients to the pose and the texture included"""
    from self6dpp_b200 import Renderer_dibr
    from self6dpp_b200.renderer import tex as T
    from self6dpp_b200.renderer.cameras import camera_params_from_RT_K
    dev = torch.device("cuda:0")
    d = _load()
    H, W = int(d["H"]), int(d["W"])
    t = lambda k: torch.tensor(d[k], device=dev)
    pick = [0, 2, 0]
    models = [{"vertices": t(f"verts{k}"), "faces": t(f"faces{k}"), "vertex_uvs": t(f"uv{k}"),
               "texture": t("tex0")[0].clone().requires_grad_(True)} for k in pick]
    Rs = t("Rs").clone().requires_grad_(True)
    ts = t("ts").clone().requires_grad_(True)
    ren = Renderer_dibr(H, W, mode="TextureBatch")
    out = ren.render_batch_tex(Rs, ts, models, Ks=t("K"), width=W, height=H, uv_type="vertex", mode=["color", "depth"])
    assert getattr(ren, "last_meta", None) is not None, "fast path not taken"
    g = torch.Generator().manual_seed(3)
    gi, gp = torch.randn(out["color"].shape, generator=g).to(dev), torch.randn(out["prob"].shape, generator=g).to(dev)
    ((out["color"] * gi).sum() + (out["prob"] * gp).sum()).backward()
    # the module path on the same inputs
    Rs2, ts2 = t("Rs").clone().requires_grad_(True), t("ts").clone().requires_grad_(True)
    tex2 = [m["texture"].detach().clone().requires_grad_(True) for m in models]
    cams = camera_params_from_RT_K(Rs2, ts2, t("K"), H, W, near=0.01, far=100.0, device=dev)
    im, prob, _, mask = T.TexRenderBatch(H, W)([[m["vertices"][None], m["faces"].long()] for m in models], cams,
                                               [m["vertex_uvs"][None] for m in models], [x[None] for x in tex2])
    ((im * gi).sum() + (prob.squeeze(-1) * gp).sum()).backward()
    Hh.assert_close("color", out["color"], im, rtol=3e-4, atol_rel=3e-4, outlier_frac=3e-3, outlier_tol=2.0)
    Hh.assert_close("prob", out["prob"], prob.squeeze(-1), rtol=1e-4, atol_rel=1e-4, outlier_frac=3e-3, outlier_tol=2.0)
    Hh.assert_close("mask", out["mask"], mask.squeeze(-1), rtol=1e-5, atol_rel=1e-5, outlier_frac=3e-3, outlier_tol=2.0)
    covered = out["mask"] > 0.5
    assert float((out["depth"][covered] - t("ts")[:, 2].view(3, 1, 1).expand(3, H, W)[covered]).abs().max()) < 0.1
    for a, b in zip(models, tex2):
        Hh.assert_close("dL/dtexture", a["texture"].grad, b.grad, rtol=1e-3, atol_rel=1e-3)
    Hh.assert_close("dL/dR", Rs.grad, Rs2.grad, rtol=2e-2, atol_rel=2e-2)
    Hh.assert_close("dL/dt", ts.grad, ts2.grad, rtol=2e-2, atol_rel=2e-2)
