"""Generates tests/golden/*.npz by running the REFERENCE's own Python layers
(/root/reference/lib/dr_utils/dib_renderer_x: VCRenderBatch, VCRenderMulti, perspective_projection,
LinearRasterizer incl. its autograd backward) on CPU, with the oracle plugged in at the
``kaolin.graphics.dib_renderer.cuda.rasterizer`` boundary (oracle/dibr_oracle.py).

Run here (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The committed .npz files hold inputs AND outputs so the GPU parity tests need nothing but numpy.

renderer_dibr.py itself cannot be imported (mmcv / detectron2-era deps are absent), so its
render_batch body (renderer_dibr.py:259-306) is restated below on top of the reference's VCRenderBatch;
base.py:131-191 (hard-coded .cuda()) is restated in oracle.camera_params_from_RT_K.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dibr_oracle as O  # noqa: E402
from self6dpp_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def small_meshes():
    return [synth.ellipsoid(9, 12, radii=(0.05, 0.04, 0.03), noise_sigma=0.002, seed=1),
            synth.icosphere(2, radius=0.045, noise_sigma=0.003, seed=2),
            synth.ellipsoid(12, 10, radii=(0.03, 0.05, 0.04), noise_sigma=0.002, seed=3)]


def to_models(meshes):
    return [{k: torch.tensor(v) for k, v in m.items()} for m in meshes]


def transform_pts_Rt_th(pts, R, t):
    """lib/pysixd/misc.py:985-1004"""
    return (R.view(1, 3, 3) @ pts.view(-1, 3, 1) + t.view(1, 3, 1)).squeeze(-1)


def golden_batch(ref, H=64, W=64, B=4, seed=0):
    meshes = small_meshes()
    models = to_models(meshes)
    ids = [0, 1, 2, 1][:B]
    batch = synth.roi_batch([meshes[i] for i in ids], B, res=H, seed=seed, fill=(0.45, 0.7))
    # roi_batch indexes meshes by i % len: we passed the per-sample list so ids are 0..B-1
    Rs = torch.tensor(batch["Rs"], requires_grad=True)
    ts = torch.tensor(batch["ts"], requires_grad=True)
    Ks = torch.tensor(batch["Ks"])
    cams = O.camera_params_from_RT_K(Rs, ts, Ks, H, W, near=0.01, far=100.0)
    for c in cams[:2]:
        c.retain_grad()
    ren = ref["VCRenderBatch"](H, W)
    cur = [models[i] for i in ids]
    points = [[m["vertices"][None], m["faces"].long()] for m in cur]
    # renderer_dibr.py:273-279 colour pass
    color, im_prob, _, im_mask = ren(points, cams, [m["colors"][None] for m in cur])
    # :281-286 normal pass
    _ren_norms, _, _, norm_im_mask = ren(points, cams, [m["normals"][None] for m in cur])
    shift = _ren_norms - _ren_norms.min()
    norm = shift / (torch.norm(shift, dim=-1, keepdim=True) + 1e-5) * norm_im_mask
    # :288-301 depth pass
    xyzs = [transform_pts_Rt_th(m["vertices"], Rs[i], ts[i])[None] for i, m in enumerate(cur)]
    ren_xyzs, _, _, _ = ren(points, cams, xyzs)
    depth = ren_xyzs[:, :, :, 2]
    g = torch.Generator().manual_seed(seed + 1)
    g_color = torch.randn(color.shape, generator=g)
    g_prob = torch.randn(im_prob.shape, generator=g)
    g_depth = torch.randn(depth.shape, generator=g)
    g_norm = torch.randn(norm.shape, generator=g)
    loss = (color * g_color).sum() + (im_prob * g_prob).sum() + (depth * g_depth).sum() + (norm * g_norm).sum()
    loss.backward()
    out = dict(H=H, W=W, ids=np.asarray(ids), Rs=batch["Rs"], ts=batch["ts"], Ks=batch["Ks"],
               color=color.detach().numpy(), prob=im_prob.detach().numpy(), mask=im_mask.detach().numpy(),
               norm=norm.detach().numpy(), depth=depth.detach().numpy(),
               g_color=g_color.numpy(), g_prob=g_prob.numpy(), g_depth=g_depth.numpy(), g_norm=g_norm.numpy(),
               grad_Rs=Rs.grad.numpy(), grad_ts=ts.grad.numpy(),
               grad_cam_rot=cams[0].grad.numpy(), grad_cam_pos=cams[1].grad.numpy())
    for i, m in enumerate(meshes):
        for k, v in m.items():
            out[f"mesh{i}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ref_batch64.npz"), **out)
    print("ref_batch64: covered", int((im_mask > 0.5).sum()), "grad_Rs max", float(Rs.grad.abs().max()))


def golden_multi(ref, H=64, W=80, seed=3):
    meshes = small_meshes()
    models = to_models(meshes)
    n = 3
    Rsn, _ = synth.random_rotations(n, seed)
    tsn = np.array([[-0.03, 0.0, 0.55], [0.02, 0.01, 0.5], [0.0, -0.02, 0.62]], np.float32)
    K = synth.crop_K(synth.K_LM, (325.0, 242.0), 200.0, W)
    Rs = torch.tensor(Rsn, requires_grad=True)
    ts = torch.tensor(tsn, requires_grad=True)
    cams = O.camera_params_from_RT_K(Rs, ts, torch.tensor(K), H, W, near=0.01, far=100.0)
    ren = ref["VCRenderMulti"](H, W)
    points = [[m["vertices"][None], m["faces"].long()] for m in models]
    color, im_prob, _, im_mask = ren(points, cams, [m["colors"][None] for m in models])
    g = torch.Generator().manual_seed(seed + 1)
    g_color = torch.randn(color.shape, generator=g)
    g_prob = torch.randn(im_prob.shape, generator=g)
    ((color * g_color).sum() + (im_prob * g_prob).sum()).backward()
    out = dict(H=H, W=W, Rs=Rsn, ts=tsn, K=K, color=color.detach().numpy(), prob=im_prob.detach().numpy(),
               mask=im_mask.detach().numpy(), g_color=g_color.numpy(), g_prob=g_prob.numpy(),
               grad_Rs=Rs.grad.numpy(), grad_ts=ts.grad.numpy())
    for i, m in enumerate(meshes):
        for k, v in m.items():
            out[f"mesh{i}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ref_multi64.npz"), **out)
    print("ref_multi64: covered", int((im_mask > 0.5).sum()))


def golden_seam(ref, H=48, W=64, seed=5):
    """the reference's own LinearRasterizer.apply (rasterizer.py:294) at the operator seam, incl. backward"""
    mesh = synth.icosphere(2, radius=0.05, noise_sigma=0.004, seed=seed)
    R, _ = synth.random_rotations(1, seed)
    t = np.array([[0.01, -0.015, 0.5]], np.float32)
    K = synth.crop_K(synth.K_LM, (325.0, 242.0), 170.0, W)
    cams = O.camera_params_from_RT_K(torch.tensor(R), torch.tensor(t), torch.tensor(K), H, W)
    v, f = torch.tensor(mesh["vertices"])[None], torch.tensor(mesh["faces"]).long()
    p3, p2, nrm = ref["perspective_projection"](v, f, cams)
    c = torch.tensor(mesh["colors"])[None]
    one = torch.ones_like(c[:, f[:, 0], :1])
    attr = torch.cat((c[:, f[:, 0]], one, c[:, f[:, 1]], one, c[:, f[:, 2]], one), dim=2)
    p2 = p2.detach().clone().requires_grad_(True)
    attr = attr.detach().clone().requires_grad_(True)
    im, prob = ref["linear_rasterizer"](W, H, p3.detach(), p2, nrm[:, :, 2:3].detach(), attr)
    g = torch.Generator().manual_seed(seed + 1)
    g_im, g_prob = torch.randn(im.shape, generator=g), torch.randn(prob.shape, generator=g)
    ((im * g_im).sum() + (prob * g_prob).sum()).backward()
    np.savez_compressed(os.path.join(OUT, "ref_seam48x64.npz"), H=H, W=W, points3d=p3.detach().numpy(),
                        points2d=p2.detach().numpy(), normalz=nrm[:, :, 2:3].detach().numpy(), attr=attr.detach().numpy(),
                        im=im.detach().numpy(), prob=prob.detach().numpy(), g_im=g_im.numpy(), g_prob=g_prob.numpy(),
                        grad_points2d=p2.grad.numpy(), grad_attr=attr.grad.numpy())
    print("ref_seam48x64: covered", int((im[..., 3] > 0.5).sum()))


if __name__ == "__main__" and not any(a in sys.argv for a in ("--nnd", "--tex", "--maskloss", "--diceloss", "--normloss")):
    import warnings
    warnings.filterwarnings("ignore")
    ref = O.import_reference()
    golden_seam(ref)
    golden_batch(ref)
    golden_multi(ref)


# ------------------------------------------------------------------------------------------------------------------
# chamfer nearest-neighbour op + depth back-projection chamfer loss: driven through the reference's OWN Python
# (core/csrc/torch_nndistance/torch_nndistance.py, core/self6dpp/losses/depth_bp_chamfer_loss.py) on top of the
# reference's OWN nnd_cpu.cpp compiled into oracle/_ref/libnnd_ref.so (make -C oracle ref).
# ------------------------------------------------------------------------------------------------------------------
def import_reference_chamfer():
    import importlib.util
    import types
    from oracle import nnd_oracle as N
    ref = N.ref_module()
    assert ref is not None, "run `make -C oracle ref` first"
    sys.modules["torch_nndistance_aten"] = ref
    spec = importlib.util.spec_from_file_location(
        "ref_torch_nndistance", "/root/reference/core/csrc/torch_nndistance/torch_nndistance.py")
    nnd_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nnd_mod)
    # the loss module's imports: only backproject_th, smooth_l1_loss and NND are used on the path
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    for pkg in ("core", "core.csrc", "core.csrc.torch_nndistance", "fvcore", "lib", "lib.pysixd", "lib.vis_utils"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    stub("core.csrc.torch_nndistance.torch_nndistance", nnd=nnd_mod.nnd)
    sys.modules["core.csrc.torch_nndistance"].torch_nndistance = sys.modules["core.csrc.torch_nndistance.torch_nndistance"]
    stub("fvcore.nn", smooth_l1_loss=lambda a, b, beta, reduction: (a - b).abs().mean())
    stub("lib.pysixd.misc", backproject_th=N.backproject_th)          # lib/pysixd/misc.py:350-367 restated (needs numba/mmcv)
    stub("lib.vis_utils.image", heatmap=lambda *a, **k: None)
    spec = importlib.util.spec_from_file_location(
        "ref_depth_bp_chamfer_loss", "/root/reference/core/self6dpp/losses/depth_bp_chamfer_loss.py")
    loss_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(loss_mod)
    return nnd_mod, loss_mod


def golden_nnd():
    nnd_mod, loss_mod = import_reference_chamfer()
    g = torch.Generator().manual_seed(11)
    x1 = (torch.randn(2, 500, 3, generator=g) * 0.05).requires_grad_(True)
    x2 = (torch.randn(2, 700, 3, generator=g) * 0.05).requires_grad_(True)
    x2.data[0, 10] = x2.data[0, 3]                                   # duplicate target: first index must win
    d1, d2 = nnd_mod.nnd(x1, x2)
    g1, g2 = torch.randn(d1.shape, generator=g), torch.randn(d2.shape, generator=g)
    ((d1 * g1).sum() + (d2 * g2).sum()).backward()
    # depth loss: two 40x48 depth maps of a tilted plane / bumpy blob with holes
    H, W, B = 40, 48, 3
    K = torch.tensor([[60.0, 0, 23.5], [0, 60.0, 19.5], [0, 0, 1]])
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    real = 0.8 + 0.002 * xx + 0.001 * yy + 0.01 * torch.randn(B, H, W, generator=g)
    real = real * (((xx - 24) ** 2 + (yy - 20) ** 2) < 15 ** 2)
    ren = (0.82 + 0.0015 * xx + 0.0012 * yy + 0.01 * torch.randn(B, H, W, generator=g))
    ren = ren * (((xx - 22) ** 2 + (yy - 21) ** 2) < 14 ** 2)
    ren[2] = 0                                                        # an empty render: the reference skips the sample (nan)
    ren = ren.requires_grad_(True)
    loss, loss_c = loss_mod.depth_bp_chamfer_loss(ren, real, K, distance_threshold=0.05, center_lw=0.5)
    (loss + loss_c).backward()
    np.savez_compressed(os.path.join(OUT, "ref_nnd.npz"), x1=x1.detach().numpy(), x2=x2.detach().numpy(),
                        d1=d1.detach().numpy(), d2=d2.detach().numpy(), g1=g1.numpy(), g2=g2.numpy(),
                        gx1=x1.grad.numpy(), gx2=x2.grad.numpy(), K=K.numpy(), real=real.numpy(), ren=ren.detach().numpy(),
                        loss=loss.detach().numpy(), loss_center=loss_c.detach().numpy(), g_ren=ren.grad.numpy())
    print("ref_nnd: loss", float(loss), "center", float(loss_c))


if __name__ == "__main__" and "--nnd" in sys.argv:
    golden_nnd()


# ------------------------------------------------------------------------------------------------------------------
# texture / SH / Phong modes: the reference's OWN TexRender, TexRenderBatch, TexRenderMulti,
# SHRender, PhongRender (+ fragment shaders) on CPU with the oracle at the kaolin boundary.
# ------------------------------------------------------------------------------------------------------------------
def import_reference_tex():
    O.install_reference_stubs()
    from lib.dr_utils.dib_renderer_x.renderer.texrender import TexRender
    from lib.dr_utils.dib_renderer_x.renderer.texrender_batch import TexRenderBatch
    from lib.dr_utils.dib_renderer_x.renderer.texrender_multi import TexRenderMulti
    from lib.dr_utils.dib_renderer_x.renderer.shrender import SHRender
    from lib.dr_utils.dib_renderer_x.renderer.phongrender import PhongRender
    return dict(TexRender=TexRender, TexRenderBatch=TexRenderBatch, TexRenderMulti=TexRenderMulti, SHRender=SHRender,
                PhongRender=PhongRender)


def sphere_uv(verts):
    """a (u, v) per vertex from the direction of the vertex; deliberately leaves [0,1] so the wrap-around is used"""
    v = verts / verts.norm(dim=1, keepdim=True)
    return torch.stack((torch.atan2(v[:, 1], v[:, 0]) / (2 * np.pi) + 0.55, torch.asin(v[:, 2].clamp(-1, 1)) / np.pi + 0.6), dim=1)


def face_uv_layout(verts, faces, seed):
    """'face uv' layout: every face owns three uv rows (a permutation of the welded ones), so ft != faces"""
    g = torch.Generator().manual_seed(seed)
    uv_v = sphere_uv(verts)
    f = faces.long()
    perm = torch.randperm(3 * f.shape[0], generator=g)
    uv_rows = torch.empty(3 * f.shape[0], 2)
    uv_rows[perm] = uv_v[f.reshape(-1)]
    return uv_rows, perm.reshape(-1, 3)


def golden_tex():
    ref = import_reference_tex()
    H, W = 48, 64
    meshes = small_meshes()
    models = to_models(meshes)
    g = torch.Generator().manual_seed(21)
    ids = [0, 1, 2]
    batch = synth.roi_batch([meshes[i] for i in ids], 3, res=H, seed=5, fill=(0.5, 0.8))
    Rs, ts = torch.tensor(batch["Rs"]), torch.tensor(batch["ts"])
    K = torch.tensor(batch["Ks"][0])
    K[0, 2] += (W - H) / 2.0
    out = {"H": H, "W": W, "Rs": Rs.numpy(), "ts": ts.numpy(), "K": K.numpy()}
    cams = O.camera_params_from_RT_K(Rs, ts, K, H, W, near=0.01, far=100.0)
    # ---- TexRenderBatch + TexRenderMulti: three objects, textures of different sizes, object 1 with face uvs
    verts = [models[i]["vertices"].clone().requires_grad_(True) for i in ids]
    faces = [models[i]["faces"].long() for i in ids]
    uvs, fts, texs = [], [], []
    for k, i in enumerate(ids):
        if k == 1:
            uv_rows, ft = face_uv_layout(models[i]["vertices"], faces[k], seed=3)
        else:
            uv_rows, ft = sphere_uv(models[i]["vertices"]), faces[k]
        uvs.append(uv_rows.clone().requires_grad_(True))
        fts.append(ft)
        texs.append(torch.rand(1, 3, 16 + 8 * k, 24 - 4 * k, generator=g).requires_grad_(True))
    points = [[v[None], f] for v, f in zip(verts, faces)]
    for name in ("TexRenderBatch", "TexRenderMulti"):
        ren = ref[name](H, W)
        if name == "TexRenderMulti":
            # the reference composites with in-place indexed writes (texrender_multi.py:124-136), which autograd rejects:
            # forward only
            with torch.no_grad():
                im, prob, normal1, mask = ren(points, cams, [u[None] for u in uvs], texs, ts=ts, ft_fx3=fts)
            out.update({f"{name}_im": im.numpy(), f"{name}_prob": prob.numpy(), f"{name}_mask": mask.numpy()})
            continue
        im, prob, normal1, mask = ren(points, cams, [u[None] for u in uvs], texs, ft_fx3=fts)
        gi, gp = torch.randn(im.shape, generator=g), torch.randn(prob.shape, generator=g)
        for t in verts + uvs + texs:
            t.grad = None
        ((im * gi).sum() + (prob * gp).sum()).backward()
        out.update({f"{name}_im": im.detach().numpy(), f"{name}_prob": prob.detach().numpy(), f"{name}_mask": mask.detach().numpy(),
                    f"{name}_gi": gi.numpy(), f"{name}_gp": gp.numpy()})
        for k in range(3):
            out[f"{name}_gv{k}"] = verts[k].grad.numpy().copy()
            out[f"{name}_gtex{k}"] = texs[k].grad.numpy().copy()
            out[f"{name}_normal1_{k}"] = normal1[k].detach().numpy()
    for k in range(3):
        out[f"verts{k}"], out[f"faces{k}"] = verts[k].detach().numpy(), faces[k].numpy()
        out[f"uv{k}"], out[f"ft{k}"], out[f"tex{k}"] = uvs[k].detach().numpy(), fts[k].numpy(), texs[k].detach().numpy()
    # ---- one topology, batch of 2 vertex sets: TexRender (bilinear), SHRender (flat and smooth), PhongRender
    m = models[1]
    f1 = m["faces"].long()
    vb = torch.stack((m["vertices"], m["vertices"] * 1.05 + 0.002)).requires_grad_(True)
    cams2 = O.camera_params_from_RT_K(Rs[:2], ts[:2], K, H, W, near=0.01, far=100.0)
    uvb = torch.stack((sphere_uv(m["vertices"]), sphere_uv(m["vertices"]) * 1.3))
    texb = torch.rand(2, 3, 20, 28, generator=g).requires_grad_(True)
    # vertex <- face averaging matrix for smooth shading (p x f)
    P_, F_ = m["vertices"].shape[0], f1.shape[0]
    pf = torch.zeros(P_, F_)
    for c in range(3):
        pf[f1[:, c], torch.arange(F_)] = 1.0
    pf = pf / pf.sum(1, keepdim=True).clamp(min=1)
    light9 = torch.randn(2, 9, generator=g) * 0.5 + 0.6
    lightdir = torch.tensor([[0.3, 0.2, 1.0], [-0.4, 0.5, 0.8]])
    material = torch.rand(2, 3, 3, generator=g) * 0.5 + 0.2
    shin = torch.tensor([[4.0], [9.0]])
    out.update(vb=vb.detach().numpy(), f1=f1.numpy(), uvb=uvb.numpy(), texb=texb.detach().numpy(), pf=pf.numpy(),
               light9=light9.numpy(), lightdir=lightdir.numpy(), material=material.numpy(), shin=shin.numpy())

    def run(tag, ren, *args):
        vb.grad = None
        texb.grad = None
        im, prob, normal1, mask = ren([vb, f1], cams2, uvb, texb, *args)
        gi, gp = torch.randn(im.shape, generator=g), torch.randn(prob.shape, generator=g)
        ((im * gi).sum() + (prob * gp).sum()).backward()
        out.update({f"{tag}_im": im.detach().numpy(), f"{tag}_prob": prob.detach().numpy(), f"{tag}_mask": mask.detach().numpy(),
                    f"{tag}_normal1": normal1.detach().numpy(), f"{tag}_gi": gi.numpy(), f"{tag}_gp": gp.numpy(),
                    f"{tag}_gv": vb.grad.numpy().copy(), f"{tag}_gtex": texb.grad.numpy().copy()})
        print(tag, "covered", int((mask > 0.5).sum()), "im mean", float(im.mean()))
    run("TexRender", ref["TexRender"](H, W, filtering="bilinear"))
    run("SHflat", ref["SHRender"](H, W), light9)
    sh = ref["SHRender"](H, W)
    sh.set_smooth(pf[None])
    run("SHsmooth", sh, light9)
    run("Phong", ref["PhongRender"](H, W), lightdir, material, shin)
    np.savez_compressed(os.path.join(OUT, "ref_tex.npz"), **out)
    print("ref_tex written:", len(out), "arrays")


if __name__ == "__main__" and "--tex" in sys.argv:
    golden_tex()


# ------------------------------------------------------------------------------------------------------------------
# re-weighted BCE on probabilities: the reference's OWN weighted_ex_loss_probs (mask_losses.py:63-108) on CPU
# ------------------------------------------------------------------------------------------------------------------
def golden_maskloss():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_mask_losses", "/root/reference/core/self6dpp/losses/mask_losses.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(31)
    out = {}
    for tag, shape, with_w in (("a", (3, 1, 40, 56), True), ("b", (2, 1, 33, 47), False), ("c", (1, 1, 8, 8), True)):
        probs = torch.rand(shape, generator=g)
        probs.view(-1)[:5] = torch.tensor([0.0, 1.0, 1e-9, 1 - 1e-9, 0.5])          # the clamp's both sides
        target = (torch.rand(shape, generator=g) > 0.6).float() * (0.5 + 0.5 * torch.rand(shape, generator=g))   # soft positives
        if tag == "c":
            target[:] = 0                                                             # no positives: the term is dropped
        weight = torch.rand(shape, generator=g) + 0.5 if with_w else None
        probs.requires_grad_(True)
        loss = mod.weighted_ex_loss_probs(probs, target, weight=weight)
        (loss * 1.7).backward()
        out.update({f"{tag}_probs": probs.detach().numpy(), f"{tag}_target": target.numpy(), f"{tag}_loss": loss.detach().numpy(),
                    f"{tag}_grad": probs.grad.numpy()})
        if with_w:
            out[f"{tag}_weight"] = weight.numpy()
        print("maskloss", tag, float(loss))
    np.savez_compressed(os.path.join(OUT, "ref_maskloss.npz"), **out)


if __name__ == "__main__" and "--maskloss" in sys.argv:
    golden_maskloss()


# ------------------------------------------------------------------------------------------------------------------
# soft dice loss: the reference's OWN soft_dice_loss (mask_losses.py:444-463) on CPU
# ------------------------------------------------------------------------------------------------------------------
def golden_diceloss():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_mask_losses", "/root/reference/core/self6dpp/losses/mask_losses.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(41)
    out = {}
    cases = (("a", (3, 1, 40, 56), 0.0, 0.002, "mean"), ("b", (2, 1, 33, 47), 1.0, 1e-7, "sum"), ("c", (4, 1, 9, 7), 0.0, 0.002, "none"),
             ("d", (2, 1, 8, 8), 0.0, 0.002, "mean"))
    for tag, shape, smooth, eps, red in cases:
        probs = torch.rand(shape, generator=g)
        labels = (torch.rand(shape, generator=g) > 0.6).float()
        if tag == "d":
            labels[1] = 0                                     # an empty label: score 0, loss contribution 1
        go = torch.rand(shape[0] if red == "none" else 1, generator=g) + 0.5
        probs.requires_grad_(True)
        loss = mod.soft_dice_loss(probs, labels, smooth=smooth, eps=eps, reduction=red)
        (loss * (go if red == "none" else go[0])).sum().backward()
        out.update({f"{tag}_probs": probs.detach().numpy(), f"{tag}_labels": labels.numpy(), f"{tag}_loss": loss.detach().numpy(),
                    f"{tag}_grad": probs.grad.numpy(), f"{tag}_go": go.numpy(), f"{tag}_cfg": np.array([smooth, eps]),
                    f"{tag}_red": np.array(red)})
        print("diceloss", tag, loss.detach().numpy())
    np.savez_compressed(os.path.join(OUT, "ref_diceloss.npz"), **out)


if __name__ == "__main__" and "--diceloss" in sys.argv:
    golden_diceloss()


# ------------------------------------------------------------------------------------------------------------------
# normal-map loss: the reference's OWN NORMLoss (vf_norm_loss.py:56-103) on CPU
# ------------------------------------------------------------------------------------------------------------------
def golden_normloss():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_vf_norm_loss", "/root/reference/core/self6dpp/losses/vf_norm_loss.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(51)
    out = {}
    for tag, shape, l1, cs in (("a", (3, 3, 16, 20), True, True), ("b", (2, 3, 9, 7), False, True), ("c", (2, 3, 8, 8), True, False)):
        b, _, h, w = shape
        o = torch.randn(shape, generator=g)
        gt = torch.nn.functional.normalize(torch.randn(shape, generator=g), dim=1)
        m = (torch.rand(b, 1, h, w, generator=g) > 0.4).float()
        o.requires_grad_(True)
        loss = mod.NORMLoss(with_l1=l1, with_cs=cs)(o, gt, m)
        (loss * 1.3).backward()
        out.update({f"{tag}_out": o.detach().numpy(), f"{tag}_gt": gt.numpy(), f"{tag}_mask": m.numpy(), f"{tag}_loss": loss.detach().numpy(),
                    f"{tag}_grad": o.grad.numpy(), f"{tag}_flags": np.array([int(l1), int(cs)])})
        print("normloss", tag, float(loss))
    np.savez_compressed(os.path.join(OUT, "ref_normloss.npz"), **out)


if __name__ == "__main__" and "--normloss" in sys.argv:
    golden_normloss()
