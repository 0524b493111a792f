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


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore")
    ref = O.import_reference()
    golden_seam(ref)
    golden_batch(ref)
    golden_multi(ref)
