"""Timing of BASELINE.json's other configurations through the Python API (CUDA events, 10 reps after 3 warm-ups, L2 flush
between reps).  bench.py's headline is cfg2; these lines go to profiles/r01_configs.md."""
import sys, os, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from self6dpp_b200 import Renderer_dibr, synth
DEV = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)

def dev_models(meshes):
    return [{"vertices": torch.tensor(m["vertices"], device=DEV), "colors": torch.tensor(m["colors"], device=DEV),
             "normals": torch.tensor(m["normals"], device=DEV), "faces": torch.tensor(m["faces"], device=DEV, dtype=torch.int32)} for m in meshes]

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)

out = []
# cfg1: one 5,120-face mesh, batch 1, 480x640, colour + depth passes fwd+bwd
mesh = synth.icosphere(4, radius=0.05, noise_sigma=0.005, seed=0); models = dev_models([mesh]); H, W = 480, 640
R, _ = synth.random_rotations(1, 0); t = np.array([[0, 0, 0.8]], np.float32)
ren = Renderer_dibr(H, W, "VertexColorBatch")
g = torch.Generator().manual_seed(1)
gc, gp, gd = torch.randn(1, H, W, 3, generator=g).to(DEV), torch.randn(1, H, W, generator=g).to(DEV), torch.randn(1, H, W, generator=g).to(DEV)
def cfg1():
    Rs = torch.tensor(R, device=DEV, requires_grad=True); ts = torch.tensor(t, device=DEV, requires_grad=True)
    ret = ren.render_batch(Rs, ts, models, Ks=torch.tensor(synth.K_LM, device=DEV), width=W, height=H, mode=["color", "depth", "mask", "prob"])
    torch.autograd.backward([ret["color"], ret["prob"], ret["depth"]], [gc, gp, gd])
ms = timeit(cfg1); out.append({"config": "cfg1 (B=1, 480x640, F=5120, colour+depth+prob fwd+bwd)", "ms": ms, "samples_per_s": 1e3 / ms})
# cfg3: 8 scenes x 8 objects (~40k faces each), 480x640, fwd+bwd
meshes = synth.lm13_meshes()[:8]; models8 = dev_models(meshes)
rng = np.random.default_rng(3)
scenes = []
for s in range(8):
    Rs, _ = synth.random_rotations(8, 40 + s)
    ts = np.stack([np.array([rng.uniform(-0.09, 0.09), rng.uniform(-0.06, 0.06), rng.uniform(0.45, 0.8)], np.float32) for _ in range(8)])
    scenes.append((Rs, ts))
sren = Renderer_dibr(H, W, "VertexColorMulti")
gsc, gsp = torch.randn(H, W, 3, generator=g).to(DEV), torch.randn(H, W, generator=g).to(DEV)
def cfg3():
    for Rs, ts in scenes:
        tR = torch.tensor(Rs, device=DEV, requires_grad=True); tt = torch.tensor(ts, device=DEV, requires_grad=True)
        ret = sren.render_scene(tR, tt, models8, K=torch.tensor(synth.K_LM, device=DEV), width=W, height=H)
        torch.autograd.backward([ret["color"], ret["prob"]], [gsc, gsp])
ms = timeit(cfg3); out.append({"config": "cfg3 (8 scenes x 8 objects, ~40k faces per scene, 480x640, fwd+bwd)", "ms": ms, "scenes_per_s": 8e3 / ms})
# cfg4: 100,352-face mesh, batch 16, 480x640, fwd+bwd
big = synth.ellipsoid(225, 224, radii=(0.06, 0.05, 0.045), noise_sigma=0.001, seed=5); mb = dev_models([big])
B = 16; Rb, _ = synth.random_rotations(B, 6); tb = np.tile(np.array([[0.0, 0.0, 0.9]], np.float32), (B, 1))
bren = Renderer_dibr(H, W, "VertexColorBatch")
gbc, gbp, gbd = torch.randn(B, H, W, 3, generator=g).to(DEV), torch.randn(B, H, W, generator=g).to(DEV), torch.randn(B, H, W, generator=g).to(DEV)
def cfg4():
    Rs = torch.tensor(Rb, device=DEV, requires_grad=True); ts = torch.tensor(tb, device=DEV, requires_grad=True)
    ret = bren.render_batch(Rs, ts, mb * B, Ks=torch.tensor(synth.K_YCBV, device=DEV), width=W, height=H, mode=["color", "depth", "mask", "prob"])
    torch.autograd.backward([ret["color"], ret["prob"], ret["depth"]], [gbc, gbp, gbd])
ms = timeit(cfg4, reps=5); out.append({"config": "cfg4 (B=16, 480x640, F=100352, colour+depth+prob fwd+bwd)", "ms": ms, "samples_per_s": 16e3 / ms})
# texture batch: 32 textured cfg2 objects, 256x256, vertex uvs, 64x64 textures, fwd + bwd to the pose and the texture
import bench
meshes2, student, _ = bench.workload(0)
def uv_of(v):
    n = v / np.linalg.norm(v, axis=1, keepdims=True)
    return np.stack((np.arctan2(n[:, 1], n[:, 0]) / (2 * np.pi) + 0.5, np.arcsin(np.clip(n[:, 2], -1, 1)) / np.pi + 0.5), 1).astype(np.float32)
tex_models = []
for m in meshes2:
    tex_models.append({"vertices": torch.tensor(m["vertices"], device=DEV), "faces": torch.tensor(m["faces"], device=DEV, dtype=torch.int32),
                       "vertex_uvs": torch.tensor(uv_of(m["vertices"]), device=DEV),
                       "texture": torch.rand(3, 64, 64, generator=g).to(DEV).requires_grad_(True)})
cur = [tex_models[int(i)] for i in student["ids"]]
tren = Renderer_dibr(256, 256, "TextureBatch")
gtc = torch.randn(32, 256, 256, 3, generator=g).to(DEV); gtp = torch.randn(32, 256, 256, generator=g).to(DEV)
Kt = torch.tensor(student["Ks"], device=DEV)
def tex():
    Rs = torch.tensor(student["Rs"], device=DEV, requires_grad=True); ts = torch.tensor(student["ts"], device=DEV, requires_grad=True)
    ret = tren.render_batch_tex(Rs, ts, cur, Ks=Kt, width=256, height=256, uv_type="vertex", mode=["color"])
    torch.autograd.backward([ret["color"], ret["prob"]], [gtc, gtp])
ms = timeit(tex, reps=5); out.append({"config": "texture batch (B=32, 256x256, cfg2 meshes, TextureBatch colour+prob fwd+bwd, torch vertex/fragment shaders)", "ms": ms, "samples_per_s": 32e3 / ms})
# RW-BCE mask loss (loss half): ours vs the reference's expression in torch (boolean indexing, host syncs)
from self6dpp_b200.losses import weighted_ex_loss_probs
mp = torch.rand(32, 1, 256, 256, generator=g).to(DEV); mt = (torch.rand(32, 1, 256, 256, generator=g) > 0.7).float().to(DEV)
mw = (torch.rand(32, 1, 256, 256, generator=g) + 0.5).to(DEV)
def ours():
    p = mp.clone().requires_grad_(True); weighted_ex_loss_probs(p, mt, weight=mw).backward()
def torch_ref():          # mask_losses.py:63-108 restated
    p = mp.clone().requires_grad_(True)
    pos, neg = mt > 0, mt == 0
    q = p.clamp(min=1e-7, max=1 - 1e-7)
    pl = -mt[pos] * torch.log(q[pos]) * mw[pos]; nl = -(torch.log(1 - q[neg])) * mw[neg]
    if torch.isnan(pl).any() or torch.isnan(nl).any(): print("nan")
    loss = 0.0; npos, nneg = pos.sum(), neg.sum()
    if npos > 0: loss = loss + 1.0 / npos.float() * pl.sum()
    if nneg > 0: loss = loss + 1.0 / nneg.float() * nl.sum()
    loss.backward()
out.append({"config": "RW-BCE mask loss fwd+bwd on [32,1,256,256]", "ms": timeit(ours), "torch_expression_ms": timeit(torch_ref)})
# photometric losses: ours vs the reference's expressions restated in torch eager on the same GPU
import torch.nn.functional as F
from self6dpp_b200.losses import lab_l1_loss
from self6dpp_b200.ssim import MS_SSIM, create_window
pg = torch.rand(32, 3, 256, 256, generator=g).to(DEV); pr = (pg + 0.1 * torch.randn(32, 3, 256, 256, generator=g).to(DEV)).clamp(0.002, 1)
pm = (torch.rand(32, 1, 256, 256, generator=g) > 0.4).float().to(DEV)
def t_rgb_to_lab_n(img):   # lab.py:16-82 + xyz.py:28-30
    r, gg, b = img[..., 0, :, :], img[..., 1, :, :], img[..., 2, :, :]
    rs = torch.where(r > 0.04045, torch.pow(((r + 0.055) / 1.055), 2.4), r / 12.92)
    gs = torch.where(gg > 0.04045, torch.pow(((gg + 0.055) / 1.055), 2.4), gg / 12.92)
    bs = torch.where(b > 0.04045, torch.pow(((b + 0.055) / 1.055), 2.4), b / 12.92)
    x = 0.412453 * rs + 0.357580 * gs + 0.180423 * bs; y = 0.212671 * rs + 0.715160 * gs + 0.072169 * bs; z = 0.019334 * rs + 0.119193 * gs + 0.950227 * bs
    xyz = torch.stack([x, y, z], -3) / torch.tensor([0.95047, 1.0, 1.08883], device=DEV)[..., :, None, None]
    f = torch.where(xyz > 0.008856, torch.pow(xyz, 1 / 3), 7.787 * xyz + 4.0 / 29.0)
    fx, fy, fz = f[..., 0, :, :], f[..., 1, :, :], f[..., 2, :, :]
    lab = torch.stack([116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz)], -3)
    mn = torch.tensor([0.0, -110.0, -110.0]).view(3, 1, 1).to(lab); mx = torch.tensor([100.0, 110.0, 110.0]).view(3, 1, 1).to(lab)
    return (lab - mn) / (mx - mn)
def lab_ours():
    r = pr.clone().requires_grad_(True); lab_l1_loss(pg, r, pm, no_l=True).backward()
def lab_torch():           # self_engine_utils.py:745-773
    r = pr.clone().requires_grad_(True)
    a = t_rgb_to_lab_n(pg[:, [2, 1, 0]]); b = t_rgb_to_lab_n(r[:, [2, 1, 0]].contiguous())
    (torch.abs(a[:, 1:] * pm - b[:, 1:] * pm).sum() / max(1, pm.sum())).backward()
out.append({"config": "Lab (a,b) L1 colour loss fwd+bwd on [32,3,256,256]", "ms": timeit(lab_ours), "torch_expression_ms": timeit(lab_torch)})
ms_mod = MS_SSIM(data_range=1.0, normalize=True).to(DEV)
win = create_window(11, 1.5).to(DEV).reshape(1, 1, 1, -1).repeat(3, 1, 1, 1); wts = ms_mod.weights
def t_filt(x):             # ssim.py:33-55
    return F.conv2d(F.conv2d(x, win, groups=3), win.transpose(2, 3), groups=3)
def ssim_ours():
    r = pr.clone().requires_grad_(True); (1 - ms_mod(pg * pm, r * pm)).mean().backward()
def ssim_torch():          # ssim.py:58-160 with normalize=True, as self_engine_utils.py:777-785 calls it
    r = pr.clone().requires_grad_(True); X, Y = pg * pm, r * pm
    css, ss = [], []
    for _ in range(5):
        mu1, mu2 = t_filt(X), t_filt(Y); s11, s22, s12 = t_filt(X * X) - mu1.pow(2), t_filt(Y * Y) - mu2.pow(2), t_filt(X * Y) - mu1 * mu2
        cs_map = (2 * s12 + 9e-4) / (s11 + s22 + 9e-4); ssim_map = ((2 * mu1 * mu2 + 1e-4) / (mu1.pow(2) + mu2.pow(2) + 1e-4)) * cs_map
        ss.append(ssim_map.mean(dim=(1, 2, 3))); css.append(cs_map.mean(dim=(1, 2, 3)))
        X, Y = F.avg_pool2d(X, 2, 2), F.avg_pool2d(Y, 2, 2)
    css = (torch.stack(css, 0) + 1) / 2; last = (ss[-1] + 1) / 2
    ms = torch.prod((css[:-1] ** wts[:-1].unsqueeze(1)) * (last ** wts[-1]), dim=0)
    (1 - ms).mean().backward()
out.append({"config": "MS-SSIM loss fwd+bwd on [32,3,256,256] (5 levels, normalize)", "ms": timeit(ssim_ours), "torch_expression_ms": timeit(ssim_torch),
            "note": "torch eager runs the depthwise convolutions through cuDNN with TF32 allowed (its default)"})
for o in out[-2:]: print(json.dumps(o))
if "--photometric-only" not in sys.argv:
    for o in out[:-2]: print(json.dumps(o))
