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
for o in out: print(json.dumps(o))
