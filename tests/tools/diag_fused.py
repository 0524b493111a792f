import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import dibr_oracle as O
from tests import helpers as Hh
from tests.test_gpu_fused_parity import to_dev_models, DEV
from self6dpp_b200 import Renderer_dibr, synth
from tests.golden.make_golden import small_meshes

def rel(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())

# ---- A: pose grads vs fp64
meshes = small_meshes(); H = W = 64; ids = [2, 0, 1, 1]; B = 4
batch = synth.roi_batch([meshes[i] for i in ids], B, res=W, seed=7, fill=(0.45, 0.7))
models = to_dev_models(meshes)
Rs = torch.tensor(batch["Rs"], device=DEV, requires_grad=True); ts = torch.tensor(batch["ts"], device=DEV, requires_grad=True)
ren = Renderer_dibr(H, W, "VertexColorBatch")
ret = ren.render_batch(Rs, ts, [models[i] for i in ids], Ks=torch.tensor(batch["Ks"], device=DEV), width=W, height=H, mode=["color", "depth", "mask", "prob"])
g = torch.Generator().manual_seed(3)
g_color = torch.randn(B, H, W, 3, generator=g, dtype=torch.float64); g_prob = torch.randn(B, H, W, 1, generator=g, dtype=torch.float64); g_depth = torch.randn(B, H, W, 1, generator=g, dtype=torch.float64)
for which in ("color", "prob", "depth", "all"):
    gc = g_color if which in ("color", "all") else torch.zeros_like(g_color)
    gp = g_prob if which in ("prob", "all") else torch.zeros_like(g_prob)
    gd = g_depth if which in ("depth", "all") else torch.zeros_like(g_depth)
    grads = {"im": torch.cat([gc, torch.zeros(B, H, W, 1, dtype=torch.float64), gd], -1), "prob": gp}
    ref = Hh.oracle_render_batch64(meshes, ids, batch["Rs"], batch["ts"], batch["Ks"], H, W, ["colors"], True, grads)
    Rs.grad = None; ts.grad = None
    loss = (ret["color"] * gc.float().to(DEV)).sum() + (ret["prob"] * gp[..., 0].float().to(DEV)).sum() + (ret["depth"] * gd[..., 0].float().to(DEV)).sum()
    loss.backward(retain_graph=True)
    print(which, "dR rel", rel(Rs.grad, ref["grad_Rs"]), "dt rel", rel(ts.grad, ref["grad_ts"]), "scale", float(ref["grad_Rs"].abs().max()))
got_im = torch.cat([ret["color"], ret["mask"].unsqueeze(-1), ret["depth"].unsqueeze(-1)], -1).detach().cpu().double()
same = ((ref["imidx"] > 0) == (ret["mask"].detach().cpu() > 0.5).unsqueeze(-1))
print("same frac", same.float().mean().item(), "im maxerr", ((got_im - ref["im"]).abs() * same).max().item(), "prob maxerr", ((ret["prob"].detach().cpu().double().unsqueeze(-1) - ref["prob"]).abs() * same).max().item())

# ---- B: golden batch, compare golden and mine vs fp64 truth for color+prob+depth only
d, meshes = Hh.load_golden("ref_batch64.npz"); H, W = int(d["H"]), int(d["W"]); ids = [int(i) for i in d["ids"]]; B = len(ids)
models = to_dev_models(meshes)
for terms in (("color",), ("prob",), ("depth",), ("norm",), ("color", "prob", "depth", "norm")):
    Rs = torch.tensor(d["Rs"], device=DEV, requires_grad=True); ts = torch.tensor(d["ts"], device=DEV, requires_grad=True)
    ret = Renderer_dibr(H, W, "VertexColorBatch").render_batch(Rs, ts, [models[i] for i in ids], Ks=torch.tensor(d["Ks"], device=DEV), width=W, height=H, mode=["color", "depth", "mask", "norm", "prob"])
    dev = lambda k: torch.tensor(d[k], device=DEV)
    loss = 0
    if "color" in terms: loss = loss + (ret["color"] * dev("g_color")).sum()
    if "prob" in terms: loss = loss + (ret["prob"] * dev("g_prob")[..., 0]).sum()
    if "depth" in terms: loss = loss + (ret["depth"] * dev("g_depth")).sum()
    if "norm" in terms: loss = loss + (ret["norm"] * dev("g_norm")).sum()
    loss.backward()
    # fp64 truth with the same terms
    dt = torch.float64
    gim = torch.zeros(B, H, W, 8, dtype=dt)   # [rgb, nxyz, one, depth]
    if "color" in terms: gim[..., 0:3] = torch.tensor(d["g_color"], dtype=dt)
    if "depth" in terms: gim[..., 7] = torch.tensor(d["g_depth"], dtype=dt)
    gpr = torch.tensor(d["g_prob"], dtype=dt) if "prob" in terms else torch.zeros(B, H, W, 1, dtype=dt)
    fwd = Hh.oracle_render_batch64(meshes, ids, d["Rs"], d["ts"], d["Ks"], H, W, ["colors", "normals"], True, None)
    if "norm" in terms:
        leaf = fwd["im"].clone().requires_grad_(True)
        rn = leaf[..., 3:6]; sh = rn - rn.min(); nm = sh / (torch.norm(sh, dim=-1, keepdim=True) + 1e-5) * leaf[..., 6:7]
        (nm * torch.tensor(d["g_norm"], dtype=dt)).sum().backward()
        gim = gim + leaf.grad
    ref = Hh.oracle_render_batch64(meshes, ids, d["Rs"], d["ts"], d["Ks"], H, W, ["colors", "normals"], True, {"im": gim, "prob": gpr})
    print(terms, "mine vs fp64: dR", rel(Rs.grad, ref["grad_Rs"]), "dt", rel(ts.grad, ref["grad_ts"]))
    if len(terms) == 4:
        print("   golden(ref fp32 python) vs fp64: dR", rel(d["grad_Rs"], ref["grad_Rs"]), "dt", rel(d["grad_ts"], ref["grad_ts"]), " mine vs golden dR", rel(Rs.grad, d["grad_Rs"]))
