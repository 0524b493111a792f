import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tests.test_gpu_full_size import dev_models, DEV
from self6dpp_b200 import Renderer_dibr, synth
meshes = synth.lm13_meshes()[:8]; models = dev_models(meshes); H, W = 480, 640
rng = np.random.default_rng(3); Rs, _ = synth.random_rotations(8, 4)
ts = np.stack([np.array([rng.uniform(-0.09, 0.09), rng.uniform(-0.06, 0.06), rng.uniform(0.45, 0.8)], np.float32) for _ in range(8)])
K = torch.tensor(synth.K_LM, device=DEV); tR, tt = torch.tensor(Rs, device=DEV), torch.tensor(ts, device=DEV)
scene = Renderer_dibr(H, W, "VertexColorMulti").render_scene(tR, tt, models, K=K, width=W, height=H)
batch = Renderer_dibr(H, W, "VertexColorBatch").render_batch(tR, tt, models, Ks=K, width=W, height=H, mode=["color", "depth", "mask", "prob"])
covered = batch["mask"] > 0.5
depth = torch.where(covered, batch["depth"], torch.full_like(batch["depth"], 1e9))
zmin, who = depth.min(dim=0); any_cov = covered.any(0)
second = depth.clone(); second.scatter_(0, who.unsqueeze(0), 1e9)
unique = any_cov & (second.min(dim=0)[0] > zmin)
pick = torch.gather(batch["color"], 0, who.view(1, H, W, 1).expand(1, H, W, 3))[0]
diff = (scene["color"] - pick).abs().max(-1)[0]
bad = unique & (diff > 0)
print("covered", int(any_cov.sum()), "unique", int(unique.sum()), "bad", int(bad.sum()), "maxdiff", float(diff[unique].max()))
ys, xs = torch.nonzero(bad, as_tuple=True)
for y, x in list(zip(ys.tolist(), xs.tolist()))[:8]:
    d = depth[:, y, x]
    order = torch.argsort(d)[:3].tolist()
    sc = scene["color"][y, x].tolist()
    match = [i for i in range(8) if covered[i, y, x] and torch.equal(batch["color"][i, y, x], scene["color"][y, x])]
    print((y, x), "depths", [(i, float(d[i])) for i in order], "scene depth", float(scene["depth"][y, x]), "scene color matches obj", match)
