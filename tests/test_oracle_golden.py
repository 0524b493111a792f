"""CPU: the oracle restatements (fixed-order vertex shader, camera set-up, LinearRasterizer fwd/bwd) against
golden vectors produced by the REFERENCE's own Python layers (tests/golden/make_golden.py).  When the
reference tree is present (build container) the fixtures are also regenerated in memory and compared."""
import os

import numpy as np
import pytest
import torch

from oracle import dibr_oracle as O
from tests import helpers as Hh


def test_seam_golden_matches_oracle_restatement():
    d, _ = Hh.load_golden("ref_seam48x64.npz")
    H, W = int(d["H"]), int(d["W"])
    t = lambda k: torch.tensor(d[k])
    fw = O.rasterize(W, H, t("points3d"), t("points2d"), t("normalz"), t("attr"))
    # same fp32 inputs, same C kernels underneath: bit-identical
    assert np.array_equal(fw["im"].numpy(), d["im"]) and np.array_equal(fw["improb"].numpy(), d["prob"])
    dp2, dc = O.rasterize_backward(fw, t("g_im"), t("g_prob"))
    assert np.allclose(dp2.numpy(), d["grad_points2d"], rtol=1e-6, atol=1e-6)
    assert np.allclose(dc.numpy(), d["grad_attr"], rtol=1e-6, atol=1e-6)


def test_batch_golden_matches_fused_restatement():
    """reference VCRenderBatch (torch.matmul vertex shader) vs oracle.project (fixed FMA order) + rasterize."""
    d, meshes = Hh.load_golden("ref_batch64.npz")
    H, W = int(d["H"]), int(d["W"])
    grads = {"im": torch.cat([torch.tensor(d["g_color"]), torch.zeros(len(d["ids"]), H, W, 1)], -1).double(),
             "prob": torch.tensor(d["g_prob"]).double()}
    out = Hh.oracle_render_batch64(meshes, d["ids"], d["Rs"], d["ts"], d["Ks"], H, W, ["colors"], False, grads)
    color, mask = out["im"][..., :3].float().numpy(), out["im"][..., 3:].float().numpy()
    # a handful of edge pixels may flip face between the two vertex-shader roundings
    frac_bad = (np.abs(color - d["color"]).max(-1) > 1e-4).mean()
    assert frac_bad < 2e-3, frac_bad
    assert (np.abs(mask - d["mask"]) > 1e-4).mean() < 2e-3
    assert (np.abs(out["prob"].float().numpy() - d["prob"]) > 1e-4).mean() < 2e-3


@pytest.mark.skipif(not os.path.isdir(O.REFERENCE_ROOT), reason="reference tree only exists in the build container")
def test_reference_python_runs_on_the_oracle_stub():
    """the reference's own VCRenderBatch + LinearRasterizer + autograd, driven through the stub, reproduce
    the committed fixture (guards against the fixture and the oracle drifting apart)."""
    import warnings
    warnings.filterwarnings("ignore")
    ref = O.import_reference()
    d, meshes = Hh.load_golden("ref_multi64.npz")
    H, W = int(d["H"]), int(d["W"])
    Rs = torch.tensor(d["Rs"], requires_grad=True)
    ts = torch.tensor(d["ts"], requires_grad=True)
    cams = O.camera_params_from_RT_K(Rs, ts, torch.tensor(d["K"]), H, W, near=0.01, far=100.0)
    models = [{k: torch.tensor(v) for k, v in m.items()} for m in meshes]
    ren = ref["VCRenderMulti"](H, W)
    color, prob, _, mask = ren([[m["vertices"][None], m["faces"].long()] for m in models], cams,
                               [m["colors"][None] for m in models])
    assert np.array_equal(color.detach().numpy(), d["color"]) and np.array_equal(prob.detach().numpy(), d["prob"])
    ((color * torch.tensor(d["g_color"])).sum() + (prob * torch.tensor(d["g_prob"])).sum()).backward()
    assert np.allclose(Rs.grad.numpy(), d["grad_Rs"], rtol=1e-5, atol=1e-5)
    assert np.allclose(ts.grad.numpy(), d["grad_ts"], rtol=1e-5, atol=1e-5)
