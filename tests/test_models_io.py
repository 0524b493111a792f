"""Model readers and the model cache against what the reference's own readers
returned for the same files (tests/golden/ref_models.npz, made by tests/golden/make_golden_models.py from
lib/pysixd/inout.py:489 load_ply and lib/dr_utils/rep/Mesh.py:186 from_obj).  Host logic only: runs on CPU."""
import os
import time

import numpy as np
import pytest
import torch

from self6dpp_b200 import models as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_models.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def _write(tmp_path, name, data):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(bytes(data))
    return p


@pytest.mark.parametrize("tag", ["ascii_full", "ascii_plain", "binary_f32", "binary_f64"])
def test_load_ply_matches_reference(gold, tmp_path, tag):
    p = _write(tmp_path, tag + ".ply", gold["ply/%s/bytes" % tag])
    got = M.load_ply(p, vertex_scale=0.001)
    want = {k.split("/")[-1]: v for k, v in gold.items() if k.startswith("ply/%s/" % tag) and not k.endswith("bytes")}
    assert set(got) == set(want)
    for k, v in want.items():
        if k == "texture_file":
            assert got[k] == str(v)
            continue
        assert got[k].dtype == np.float64 and got[k].shape == v.shape, k
        # the reference goes value -> python float -> float64, so do we: exact
        np.testing.assert_array_equal(got[k], v, err_msg=k)


@pytest.mark.parametrize("style", ["slash", "dslash", "plain"])
def test_load_obj_matches_reference(gold, tmp_path, style):
    p = _write(tmp_path, style + ".obj", gold["obj/%s/bytes" % style])
    got = M.load_obj(p)
    np.testing.assert_array_equal(got["vertices"].numpy(), gold["obj/%s/vertices" % style])
    np.testing.assert_array_equal(got["faces"].numpy(), gold["obj/%s/faces" % style])
    assert got["faces"].dtype == torch.int64
    for k in ("uvs", "face_textures"):
        key = "obj/%s/%s" % (style, k)
        if key in gold:
            np.testing.assert_array_equal(got[k].numpy(), gold[key])
        else:
            assert got[k] is None


def test_load_ply_rejects_non_triangles(tmp_path):
    txt = "\n".join(["ply", "format ascii 1.0", "element vertex 4", "property float x", "property float y",
                     "property float z", "element face 1", "property list uchar int vertex_indices",
                     "end_header", "0 0 0", "1 0 0", "1 1 0", "0 1 0", "4 0 1 2 3"]) + "\n"
    p = _write(tmp_path, "quad.ply", txt.encode())
    with pytest.raises(ValueError):
        M.load_ply(p)


def test_load_ply_truncated(gold, tmp_path):
    data = bytes(gold["ply/binary_f32/bytes"])[:-7]
    with pytest.raises(ValueError):
        M.load_ply(_write(tmp_path, "cut.ply", data))


def test_load_ply_models_layout(gold, tmp_path):
    """renderer_dibr.py:58-72: global-middle centring, colours split off, int32 faces, texture dict keys."""
    p = _write(tmp_path, "textured.obj", gold["obj/slash/bytes"])
    v = torch.from_numpy(gold["obj/slash/vertices"])
    cv2 = pytest.importorskip("cv2")
    tex = (np.random.default_rng(0).random((12, 10, 3)) * 255).astype(np.uint8)
    tp = str(tmp_path / "texture_map.png")
    cv2.imwrite(tp, tex)
    (m,) = M.load_ply_models([p], texture_paths=[tp], device="cpu")
    mid = (v[:, :3].max() + v[:, :3].min()) / 2.0
    assert torch.equal(m["vertices"], v[:, :3] - mid)
    assert torch.equal(m["colors"], v[:, 3:6])
    assert m["faces"].dtype == torch.int32
    assert torch.equal(m["faces"].long(), torch.from_numpy(gold["obj/slash/faces"]))
    assert torch.equal(m["face_uv_ids"], torch.from_numpy(gold["obj/slash/face_textures"]))
    assert m["texture_uv"] is None
    want = torch.from_numpy(tex[:, :, ::-1].astype(np.float32).transpose(2, 0, 1) / 255.0)
    assert torch.equal(m["texture"], want)
    (r,) = M.load_ply_models([p], texture_paths=[tp], device="cpu", tex_resize=True, width=6, height=4)
    assert tuple(r["texture"].shape) == (3, 4, 6)
    (plain,) = M.load_ply_models([p], device="cpu")
    assert "texture" not in plain
    with pytest.raises(AssertionError):
        M.load_ply_models([str(tmp_path / "a.ply")], device="cpu")


def test_model_cache_roundtrip_and_invalidation(gold, tmp_path):
    a = _write(tmp_path, "obj_000001.ply", gold["ply/binary_f32/bytes"])
    b = _write(tmp_path, "obj_000005.ply", gold["ply/ascii_full/bytes"])
    cache = M.ModelCache(str(tmp_path / "models.npz"))
    first = cache.load({"ape": a, "can": b}, vertex_scale=0.001)
    assert os.path.exists(cache.cache_path)
    stamp = os.path.getmtime(cache.cache_path)
    again = cache.load({"ape": a, "can": b}, vertex_scale=0.001)
    assert os.path.getmtime(cache.cache_path) == stamp  # served from the cache, not rewritten
    for n in first:
        assert set(first[n]) == set(again[n])
        for k in first[n]:
            np.testing.assert_array_equal(first[n][k], again[n][k])
    np.testing.assert_array_equal(first["ape"]["pts"], gold["ply/binary_f32/pts"])
    # a changed source file or scale invalidates the cache
    time.sleep(0.01)
    _write(tmp_path, "obj_000001.ply", gold["ply/binary_f64/bytes"])
    os.utime(a, (time.time() + 5, time.time() + 5))
    changed = cache.load({"ape": a, "can": b}, vertex_scale=0.001)
    np.testing.assert_array_equal(changed["ape"]["pts"], gold["ply/binary_f64/pts"])
    scaled = cache.load({"ape": a, "can": b}, vertex_scale=1.0)
    np.testing.assert_allclose(scaled["ape"]["pts"], gold["ply/binary_f64/pts"] * 1000.0, rtol=1e-12)
    # a corrupt cache is rebuilt, not trusted
    with open(cache.cache_path, "wb") as f:
        f.write(b"not an npz")
    rebuilt = cache.load({"ape": a, "can": b}, vertex_scale=1.0)
    np.testing.assert_array_equal(rebuilt["ape"]["pts"], scaled["ape"]["pts"])


def test_get_dibr_models_renderer(gold, tmp_path):
    """self_engine_utils.py:1333-1380: name lookup by the id in the file name, colour range, float32 dicts."""
    _write(tmp_path, "obj_000001.ply", gold["ply/binary_f32/bytes"])
    _write(tmp_path, "obj_000005.ply", gold["ply/binary_f64/bytes"])
    _write(tmp_path, "readme.txt", b"x")
    sel, ren = M.get_dibr_models_renderer(str(tmp_path), ["can", "ape", "can"], {1: "ape", 5: "can"}, height=64,
                                          width=64, mode="VertexColorBatch", color_range=255, device="cpu")
    assert [tuple(m["vertices"].shape) for m in sel] == [(9, 3)] * 3
    assert all(m[k].dtype == torch.float32 for m in sel for k in ("vertices", "colors", "normals", "faces"))
    np.testing.assert_allclose(sel[1]["colors"].numpy(), gold["ply/binary_f32/colors"] / 255.0, rtol=1e-6)
    np.testing.assert_allclose(sel[0]["vertices"].numpy(), gold["ply/binary_f64/pts"], rtol=1e-6)
    assert ren.dib_ren.mode == "VertexColorBatch"
    assert os.path.exists(str(tmp_path / "models_all_w_name.npz"))
    with pytest.raises(KeyError):
        M.get_dibr_models_renderer(str(tmp_path), ["duck"], {1: "ape", 5: "can"}, height=8, width=8,
                                   mode="VertexColorBatch", device="cpu")


def test_model_registry_fast_path_still_sees_changed_tensors():
    """_ModelRegistry.slots answers an unchanged list of model dicts from its last result (one identity + version look per
    distinct model), and must still notice a geometry tensor modified in place, replaced, or switched to requires_grad."""
    import torch
    from self6dpp_b200.renderer_dibr import _ModelRegistry

    def mk(n):
        return {"vertices": torch.rand(n, 3), "faces": torch.randint(0, n, (2 * n, 3), dtype=torch.int32), "colors": torch.rand(n, 3)}
    ms = [mk(10), mk(12), mk(8)]
    reg = _ModelRegistry()
    batch = [ms[0], ms[2], ms[0], ms[1]]
    a = reg.slots(batch)
    g0 = reg.generation
    assert a.tolist() == [0, 1, 0, 2]
    assert reg.slots(batch) is a and reg.slots(list(batch)) is a and reg.generation == g0      # nothing changed: cached answer
    ms[2]["vertices"].add_(1.0)                                                                  # in place: version counter
    assert reg.slots(batch).tolist() == [0, 1, 0, 2] and reg.generation == g0 + 1
    ms[1]["faces"] = ms[1]["faces"].clone()                                                      # replaced: another object
    reg.slots(batch)
    assert reg.generation == g0 + 2
    ms[0]["vertices"].requires_grad_(True)                                                       # cannot be cached any more
    assert reg.slots(batch) is None
