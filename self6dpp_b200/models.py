"""Mesh model loading and the model cache in front of ``Renderer_dibr``.

Host-side mirror of
  * ``lib/pysixd/inout.py:489-700``            ``load_ply``   (BOP PLY models: ascii and binary_little_endian)
  * ``lib/dr_utils/rep/Mesh.py:186-267``       ``Mesh.from_obj`` (Wavefront OBJ: v / vt / f lines)
  * ``lib/dr_utils/dib_renderer_x/renderer_dibr.py:20-92``   ``load_ply_models``
  * ``core/self6dpp/engine/self_engine_utils.py:1333-1380`` ``my_get_DIBR_models_renderer``

Same dictionary keys, dtypes and value conventions as the reference so a model list built here can be
handed to either renderer.  What differs is how the work is done: the reference walks every vertex and
every face in a Python loop with one ``struct.unpack`` per property (seconds per BOP model, every
process start unless a pickle happens to exist); here the element tables are read in one
``numpy.frombuffer`` / ``numpy.loadtxt`` call with a structured dtype built from the header, the cache
is an ``.npz`` keyed by the source files' size+mtime (a stale cache is rebuilt, the reference's pickle is
trusted forever), and the device tensors are made once in the layout the CUDA path reads (int32 faces;
``Renderer_dibr`` pads vertices to 16 B rows itself).

Nothing here touches the GPU kernels; it only feeds them.
"""
import os
import os.path as osp
import re

import numpy as np
import torch

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1",
    "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def _ply_header(f):
    """Parse the PLY header.  Returns (format, [(element, count, [(kind, name, types...)])], texture_file)."""
    first = f.readline().strip()
    if first != b"ply":
        raise ValueError("not a PLY file")
    fmt = None
    texture_file = None
    elements = []
    while True:
        raw = f.readline()
        if not raw:
            raise ValueError("PLY header is not terminated by end_header")
        line = raw.decode("utf-8", "replace").strip()
        if line.startswith("comment TextureFile"):
            texture_file = line.split()[-1]
        elif line.startswith("format"):
            fmt = line.split()[1]
        elif line.startswith("element"):
            _, name, count = line.split()[:3]
            elements.append((name, int(count), []))
        elif line.startswith("property list"):
            p = line.split()
            elements[-1][2].append(("list", p[-1], p[2], p[3]))
        elif line.startswith("property"):
            p = line.split()
            elements[-1][2].append(("scalar", p[-1], p[-2]))
        elif line.startswith("end_header"):
            break
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError("unsupported PLY format: {}".format(fmt))
    return fmt, elements, texture_file


def _face_dtype(props, order):
    """Structured dtype of one triangular face row (every list is fixed-length: 3 indices / 6 uvs)."""
    fields = []
    for p in props:
        if p[0] == "scalar":
            fields.append((p[1], order + _PLY_TYPES[p[2]]))
            continue
        _, name, cnt_t, val_t = p
        if name in ("vertex_indices", "vertex_index"):
            fields.append(("n_corners", order + _PLY_TYPES[cnt_t]))
            fields.append(("ind", order + _PLY_TYPES[val_t], (3,)))
        elif name == "texcoord":
            fields.append(("n_texcoord", order + _PLY_TYPES[cnt_t]))
            fields.append(("texcoord", order + _PLY_TYPES[val_t], (6,)))
        else:
            raise ValueError("Not supported face property: " + name)
    return np.dtype(fields)


def load_ply(path, vertex_scale=1.0):
    """Loads a 3D mesh model from a PLY file (``lib/pysixd/inout.py:489``).

    Returns a dict with 'pts' (n,3), and when present 'normals' (n,3), 'colors' (n,3), 'faces' (m,3),
    'texture_uv' (n,2), 'texture_uv_face' (m,6), 'texture_file' -- all float64 arrays like the reference
    (faces too: the reference allocates them with ``np.float``).  ``pts`` are multiplied by ``vertex_scale``.
    Only triangular faces are supported (``ValueError`` otherwise, as in the reference).
    """
    with open(path, "rb") as f:
        fmt, elements, texture_file = _ply_header(f)
        body = f.read()
    order = ">" if fmt == "binary_big_endian" else "<"
    tables = {}
    off = 0
    lines = None
    line_at = 0
    if fmt == "ascii":
        lines = body.decode("utf-8", "replace").split("\n")
        lines = [ln for ln in lines if ln.strip()]
    for name, count, props in elements:
        if name == "face":
            dt = _face_dtype(props, order)
        else:
            if any(p[0] == "list" for p in props):
                if count == 0:
                    continue
                raise ValueError("list property in element '{}' is not supported".format(name))
            dt = np.dtype([(p[1], order + _PLY_TYPES[p[2]]) for p in props])
        if fmt == "ascii":
            rows = lines[line_at:line_at + count]
            line_at += count
            if len(rows) != count:
                raise ValueError("PLY file is truncated")
            width = sum(int(np.prod(dt[n].shape)) if dt[n].shape else 1 for n in dt.names)
            flat = np.array([r.split()[:width] for r in rows], dtype=np.float64).reshape(count, width)
            tab = {}
            c = 0
            for n in dt.names:
                w = int(np.prod(dt[n].shape)) if dt[n].shape else 1
                tab[n] = flat[:, c] if not dt[n].shape else flat[:, c:c + w]
                c += w
            tables[name] = tab
        else:
            nbytes = dt.itemsize * count
            if off + nbytes > len(body):
                raise ValueError("PLY file is truncated")
            arr = np.frombuffer(body, dtype=dt, count=count, offset=off)
            off += nbytes
            tables[name] = {n: arr[n] for n in dt.names}

    model = {}
    if texture_file is not None:
        model["texture_file"] = texture_file
    v = tables.get("vertex", {})
    n_pts = next((c for n, c, _ in elements if n == "vertex"), 0)
    names = set(v.keys())
    # the reference renames s/t to texture_u/texture_v (inout.py:547-550)
    if "s" in names:
        v["texture_u"] = v["s"]
    if "t" in names:
        v["texture_v"] = v["t"]
    names = set(v.keys())

    def cols(keys):
        return np.stack([np.asarray(v[k], dtype=np.float64) for k in keys], axis=1) if n_pts else \
            np.zeros((0, len(keys)), np.float64)

    model["pts"] = cols(("x", "y", "z"))
    fc = tables.get("face")
    n_faces = next((c for n, c, _ in elements if n == "face"), 0)
    if n_faces > 0:
        if np.any(np.asarray(fc["n_corners"]) != 3):
            raise ValueError("Only triangular faces are supported.")
        model["faces"] = np.asarray(fc["ind"], dtype=np.float64).reshape(n_faces, 3)
    if {"nx", "ny", "nz"} <= names:
        model["normals"] = cols(("nx", "ny", "nz"))
    if {"red", "green", "blue"} <= names:
        model["colors"] = cols(("red", "green", "blue"))
    if {"texture_u", "texture_v"} <= names:
        model["texture_uv"] = cols(("texture_u", "texture_v"))
    if fc is not None and "texcoord" in fc:
        if n_faces and np.any(np.asarray(fc["n_texcoord"]) != 6):
            raise ValueError("Wrong number of UV face coordinates.")
        model["texture_uv_face"] = np.asarray(fc["texcoord"], dtype=np.float64).reshape(n_faces, 6)
    model["pts"] = model["pts"] * vertex_scale
    return model


def load_obj(path):
    """Wavefront OBJ reader with the semantics of ``Mesh.from_obj`` (``lib/dr_utils/rep/Mesh.py:213-267``).

    Returns a dict of torch CPU tensors:
      'vertices'       float32 [n, 3] or [n, 6] (x y z [r g b] -- whatever follows ``v``),
      'faces'          int64 [m, k] zero-based,
      'uvs'            float32 [t, 2] or None,
      'face_textures'  int64 [m, k] zero-based or None.
    ``f a/b/c`` and ``f a//c`` take the SECOND field as the texture index (the reference does the same for
    both spellings); ``f a`` lines contribute no texture index.
    """
    vertices, faces, face_textures, uvs = [], [], [], []
    with open(path, "r") as fh:
        for line in fh:
            data = line.split()
            if not data:
                continue
            tag = data[0]
            if tag == "v":
                vertices.append([float(d) for d in data[1:]])
            elif tag == "vt":
                uvs.append([float(d) for d in data[1:]])
            elif tag == "f":
                if "//" in data[1]:
                    parts = [d.split("//") for d in data[1:]]
                    faces.append([int(p[0]) for p in parts])
                    face_textures.append([int(p[1]) for p in parts])
                elif "/" in data[1]:
                    parts = [d.split("/") for d in data[1:]]
                    faces.append([int(p[0]) for p in parts])
                    face_textures.append([int(p[1]) for p in parts])
                else:
                    faces.append([int(d) for d in data[1:]])
    out = {
        "vertices": torch.from_numpy(np.array(vertices, dtype=np.float32)),
        "faces": torch.from_numpy(np.array(faces, dtype=np.int64)) - 1,
        "uvs": None,
        "face_textures": None,
    }
    if uvs:
        out["uvs"] = torch.from_numpy(np.array([u for row in uvs for u in row], dtype=np.float32)).view(-1, 2)
    if face_textures:
        out["face_textures"] = torch.from_numpy(np.array(face_textures, dtype=np.int64)) - 1
    return out


def _read_texture(path, width, height, tex_resize):
    """BGR->RGB float32 [0,1] CHW; INTER_AREA resize when asked (renderer_dibr.py:76-80)."""
    import cv2  # only the textured path needs it

    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    tex = img[:, :, ::-1].astype(np.float32) / 255.0
    if tex_resize:
        tex = cv2.resize(tex, (width, height), interpolation=cv2.INTER_AREA)
    return torch.from_numpy(np.ascontiguousarray(tex.transpose(2, 0, 1)))


def load_ply_models(obj_paths, texture_paths=None, vertex_scale=0.001, device="cuda", width=512, height=512,
                    tex_resize=False):
    """``renderer_dibr.py:20-92``: a list of model dicts from ``.obj`` files (the name is the reference's).

    Each dict: 'vertices' [n,3] centred on the middle of the GLOBAL min/max coordinate (one scalar, as the
    reference does), 'colors' [n,3], 'faces' int32 [m,3]; with ``texture_paths`` also 'face_uvs',
    'face_uv_ids', 'texture' (CHW RGB in [0,1]) and 'texture_uv' = None.  ``vertex_scale`` is accepted and
    unused, like the reference.  The reference's dataset-specific pickle short-cut (``models_s6dpp.pkl`` +
    ``LM_DICT``) is replaced by ``ModelCache`` below.
    """
    if not all(".obj" in p for p in obj_paths):
        raise AssertionError("load_ply_models expects .obj files")
    models = []
    for i, obj_path in enumerate(obj_paths):
        mesh = load_obj(obj_path)
        v = mesh["vertices"]
        vertices, colors = v[:, :3], v[:, 3:6]
        middle = (vertices.max() + vertices.min()) / 2.0
        model = {
            "vertices": (vertices - middle).to(device),
            "colors": colors.contiguous().to(device),
            "faces": mesh["faces"].int().to(device),
        }
        if texture_paths is not None:
            model["face_uvs"] = mesh["uvs"].to(device)
            model["face_uv_ids"] = mesh["face_textures"].to(device)
            model["texture"] = _read_texture(texture_paths[i], width, height, tex_resize).to(device)
            model["texture_uv"] = None
        models.append(model)
    return models


class ModelCache(object):
    """Parsed-model cache: one ``.npz`` next to the models, invalidated by (name, size, mtime) of the sources.

    Replaces the ``mmcv.dump``/``mmcv.load`` pickle of ``self_engine_utils.py:1340-1356`` (which is never
    invalidated and executes arbitrary pickled code on load).
    """

    KEYS = ("pts", "normals", "colors", "faces", "texture_uv", "texture_uv_face")

    def __init__(self, cache_path):
        self.cache_path = cache_path

    @staticmethod
    def _stamp(paths):
        return "|".join("{}:{}:{}".format(osp.basename(p), osp.getsize(p), int(osp.getmtime(p) * 1000))
                        for p in paths)

    def load(self, named_paths, vertex_scale=0.001):
        """``named_paths``: {name: ply_path}.  Returns {name: load_ply dict}."""
        names = sorted(named_paths)
        stamp = self._stamp([named_paths[n] for n in names]) + "|scale={!r}".format(float(vertex_scale))
        if osp.exists(self.cache_path):
            try:
                with np.load(self.cache_path, allow_pickle=False) as z:
                    if str(z["__stamp__"]) == stamp:
                        out = {}
                        for n in names:
                            out[n] = {k: z["{}/{}".format(n, k)] for k in self.KEYS
                                      if "{}/{}".format(n, k) in z.files}
                        return out
            except Exception:
                pass  # unreadable cache: rebuild
        out = {n: load_ply(named_paths[n], vertex_scale=vertex_scale) for n in names}
        flat = {"__stamp__": np.array(stamp)}
        for n in names:
            for k in self.KEYS:
                if k in out[n]:
                    flat["{}/{}".format(n, k)] = out[n][k]
        tmp = self.cache_path + ".tmp.{}.npz".format(os.getpid())
        try:
            np.savez(tmp, **flat)
            os.replace(tmp, self.cache_path)
        except OSError:
            if osp.exists(tmp):
                os.remove(tmp)
        return {n: {k: v for k, v in out[n].items() if k in self.KEYS} for n in names}


def get_dibr_models_renderer(model_dir, obj_names, id2obj, *, height, width, mode, color_range=1,
                             vertex_scale=0.001, device="cuda", cache_name="models_all_w_name.npz"):
    """``my_get_DIBR_models_renderer`` (``self_engine_utils.py:1333-1380``) without the cfg/data_ref objects.

    Scans ``model_dir`` for ``*.ply``, maps the first integer in each file name through ``id2obj`` (the
    reference's ``LM_DICT``), caches the parsed models and returns ``(models_selected, Renderer_dibr)`` where
    each model is {'vertices','colors' (/color_range),'normals','faces'} float32 on ``device`` (faces stay
    float32 like the reference; ``Renderer_dibr`` converts them to int32 once per model, not per call).
    """
    from .renderer_dibr import Renderer_dibr

    ply = sorted(osp.abspath(osp.join(model_dir, f)) for f in os.listdir(model_dir)
                 if osp.splitext(f)[-1] == ".ply")
    named = {}
    for p in ply:
        m = re.search(r"\d+", osp.basename(p))
        if m is None or int(m.group()) not in id2obj:
            continue
        named[id2obj[int(m.group())]] = p
    missing = [n for n in obj_names if n not in named]
    if missing:
        raise KeyError("no PLY model for {}".format(missing))
    models = ModelCache(osp.join(model_dir, cache_name)).load(named, vertex_scale=vertex_scale)
    selected = []
    for name in obj_names:
        m = models[name]
        selected.append({
            "vertices": torch.tensor(m["pts"], device=device, dtype=torch.float32),
            "colors": torch.tensor(m["colors"] / color_range, device=device, dtype=torch.float32),
            "normals": torch.tensor(m["normals"], device=device, dtype=torch.float32),
            "faces": torch.tensor(m["faces"], device=device, dtype=torch.float32),
        })
    return selected, Renderer_dibr(height=height, width=width, mode=mode)
