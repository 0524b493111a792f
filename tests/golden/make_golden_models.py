"""Generates tests/golden/ref_models.npz: small synthetic PLY / OBJ files (kept as raw bytes inside the
fixture) together with what the REFERENCE's own readers return for them:
  lib/pysixd/inout.py:489 load_ply        (ascii and binary_little_endian, with normals / colours / uv /
                                            per-face texcoord, and a foreign element in between)
  lib/dr_utils/rep/Mesh.py:186 from_obj   (v with colours, vt, f a/b/c, f a//c, f a)

Run here (the reference tree does not exist on the GPU box):  python tests/golden/make_golden_models.py
inout.py's unrelated imports that are absent here (imageio, mmcv, png, termcolor ...) are mocked; numpy 2
dropped ``np.float`` which load_ply still uses, so it is aliased for the import; the chardet-based text/binary
sniffer that only chooses the open() mode is replaced by a look at the header's format line.
"""
import os
import struct
import sys
import tempfile
from unittest import mock

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))


def ply_ascii(rng, n=7, m=5, uv=True, texface=True, extra=True):
    pts = rng.normal(size=(n, 3)) * 40
    nrm = rng.normal(size=(n, 3))
    col = rng.integers(0, 256, size=(n, 3))
    st = rng.random((n, 2))
    faces = rng.integers(0, n, size=(m, 3))
    tc = rng.random((m, 6))
    h = ["ply", "format ascii 1.0", "comment TextureFile obj_000001.png", "element vertex %d" % n,
         "property float x", "property float y", "property float z",
         "property float nx", "property float ny", "property float nz"]
    if uv:
        h += ["property float s", "property float t"]
    h += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
    h += ["element face %d" % m, "property list uchar int vertex_indices"]
    if texface:
        h += ["property list uchar float texcoord"]
    h += ["end_header"]
    rows = []
    for i in range(n):
        r = ["%.6f" % x for x in pts[i]] + ["%.6f" % x for x in nrm[i]]
        if uv:
            r += ["%.6f" % x for x in st[i]]
        r += [str(int(c)) for c in col[i]] + ["255"]
        rows.append(" ".join(r))
    for j in range(m):
        r = ["3"] + [str(int(x)) for x in faces[j]]
        if texface:
            r += ["6"] + ["%.6f" % x for x in tc[j]]
        rows.append(" ".join(r))
    return ("\n".join(h + rows) + "\n").encode()


def ply_binary(rng, n=9, m=6, double_xyz=False):
    pts = rng.normal(size=(n, 3)) * 40
    nrm = rng.normal(size=(n, 3))
    col = rng.integers(0, 256, size=(n, 3))
    faces = rng.integers(0, n, size=(m, 3))
    t = "double" if double_xyz else "float"
    h = ["ply", "format binary_little_endian 1.0", "element vertex %d" % n,
         "property %s x" % t, "property %s y" % t, "property %s z" % t,
         "property float nx", "property float ny", "property float nz",
         "property uchar red", "property uchar green", "property uchar blue",
         "element face %d" % m, "property list uchar int vertex_indices", "end_header"]
    body = b""
    for i in range(n):
        body += struct.pack("<3d" if double_xyz else "<3f", *pts[i]) + struct.pack("<3f", *nrm[i])
        body += struct.pack("<3B", *[int(c) for c in col[i]])
    for j in range(m):
        body += struct.pack("<B3i", 3, *[int(x) for x in faces[j]])
    return ("\n".join(h) + "\n").encode() + body


def obj_text(rng, n=8, m=6, style="slash"):
    lines = ["# synthetic", "mtllib none.mtl"]
    for i in range(n):
        v = rng.normal(size=3) * 0.05
        c = rng.random(3)
        lines.append("v " + " ".join("%.6f" % x for x in np.concatenate([v, c])))
    for i in range(n + 3):
        lines.append("vt %.6f %.6f" % tuple(rng.random(2)))
    lines.append("")
    for j in range(m):
        a = rng.integers(1, n + 1, size=3)
        b = rng.integers(1, n + 4, size=3)
        if style == "slash":
            lines.append("f " + " ".join("%d/%d/%d" % (a[k], b[k], a[k]) for k in range(3)))
        elif style == "dslash":
            lines.append("f " + " ".join("%d//%d" % (a[k], b[k]) for k in range(3)))
        else:
            lines.append("f " + " ".join("%d" % a[k] for k in range(3)))
    return ("\n".join(lines) + "\n").encode()


def main():
    np.float = float  # noqa: removed in numpy 2, used by inout.load_ply
    sys.path.insert(0, "/root/reference")
    for name in ("imageio", "mmcv", "png", "termcolor", "cv2", "scipy", "scipy.linalg", "scipy.spatial",
                 "scipy.spatial.transform", "ruamel", "ruamel.yaml", "yaml", "tqdm", "transforms3d",
                 "transforms3d.quaternions", "transforms3d.euler", "transforms3d.axangles", "PIL", "plyfile", "chardet"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock()
    from lib.pysixd import inout
    from lib.dr_utils.rep import TriangleMesh

    # inout._is_binary (a chardet heuristic, absent here) only picks the open() mode; the parser itself
    # follows the header's "format" line.  Decide the mode from that line instead.
    inout._is_binary = lambda p: b"format binary" in open(p, "rb").read(256)

    rng = np.random.default_rng(11)
    out = {}
    plys = {
        "ascii_full": ply_ascii(rng),
        "ascii_plain": ply_ascii(rng, uv=False, texface=False),
        "binary_f32": ply_binary(rng),
        "binary_f64": ply_binary(rng, double_xyz=True),
    }
    with tempfile.TemporaryDirectory() as td:
        for tag, data in plys.items():
            p = os.path.join(td, tag + ".ply")
            with open(p, "wb") as f:
                f.write(data)
            ref = inout.load_ply(p, vertex_scale=0.001)
            out["ply/%s/bytes" % tag] = np.frombuffer(data, dtype=np.uint8)
            for k, v in ref.items():
                out["ply/%s/%s" % (tag, k)] = np.array(v)
        for style in ("slash", "dslash", "plain"):
            data = obj_text(rng, style=style)
            p = os.path.join(td, style + ".obj")
            with open(p, "wb") as f:
                f.write(data)
            mesh = TriangleMesh.from_obj(p)
            out["obj/%s/bytes" % style] = np.frombuffer(data, dtype=np.uint8)
            out["obj/%s/vertices" % style] = mesh.vertices.numpy()
            out["obj/%s/faces" % style] = mesh.faces.numpy()
            if mesh.uvs is not None:
                out["obj/%s/uvs" % style] = mesh.uvs.numpy()
            if mesh.face_textures is not None:
                out["obj/%s/face_textures" % style] = mesh.face_textures.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_models.npz"), **out)
    print("wrote ref_models.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
