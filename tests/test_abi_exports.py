"""The C-ABI library loads on a box without a GPU and exports exactly what include/dibr_b200.h declares.

No compute is launched here: only symbol resolution, the struct-size handshakes and the argument validation that
runs before any CUDA call.
"""
import ctypes
import os
import re

import pytest

from self6dpp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dibr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # drop comments
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|long long|size_t|char|void)\s*\*?\s*(dibr_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_something():
    names = declared_symbols()
    assert len(names) >= 16, names
    for must in ("dibr_forward", "dibr_backward_meshes", "dibr_render_step", "dibr_render_forward", "dibr_render_backward", "dibr_nnd_forward"):
        assert must in names


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_list_matches_header():
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_no_undeclared_dibr_exports():
    """`nm -D` view: every exported dibr_* function is declared in the header (the debug phase counter aside)."""
    import subprocess
    path = _lib.lib_path() if hasattr(_lib, "lib_path") else os.path.join(ROOT, "self6dpp_b200", "lib", "libdibr_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("dibr_")}
    extra = exported - set(declared_symbols()) - {"dibr_debug_phase_cycles", "dibr_debug_item_cycles"}
    assert not extra, extra


def test_struct_size_handshake():
    lib = _lib.load()
    assert lib.dibr_sizeof_pass() == ctypes.sizeof(_lib.DibrPass)
    assert lib.dibr_sizeof_step() == ctypes.sizeof(_lib.DibrStep)
    import re
    want = int(re.search(r"#define\s+DIBR_ABI_VERSION\s+(\d+)", open(os.path.join(ROOT, "include", "dibr_b200.h")).read()).group(1))
    assert lib.dibr_abi_version() == want == 3
    # the driver's build check must expect what the header says (it once kept a stale literal)
    src = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "DIBR_ABI_VERSION" in src and not re.search(r"dibr_abi_version\(\)\s*==\s*\d", src)


def test_validation_runs_before_any_cuda_call():
    """A zeroed pass is rejected with a message; nothing touches the device."""
    lib = _lib.load()
    p = _lib.DibrPass()
    n = ctypes.c_size_t(0)
    rc = lib.dibr_workspace_bytes(ctypes.byref(p), ctypes.byref(n))
    assert rc != 0
    lib.dibr_last_error.restype = ctypes.c_char_p
    assert lib.dibr_last_error()


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the operator raises on CPU inputs instead of computing something."""
    import torch
    from self6dpp_b200 import linear_rasterizer
    z = torch.zeros(1, 2, 9)
    with pytest.raises(Exception):
        linear_rasterizer(8, 8, z, torch.zeros(1, 2, 6), torch.ones(1, 2, 1), torch.zeros(1, 2, 9))


def test_pass_templates_copy_the_filled_struct():
    """fused._pass_from_template: the size / option fields are filled once per configuration, later calls get a COPY (their
    pointer fields must not leak into the template) and the cached workspace size (host arithmetic, no GPU)."""
    from self6dpp_b200 import fused
    from self6dpp_b200.rasterizer import _base_pass
    calls = []

    def build():
        calls.append(1)
        q = _base_pass(2, 32, 48, 5, 30, 1000, 7000, 0.02, 100, 0)
        q.num_instances = 2
        q.num_outputs = 2
        q.out_channels[0], q.out_channels[1] = 3, 2
        return q
    key = ("test_pass_templates", 2, 32, 48)
    p1, n1 = fused._pass_from_template(key, build)
    p1.improb = 0x1234                                  # a pointer set by one call ...
    p2, n2 = fused._pass_from_template(key, build)
    assert len(calls) == 1 and n1 == n2 == _lib.workspace_bytes(build())
    assert p2.improb is None and p1 is not p2           # ... is not in the next call's struct
    assert (p2.batch, p2.height, p2.width, p2.num_attr, p2.total_faces, p2.num_outputs, p2.out_channels[1]) == (2, 32, 48, 5, 100, 2, 2)
