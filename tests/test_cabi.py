"""CPU-only checks of the drop-in boundary: the library builds, loads, exports every symbol the
headers declare, and its structs have the reference's LP64 layout."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from piplib_b200 import build
    return build.build()


def _declared_functions():
    names = set()
    for hdr in ("include/piplib/piplib.h", "include/piplib_b200.h"):
        text = open(os.path.join(ROOT, hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = "\n".join(ln for ln in text.split("\n") if not ln.strip().startswith("#"))
        for m in re.finditer(r"\b(pip_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_exports_every_declared_symbol(libpath):
    lib = C.CDLL(libpath)
    names = _declared_functions()
    assert "pip_solve_dp" in names and "pip_solve_batch_dp" in names and len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_reference():
    # include/piplib/piplib.h:194-329 of the reference on LP64: 32/24/32/16/48/32 bytes
    from piplib_b200 import api
    assert C.sizeof(api.PipMatrix) == 32
    assert C.sizeof(api.PipOptions) == 32

    class V(C.Structure):
        _fields_ = [("n", C.c_int), ("a", C.c_void_p), ("b", C.c_void_p)]
    assert C.sizeof(V) == 24


def test_host_side_objects_without_gpu(libpath):
    """matrix / options allocation and the option defaults never touch the device."""
    from piplib_b200 import api
    L = api.lib()
    m = L.pip_matrix_alloc_dp(3, 4)
    assert m.contents.NbRows == 3 and m.contents.NbColumns == 4 and m.contents.p_Init_size == 12
    assert m.contents.p[2][3] == 0
    L.pip_matrix_free_dp(m)
    o = L.pip_options_init_dp()
    assert (o.contents.Nq, o.contents.Verbose, o.contents.Simplify, o.contents.Maximize) == (1, 0, 0, 0)
    L.pip_options_free_dp(o)
    assert b"sm_100a" in L.pip_b200_version()


def test_no_cpu_solver_in_product():
    """the product tree must not reference the oracle (parity claims depend on it)."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "piplib_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cpp", ".cu")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle/|piporacle|pipref|libpipemu", t) and f != "pip_host.cpp":
                    bad.append(f)
    assert not bad, bad
