"""The C-ABI library loads on a machine without a GPU and exports every symbol include/dgx.h declares; argument
validation that happens before any CUDA call reports the reference's error texts. CPU only (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dgx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgx_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from galaexi_b200 import dg
    if not os.path.exists(dg.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return dg.load_library()


def test_exports_match_header(lib):
    from galaexi_b200 import dg
    decl = _declared()
    assert len(decl) >= 16
    for sym in decl:
        assert hasattr(lib, sym), f"{sym} declared in include/dgx.h but not exported by libdgx.so"
    assert sorted(dg.EXPORTS) == decl


def test_config_struct_mirror(lib):
    from galaexi_b200 import dg
    assert lib.dgx_sizeof_config() == C.sizeof(dg.DgxConfig)


@pytest.mark.parametrize("field,value,msg", [
    ("N", 0, "polynomial degree"),
    ("N", 12, "polynomial degree"),
    ("nodeType", 3, "nodeType"),
    ("riemann", 8, "Riemann solver"),
    ("riemann", 10, "Riemann solver"),
    ("splitDG", 7, "SplitDG variant"),
])
def test_create_rejects_bad_config_before_touching_the_gpu(lib, field, value, msg):
    from galaexi_b200 import dg
    c = dg.DgxConfig()
    c.N, c.nodeType, c.splitDG, c.riemann, c.nRKStages = 3, 2, 4, 3, 5
    setattr(c, field, value)
    h = C.c_void_p()
    rc = lib.dgx_create(C.byref(h), C.byref(c))
    assert rc != 0 and h
    assert msg in lib.dgx_last_error(h).decode()
    lib.dgx_destroy(h)


def test_split_dg_needs_gauss_lobatto_and_no_hllc(lib):
    """splitflux.f90:116-119 and src/CMakeLists.txt:113-117."""
    from galaexi_b200 import dg
    for nt, riem, msg in ((1, 3, "Gauss-Lobatto-Points are mandatory"), (2, 5, "HLL-type Riemann solvers are not supported for SPLIT_DG=ON")):
        c = dg.DgxConfig()
        c.N, c.nodeType, c.splitDG, c.riemann, c.nRKStages = 3, nt, 4, riem, 5
        h = C.c_void_p()
        assert lib.dgx_create(C.byref(h), C.byref(c)) != 0
        assert msg in lib.dgx_last_error(h).decode()
        lib.dgx_destroy(h)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under galaexi_b200/ may import, load or link it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|dgoracle|dg_oracle|oracle/_ref", re.M)
    for dp, _, fns in os.walk(os.path.join(ROOT, "galaexi_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                assert not pat.search(open(os.path.join(dp, fn)).read()), (dp, fn)
