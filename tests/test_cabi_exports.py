"""The C-ABI library loads on a CPU-only box, exports every symbol include/b200lp.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from simplex_solver_b200 import native

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200lp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200lp_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    native.build()
    L = native.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/b200lp.h but not exported"
    assert sorted(native.EXPORTS) == declared


def test_version_and_default_opts():
    L = native.lib()
    assert L.b200lp_version() == 200
    o = native.make_opts()
    assert o.rule == native.RULE_DANTZIG and o.eps_cost == 1e-9 and o.eps_pivot == 1e-9 and o.eps_feas == 1e-7
    assert o.max_pivots == 1 << 40 and o.loop_mode == native.LOOP_AUTO and o.time_limit_s == 0.0
    assert native.make_opts(time_limit=10).time_limit_s == 10.0


def test_struct_layouts_match_the_header():
    # sizes implied by the field lists of include/b200lp.h (LP64)
    assert C.sizeof(native.Opts) == 4 + 4 + 8 + 3 * 8 + 4 + 4 + 8
    assert C.sizeof(native.Problem) == 3 * 8 + 4 * 8 + 4 + 4
    assert C.sizeof(native.Result) == 4 + 4 + 8 + 8 + 8 + 8 + 8 + 4 * 8 + 8 + 8 + 8


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="checks the behaviour of a box without a GPU")
def test_no_gpu_means_a_loud_error_not_a_fallback():
    with pytest.raises(native.B200LPError) as e:
        native.Solver(0)
    assert "no CPU fallback" in str(e.value)
    from simplex_solver_b200.linprog import linprog
    with pytest.raises(native.B200LPError):
        linprog([1.0, 1.0], A_ub=[[1.0, 1.0]], b_ub=[1.0])


def test_null_arguments_are_rejected_without_a_device():
    L = native.lib()
    assert L.b200lp_create(None, 0) == -1
    assert b"NULL" in L.b200lp_last_error()
    assert L.b200lp_destroy(None) == 0


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "simplex_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_integration_stub_matches_the_library_structs():
    """INTEGRATION.md shows the ctypes stub a maintainer would write; its struct fields must be the ones native.py (and
    therefore include/b200lp.h, checked above) declares -- a field missing from b200lp_opts makes the library read past
    the caller's struct."""
    import re
    from simplex_solver_b200 import native
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cls, ref in (("Opts", native.Opts), ("Problem", native.Problem), ("Result", native.Result)):
        m = re.search(r"class %s\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n" % cls, text, re.S)
        assert m, cls
        names = re.findall(r'\("(\w+)"', m.group(1))
        assert names == [f[0] for f in ref._fields_], (cls, names)
