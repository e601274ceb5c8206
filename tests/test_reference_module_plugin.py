"""The plugin on the REAL reference module.  `/root/reference` exists only in the build container, so these tests skip
on the GPU box; what they establish there is carried to the GPU by the committed fixture
tests/golden/reference_strings.json (the exact strings the reference's `_run_simple_simplex` emits,
solver_controller.py:297-318, recorded by this file's `python tests/test_reference_module_plugin.py --write`).

The reference is imported unmodified with gilp / simple_simplex / reportlab stubbed in sys.modules (they are not
installable offline; same stubs as tests/golden/make_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from simplex_solver_b200 import plugin, simple_simplex as ss, workloads as W  # noqa: E402

REFERENCE = "/root/reference"
FIXTURE = os.path.join(ROOT, "tests", "golden", "reference_strings.json")
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "app")), reason="needs /root/reference")


def _reference_module():
    """Import the reference controller once; sys.path is restored afterwards (the reference has a `tests` package of
    its own that must not shadow this repo's in processes spawned later)."""
    saved = list(sys.path)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden
        return make_golden.import_reference_controller()
    finally:
        sys.path[:] = saved


def _problems():
    out = dict(W.known_answer_problems())
    # values whose str(float) is not plain decimal: 1e-05, 1e+16, negative zero, integers given as int
    out["X1_repr"] = {"problema_definicion": {
        "funcion_objetivo": {"type": "minimize", "coefficients": {"x1": 1e-05, "x2": 2.5e16, "x10": 3}},
        "restricciones": [{"coefficients": {"x1": 1e-05, "x2": -0.0}, "operator": ">=", "rhs": 1e-07},
                          {"coefficients": {"x10": 7, "x2": 1.5}, "operator": "=", "rhs": 12},
                          {"coefficients": {"x1": 1, "x2": 1, "x10": 1}, "operator": "<=", "rhs": 1e+22}]}}
    return out


def _record_strings(sc, wrapper):
    """Run the reference's own _run_simple_simplex with recording stand-ins for the four simple_simplex names."""
    calls = {"constraints": [], "objective": None}
    saved = {k: getattr(sc, k) for k in ("create_tableau", "add_constraint", "add_objective", "optimize_json_format")}
    sc.create_tableau = lambda number_of_variables, number_of_constraints: calls.update(
        n=number_of_variables, m=number_of_constraints) or calls
    sc.add_constraint = lambda t, s: calls["constraints"].append(s)
    sc.add_objective = lambda t, s: calls.update(objective=s)
    sc.optimize_json_format = lambda t, maximize: calls.update(maximize=bool(maximize)) or {"pivotSteps": []}
    try:
        sc.SolverController(wrapper)._run_simple_simplex()
    finally:
        for k, v in saved.items():
            setattr(sc, k, v)
    return calls


@needs_reference
def test_install_patches_the_real_module_and_uninstall_restores_it():
    sc = _reference_module()
    before = {k: getattr(sc, k) for k in ("linprog", "create_tableau", "add_constraint", "add_objective",
                                           "optimize_json_format")}
    assert plugin.install() is sc  # default argument: imports app.controllers.solver_controller by name
    try:
        from simplex_solver_b200.linprog import linprog
        assert sc.linprog is linprog and sc.create_tableau is ss.create_tableau
        assert sc.add_constraint is ss.add_constraint and sc.add_objective is ss.add_objective
        assert sc.optimize_json_format is ss.optimize_json_format
        # the names are looked up at call time inside the reference's methods (module globals), so the patch is live
        assert sc.SolverController._run_simple_simplex.__globals__["optimize_json_format"] is ss.optimize_json_format
        assert sc.SolverController.run.__globals__["linprog"] is linprog
    finally:
        plugin.uninstall(sc)
    assert all(getattr(sc, k) is v for k, v in before.items())


@needs_reference
def test_strings_the_reference_emits_parse_and_match_the_fixture():
    """What the reference really hands to add_constraint / add_objective (str(float) such as '1e-05') goes through the
    product's parsers with the reference's own call signature (keyword arguments, maximize=)."""
    sc = _reference_module()
    with open(FIXTURE) as f:
        fixture = json.load(f)
    for name, wrapper in _problems().items():
        calls = _record_strings(sc, wrapper)
        assert fixture[name] == {"n": calls["n"], "m": calls["m"], "constraints": calls["constraints"],
                                 "objective": calls["objective"], "maximize": calls["maximize"]}, name
        # feed them to the product's string API exactly as the reference would after plugin.install()
        plugin.install(sc)
        try:
            t = sc.create_tableau(number_of_variables=calls["n"], number_of_constraints=calls["m"])
            for s in calls["constraints"]:
                sc.add_constraint(t, s)
            sc.add_objective(t, calls["objective"])
        finally:
            plugin.uninstall(sc)
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(wrapper)
        assert [r[0] for r in t["rows"]] == A.tolist() and [r[2] for r in t["rows"]] == b.tolist(), name
        assert [r[1] for r in t["rows"]] == ops.tolist() and t["objective"] == c.tolist(), name
        assert t["maximize_flag"] == mx == calls["maximize"], name


def test_fixture_strings_parse_without_the_reference():
    """Runs everywhere (also on the GPU box): the committed strings parse to the arrays of the same problems."""
    with open(FIXTURE) as f:
        fixture = json.load(f)
    for name, wrapper in _problems().items():
        fx = fixture[name]
        t = ss.create_tableau(number_of_variables=fx["n"], number_of_constraints=fx["m"])
        for s in fx["constraints"]:
            ss.add_constraint(t, s)
        ss.add_objective(t, fx["objective"])
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(wrapper)
        assert np.array_equal(np.array([r[0] for r in t["rows"]]).reshape(A.shape), A), name
        assert t["objective"] == c.tolist() and t["maximize_flag"] == mx


@pytest.mark.gpu
def test_fixture_strings_through_the_gpu_producer_equal_the_oracle(oracle):
    from tests.helpers import assert_steps_equal_full_oracle, to_min_form
    with open(FIXTURE) as f:
        fixture = json.load(f)
    for name, wrapper in _problems().items():
        fx = fixture[name]
        t = ss.create_tableau(number_of_variables=fx["n"], number_of_constraints=fx["m"])
        for s in fx["constraints"]:
            ss.add_constraint(t, s)
        ss.add_objective(t, fx["objective"])
        js = ss.optimize_json_format(t, maximize=fx["maximize"])
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(wrapper)
        full = oracle.full_steps(A, b, to_min_form(c, mx), ops, oracle.make_opts(rule=0),
                                 cap=max(len(js["pivotSteps"]) - 1, 1))
        assert_steps_equal_full_oracle(js, full, name)


if __name__ == "__main__" and "--write" in sys.argv:
    sc = _reference_module()
    out = {}
    for name, wrapper in _problems().items():
        c = _record_strings(sc, wrapper)
        out[name] = {"n": c["n"], "m": c["m"], "constraints": c["constraints"], "objective": c["objective"],
                     "maximize": c["maximize"]}
    with open(FIXTURE, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {FIXTURE}")
