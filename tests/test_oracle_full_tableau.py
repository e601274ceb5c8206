"""The two CPU restatements check each other: the textbook FULL-tableau simplex (orc_full_steps: explicit column per
variable, no labels) against the condensed (Tucker) oracle whose tableau is expanded for display by the product's
`expand_full` -- same pivots, same displayed cells bit for bit, on the reference's fixtures K1-K10, the mixed-operator
family, the ragged fuzz family and redundant-equality cases.  The full-tableau oracle is what the GPU's `pivotSteps`
(solver_controller.py:332-362) are compared with in tests/test_gpu_pivot_steps.py."""
import numpy as np
import pytest

from simplex_solver_b200 import workloads as W
from simplex_solver_b200.simple_simplex import expand_full
from tests.helpers import assert_bit_equal, to_min_form

CAP = 512


def _compare(O, A, b, cmin, ops, rule, what):
    m = len(b)
    full = O.full_steps(A, b, cmin, ops, O.make_opts(rule=rule), cap=CAP)
    ref = O.solve_lp(A, b, cmin, ops, O.make_opts(rule=rule), hist_cap=CAP)
    assert full["status"] == ref["status"], what
    assert full["n_pivots"] == ref["n_pivots"] and full["n_phase1"] == ref["n_phase1"], what
    t = O.OracleTableau.build(A, b, cmin, ops)
    var_ids = sorted(set(int(v) for v in t.rowlab[:m]) | set(int(v) for v in t.collab[:-1]))
    assert list(full["var_ids"]) == var_ids, what
    where = {v: i for i, v in enumerate(var_ids)}
    assert_bit_equal(expand_full(t.T.copy(), t.rowlab.copy(), t.collab.copy(), var_ids, m), full["steps"][0][0], what)
    for k in range(min(ref["n_pivots"], CAP)):
        r, s, enter = int(ref["piv_row"][k]), int(ref["piv_col"][k]), int(ref["enter_lab"][k])
        t.pivot(r, s)
        Tk, fr, fc = full["steps"][k + 1]
        assert (fr, fc) == (r, where[enter]), f"{what} step {k + 1}"
        assert_bit_equal(expand_full(t.T.copy(), t.rowlab.copy(), t.collab.copy(), var_ids, m), Tk, f"{what} step {k + 1}")
    return full, ref


@pytest.mark.parametrize("rule", [0, 1])
def test_full_tableau_oracle_equals_condensed_oracle_on_reference_fixtures(oracle, golden, rule):
    for name, g in golden["kat"].items():
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(g["problem"])
        full, _ = _compare(oracle, A, b, to_min_form(c, mx), ops, rule, name)
        assert full["status"] == g["scipy_status"], name
        if g["z"] is not None:  # last cell of the objective row of the last displayed tableau = -c'x of the min form
            z = -full["steps"][-1][0][len(b), -1]
            z = -z if mx else z
            assert abs(z - g["z"]) <= 1e-9 * max(1.0, abs(g["z"])), name


def test_full_tableau_oracle_on_mixed_and_fuzz_families(oracle, golden):
    for k, g in enumerate(golden["mixed"]):
        A = np.array(g["A"], dtype=np.float64).reshape(len(g["b"]), len(g["c"]))
        cmin = to_min_form(g["c"], g["maximize"])
        for rule in (0, 1):
            full, _ = _compare(oracle, A, np.array(g["b"]), cmin, np.array(g["ops"], dtype=np.int8), rule, f"mixed {k}")
            assert full["status"] == g["scipy_status"]
    for k in range(400):
        A, b, c, ops = W.fuzz_lp(k)
        _compare(oracle, A, b, c, ops, k & 1, f"fuzz {k}")


def test_full_tableau_oracle_redundant_rows_and_limit(oracle):
    # x1 + x2 = 4 stated twice and once scaled: two redundant rows keep their artificial (flagged), z* = 4
    A = np.array([[1.0, 1.0], [1.0, 1.0], [2.0, 2.0], [1.0, 0.0]])
    b = np.array([4.0, 4.0, 8.0, 3.0])
    ops = np.array([2, 2, 2, 0], dtype=np.int8)
    full, ref = _compare(oracle, A, b, np.array([1.0, 1.0]), ops, 0, "redundant")
    assert full["status"] == 0 and (full["basis"] < 0).sum() == 2
    # explicit budget: LIMIT after 1 pivot, one recorded step
    A, b, c, ops, mx = W.dense_feasible_lp(12, seed=5)
    full = oracle.full_steps(A, b, -c, ops, oracle.make_opts(rule=0, max_pivots=1), cap=8)
    assert full["status"] == 1 and full["n_pivots"] == 1 and len(full["steps"]) == 2
    # recording cap smaller than the solve: truncated, decisions unaffected
    f2 = oracle.full_steps(A, b, -c, ops, oracle.make_opts(rule=0), cap=3)
    f3 = oracle.full_steps(A, b, -c, ops, oracle.make_opts(rule=0), cap=CAP)
    assert f2["truncated"] and len(f2["steps"]) == 4 and f2["n_pivots"] == f3["n_pivots"]
    assert_bit_equal(f2["steps"][3][0], f3["steps"][3][0])
