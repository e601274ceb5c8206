"""`pivotSteps` of the GPU path (simple_simplex.optimize_json_format, the producer behind solver_controller.py:290-363)
against the textbook full-tableau oracle: every displayed cell bit-equal, same 0-based pivot (row, column), same number
of steps -- on the reference's fixtures K1-K10, the mixed >= / = family, ragged fuzz LPs, redundant rows and a truncated
recording.  Plus the workspace-reuse sequences that used to replay a stale CUDA graph."""
import numpy as np
import pytest

from simplex_solver_b200 import native, simple_simplex as ss, workloads as W
from tests.helpers import assert_steps_equal_full_oracle, run_pivot_steps, to_min_form

pytestmark = pytest.mark.gpu


def _check(oracle, A, b, c_user, ops, mx, rule, what, cap=None):
    js = run_pivot_steps(A, b, c_user, ops, mx, rule="bland" if rule else "dantzig")
    cap = cap if cap is not None else max(len(js["pivotSteps"]) - 1, 1)
    full = oracle.full_steps(A, b, to_min_form(c_user, mx), ops, oracle.make_opts(rule=rule), cap=cap)
    assert_steps_equal_full_oracle(js, full, what)
    assert [int(v) for v in full["var_ids"]] == [  # column order of the display: variable ids ascending
        int(n[1:]) - 1 + {"x": 0, "s": len(c_user), "a": len(c_user) + len(b)}[n[0]] for n in js["columns"][:-1]]
    return js, full


@pytest.mark.parametrize("rule", [0, 1])
def test_pivot_steps_equal_full_tableau_oracle_on_reference_fixtures(oracle, golden, rule):
    for name, g in golden["kat"].items():
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(g["problem"])
        js, _ = _check(oracle, A, b, c, ops, mx, rule, name)
        if g["z"] is not None:
            assert abs(js["optimalValue"] - g["z"]) <= 1e-9 * max(1.0, abs(g["z"])), name


def test_wyndor_pivot_steps_are_the_published_iterations():
    """The GPU's `pivotSteps` for the Wyndor LP of the reference's tests (tests/test_visualization_integration.py:38-48)
    against the iterations printed in textbooks (Hillier & Lieberman), not against anything in this repo: the same
    arrays pin the CPU oracle in tests/test_oracle_textbook.py."""
    A = np.array([[1.0, 0.0], [0.0, 2.0], [3.0, 2.0]])
    js = run_pivot_steps(A, np.array([4.0, 12.0, 18.0]), np.array([3.0, 5.0]), np.zeros(3, dtype=np.int8), True)
    published = [
        np.array([[1, 0, 1, 0, 0, 4], [0, 2, 0, 1, 0, 12], [3, 2, 0, 0, 1, 18], [-3, -5, 0, 0, 0, 0]], dtype=float),
        np.array([[1, 0, 1, 0, 0, 4], [0, 1, 0, 0.5, 0, 6], [3, 0, 0, -1, 1, 6], [-3, 0, 0, 2.5, 0, 30]], dtype=float),
        np.array([[0, 0, 1, 1 / 3, -1 / 3, 2], [0, 1, 0, 0.5, 0, 6], [1, 0, 0, -1 / 3, 1 / 3, 2], [0, 0, 0, 1.5, 1, 36]]),
    ]
    pivots = [(None, None), (1, 1), (2, 0)]
    assert len(js["pivotSteps"]) == 3 and js["status"] == 0 and js["optimalValue"] == 36.0
    for step, want, (wr, wc) in zip(js["pivotSteps"], published, pivots):
        assert (step["pivotRowIndex"], step["pivotColIndex"]) == (wr, wc)
        np.testing.assert_allclose(np.array(step["tableau"], dtype=float), want, rtol=0, atol=1e-14)


def test_pivot_steps_equal_full_tableau_oracle_on_mixed_and_fuzz(oracle, golden):
    for k, g in enumerate(golden["mixed"]):
        A = np.array(g["A"], dtype=np.float64).reshape(len(g["b"]), len(g["c"]))
        _check(oracle, A, np.array(g["b"]), np.array(g["c"], dtype=np.float64), np.array(g["ops"], dtype=np.int8),
               g["maximize"], k & 1, f"mixed {k}")
    for k in range(0, 200):
        A, b, c, ops = W.fuzz_lp(k)
        _check(oracle, A, b, c, ops, False, k & 1, f"fuzz {k}")


def test_pivot_steps_redundant_rows_and_truncation(oracle, monkeypatch):
    A = np.array([[1.0, 1.0], [1.0, 1.0], [2.0, 2.0], [1.0, 0.0]])
    b = np.array([4.0, 4.0, 8.0, 3.0])
    ops = np.array([2, 2, 2, 0], dtype=np.int8)
    js, full = _check(oracle, A, b, np.array([1.0, 1.0]), ops, False, 0, "redundant")
    assert (full["basis"] < 0).sum() == 2 and js["status"] == 0
    # a recording shorter than the solve: the first `cap` steps are shown, `truncated` is set, the result is unchanged
    monkeypatch.setattr(ss, "MAX_RECORDED_STEPS", 3)
    A, b, c, ops, mx = W.dense_feasible_lp(40, seed=7)
    js = run_pivot_steps(A, b, c, ops, mx)
    cap = len(js["pivotSteps"]) - 1
    full = oracle.full_steps(A, b, -c, ops, oracle.make_opts(rule=0), cap=cap)
    assert full["n_pivots"] > cap, "the LP must need more pivots than are recorded"
    assert_steps_equal_full_oracle(js, full, "truncated")
    ref = oracle.solve_lp(A, b, -c, ops)
    assert js["truncated"] and js["optimalValue"] == -ref["fun"]


def test_workspace_reuse_does_not_replay_a_stale_graph(oracle):
    """One thread-local workspace serves unrelated problems (thread_solver).  Sequences whose stored tableaux share
    R, C and the tableau address but differ in m / art_base / buffers must each match the oracle."""
    kat = W.known_answer_problems()

    def steps(A, b, c, ops, mx, what):
        _check(oracle, np.asarray(A, float), np.asarray(b, float), np.asarray(c, float), np.asarray(ops, np.int8), mx, 0, what)

    wy = W.problem_dict_to_arrays(kat["K1_wyndor"])[:5]
    # (a) Wyndor (all <=: R = 4, C = 3) then '=' + '<=' in 2 variables (m = 2, two objective rows: R = 4, C = 3)
    for _ in range(2):
        steps(*wy, "wyndor")
        steps([[1, 1], [2, 1]], [10, 15], [1, 1], [2, 0], True, "K5-shaped two-phase after wyndor")
    # (b) an infeasible LP leaves its phase-1 graph behind; same m and C, larger n follows (smaller art_base inherited)
    steps([[1, 0], [1, 0]], [5, 10], [1, 0], [0, 1], True, "infeasible: n = 2, one >= row, C = 4")
    steps([[1, 1, 0], [1, 0, 1]], [4, 6], [1, 2, 1], [0, 2], True, "n = 3, one = row, C = 4")
    # (c) a taller problem in between reallocates label / column buffers
    steps(*wy, "wyndor again")
    A, b, c, ops, mx = W.dense_feasible_lp(24, seed=3)
    steps(A, b, c, ops, mx, "taller")
    steps(*wy, "wyndor after taller")
    # and the same through solve_dense with explicit graph loops on one Solver
    s = native.Solver(0)
    for loop in (native.LOOP_GRAPH, native.LOOP_BLOCKED):
        for (A, b, c, ops, mx) in (wy, (np.array([[1.0, 1], [2, 1]]), np.array([10.0, 15]), np.array([1.0, 1]),
                                        np.array([2, 0], np.int8), True), wy):
            got = s.solve_dense(A, b, -c if mx else c, ops, native.make_opts(loop_mode=loop), hist_cap=16)
            ref = oracle.solve_lp(A, b, -c if mx else c, ops, hist_cap=16)
            assert got["status"] == ref["status"] and got["fun"] == ref["fun"]
            assert np.array_equal(got["piv_row"], ref["piv_row"]) and np.array_equal(got["enter_lab"], ref["enter_lab"])
    s.close()
