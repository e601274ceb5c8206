"""Pins the CPU oracle (oracle/simplex_oracle.c) against the golden vectors produced by the reference's own solve
path -- SolverController.run() with real scipy HiGHS, tests/golden/make_golden.py -- before anything trusts it.
Status must be identical; z* (and x* where the optimum is unique) within 1e-9 relative."""
import numpy as np
import pytest

from simplex_solver_b200 import workloads as W
from tests.helpers import to_min_form, z_from_fun

REL = 1e-9


@pytest.mark.parametrize("rule", [0, 1])
def test_known_answer_problems(oracle, golden, rule):
    assert len(golden["kat"]) == 10
    for name, g in golden["kat"].items():
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(g["problem"])
        r = oracle.solve_lp(A, b, to_min_form(c, mx), ops, oracle.make_opts(rule=rule))
        assert r["status"] == g["scipy_status"], name
        if g["z"] is None:
            continue
        z = z_from_fun(r["fun"], mx)
        assert abs(z - g["z"]) <= REL * max(1.0, abs(g["z"])), name
        if g["x_unique"]:
            np.testing.assert_allclose(r["x"], g["x"], rtol=REL, atol=REL, err_msg=name)


def test_survey_known_values(oracle, golden):
    """The values SURVEY.md 8c lists (including the three places where the reference's own mocks are wrong)."""
    k = golden["kat"]
    assert k["K1_wyndor"]["z"] == 36.0 and k["K1_wyndor"]["x"] == [2.0, 6.0]
    assert abs(k["K3_min_ge"]["z"] - 153.33333333333337) < 1e-9
    assert k["K4_min_ge2"]["z"] == 10.0
    assert k["K6_infeasible"]["status_text"] == "Sin Solucion Factible"
    assert k["K7_unbounded"]["status_text"] == "Error" and k["K7_unbounded"]["scipy_status"] == 3
    assert k["K9_three_var"]["x"] == [0.0, 0.0, 10.0]
    assert abs(golden["dense"]["1024"]["z"] - 447.03288201572923) < 1e-9


@pytest.mark.parametrize("n", [16, 64, 128, 256, 512, 1024])
def test_dense_family_config2(oracle, golden, n):
    A, b, c, ops, mx = W.dense_feasible_lp(n, seed=0)
    r = oracle.solve_lp(A, b, to_min_form(c, mx), ops, oracle.make_opts(rule=0, threads=min(8, oracle.max_threads())))
    zg = golden["dense"][str(n)]["z"]
    assert r["status"] == 0
    assert abs(-r["fun"] - zg) <= REL * abs(zg)
    x = r["x"]
    assert (x >= -1e-9).all() and (A @ x <= b + 1e-7 * np.maximum(1.0, np.abs(b))).all()


def test_batched_family_config3(oracle, golden):
    g = golden["batched"]
    A, b, c, ops = W.batched_small_lps(0, g["count"], g["m"], g["n"], g["base_seed"])
    for rule in (0, 1):
        r = oracle.solve_batched(A, b, c, ops, oracle.make_opts(rule=rule), threads=4)
        for k, row in enumerate(g["results"]):
            assert r["status"][k] == row["scipy_status"], (rule, k)
            if row["z"] is not None:
                assert abs(r["fun"][k] - row["z"]) <= REL * max(1.0, abs(row["z"])), (rule, k)
    seen = {row["scipy_status"] for row in g["results"]}
    assert seen == {0, 2, 3}


def test_mixed_operator_family(oracle, golden):
    for k, g in enumerate(golden["mixed"]):
        A = np.array(g["A"], dtype=np.float64).reshape(len(g["b"]), len(g["c"]))
        r = oracle.solve_lp(A, np.array(g["b"]), to_min_form(g["c"], g["maximize"]), np.array(g["ops"], dtype=np.int8))
        assert r["status"] == g["scipy_status"], k
        if g["z"] is not None:
            z = z_from_fun(r["fun"], g["maximize"])
            assert abs(z - g["z"]) <= REL * max(1.0, abs(g["z"])), k


def test_generators_are_stable(golden):
    """The committed expectations belong to these exact inputs (numpy RNG stream check)."""
    import hashlib

    def sha(*arrays):
        h = hashlib.sha256()
        for a in arrays:
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()[:16]

    A, b, c, ops, _ = W.dense_feasible_lp(128, seed=0)
    assert sha(A, b, c) == golden["dense"]["128"]["inputs_sha"]
    g = golden["batched"]
    A, b, c, ops = W.batched_small_lps(0, g["count"], g["m"], g["n"], g["base_seed"])
    assert sha(A, b, c, ops) == g["inputs_sha"]


def test_pivot_identities(oracle):
    """Properties of one pivot that hold at any size: the pivot column becomes -col/p (1/p at the pivot), the
    pivot row is divided by p, pivoting back on the same cell restores the basis labels."""
    t = oracle.OracleTableau.generate(7, 40, 60)
    T0 = t.T.copy()
    rl0, cl0 = t.rowlab.copy(), t.collab.copy()
    s = t.price(rule=0)
    col = t.extract_col(s)
    r = t.ratio(col)
    p = T0[r, s]
    t.pivot(r, s)
    exp_col = -(T0[:, s] * (1.0 / p))
    exp_col[r] = 1.0 / p
    np.testing.assert_array_equal(t.T[:, s], exp_col)
    exp_row = T0[r] / p
    exp_row[s] = 1.0 / p
    np.testing.assert_array_equal(t.T[r], exp_row)
    assert t.rowlab[r] == cl0[s] and t.collab[s] == rl0[r]
    t.pivot(r, s)
    np.testing.assert_array_equal(t.rowlab, rl0)
    np.testing.assert_array_equal(t.collab, cl0)
    np.testing.assert_allclose(t.T, T0, rtol=1e-12, atol=1e-12)


def test_bland_and_dantzig_reach_the_same_optimum(oracle):
    A, b, c, ops, mx = W.dense_feasible_lp(96, seed=3)
    r0 = oracle.solve_lp(A, b, -c, ops, oracle.make_opts(rule=0))
    r1 = oracle.solve_lp(A, b, -c, ops, oracle.make_opts(rule=1))
    assert r0["status"] == r1["status"] == 0
    assert abs(r0["fun"] - r1["fun"]) <= 1e-9 * abs(r0["fun"])


def test_threads_do_not_change_bits(oracle):
    A, b, c, ops, mx = W.dense_feasible_lp(128, seed=2)
    r1 = oracle.solve_lp(A, b, -c, ops, oracle.make_opts(threads=1), hist_cap=4096)
    r4 = oracle.solve_lp(A, b, -c, ops, oracle.make_opts(threads=4), hist_cap=4096)
    assert r1["n_pivots"] == r4["n_pivots"] and r1["fun"] == r4["fun"]
    np.testing.assert_array_equal(r1["piv_row"], r4["piv_row"])
    np.testing.assert_array_equal(r1["x"], r4["x"])


BEALE_A = np.array([[0.25, -8.0, -1.0, 9.0], [0.5, -12.0, -0.5, 3.0], [0.0, 0.0, 1.0, 0.0]])
BEALE_B = np.array([0.0, 0.0, 1.0])
BEALE_C = np.array([-0.75, 20.0, -0.5, 6.0])


def test_beale_cycling_example_terminates(oracle):
    """Beale's degenerate LP cycles under Dantzig's rule with lowest-id tie-breaking.  An explicit budget ends in
    LIMIT; the automatic budget continues under Bland's rule and reaches the optimum z* = -1.25 (HiGHS: -1.25)."""
    ops = np.zeros(3, dtype=np.int8)
    r = oracle.solve_lp(BEALE_A, BEALE_B, BEALE_C, ops, oracle.make_opts(rule=0, max_pivots=1000), hist_cap=16)
    assert r["status"] == oracle.LIMIT and r["n_pivots"] == 1000
    assert list(r["piv_col"][:8]) == [0, 1, 2, 3, 0, 1, 2, 3]          # the cycle
    r = oracle.solve_lp(BEALE_A, BEALE_B, BEALE_C, ops, oracle.make_opts(rule=1))
    assert r["status"] == 0 and r["n_pivots"] == 6 and r["fun"] == -1.25
    r = oracle.solve_lp(BEALE_A, BEALE_B, BEALE_C, ops, oracle.make_opts(rule=0))      # automatic budget
    assert r["status"] == 0 and abs(r["fun"] + 1.25) < 1e-9
    np.testing.assert_allclose(r["x"], [1.0, 0.0, 1.0, 0.0], atol=1e-9)


def test_fuzz_family_against_reference(oracle, golden):
    """1000 ragged LPs (degenerate, sparse, badly scaled; infeasible and unbounded ones included) solved by the
    reference path: status identical, z* within 1e-9, for both entering rules."""
    g = golden["fuzz"]
    import hashlib
    h = hashlib.sha256()
    for a in W.fuzz_lp(7):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest()[:16] == g["inputs_sha"]
    seen = set()
    for k, (st, z) in enumerate(g["results"]):
        A, b, c, ops = W.fuzz_lp(k, g["seed"])
        for rule in (0, 1):
            r = oracle.solve_lp(A, b, c, ops, oracle.make_opts(rule=rule))
            assert r["status"] == st, (k, rule)
            if z is not None:
                assert abs(r["fun"] - z) <= REL * max(1.0, abs(z)), (k, rule)
        seen.add(st)
    assert seen == {0, 2, 3}


def test_committed_pivot_histories_of_configs_4_and_5(oracle):
    """bench.py compares the pivot history of its timed region with tests/golden/pivot_history_config{4,5}.npy.  The
    files come from the oracle run on a column SLAB (make_pivot_history.py explains why that is exact under Bland);
    here the slab method is checked against the full tableau, and the head of both files is recomputed."""
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")

    def hist(t, k):
        r = t.solve(oracle.make_opts(rule=1, max_pivots=k, threads=4), hist_cap=k)
        return np.stack([r["piv_row"], r["enter_lab"], r["leave_lab"]], axis=1)

    # slab == full tableau while every entering id is inside the slab (small instance, 300 pivots)
    full = hist(oracle.OracleTableau.generate(4, 400, 500), 300)
    slab = hist(oracle.OracleTableau.generate(4, 400, 500, 0, 128), 300)
    assert full[:, 1].max() < 128 and np.array_equal(full, slab)
    # config 4: the FULL 16384 x 16384 tableau for the first pivots; config 5: its slab
    h4 = np.load(os.path.join(here, "pivot_history_config4.npy"))
    h5 = np.load(os.path.join(here, "pivot_history_config5.npy"))
    assert h4.shape == (12288, 3) and h5.shape == (1024, 3) and h4.dtype == np.int32
    assert np.array_equal(hist(oracle.OracleTableau.generate(4, 16383, 16383), 12), h4[:12])
    assert np.array_equal(hist(oracle.OracleTableau.generate(4, 131071, 131071, 0, 256), 48), h5[:48])
    assert h4[:, 1].max() < 4096 and h5[:, 1].max() < 1024   # the slabs that produced them held every entering id
