"""GPU parity: the CUDA path, called through the C ABI (libb200lp.so), against
  (i)  the CPU oracle (oracle/simplex_oracle.c) bit for bit: pivot sequence, tableau entries, z, x;
  (ii) the golden vectors produced by the reference's own solve path (tests/golden/reference_golden.json),
       status identical and z* within 1e-9 relative (the tolerance BASELINE.json's north_star states).
"""
import numpy as np
import pytest

from simplex_solver_b200 import native, workloads as W
from tests.helpers import assert_bit_equal, to_min_form, z_from_fun

pytestmark = pytest.mark.gpu

REL = 1e-9  # north_star: z* and x* within 1e-9 relative in fp64


def _torch():
    import torch
    return torch


def _device_tableau(solver, m, n_obj, C, n_struct, art_base):
    torch = _torch()
    ld = (C + 15) // 16 * 16
    T = torch.empty((m + n_obj) * ld, dtype=torch.float64, device="cuda:0")
    solver.attach(T.data_ptr(), m, n_obj, C, ld, n_struct, art_base, keep=T)
    return T, ld


# ---------------------------------------------------------------------------------------------------
# reference known-answer problems through the linprog seam
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rule", [native.RULE_DANTZIG, native.RULE_BLAND])
def test_kat_against_reference_golden_and_oracle(solver, oracle, golden, rule):
    for name, g in golden["kat"].items():
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(g["problem"])
        cmin = to_min_form(c, mx)
        got = solver.solve_dense(A, b, cmin, ops, native.make_opts(rule=rule), hist_cap=64)
        ref = oracle.solve_lp(A, b, cmin, ops, oracle.make_opts(rule=rule), hist_cap=64)
        assert got["status"] == g["scipy_status"], name
        assert got["status"] == ref["status"], name
        assert got["n_pivots"] == ref["n_pivots"] and got["n_phase1"] == ref["n_phase1"], name
        np.testing.assert_array_equal(got["piv_row"], ref["piv_row"], err_msg=name)
        np.testing.assert_array_equal(got["piv_col"], ref["piv_col"], err_msg=name)
        np.testing.assert_array_equal(got["enter_lab"], ref["enter_lab"], err_msg=name)
        np.testing.assert_array_equal(got["leave_lab"], ref["leave_lab"], err_msg=name)
        if got["status"] == 0:
            assert_bit_equal(got["fun"], ref["fun"], name + " fun")
            assert_bit_equal(got["x"], ref["x"], name + " x")
            z = z_from_fun(got["fun"], mx)
            assert abs(z - g["z"]) <= REL * max(1.0, abs(g["z"])), name
            if g["x_unique"]:
                np.testing.assert_allclose(got["x"], g["x"], rtol=REL, atol=REL, err_msg=name)


def test_mixed_operator_lps_against_reference_golden(solver, oracle, golden):
    for k, g in enumerate(golden["mixed"]):
        A = np.array(g["A"], dtype=np.float64).reshape(len(g["b"]), len(g["c"]))
        b, c, ops = np.array(g["b"]), np.array(g["c"]), np.array(g["ops"], dtype=np.int8)
        cmin = to_min_form(c, g["maximize"])
        got = solver.solve_dense(A, b, cmin, ops, hist_cap=256)
        ref = oracle.solve_lp(A, b, cmin, ops, hist_cap=256)
        assert got["status"] == g["scipy_status"], k
        assert got["status"] == ref["status"] and got["n_pivots"] == ref["n_pivots"], k
        np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
        np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
        if got["status"] == 0:
            assert_bit_equal(got["fun"], ref["fun"], f"mixed {k} fun")
            assert_bit_equal(got["x"], ref["x"], f"mixed {k} x")
            z = z_from_fun(got["fun"], g["maximize"])
            assert abs(z - g["z"]) <= REL * max(1.0, abs(g["z"])), k


# ---------------------------------------------------------------------------------------------------
# BASELINE config 2 family: dense feasible LPs, Dantzig
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [16, 64, 128, 256])
def test_dense_lp_exact_vs_oracle_and_golden(solver, oracle, golden, n):
    A, b, c, ops, mx = W.dense_feasible_lp(n, seed=0)
    cmin = to_min_form(c, mx)
    cap = 1 << 14
    got = solver.solve_dense(A, b, cmin, ops, native.make_opts(rule=native.RULE_DANTZIG), hist_cap=cap)
    ref = oracle.solve_lp(A, b, cmin, ops, oracle.make_opts(rule=oracle.RULE_DANTZIG), hist_cap=cap)
    assert got["status"] == 0 == ref["status"]
    assert got["n_pivots"] == ref["n_pivots"]
    np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
    np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
    assert_bit_equal(got["fun"], ref["fun"], "fun")
    assert_bit_equal(got["x"], ref["x"], "x")
    zg = golden["dense"][str(n)]["z"]
    assert abs(-got["fun"] - zg) <= REL * abs(zg)


def test_dense_1024_config2_vs_reference_golden(solver, golden):
    """BASELINE config 2 at full size: z* within 1e-9 of the reference path (HiGHS) -- 447.0328820157."""
    A, b, c, ops, mx = W.dense_feasible_lp(1024, seed=0)
    got = solver.solve_dense(A, b, to_min_form(c, mx), ops, native.make_opts(rule=native.RULE_DANTZIG))
    zg = golden["dense"]["1024"]["z"]
    assert got["status"] == 0
    assert abs(-got["fun"] - zg) <= REL * abs(zg)
    x = got["x"]
    assert (x >= -1e-9).all() and (A @ x <= b + 1e-7).all()
    assert abs(c @ x - zg) <= 1e-8 * abs(zg)


def test_dense_update_variants_agree(solver, oracle):
    A, b, c, ops, mx = W.dense_feasible_lp(192, seed=5)
    cmin = to_min_form(c, mx)
    ref = oracle.solve_lp(A, b, cmin, ops, hist_cap=1 << 13)
    for variant in (native.UPDATE_LDG, native.UPDATE_TMA):
        for mode in (native.LOOP_LAUNCHES, native.LOOP_GRAPH, native.LOOP_AUTO, native.LOOP_BLOCKED):
            got = solver.solve_dense(A, b, cmin, ops, native.make_opts(update_variant=variant, loop_mode=mode,
                                                                     check_every=16), hist_cap=1 << 13)
            assert got["n_pivots"] == ref["n_pivots"], (variant, mode)
            np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
            np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
            np.testing.assert_array_equal(got["enter_lab"], ref["enter_lab"])
            np.testing.assert_array_equal(got["leave_lab"], ref["leave_lab"])
            assert_bit_equal(got["fun"], ref["fun"], f"variant {variant} loop mode {mode}")
            assert_bit_equal(got["x"], ref["x"], f"variant {variant} loop mode {mode}")


# ---------------------------------------------------------------------------------------------------
# device-resident generated tableau (configs 4/5 generator), one phase at a time and as a loop
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(37, 53), (200, 300), (513, 1030)])
def test_generated_tableau_and_single_phases_bit_exact(solver, oracle, shape):
    m, n = shape
    seed = 4
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    solver.generate(seed, n, 0)
    ot = oracle.OracleTableau.generate(seed, m, n)
    assert_bit_equal(solver.read_tableau(), ot.T, "generated tableau")
    rl, cl = solver.get_labels()
    np.testing.assert_array_equal(rl, ot.rowlab)
    np.testing.assert_array_equal(cl, ot.collab)
    for it in range(6):
        rule = native.RULE_BLAND if it % 2 else native.RULE_DANTZIG
        s = solver.select_entering(rule=rule)
        assert s == ot.price(rule=rule)
        r = solver.ratio_test(s)
        col = ot.extract_col(s)
        assert r == ot.ratio(col)
        variant = native.UPDATE_TMA if it % 3 == 2 else native.UPDATE_LDG
        solver.pivot(r, s, variant)
        ot.pivot(r, s)
        assert_bit_equal(solver.read_tableau(), ot.T, f"tableau after pivot {it}")
        rl, cl = solver.get_labels()
        np.testing.assert_array_equal(rl, ot.rowlab)
        np.testing.assert_array_equal(cl, ot.collab)


@pytest.mark.parametrize("variant", [native.UPDATE_LDG, native.UPDATE_TMA, "onchip", "blocked3", "blocked8", "blocked16", "blocked32"])
@pytest.mark.parametrize("rule", [native.RULE_BLAND, native.RULE_DANTZIG])
def test_device_loop_fixed_budget_bit_exact(solver, oracle, rule, variant):
    """Config 4 in miniature: a fixed pivot budget on a generated square tableau, Bland and Dantzig, through the
    multi-kernel graph loop (both update kernels) and through the on-chip persistent loop."""
    m, n, budget = 255, 255, 150
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    solver.generate(4, n, 0)
    if variant == "onchip":
        o = native.make_opts(rule=rule, max_pivots=budget, loop_mode=native.LOOP_AUTO)
    elif isinstance(variant, str):  # look-ahead blocks of K pivots (150 is not a multiple of 8 or 16: ragged last block)
        o = native.make_opts(rule=rule, max_pivots=budget, loop_mode=native.LOOP_BLOCKED, check_every=int(variant[7:]))
    else:
        o = native.make_opts(rule=rule, max_pivots=budget, update_variant=variant, loop_mode=native.LOOP_GRAPH)
    got = solver.run(o, hist_cap=budget)
    ot = oracle.OracleTableau.generate(4, m, n)
    ref = ot.solve(oracle.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
    assert got["status"] == ref["status"] == 1 and got["n_pivots"] == ref["n_pivots"] == budget
    np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
    np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
    np.testing.assert_array_equal(got["enter_lab"], ref["enter_lab"])
    np.testing.assert_array_equal(got["leave_lab"], ref["leave_lab"])
    assert_bit_equal(solver.read_tableau(), ot.T, "tableau after the budget")
    assert_bit_equal(got["fun"], ref["fun"], "fun")


@pytest.mark.parametrize("shape", [(40, 6200), (3, 5), (2, 1), (1025, 300), (512, 2047)])
@pytest.mark.parametrize("rule", [native.RULE_BLAND, native.RULE_DANTZIG])
def test_onchip_loop_shapes_bit_exact(solver, oracle, rule, shape):
    """The on-chip persistent loop on the shapes that take its other code paths: slices wider than 32 columns (priced by
    the whole CTA instead of per warp; 6200 / 148 = 42 columns per SM), fewer columns than SMs, a single column, row counts
    that leave the last thread of the update / ratio / column-copy loops with a ragged trip."""
    m, n = shape
    budget = 90
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    solver.generate(11, n, 0)
    got = solver.run(native.make_opts(rule=rule, max_pivots=budget, loop_mode=native.LOOP_AUTO), hist_cap=budget)
    assert got["kernel_launches"] < 16, "AUTO should have taken the single-launch on-chip loop (plus a few state kernels)"
    ot = oracle.OracleTableau.generate(11, m, n)
    ref = ot.solve(oracle.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
    assert got["status"] == ref["status"] and got["n_pivots"] == ref["n_pivots"]
    np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
    np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
    assert_bit_equal(solver.read_tableau(), ot.T, "tableau")
    assert_bit_equal(got["fun"], ref["fun"], "fun")


@pytest.mark.parametrize("shape", [(300, 1100, 0), (700, 520, 3), (129, 513, 1)])
@pytest.mark.parametrize("K", [5, 32])
def test_lookahead_ragged_strips_and_row_blocks(solver, oracle, shape, K):
    """The look-ahead flush walks 512-column strips in 128-row tiles: shapes with several strips / row blocks whose last
    ones are ragged, an odd number of stored columns, and a row stride with (pad > 0) and without padding."""
    m, n, pad = shape
    budget = 70
    torch = _torch()
    C = n + 1
    ld = C + pad + ((C + pad) & 1)
    T = torch.empty((m + 1) * ld, dtype=torch.float64, device="cuda:0")
    solver.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=T)
    solver.generate(9, n, 0)
    for rule in (native.RULE_BLAND, native.RULE_DANTZIG):
        solver.generate(9, n, 0)
        got = solver.run(native.make_opts(rule=rule, max_pivots=budget, loop_mode=native.LOOP_BLOCKED, check_every=K),
                         hist_cap=budget)
        ot = oracle.OracleTableau.generate(9, m, n)
        ref = ot.solve(oracle.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
        assert got["status"] == ref["status"] and got["n_pivots"] == ref["n_pivots"]
        np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
        np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
        assert_bit_equal(solver.read_tableau(), ot.T, f"tableau, rule {rule}")


def test_lookahead_dmma_flush_variant_bit_exact(oracle, monkeypatch):
    """The DMMA flush (kernels_flush_mma.cuh: four pending steps of an 8 x 8 tile per mma.sync.m8n8k4.f64) is kept as a
    diagnostic variant behind B200LP_FLUSH_MMA (the DFMA kernel is faster on B200).  Its claim -- the instruction evaluates
    the sequential fma chain, k = 0 first -- is checked here against the oracle: ragged strips, K not a multiple of 4."""
    monkeypatch.setenv("B200LP_FLUSH_MMA", "1")
    s = native.Solver(0)
    torch = _torch()
    try:
        for (m, n, K) in ((300, 1100, 32), (129, 513, 7), (700, 520, 13)):
            budget = 70
            C = n + 1
            ld = C + (C & 1)
            T = torch.empty((m + 1) * ld, dtype=torch.float64, device="cuda:0")
            s.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=T)
            s.generate(9, n, 0)
            got = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=budget, loop_mode=native.LOOP_BLOCKED,
                                         check_every=K), hist_cap=budget)
            ot = oracle.OracleTableau.generate(9, m, n)
            ref = ot.solve(oracle.make_opts(rule=oracle.RULE_BLAND, max_pivots=budget), hist_cap=budget)
            assert got["n_pivots"] == ref["n_pivots"]
            np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
            assert_bit_equal(s.read_tableau(), ot.T, f"tableau {m}x{n} K={K}")
    finally:
        s.close()


def test_config4_lookahead_equals_rank1_loop_at_full_size(solver):
    """BASELINE config 4 at full size: 70 Bland pivots through the look-ahead loop (K = 32: two full blocks and a ragged
    one) leave the 2 GB tableau bit-identical to the rank-1 graph loop's, with the same pivot sequence."""
    torch = _torch()
    m = n = 16383
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    solver.generate(4, n, 0)
    a = solver.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=70, loop_mode=native.LOOP_GRAPH), hist_cap=70)
    torch.cuda.synchronize()
    Ta = T.clone()
    torch.cuda.synchronize()  # the copy runs on torch's stream, the generator on the solver's
    solver.generate(4, n, 0)
    b = solver.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=70, loop_mode=native.LOOP_BLOCKED, check_every=32),
                   hist_cap=70)
    torch.cuda.synchronize()
    assert a["n_pivots"] == b["n_pivots"] == 70
    np.testing.assert_array_equal(a["piv_row"], b["piv_row"])
    np.testing.assert_array_equal(a["piv_col"], b["piv_col"])
    assert a["fun"] == b["fun"]
    # bit patterns, not values: NaN-safe and distinguishes -0.0
    assert torch.equal(Ta.view(torch.int64), T.view(torch.int64))
    del T, Ta


def test_device_loop_to_optimality(solver, oracle):
    m, n = 96, 160
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    solver.generate(11, n, 0)
    got = solver.run(native.make_opts(rule=native.RULE_DANTZIG), hist_cap=1 << 13)
    ot = oracle.OracleTableau.generate(11, m, n)
    ref = ot.solve(oracle.make_opts(rule=oracle.RULE_DANTZIG), hist_cap=1 << 13)
    assert got["status"] == ref["status"] == 0
    assert got["n_pivots"] == ref["n_pivots"]
    np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
    assert_bit_equal(got["x"], ot.read_x(), "x")
    assert_bit_equal(solver.read_tableau(), ot.T, "final tableau")


def test_config4_full_size_properties(solver):
    """BASELINE config 4 at full size (16384 x 16384, Bland): size-independent properties of a pivot.
    After pivoting on (r, s): column s holds -col/p (1/p at row r), row r holds row/p, the objective value
    moved by -d_s * rhs_r / p, the basis swap is recorded, and the next Bland choice is deterministic."""
    torch = _torch()
    m = n = 16383
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    solver.generate(4, n, 0)
    Tt = T.view(m + 1, ld)
    s = solver.select_entering(rule=native.RULE_BLAND)
    assert s == 0  # every reduced cost is negative at the start: Bland takes variable 0
    r = solver.ratio_test(s)
    col = Tt[:, s].clone()
    rowr = Tt[r, : n + 1].clone()
    rhs = Tt[:, n].clone()
    ratios = torch.where(col[:m] > 1e-9, rhs[:m] / col[:m], torch.full_like(col[:m], float("inf")))
    assert int(torch.argmin(ratios)) == r
    p = col[r].item()
    z0 = Tt[m, n].item()
    some = Tt[5:9, 7:11].clone()
    torch.cuda.synchronize()  # torch's copies above run on torch's stream, the pivot on the solver's
    solver.pivot(r, s, native.UPDATE_AUTO)
    torch.cuda.synchronize()
    inv_p = 1.0 / p
    exp_col = -(col * inv_p)
    exp_col[r] = inv_p
    assert torch.equal(Tt[:, s], exp_col)
    # torch turns "tensor / python_scalar" into a multiplication by the reciprocal; divide by a device tensor to get
    # the IEEE division the kernel performs
    p_dev = col[r].clone()
    exp_row = rowr / p_dev
    exp_row[s] = inv_p
    assert torch.equal(Tt[r, : n + 1], exp_row)
    q = rowr[7:11] / p_dev
    exp_some = torch.stack([torch.addcmul(some[k], -col[5 + k], q) for k in range(4)])  # not fused: 1 ulp slack
    assert torch.allclose(Tt[5:9, 7:11], exp_some, rtol=1e-15, atol=1e-15)
    zexp = z0 - col[m].item() * (rhs[r].item() / p)
    assert abs(Tt[m, n].item() - zexp) <= 1e-12 * max(1.0, abs(zexp))
    rl, cl = solver.get_labels()
    assert rl[r] == 0 and cl[s] == n + r
    # a short Bland loop keeps the objective monotone (maximisation form: -T[m][n] decreases)
    fun_before = -Tt[m, n].item()
    res = solver.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=8), hist_cap=8)
    assert res["n_pivots"] == 8 and res["status"] == 1
    assert res["fun"] <= fun_before + 1e-9 * abs(fun_before)
    assert (np.diff(res["enter_lab"]) != 0).all()
    del T


# ---------------------------------------------------------------------------------------------------
# BASELINE config 3: batched small LPs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rule", [native.RULE_DANTZIG, native.RULE_BLAND])
def test_batched_bit_exact_vs_oracle(solver, oracle, rule):
    A, b, c, ops = W.batched_small_lps(0, 1000)
    got = solver.solve_batched(A, b, c, ops, native.make_opts(rule=rule), log_cap=96)
    ref = oracle.solve_batched(A, b, c, ops, oracle.make_opts(rule=rule), log_cap=96, threads=8)
    np.testing.assert_array_equal(got["status"], ref["status"])
    np.testing.assert_array_equal(got["n_pivots"], ref["n_pivots"])
    np.testing.assert_array_equal(got["piv_log"], ref["piv_log"])
    ok = got["status"] == 0
    assert_bit_equal(got["fun"][ok], ref["fun"][ok], "fun")
    assert_bit_equal(got["x"][ok], ref["x"][ok], "x")
    assert set(np.unique(got["status"])) == {0, 2, 3}


def test_batched_against_reference_golden(solver, golden):
    g = golden["batched"]
    A, b, c, ops = W.batched_small_lps(0, g["count"], g["m"], g["n"], g["base_seed"])
    got = solver.solve_batched(A, b, c, ops)
    for k, row in enumerate(g["results"]):
        assert got["status"][k] == row["scipy_status"], k
        if row["z"] is not None:
            assert abs(got["fun"][k] - row["z"]) <= REL * max(1.0, abs(row["z"])), k


@pytest.mark.parametrize("shape", [(1, 1), (3, 2), (7, 12), (33, 40), (64, 64)])
def test_batched_ragged_shapes(solver, oracle, shape):
    m, n = shape
    A, b, c, ops = W.batched_small_lps(0, 257, m, n, base_seed=9)
    got = solver.solve_batched(A, b, c, ops, log_cap=32)
    ref = oracle.solve_batched(A, b, c, ops, log_cap=32, threads=8)
    np.testing.assert_array_equal(got["status"], ref["status"])
    np.testing.assert_array_equal(got["n_pivots"], ref["n_pivots"])
    np.testing.assert_array_equal(got["piv_log"], ref["piv_log"])
    ok = got["status"] == 0
    assert_bit_equal(got["fun"][ok], ref["fun"][ok], "fun")
    assert_bit_equal(got["x"][ok], ref["x"][ok], "x")


def test_batched_matches_single_lp_path(solver):
    """The warp-per-LP kernel and the multi-kernel single-LP path take the same pivots."""
    A, b, c, ops = W.batched_small_lps(0, 40)
    got = solver.solve_batched(A, b, c, ops, log_cap=96)
    for k in range(40):
        one = solver.solve_dense(A[k], b[k], c[k], ops[k], hist_cap=96)
        assert one["status"] == got["status"][k]
        assert one["n_pivots"] == got["n_pivots"][k]
        np.testing.assert_array_equal(one["piv_row"], got["piv_log"][k, : one["n_pivots"], 0])
        np.testing.assert_array_equal(one["piv_col"], got["piv_log"][k, : one["n_pivots"], 1])
        if one["status"] == 0:
            assert_bit_equal(one["fun"], got["fun"][k], "fun")
            assert_bit_equal(one["x"], got["x"][k], "x")


# ---------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------
def test_empty_constraint_set_is_unbounded_or_trivial(solver):
    # K7: no constraints, max x1+x2  -> unbounded (status 3 -> "Error" in the reference's mapping)
    got = solver.solve_dense(np.zeros((0, 2)), np.zeros(0), np.array([-1.0, -1.0]), np.zeros(0, dtype=np.int8))
    assert got["status"] == native.STATUS_UNBOUNDED
    # no constraints, min x1+x2 -> optimal at 0
    got = solver.solve_dense(np.zeros((0, 2)), np.zeros(0), np.array([1.0, 1.0]), np.zeros(0, dtype=np.int8))
    assert got["status"] == 0 and got["fun"] == 0.0 and (got["x"] == 0).all()


def test_pivot_limit_status(solver):
    A, b, c, ops, mx = W.dense_feasible_lp(64, seed=1)
    got = solver.solve_dense(A, b, to_min_form(c, mx), ops, native.make_opts(max_pivots=5), hist_cap=16)
    assert got["status"] == native.STATUS_LIMIT and got["n_pivots"] == 5


def test_redundant_equalities_and_degenerate_rows(solver, oracle):
    # duplicated equality rows leave an artificial basic at level zero: exercises the drive-out path
    A = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [2.0, 2.0, 0.0], [0.0, 1.0, 1.0]])
    b = np.array([4.0, 4.0, 8.0, 3.0])
    ops = np.array([2, 2, 2, 0], dtype=np.int8)
    c = np.array([1.0, 2.0, -1.0])
    got = solver.solve_dense(A, b, c, ops, hist_cap=32)
    ref = oracle.solve_lp(A, b, c, ops, hist_cap=32)
    assert got["status"] == ref["status"] == 0
    assert got["n_pivots"] == ref["n_pivots"]
    np.testing.assert_array_equal(got["piv_row"], ref["piv_row"])
    np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
    assert_bit_equal(got["fun"], ref["fun"], "fun")
    assert_bit_equal(got["x"], ref["x"], "x")


def test_invalid_arguments_fail_loudly(solver):
    with pytest.raises(ValueError):  # the Python host validates before the C ABI is entered
        solver.solve_dense(np.zeros((1, 1)), np.zeros(1), np.zeros(1), np.array([7], dtype=np.int8))
    # the C ABI itself rejects the same problem (a caller that binds the library directly)
    import ctypes as C
    A, b, c, ops = np.zeros((1, 1)), np.zeros(1), np.zeros(1), np.array([7], dtype=np.int8)
    prob = native.Problem(1, 1, 1, native._ptr(A), native._ptr(b), native._ptr(c), native._ptr(ops), 0, 0)
    res, _keep = native.Solver._result(1, 0)
    o = native.make_opts()
    assert native.lib().b200lp_solve_dense(solver._h, C.byref(prob), C.byref(o), C.byref(res)) == -1
    assert b"not L/G/E" in native.lib().b200lp_last_error()
    with pytest.raises(native.B200LPError):
        solver.solve_dense(np.zeros((1, 1)), np.zeros(1), np.zeros(1), np.zeros(1, dtype=np.int8),
                           native.make_opts(rule=9))


def test_beale_cycling_example_terminates_like_the_oracle(solver, oracle):
    """A degenerate LP that cycles under Dantzig: the device loop must stop (explicit budget -> LIMIT) or recover
    (automatic budget -> Bland continuation), with the same pivots as the oracle, on every path."""
    from tests.test_oracle_golden import BEALE_A, BEALE_B, BEALE_C
    ops = np.zeros(3, dtype=np.int8)
    for mode in (native.LOOP_AUTO, native.LOOP_GRAPH):
        got = solver.solve_dense(BEALE_A, BEALE_B, BEALE_C, ops, native.make_opts(max_pivots=500, loop_mode=mode),
                                 hist_cap=500)
        ref = oracle.solve_lp(BEALE_A, BEALE_B, BEALE_C, ops, oracle.make_opts(max_pivots=500), hist_cap=500)
        assert got["status"] == ref["status"] == native.STATUS_LIMIT and got["n_pivots"] == 500
        np.testing.assert_array_equal(got["piv_col"], ref["piv_col"])
    got = solver.solve_dense(BEALE_A, BEALE_B, BEALE_C, ops, hist_cap=1 << 14)
    ref = oracle.solve_lp(BEALE_A, BEALE_B, BEALE_C, ops, hist_cap=1 << 14)
    assert got["status"] == ref["status"] == 0 and got["n_pivots"] == ref["n_pivots"]
    np.testing.assert_array_equal(got["enter_lab"], ref["enter_lab"])
    assert_bit_equal(got["fun"], ref["fun"], "fun")
    assert abs(got["fun"] + 1.25) < 1e-9
    A3, b3, c3, o3 = (np.stack([a] * 5) for a in (BEALE_A, BEALE_B, BEALE_C, ops))
    gb = solver.solve_batched(A3, b3, c3, o3)
    assert (gb["status"] == 0).all() and (gb["n_pivots"] == ref["n_pivots"]).all()
    assert_bit_equal(gb["fun"], np.full(5, ref["fun"]), "batched fun")


def test_midsize_two_phase_auto_takes_lookahead_and_equals_rank1_loop(solver):
    """A two-phase LP (<=, >=, = rows) whose tableau (~48 MB) is beyond the on-chip loop: loop_mode AUTO runs it through
    the look-ahead loop (phase 1, drive-out, phase 2).  Same pivots, same x*, same z* bit for bit as the rank-1 graph
    loop, which the smaller cases pin against the oracle."""
    rng = np.random.default_rng(21)
    m, n = 2000, 3000
    A = rng.uniform(-1.0, 1.0, (m, n))
    x0 = rng.random(n)
    u = rng.random(m)
    ops = np.where(u < 0.6, 0, np.where(u < 0.85, 1, 2)).astype(np.int8)
    slack = rng.uniform(0.1, 1.0, m)
    b = A @ x0 + np.where(ops == 0, slack, np.where(ops == 1, -slack, 0.0))
    c = rng.uniform(0.1, 1.0, n)
    cap = 1 << 19  # Dantzig needs ~250k pivots here
    a = solver.solve_dense(A, b, c, ops, native.make_opts(rule=native.RULE_DANTZIG, loop_mode=native.LOOP_AUTO), hist_cap=cap)
    g = solver.solve_dense(A, b, c, ops, native.make_opts(rule=native.RULE_DANTZIG, loop_mode=native.LOOP_GRAPH), hist_cap=cap)
    assert a["status"] == g["status"] == 0
    assert a["n_pivots"] == g["n_pivots"] and a["n_phase1"] == g["n_phase1"] and a["n_pivots"] < cap
    np.testing.assert_array_equal(a["piv_row"], g["piv_row"])
    np.testing.assert_array_equal(a["piv_col"], g["piv_col"])
    np.testing.assert_array_equal(a["enter_lab"], g["enter_lab"])
    assert_bit_equal(a["fun"], g["fun"], "fun")
    assert_bit_equal(a["x"], g["x"], "x")
    assert np.all(a["x"] >= 0) and np.all(A[ops == 0] @ a["x"] <= b[ops == 0] + 1e-7)
    assert np.all(A[ops == 1] @ a["x"] >= b[ops == 1] - 1e-7) and np.allclose(A[ops == 2] @ a["x"], b[ops == 2], atol=1e-7)
    assert a["fun"] <= float(c @ x0) + 1e-9


def test_config3_full_size_properties(solver):
    """BASELINE config 3 at full size (100 000 LPs of 20 x 30, two-phase): size-independent properties.
    Status pattern of the generator (index = 0 mod 100 infeasible, = 1 mod 100 unbounded, rest optimal), primal
    feasibility and objective consistency of every optimal LP, non-negativity; checked for both rules."""
    B = 100000
    A, b, c, ops = W.batched_small_lps(0, B)
    idx = np.arange(B)
    funs = []
    for rule in (native.RULE_DANTZIG, native.RULE_BLAND):
        r = solver.solve_batched(A, b, c, ops, native.make_opts(rule=rule))
        st = r["status"]
        assert (st[idx % 100 == 0] == native.STATUS_INFEASIBLE).all()
        assert (st[idx % 100 == 1] == native.STATUS_UNBOUNDED).all()
        assert (st[idx % 100 > 1] == native.STATUS_OPTIMAL).all()
        ok = st == 0
        x = r["x"][ok]
        assert (x >= -1e-9).all()
        Ax = np.einsum("kij,kj->ki", A[ok], x)
        tol = 1e-7 * np.maximum(1.0, np.abs(b[ok]))
        o = ops[ok]
        assert (Ax[o == 0] <= (b[ok] + tol)[o == 0]).all()
        assert (Ax[o == 1] >= (b[ok] - tol)[o == 1]).all()
        assert (np.abs(Ax - b[ok])[o == 2] <= tol[o == 2]).all()
        cx = np.einsum("kj,kj->k", c[ok], x)
        assert np.allclose(cx, r["fun"][ok], rtol=1e-9, atol=1e-9)
        assert (r["n_pivots"][ok] > 0).all() and r["n_pivots"].max() < 400
        funs.append(r["fun"][ok])
    # both rules reach the same optimum
    assert np.allclose(funs[0], funs[1], rtol=1e-9, atol=1e-9)


def test_fuzz_family_gpu_vs_oracle_and_reference(solver, oracle, golden):
    """The ragged stress family through every single-LP path of the library (on-chip loop, graph loop with both update
    kernels) and through the batched kernel: bit-identical to the oracle, status / z* as the reference path."""
    g = golden["fuzz"]
    modes = [dict(loop_mode=native.LOOP_AUTO), dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_LDG),
             dict(loop_mode=native.LOOP_LAUNCHES, update_variant=native.UPDATE_TMA),
             dict(loop_mode=native.LOOP_BLOCKED, check_every=5)]
    for k in range(0, 400):
        A, b, c, ops = W.fuzz_lp(k, g["seed"])
        st, z = g["results"][k]
        rule = k % 2
        ref = oracle.solve_lp(A, b, c, ops, oracle.make_opts(rule=rule), hist_cap=512)
        got = solver.solve_dense(A, b, c, ops, native.make_opts(rule=rule, **modes[k % 4]), hist_cap=512)
        assert got["status"] == st == ref["status"], k
        assert got["n_pivots"] == ref["n_pivots"] and got["n_phase1"] == ref["n_phase1"], k
        np.testing.assert_array_equal(got["piv_row"], ref["piv_row"], err_msg=str(k))
        np.testing.assert_array_equal(got["enter_lab"], ref["enter_lab"], err_msg=str(k))
        if st == 0:
            assert_bit_equal(got["fun"], ref["fun"], f"fuzz {k} fun")
            assert_bit_equal(got["x"], ref["x"], f"fuzz {k} x")
            assert abs(got["fun"] - z) <= REL * max(1.0, abs(z)), k
    # batched kernel: group LPs of equal shape
    shapes = {}
    for k in range(g["count"]):
        A, b, c, ops = W.fuzz_lp(k, g["seed"])
        shapes.setdefault(A.shape, []).append((k, A, b, c, ops))
    checked = 0
    for shape, items in shapes.items():
        if len(items) < 3:
            continue
        Ab = np.stack([it[1] for it in items])
        bb = np.stack([it[2] for it in items])
        cb = np.stack([it[3] for it in items])
        ob = np.stack([it[4] for it in items])
        got = solver.solve_batched(Ab, bb, cb, ob)
        for i, it in enumerate(items):
            st, z = g["results"][it[0]]
            assert got["status"][i] == st, it[0]
            if st == 0:
                assert abs(got["fun"][i] - z) <= REL * max(1.0, abs(z)), it[0]
            checked += 1
    assert checked > 500


# ---------------------------------------------------------------------------------------------------
# full-size configs against the ORACLE itself (not only properties)
# ---------------------------------------------------------------------------------------------------
def test_config4_full_size_vs_oracle(solver, oracle):
    """BASELINE config 4 at full size against the oracle: 24 Bland pivots of the 16384 x 16384 tableau -- pivot history
    and every one of the 268 M stored doubles equal to oracle.OracleTableau's, for the rank-1 loop (both update kernels)
    and the look-ahead loop; the history is also the head of the committed sequence bench.py checks its timed run with."""
    import os
    torch = _torch()
    m = n = 16383
    budget = 24
    ot = oracle.OracleTableau.generate(4, m, n)
    ref = ot.solve(oracle.make_opts(rule=oracle.RULE_BLAND, max_pivots=budget, threads=oracle.host_cores()), hist_cap=budget)
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "pivot_history_config4.npy"))[:budget]
    assert np.array_equal(np.stack([ref["piv_row"], ref["enter_lab"], ref["leave_lab"]], axis=1), want)
    T, ld = _device_tableau(solver, m, 1, n + 1, n, n + m)
    for kw in (dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_LDG),
               dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_TMA),
               dict(loop_mode=native.LOOP_BLOCKED, check_every=16)):
        solver.generate(4, n, 0)
        got = solver.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=budget, **kw), hist_cap=budget)
        assert got["n_pivots"] == budget
        for key in ("piv_row", "piv_col", "enter_lab", "leave_lab"):
            np.testing.assert_array_equal(got[key], ref[key], err_msg=f"{key} {kw}")
        assert got["fun"] == ref["fun"]
        assert_bit_equal(solver.read_tableau(), ot.T, f"config 4 tableau {kw}")
    del T


def test_config2_full_size_pivot_sequence_vs_oracle(solver, oracle):
    """BASELINE config 2 at full size (1024 x 1024, Dantzig, to optimality): the GPU's default path (on-chip loop) and
    the look-ahead loop take the oracle's ~10^4 pivots one for one and end on the oracle's tableau bits, x* and z*."""
    A, b, c, ops, mx = W.dense_feasible_lp(1024, seed=0)
    cmin = to_min_form(c, mx)
    cap = 1 << 15
    ref = oracle.solve_lp(A, b, cmin, ops, oracle.make_opts(rule=oracle.RULE_DANTZIG, threads=oracle.host_cores()), hist_cap=cap)
    assert ref["status"] == 0 and 1000 < ref["n_pivots"] < cap
    ot = oracle.OracleTableau.build(A, b, cmin, ops)
    ot.solve(oracle.make_opts(rule=oracle.RULE_DANTZIG, threads=oracle.host_cores()))
    for kw in (dict(loop_mode=native.LOOP_AUTO), dict(loop_mode=native.LOOP_BLOCKED)):
        got = solver.solve_dense(A, b, cmin, ops, native.make_opts(rule=native.RULE_DANTZIG, **kw), hist_cap=cap)
        assert got["status"] == 0 and got["n_pivots"] == ref["n_pivots"], kw
        for key in ("piv_row", "piv_col", "enter_lab", "leave_lab"):
            np.testing.assert_array_equal(got[key], ref[key], err_msg=f"{key} {kw}")
        assert got["fun"] == ref["fun"]
        assert_bit_equal(got["x"], ref["x"], "x*")
        assert_bit_equal(solver.read_tableau(), ot.T, f"config 2 final tableau {kw}")


@pytest.mark.parametrize("rule", [native.RULE_DANTZIG, native.RULE_BLAND])
def test_config3_full_size_vs_oracle(solver, oracle, rule):
    """BASELINE config 3 at full size: all 100 000 LPs against orc_solve_batched -- status, pivot count, the first 48
    (row, column) pivots of every LP, z* and x* bit for bit."""
    B = 100000
    A, b, c, ops = W.batched_small_lps(0, B)
    got = solver.solve_batched(A, b, c, ops, native.make_opts(rule=rule), log_cap=48)
    ref = oracle.solve_batched(A, b, c, ops, oracle.make_opts(rule=rule), log_cap=48, threads=oracle.host_cores())
    np.testing.assert_array_equal(got["status"], ref["status"])
    np.testing.assert_array_equal(got["n_pivots"], ref["n_pivots"])
    np.testing.assert_array_equal(got["piv_log"], ref["piv_log"])
    ok = got["status"] == 0
    assert_bit_equal(got["fun"][ok], ref["fun"][ok], "z*")
    assert_bit_equal(got["x"][ok], ref["x"][ok], "x*")
