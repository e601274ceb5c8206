"""GPU tests of the reference-facing host API: the linprog seam, the SolverController mirror and its report,
the pivotSteps producer, the plugin patching, and the column-shard engine (world = 1) against the oracle."""
import sys
import types

import numpy as np
import pytest

from simplex_solver_b200 import native, workloads as W
from simplex_solver_b200.linprog import linprog
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau
from simplex_solver_b200.solver_controller import BulkSolverController, SolverController
from tests.helpers import assert_bit_equal

pytestmark = pytest.mark.gpu
REL = 1e-9


def test_linprog_seam_on_reference_fixtures(golden):
    for name, g in golden["kat"].items():
        ctl = SolverController(g["problem"])
        c, A_ub, b_ub, A_eq, b_eq, bounds = ctl._prepare_model_for_scipy(ctl.objective_data, ctl.constraints_data,
                                                                        ctl.variables)
        r = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method="highs-ds",
                    options={"presolve": True, "time_limit": 10})
        assert r.status == g["scipy_status"], name
        assert r.success == (g["scipy_status"] == 0)
        assert isinstance(r.message, str) and r.message
        if r.success:
            z = -r.fun if ctl.objective_data["type"] == "maximize" else r.fun
            assert abs(z - g["z"]) <= REL * max(1.0, abs(g["z"])), name
            assert len(r.x) == len(ctl.variables)
        else:
            assert r.x is None and r.fun is None


def test_solver_controller_report_wyndor(golden, tmp_path):
    """BASELINE config 1 through the mirror class: same report keys / strings as the reference (:417-422)."""
    wrapper = golden["kat"]["K1_wyndor"]["problem"]
    rep = SolverController(wrapper, output_dir=str(tmp_path)).run()
    assert set(rep) == {"problema_definicion", "solucion_encontrada", "visualizacion_gilp_html", "tablas_intermedias"}
    sol = rep["solucion_encontrada"]
    assert sol["status"] == "Solucion Factible"
    assert abs(sol["valor_optimo_z"] - 36.0) < 1e-9
    assert abs(sol["valores_variables"]["x1"] - 2.0) < 1e-9 and abs(sol["valores_variables"]["x2"] - 6.0) < 1e-9
    assert f"{sol['valor_optimo_z']:.4f}" == "36.0000"
    tabs = rep["tablas_intermedias"]
    assert tabs[0]["title"] == "Iteración 0 (Tabla Inicial)" and tabs[0]["pivot"] is None
    assert len(tabs) == 3 and tabs[1]["pivot"] is not None                      # Dantzig: 2 pivots
    assert tabs[0]["table"][0][0] == "Base" and tabs[0]["table"][1][0] == "F0"
    # final tableau: objective row's last cell is z*, basic columns are unit vectors
    last = tabs[-1]["table"]
    assert last[4][-1] == 36.0
    assert "<table" in rep["visualizacion_gilp_html"] and "Iteración 2" in rep["visualizacion_gilp_html"]
    assert (tmp_path / "solucion_1.json").exists()
    import json
    json.load(open(tmp_path / "solucion_1.json"))


def test_solver_controller_statuses(golden):
    rep = SolverController(golden["kat"]["K6_infeasible"]["problem"]).run()
    assert rep["solucion_encontrada"]["status"] == "Sin Solucion Factible"
    assert rep["solucion_encontrada"]["valores_variables"] is None and rep["tablas_intermedias"] == []
    rep = SolverController(golden["kat"]["K7_unbounded"]["problem"]).run()
    assert rep["solucion_encontrada"]["status"] == "Error"


def test_pivot_steps_full_tableau_is_consistent(golden):
    """Every displayed tableau satisfies the invariant of a simplex tableau: basic columns are unit vectors and
    B^-1-consistency: T_k = E_k T_{k-1} for the elementary pivot matrix (checked through the RHS/objective)."""
    from simplex_solver_b200.simple_simplex import add_constraint, add_objective, create_tableau, optimize_json_format
    for name in ("K1_wyndor", "K3_min_ge", "K5_eq", "K9_three_var"):
        A, b, c, ops, mx, _ = W.problem_dict_to_arrays(golden["kat"][name]["problem"])
        t = create_tableau(len(c), len(b))
        for i in range(len(b)):
            add_constraint(t, ",".join(str(v) for v in A[i]) + f",{'LGE'[ops[i]]},{b[i]}")
        add_objective(t, ",".join(str(v) for v in c) + f",{1 if mx else 0}")
        js = optimize_json_format(t, maximize=mx)
        steps = js["pivotSteps"]
        assert steps[0]["step"] == 0 and steps[0]["pivotRowIndex"] is None
        m = len(b)
        for k, st in enumerate(steps):
            T = np.array(st["tableau"])
            assert st["step"] == k
            for i, lab in enumerate(st["basis"]):
                colnames = js["columns"]
                assert len(colnames) == T.shape[1]
            if k > 0:
                r, cidx = st["pivotRowIndex"], st["pivotColIndex"]
                prev = np.array(steps[k - 1]["tableau"])
                p = prev[r, cidx]
                exp = prev - np.outer(prev[:, cidx], prev[r] / p)
                exp[r] = prev[r] / p
                np.testing.assert_allclose(T, exp, rtol=1e-12, atol=1e-12, err_msg=f"{name} step {k}")
                assert abs(T[r, cidx] - 1.0) < 1e-12
        z = js["optimalValue"]
        assert abs(z - golden["kat"][name]["z"]) <= REL * max(1.0, abs(z))
        final = np.array(steps[-1]["tableau"])
        assert abs(abs(final[m, -1]) - abs(z)) <= 1e-9 * max(1.0, abs(z))


def test_plugin_patches_the_reference_seam(golden):
    """patch the five names in a stand-in for app.controllers.solver_controller."""
    from simplex_solver_b200 import plugin
    fake = types.ModuleType("app.controllers.solver_controller")
    fake.linprog = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("cpu path"))
    plugin.install(fake)
    try:
        r = fake.linprog([-3.0, -5.0], A_ub=[[1, 0], [0, 2], [3, 2]], b_ub=[4, 12, 18], bounds=[(0, None)] * 2,
                         method="highs-ds", options={"presolve": True, "time_limit": 10})
        assert r.success and abs(r.fun + 36.0) < 1e-9
        tab = fake.create_tableau(number_of_variables=2, number_of_constraints=3)
        fake.add_constraint(tab, "1.0,0.0,L,4.0")
        fake.add_constraint(tab, "0.0,2.0,L,12.0")
        fake.add_constraint(tab, "3.0,2.0,L,18.0")
        fake.add_objective(tab, "3.0,5.0,1")
        js = fake.optimize_json_format(tab, maximize=True)
        assert "pivotSteps" in js and len(js["pivotSteps"]) == 3
    finally:
        plugin.uninstall(fake)
    with pytest.raises(RuntimeError):
        fake.linprog([1.0])


def test_bulk_controller_config2_small(golden):
    A, b, c, ops, mx = W.dense_feasible_lp(128, seed=0)
    out = BulkSolverController(A, b, c, ops, mx).run()
    assert out["status"] == "Solucion Factible"
    zg = golden["dense"]["128"]["z"]
    assert abs(out["valor_optimo_z"] - zg) <= REL * abs(zg)


@pytest.mark.parametrize("n_total", [160, 158])  # 81 stored columns per shard (one padding element per row) / 80 (none)
@pytest.mark.parametrize("rule", [native.RULE_BLAND, native.RULE_DANTZIG])
def test_cuda_shard_engine_world1_and_emulated_world2(oracle, rule, n_total):
    """The shard kernels (candidate / winner / ratio on an external column / update with a remote column) against
    the oracle.  Two shards are emulated in ONE process on one GPU (two solvers, gather by copy), which exercises
    the remote-column path without needing two GPUs."""
    import torch
    m, seed, budget = 96, 4, 80
    one = oracle.OracleTableau.generate(seed, m, n_total)
    ref = one.solve(oracle.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
    opts = native.make_opts(rule=rule, max_pivots=budget)

    eng = CudaShardEngine(m, n_total, 0, n_total, seed)
    status, n = ShardedTableau(eng, 1, 0).run(opts, budget, check_every=16)
    assert status == ref["status"] and n == ref["n_pivots"]
    h = eng.history(budget)
    np.testing.assert_array_equal(h["piv_row"], ref["piv_row"])
    np.testing.assert_array_equal(h["enter_lab"], ref["enter_lab"])
    assert_bit_equal(eng.tableau(), one.T, "world-1 shard tableau")

    # world = 1 again through the look-ahead loop (K = 7: ragged last block)
    eng = CudaShardEngine(m, n_total, 0, n_total, seed)
    status, n = ShardedTableau(eng, 1, 0).run(opts, budget, check_every=21, lookahead=7)
    assert status == ref["status"] and n == ref["n_pivots"]
    np.testing.assert_array_equal(eng.history(budget)["piv_row"], ref["piv_row"])
    assert_bit_equal(eng.tableau(), one.T, "world-1 shard tableau, look-ahead")

    for lookahead in (0, 6):
        for p2p in (False, True):
            _emulated_two_shards(oracle, one, ref, opts, m, n_total, seed, budget, lookahead, p2p)
    _emulated_two_shards(oracle, one, ref, opts, m, n_total, seed, budget, 0, True, world=3)
    _emulated_two_shards(oracle, one, ref, opts, m, n_total, seed, budget, 9, True, world=4)


@pytest.mark.parametrize("p2p", [False, True])
def test_shard_driver_recaptures_chunks_when_the_history_grows(oracle, p2p):
    """A chunk captured by ShardedTableau bakes the pivot-history arrays (sized by max_pivots).  A later run with a larger
    budget and the same chunk size reallocates them: the cached graph must not be replayed (bench.py's e2e step did
    exactly that: 16-pivot steps, then one 128-pivot run -- its history was garbage and the old arrays were written after
    being freed).  b200lp_binding_epoch keys the chunks."""
    import torch
    m, n_total, seed = 96, 300, 4
    eng = CudaShardEngine(m, n_total, 0, n_total, seed)
    if p2p:
        region = torch.zeros(native.Solver.p2p_bytes(m + 1, 1) // 8, dtype=torch.float64, device="cuda:0")
        eng.enable_p2p(1, 0, bases=[region.data_ptr()], region=region)
    drv = ShardedTableau(eng, 1, 0)
    e0 = eng.epoch()
    for budget in (16, 16, 128, 16, 128):
        eng.regenerate()
        opts = native.make_opts(rule=native.RULE_BLAND, max_pivots=budget)
        status, n = drv.run(opts, budget, check_every=16)
        one = oracle.OracleTableau.generate(seed, m, n_total)
        ref = one.solve(oracle.make_opts(rule=oracle.RULE_BLAND, max_pivots=budget), hist_cap=budget)
        assert status == ref["status"] and n == ref["n_pivots"], budget
        h = eng.history(budget)
        np.testing.assert_array_equal(h["piv_row"], ref["piv_row"])
        np.testing.assert_array_equal(h["enter_lab"], ref["enter_lab"])
        np.testing.assert_array_equal(h["leave_lab"], ref["leave_lab"])
        assert_bit_equal(eng.tableau(), one.T, f"tableau, budget {budget}")
    assert eng.epoch() > e0, "growing the history must move the binding epoch"


def _emulated_two_shards(oracle, one, ref, opts, m, n_total, seed, budget, lookahead, p2p, world=2):
    """Two shards in ONE process on one GPU.  p2p: the fused pick kernel with the peer-memory exchange, the regions wired
    by hand, both shards in ONE launch (one cluster each, b200lp_shard_fused_multi): kernels that wait for one another
    are never separate launches on one GPU."""
    import torch
    engs = []
    for r in range(world):
        lo, hi = ShardedTableau.columns_of(n_total, world, r)
        engs.append(CudaShardEngine(m, n_total, lo, hi - lo, seed))
    gathered = engs[0].new_buffer(world)
    if p2p:
        n = native.Solver.p2p_bytes(m + 1, world) // 8
        regions = [torch.zeros(n, dtype=torch.float64, device="cuda:0") for _ in range(world)]
        for r, e in enumerate(engs):
            e.enable_p2p(world, r, bases=[t.data_ptr() for t in regions], region=regions[r])
    for e in engs:
        e.reset(budget)
        if lookahead:
            e.lookahead_begin()
    stride = m + 1 + 2
    for it in range(budget + 2):
        if p2p:  # ONE launch, one cluster per shard: the shards' waits for each other resolve inside the kernel
            native.Solver.shard_fused_multi([e.solver for e in engs], opts, m, lookahead > 0)
        else:
            for r, e in enumerate(engs):
                gathered[r * stride:(r + 1) * stride].copy_(e.candidate(opts, lookahead > 0))
            torch.cuda.synchronize()
            for r, e in enumerate(engs):
                e.pivot(opts, gathered, world, r, lookahead > 0)
        for e in engs:
            if lookahead and (it + 1) % lookahead == 0:
                e.lookahead_flush()
        torch.cuda.synchronize()
    if lookahead:
        for e in engs:
            e.lookahead_flush()
    for r, e in enumerate(engs):
        done, status, n = e.state()
        assert done and status == ref["status"] and n == ref["n_pivots"]
        h = e.history(budget)
        np.testing.assert_array_equal(h["piv_row"], ref["piv_row"])
        np.testing.assert_array_equal(h["enter_lab"], ref["enter_lab"])
        np.testing.assert_array_equal(h["leave_lab"], ref["leave_lab"])
        T = e.tableau()
        rl, cl = e.labels()
        np.testing.assert_array_equal(rl, one.rowlab)
        pos = {int(lab): j for j, lab in enumerate(one.collab[:-1])}
        for j, lab in enumerate(cl[:-1]):
            assert_bit_equal(T[:, j], one.T[:, pos[int(lab)]], f"shard {r} column of variable {lab}")
        assert_bit_equal(T[:, -1], one.T[:, -1], f"shard {r} rhs replica")


def test_thirty_threads_share_the_library(golden):
    """The reference's stress test drives one app from 30 threads (tests/test_performance_load.py:186-199): one
    workspace per thread (native.thread_solver), ctypes releases the GIL, results must all be right."""
    import threading
    wrapper = golden["kat"]["K1_wyndor"]["problem"]
    A, b, c, ops, mx = W.dense_feasible_lp(96, seed=1)
    expect = native.thread_solver(0).solve_dense(A, b, -c, ops)
    errors, results = [], []

    def work(k):
        try:
            for _ in range(5):
                rep = SolverController(wrapper).run()
                z = rep["solucion_encontrada"]["valor_optimo_z"]
                r = linprog(-c, A_ub=A, b_ub=b, bounds=[(0, None)] * 96)
                results.append((abs(z - 36.0) < 1e-9, r.fun == expect["fun"], r.nit == expect["n_pivots"]))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(30)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert len(results) == 150 and all(all(r) for r in results)


def test_concurrent_cooperative_loops_from_threads():
    """The on-chip loop and the look-ahead picks are cooperative launches that want every SM.  Six threads, each with its
    own workspace and stream, run them at the same time: the launches must serialise (not deadlock on each other's grid
    barriers) and every thread must get the pivots it gets when alone."""
    import threading
    A, b, c, ops, mx = W.dense_feasible_lp(160, seed=3)
    modes = [dict(loop_mode=native.LOOP_BLOCKED, check_every=16), dict(loop_mode=native.LOOP_AUTO),
             dict(loop_mode=native.LOOP_BLOCKED, check_every=32)]
    alone = [native.thread_solver(0).solve_dense(A, b, -c, ops, native.make_opts(**mo), hist_cap=4096) for mo in modes]
    assert alone[0]["n_pivots"] == alone[1]["n_pivots"] == alone[2]["n_pivots"] and alone[0]["status"] == 0
    errors, same = [], []

    def work(k):
        try:
            for rep in range(4):
                r = native.thread_solver(0).solve_dense(A, b, -c, ops, native.make_opts(**modes[k % 3]), hist_cap=4096)
                same.append(r["fun"] == alone[0]["fun"] and np.array_equal(r["piv_row"], alone[0]["piv_row"]))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "a thread is stuck"
    assert not errors, errors[:3]
    assert len(same) == 24 and all(same)


def test_time_limit_is_honoured_like_the_reference_maps_it(oracle):
    """solver_controller.py:76 bounds every solve by time_limit and :404 maps the limit status (1) to "Error".  The GPU
    path stops at the bound with status 1 and a CONSISTENT state: the pivots taken so far are the oracle's first pivots and
    the tableau is the oracle's tableau after them -- in the look-ahead loop (2000 x 3000, host check between replays) and
    in the on-chip loop (1024 x 1024, %globaltimer check inside the persistent kernel)."""
    import time
    from simplex_solver_b200.solver_controller import status_text
    for (m, n, what) in ((2000, 3000, "look-ahead loop"), (1024, 1024, "on-chip loop")):
        A, b, c, ops, mx = W.dense_feasible_lp(n, seed=1, m=m)
        s = native.Solver(0)
        # warm-up with the history capacity of the bounded call: the bound covers everything a call does, one-time
        # allocations and the graph capture included (measured: a first call with a new hist_cap costs 30-80 ms)
        s.solve_dense(A, b, -c, ops, hist_cap=1 << 16)
        t0 = time.perf_counter()
        full = s.solve_dense(A, b, -c, ops, hist_cap=1 << 16)
        t_full = time.perf_counter() - t0
        assert full["status"] == 0
        limit = t_full / 3
        t0 = time.perf_counter()
        cut = s.solve_dense(A, b, -c, ops, native.make_opts(time_limit=limit), hist_cap=1 << 16)
        t_cut = time.perf_counter() - t0
        assert cut["status"] == native.STATUS_LIMIT, what
        assert 0 < cut["n_pivots"] < full["n_pivots"], what
        assert t_cut < 0.75 * t_full, f"{what}: {t_cut:.4f} s with time_limit {limit:.4f} s (unbounded solve {t_full:.4f} s)"
        k = cut["n_pivots"]
        ot = oracle.OracleTableau.build(A, b, -c, ops)
        ref = ot.solve(oracle.make_opts(max_pivots=k), hist_cap=k)
        assert np.array_equal(cut["piv_row"][:k], ref["piv_row"]) and np.array_equal(cut["enter_lab"][:k], ref["enter_lab"])
        assert_bit_equal(s.read_tableau(), ot.T, what)
        s.close()
    # through the seam with the reference's own option name
    A, b, c, ops, mx = W.dense_feasible_lp(3000, seed=1, m=2000)
    r = linprog(-c, A_ub=A, b_ub=b, bounds=[(0, None)] * 3000, method="highs-ds",
                options={"presolve": True, "time_limit": 0.005})
    assert r.status == 1 and not r.success and r.x is None and status_text(r) == "Error"
    # and the reference's real value leaves a small problem alone
    r = linprog([-3.0, -5.0], A_ub=[[1, 0], [0, 2], [3, 2]], b_ub=[4, 12, 18], options={"presolve": True, "time_limit": 10})
    assert r.success and abs(r.fun + 36.0) < 1e-9
