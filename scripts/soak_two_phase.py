"""Soak: b200lp_solve_dense on random two-phase LPs (<=, >=, = rows; some infeasible, some unbounded) through the three
device loops.  Status, pivot history, x and fun must be identical bit for bit."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from simplex_solver_b200 import native

rng = np.random.default_rng(int(os.environ.get("SOAK_SEED", "1")))
s = native.Solver(0)
N = int(os.environ.get("SOAK_N", "60"))
bad = 0
stats = {}
for case in range(N):
    m = int(rng.integers(2, 700))
    n = int(rng.integers(2, 900))
    A = rng.uniform(-1.0, 1.0, (m, n))
    x0 = rng.random(n)
    u = rng.random(m)
    ops = np.where(u < 0.6, 0, np.where(u < 0.85, 1, 2)).astype(np.int8)
    slack = rng.uniform(0.1, 1.0, m)
    b = A @ x0 + np.where(ops == 0, slack, np.where(ops == 1, -slack, 0.0))
    c = rng.uniform(0.1, 1.0, n)
    if case % 9 == 3:   # infeasible pair
        A[0] = 0.0; A[1] = 0.0; A[0, 0] = 1.0; A[1, 0] = 1.0; ops[0] = 0; ops[1] = 1; b[0] = 5.0; b[1] = 10.0
    if case % 9 == 5:   # unbounded
        A = np.abs(A); ops[:] = 1; c = -c
    rule = int(rng.integers(0, 2))
    K = int(rng.integers(1, 33))
    cap = 1 << 16
    res = []
    for o in (dict(loop_mode=native.LOOP_GRAPH), dict(loop_mode=native.LOOP_AUTO),
              dict(loop_mode=native.LOOP_BLOCKED, check_every=K)):
        res.append(s.solve_dense(A, b, c, ops, native.make_opts(rule=rule, **o), hist_cap=cap))
    g = res[0]
    stats[g["status"]] = stats.get(g["status"], 0) + 1
    for name, r in zip(("auto", "blocked"), res[1:]):
        ok = (r["status"] == g["status"] and r["n_pivots"] == g["n_pivots"] and r["n_phase1"] == g["n_phase1"]
              and np.array_equal(r["piv_row"], g["piv_row"]) and np.array_equal(r["piv_col"], g["piv_col"]))
        if ok and g["status"] == 0:
            ok = r["fun"] == g["fun"] and np.array_equal(r["x"].view(np.int64), g["x"].view(np.int64))
        if not ok:
            bad += 1
            print("MISMATCH", name, dict(case=case, m=m, n=n, rule=rule, K=K, st=(g["status"], r["status"]),
                                         np=(g["n_pivots"], r["n_pivots"])), flush=True)
print(f"soak two-phase: {N} cases, statuses {stats}, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
