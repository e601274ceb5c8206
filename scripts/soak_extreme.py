"""Soak: extreme aspect ratios and degenerate sizes through b200lp_solve_dense in every loop mode, against the CPU oracle
(status, pivot history, x, fun bit for bit).  Shapes the regular soaks do not draw: one row, one column, very wide, very
tall, mid-size two-phase beyond the on-chip limit."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from oracle import oracle as O
from simplex_solver_b200 import native

rng = np.random.default_rng(7)
s = native.Solver(0)
shapes = [(1, 1), (1, 2), (2, 1), (1, 5000), (5000, 1), (3, 40000), (40000, 3), (7, 100001), (30011, 17), (64, 64),
          (1500, 2500), (2500, 1500), (33, 70000)]
bad = 0
for (m, n) in shapes:
    for kind in ("le", "mixed"):
        A = rng.uniform(-1.0 if kind == "mixed" else 0.0, 1.0, (m, n))
        x0 = rng.random(n)
        if kind == "mixed":
            u = rng.random(m)
            ops = np.where(u < 0.6, 0, np.where(u < 0.85, 1, 2)).astype(np.int8)
        else:
            ops = np.zeros(m, dtype=np.int8)
        slack = rng.uniform(0.1, 1.0, m)
        b = A @ x0 + np.where(ops == 0, slack, np.where(ops == 1, -slack, 0.0))
        c = rng.uniform(0.1, 1.0, n) * (-1.0 if kind == "le" else 1.0)   # le: maximise a bounded LP; mixed: minimise
        budget = 400
        ref = O.solve_lp(A, b, c, ops, O.make_opts(rule=O.RULE_BLAND, max_pivots=budget), hist_cap=budget)
        for name, o in (("graph", dict(loop_mode=native.LOOP_GRAPH)), ("auto", dict(loop_mode=native.LOOP_AUTO)),
                        ("blocked", dict(loop_mode=native.LOOP_BLOCKED, check_every=int(rng.integers(1, 33))))):
            r = s.solve_dense(A, b, c, ops, native.make_opts(rule=native.RULE_BLAND, max_pivots=budget, **o), hist_cap=budget)
            ok = (r["status"] == ref["status"] and r["n_pivots"] == ref["n_pivots"]
                  and np.array_equal(r["piv_row"], ref["piv_row"]) and np.array_equal(r["enter_lab"], ref["enter_lab"]))
            if ok and ref["status"] == 0:
                ok = r["fun"] == ref["fun"] and np.array_equal(r["x"], ref["x"])
            if not ok:
                bad += 1
                print("MISMATCH", (m, n), kind, name, (ref["status"], r["status"]), (ref["n_pivots"], r["n_pivots"]), flush=True)
        print((m, n), kind, "status", ref["status"], "pivots", ref["n_pivots"], flush=True)
print(f"soak extreme: {len(shapes) * 2} LPs x 3 loops, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
