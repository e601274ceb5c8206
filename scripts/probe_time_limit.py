"""time_limit on the on-chip loop: where does the time of a bounded solve_dense go (diagnostic for the GPU test)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from simplex_solver_b200 import native, workloads as W
for (m, n) in ((1024, 1024), (2000, 3000)):
    A, b, c, ops, mx = W.dense_feasible_lp(n, seed=1, m=m)
    s = native.Solver(0)
    s.solve_dense(A, b, -c, ops)
    for rep in range(2):
        t0 = time.perf_counter()
        full = s.solve_dense(A, b, -c, ops)
        t_full = time.perf_counter() - t0
        print(m, n, "full", round(t_full * 1e3, 2), "ms wall", round(full["device_ms"], 2), "ms device", full["n_pivots"], "pivots", flush=True)
    t0 = time.perf_counter()
    s.build_dense(A, b, -c, ops)
    print("  build alone", round((time.perf_counter() - t0) * 1e3, 2), "ms")
    for frac in (0.1, 0.25, 0.5, 0.75):
        for hc in (0, 1 << 16):
            t0 = time.perf_counter()
            cut = s.solve_dense(A, b, -c, ops, native.make_opts(time_limit=t_full * frac), hist_cap=hc)
            t_cut = time.perf_counter() - t0
            print("  limit", round(t_full * frac * 1e3, 2), "ms hist_cap", hc, "-> status", cut["status"], "pivots", cut["n_pivots"],
                  "wall", round(t_cut * 1e3, 2), "ms", flush=True)
    s.close()
