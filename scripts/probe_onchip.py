import os, sys, time, subprocess, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
if len(sys.argv) > 1:
    import numpy as np
    from simplex_solver_b200 import native, workloads as W
    s = native.Solver(0)
    for n in (256, 1024):
        A, b, c, ops, mx = W.dense_feasible_lp(n, 0)
        s.solve_dense(A, b, -c, ops, native.make_opts(max_pivots=50))
        r = s.solve_dense(A, b, -c, ops)
        print(sys.argv[1], n, "us/pivot", r["device_ms"] * 1e3 / r["n_pivots"], flush=True)
else:
    for g in (32, 48, 74, 96, 128, 148):
        env = dict(os.environ, B200LP_ONCHIP_CTAS=str(g))
        subprocess.run([sys.executable, __file__, str(g)], env=env)
