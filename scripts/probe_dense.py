"""On-GPU probe: BASELINE config 2 family through every loop driver."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from simplex_solver_b200 import native, workloads as W
s = native.Solver(0)
out = {}
for n in (256, 512, 1024, 1536):
    A, b, c, ops, mx = W.dense_feasible_lp(n, 0)
    for name, mode in (("auto/onchip", native.LOOP_AUTO), ("graph", native.LOOP_GRAPH)):
        s.solve_dense(A, b, -c, ops, native.make_opts(loop_mode=mode, max_pivots=50))
        t = time.perf_counter(); r = s.solve_dense(A, b, -c, ops, native.make_opts(loop_mode=mode)); dt = time.perf_counter() - t
        out[f"{n}_{name}"] = {"wall_ms": dt * 1e3, "device_ms": r["device_ms"], "pivots": r["n_pivots"], "z": -r["fun"],
                              "us_per_pivot": r["device_ms"] * 1e3 / r["n_pivots"], "launches": r["kernel_launches"]}
        print(n, name, out[f"{n}_{name}"], flush=True)
json.dump(out, open("gpurun_out/probe_dense.json", "w"), indent=1)
