import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from simplex_solver_b200 import native, workloads as W
s = native.Solver(0)
B = 100000
for (m, n) in ((20, 30), (6, 10), (12, 20)):
    A, b, c, ops = W.batched_small_lps(0, B, m, n)
    dev = [torch.from_numpy(a).cuda() for a in (A, b, c, ops)]
    out = [torch.empty(B, dtype=torch.int32, device="cuda"), torch.empty(B, dtype=torch.float64, device="cuda"),
           torch.empty((B, n), dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda")]
    ms = min(s.solve_batched_device(B, m, n, *[d.data_ptr() for d in dev], *[o.data_ptr() for o in out]) for _ in range(3))
    piv = int(out[3].sum().item())
    print(os.environ.get("B200LP_BATCHED_SMEM", "reg"), (m, n), "kernel ms", round(ms, 3), "MLPs/s", round(B / ms / 1e3, 2), "Mpivots/s", round(piv / ms / 1e3, 1), flush=True)
