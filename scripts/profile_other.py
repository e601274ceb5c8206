"""ncu targets besides the config-4 update kernel:  python scripts/profile_other.py [batched|onchip|tma]"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from simplex_solver_b200 import native, workloads as W
what = sys.argv[1]
s = native.Solver(0)
if what == "batched":
    B = 100000
    A, b, c, ops = W.batched_small_lps(0, B)
    dev = [torch.from_numpy(a).cuda() for a in (A, b, c, ops)]
    out = [torch.empty(B, dtype=torch.int32, device="cuda"), torch.empty(B, dtype=torch.float64, device="cuda"),
           torch.empty((B, 30), dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda")]
    for _ in range(2):
        ms = s.solve_batched_device(B, 20, 30, *[d.data_ptr() for d in dev], *[o.data_ptr() for o in out])
    print("batched kernel ms", ms)
elif what == "onchip":
    A, b, c, ops, mx = W.dense_feasible_lp(512, 0)
    r = s.solve_dense(A, b, -c, ops)
    print("onchip", r["n_pivots"], r["device_ms"])
else:
    R = 16384
    T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
    s.generate(4, R - 1, 0)
    r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=6, update_variant=native.UPDATE_TMA, check_every=6,
                               loop_mode=native.LOOP_LAUNCHES))
    print(r["n_pivots"], r["device_ms"])
