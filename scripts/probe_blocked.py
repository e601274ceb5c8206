import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
s = native.Solver(0)
R = 16384
T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
for name, o in (("rank-1 graph", dict(loop_mode=native.LOOP_GRAPH)),
                ("blocked K=4", dict(loop_mode=native.LOOP_BLOCKED, check_every=4)),
                ("blocked K=8", dict(loop_mode=native.LOOP_BLOCKED, check_every=8)),
                ("blocked K=16", dict(loop_mode=native.LOOP_BLOCKED, check_every=16)),
                ("blocked K=24", dict(loop_mode=native.LOOP_BLOCKED, check_every=24)),
                ("blocked K=32", dict(loop_mode=native.LOOP_BLOCKED, check_every=32))):
    s.generate(4, R - 1, 0)
    s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=64, **o))
    r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=768, **o))
    pps = r["n_pivots"] / (r["device_ms"] * 1e-3)
    print(name, "pivots/s", round(pps, 1), "ms/pivot", round(r["device_ms"] / r["n_pivots"], 4), "fun", r["fun"], flush=True)
