"""Look-ahead loop pivots/s (K = 32) and a checksum; used to A/B changes of the pick and flush kernels (PROBE_K=8,16,32)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
s = native.Solver(0)
R = 16384
T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
for K in [int(x) for x in os.environ.get("PROBE_K", "8,16,32").split(",")]:
    o = dict(loop_mode=native.LOOP_BLOCKED, check_every=K)
    s.generate(4, R - 1, 0)
    s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=64, **o))
    best = 0.0
    for rep in range(3):
        r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=512, **o))
        best = max(best, r["n_pivots"] / (r["device_ms"] * 1e-3))
    print("K", K, "pivots/s", round(best, 1), "us/pivot", round(1e6 / best, 2), "fun", repr(r["fun"]), flush=True)
