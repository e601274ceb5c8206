"""Look-ahead loop with the flush variant selected by B200LP_FLUSH (read once per process): pivots/s and a checksum."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
s = native.Solver(0)
R = int(os.environ.get("PROBE_R", "16384"))
T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
for K in (16, 32):
    o = dict(loop_mode=native.LOOP_BLOCKED, check_every=K)
    s.generate(4, R - 1, 0)
    s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=64, **o))
    r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=768, **o))
    pps = r["n_pivots"] / (r["device_ms"] * 1e-3)
    torch.cuda.synchronize()
    print("flush", os.environ.get("B200LP_FLUSH", "0"), "K", K, "pivots/s", round(pps, 1), "us/pivot",
          round(1e3 * r["device_ms"] / r["n_pivots"], 2), "fun", repr(r["fun"]), "sum", repr(float(T.sum())), flush=True)
