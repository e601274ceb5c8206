import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
s = native.Solver(0)
R = 16384
T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
s.generate(4, R - 1, 0)
r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=4 * K, loop_mode=native.LOOP_BLOCKED, check_every=K))
print(r["n_pivots"], r["device_ms"])
