"""ncu target: a few REAL pivots of the device loop on the BASELINE config 4 tableau (no no-op launches).

    python scripts/profile_update.py [ldg|tma] [rows]
"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native

variant = {"ldg": native.UPDATE_LDG, "tma": native.UPDATE_TMA}[sys.argv[1] if len(sys.argv) > 1 else "ldg"]
R = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
s = native.Solver(0)
T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
s.generate(4, R - 1, 0)
r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=6, update_variant=variant, check_every=6, loop_mode=native.LOOP_LAUNCHES))
print(r["n_pivots"], r["device_ms"])
