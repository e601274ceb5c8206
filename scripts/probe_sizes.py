"""Pivots/s of the three device loops over tableau sizes: where should loop_mode = AUTO switch?"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
s = native.Solver(0)
for R in (1536, 2048, 3072, 4096, 6144, 8192, 12288):
    T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
    row = [f"R={R:6d} ({8 * R * R / 2**20:7.0f} MB)"]
    for name, o in (("graph", dict(loop_mode=native.LOOP_GRAPH)), ("auto", dict(loop_mode=native.LOOP_AUTO)),
                    ("blk8", dict(loop_mode=native.LOOP_BLOCKED, check_every=8)),
                    ("blk16", dict(loop_mode=native.LOOP_BLOCKED, check_every=16)),
                    ("blk32", dict(loop_mode=native.LOOP_BLOCKED, check_every=32))):
        best = 0.0
        for rep in range(2):
            s.generate(4, R - 1, 0)
            r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=256, **o))
            best = max(best, r["n_pivots"] / (r["device_ms"] * 1e-3))
        row.append(f"{name} {best:9.0f}")
    print("  ".join(row), flush=True)
    del T
