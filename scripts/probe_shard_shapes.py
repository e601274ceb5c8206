"""On-GPU probe: pivot-update kernel on the shard shapes of BASELINE config 5 at 2 / 4 / 8 GPUs -- the shards that store
exactly cols_total / world columns and the LAST shard, which stores world - 1 more (ragged last strip, row stride not a
power of two)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native

R = 131072
s = native.Solver(0)
for world in (2, 4, 8):
    per = R // world
    for C, ld in ((per, per), (per + world - 1, (per + world - 1 + 15) // 16 * 16), (per + world - 1, per + 256)):
        T = torch.empty(R * ld, dtype=torch.float64, device="cuda:0")
        s.attach(T.data_ptr(), R - 1, 1, C, ld, R - 1, 2 * R - 2, keep=T)
        s.generate(4, R - 1, 0)
        out = []
        for name, v in (("ldg", native.UPDATE_LDG), ("tma", native.UPDATE_TMA)):
            ms = s.time_update(1, 1, v, 5)
            out.append(f"{name} {ms:.3f} ms = {16.0 * R * C / ms / 1e6:.0f} GB/s")
        print(f"world {world}: C = {C}, ld = {ld}: " + ", ".join(out), flush=True)
        del T
        torch.cuda.empty_cache()
