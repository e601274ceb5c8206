"""On-GPU probe: update kernel bandwidth vs tableau shape / footprint, next to a torch copy of the same footprint."""
import os, sys, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
s = native.Solver(0)
out = {}
for R, C in ((16384, 16384), (131072, 16384), (16384, 131072), (65536, 65536), (131072, 65536)):
    T = torch.empty(R * C, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), R - 1, 1, C, C, C - 1, C + R - 2, keep=T)
    s.generate(4, C - 1, 0); s.synchronize()
    b = 16.0 * R * C
    res = {}
    for name, v in (("ldg", native.UPDATE_LDG), ("tma", native.UPDATE_TMA)):
        ms = min(s.time_update(100, 200, v, 4) for _ in range(2))
        res[name] = b / (ms * 1e-3) / 1e9
    if R * C * 8 <= 40e9:
        U = torch.empty(R * C, dtype=torch.float64, device="cuda:0")
        U.copy_(T); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): U.copy_(T)
        e1.record(); torch.cuda.synchronize()
        res["torch_copy"] = b / (e0.elapsed_time(e1) / 4 * 1e-3) / 1e9
        del U
    out[f"{R}x{C}"] = res
    print(R, C, res, flush=True)
    del T
    s.attach(torch.empty(64, dtype=torch.float64, device="cuda:0").data_ptr(), 1, 1, 2, 2, 1, 2)
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/probe_shapes.json", "w"), indent=1)
