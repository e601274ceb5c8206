"""Quick on-GPU probe: pivot-update kernel variants at BASELINE config 4 size, and the config 2 solve."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from simplex_solver_b200 import native, workloads as W

def main():
    out = {}
    s = native.Solver(0)
    for R in (16384,):
        m = n = R - 1
        C = n + 1; ld = C
        T = torch.empty(R * ld, dtype=torch.float64, device="cuda:0")
        s.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=T)
        s.generate(4, n, 0); s.synchronize()
        bytes_per = 2.0 * R * C * 8
        for name, v in (("ldg", native.UPDATE_LDG), ("tma", native.UPDATE_TMA)):
            ms = [s.time_update(100, 200, v, 10) for _ in range(3)]
            out[f"update_{name}_{R}"] = {"ms": ms, "GBps": [bytes_per / (t * 1e-3) / 1e9 for t in ms]}
            print(name, ms, out[f"update_{name}_{R}"]["GBps"], flush=True)
        s.generate(4, n, 0)
        for name, v in (("ldg", native.UPDATE_LDG), ("tma", native.UPDATE_TMA)):
            r = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=200, update_variant=v))
            out[f"loop_{name}_{R}"] = {"ms": r["device_ms"], "pivots_per_s": 200 / (r["device_ms"] * 1e-3),
                                      "GBps": 200 * bytes_per / (r["device_ms"] * 1e-3) / 1e9, "launches": r["kernel_launches"]}
            print("loop", name, out[f"loop_{name}_{R}"], flush=True)
        # torch copy reference on the same memory size
        a = torch.empty(R * ld, dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.copy_(T); torch.cuda.synchronize()
        e0.record()
        for _ in range(10): a.copy_(T)
        e1.record(); torch.cuda.synchronize()
        out["torch_copy_GBps"] = bytes_per / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9
        print("torch copy GB/s", out["torch_copy_GBps"], flush=True)
        del a, T
    A, b, c, ops, mx = W.dense_feasible_lp(1024, 0)
    for name, v in (("ldg", native.UPDATE_LDG), ("tma", native.UPDATE_TMA)):
        for graph in (True, False):
            t = time.time()
            r = s.solve_dense(A, b, -c, ops, native.make_opts(update_variant=v, use_graph=graph))
            dt = time.time() - t
            out[f"dense1024_{name}_graph{int(graph)}"] = {"wall_s": dt, "device_ms": r["device_ms"], "pivots": r["n_pivots"], "z": -r["fun"],
                                                          "us_per_pivot": r["device_ms"] * 1e3 / max(r["n_pivots"], 1)}
            print(name, graph, out[f"dense1024_{name}_graph{int(graph)}"], flush=True)
    Ab, bb, cb, ob = W.batched_small_lps(0, 100000)
    for _ in range(2):
        t = time.time(); r = s.solve_batched(Ab, bb, cb, ob, want_x=True); dt = time.time() - t
        out["batched_100k"] = {"wall_s": dt, "device_ms": r["device_ms"], "LPs_per_s_kernel": 1e5 / (r["device_ms"] * 1e-3),
                               "pivots": int(r["n_pivots"].sum())}
        print("batched", out["batched_100k"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_update.json", "w"), indent=1)

main()
