"""On-GPU probe: pivot-update kernel variants at BASELINE config 4 size (kernel alone, CUDA events)."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from simplex_solver_b200 import native
from oracle import oracle as O

out = {}
s = native.Solver(0)
# correctness of every variant on a small ragged tableau first
m, n = 333, 777
for v in (1, 2):
    ld = (n + 1 + 15) // 16 * 16
    T = torch.empty((m + 1) * ld, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), m, 1, n + 1, ld, n, n + m, keep=T)
    s.generate(4, n, 0)
    got = s.run(native.make_opts(rule=native.RULE_DANTZIG, max_pivots=40, update_variant=v), hist_cap=40)
    ot = O.OracleTableau.generate(4, m, n)
    ref = ot.solve(O.make_opts(rule=0, max_pivots=40), hist_cap=40)
    ok = bool((got["piv_row"] == ref["piv_row"]).all() and np.array_equal(s.read_tableau(), ot.T))
    print("variant", v, "bit-exact:", ok, flush=True)
    out[f"exact_v{v}"] = ok
R = 16384
T = torch.empty(R * R, dtype=torch.float64, device="cuda:0")
s.attach(T.data_ptr(), R - 1, 1, R, R, R - 1, 2 * R - 2, keep=T)
s.generate(4, R - 1, 0); s.synchronize()
bytes_per = 16.0 * R * R
for v in (1, 2):
    ms = [s.time_update(100, 200, v, 10) for _ in range(3)]
    out[f"v{v}"] = {"ms": min(ms), "GBps": bytes_per / (min(ms) * 1e-3) / 1e9}
    print("variant", v, out[f"v{v}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_variants.json", "w"), indent=1)
