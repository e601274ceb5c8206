"""Soak: the look-ahead loop against the rank-1 graph loop on random shapes / block sizes / rules / row strides.
Both run on the GPU from the same generated tableau; pivots and every tableau bit must agree."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
from simplex_solver_b200 import native

rng = np.random.default_rng(int(os.environ.get("SOAK_SEED", "1")))
s = native.Solver(0)
bad = 0
N = int(os.environ.get("SOAK_N", "80"))
for case in range(N):
    m = int(rng.integers(3, 2600))
    n = int(rng.integers(3, 2600))
    if case % 7 == 0:
        m, n = int(rng.integers(3, 200)), int(rng.integers(3, 200))
    if case % 11 == 0:
        n = int(rng.integers(2600, 9000))
    K = int(rng.integers(1, 33))
    rule = int(rng.integers(0, 2))
    pad = int(rng.integers(0, 3)) * 2
    budget = int(rng.integers(1, 140))
    C = n + 1
    ld = C + (C & 1) + pad
    T = torch.empty((m + 1) * ld, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=T)
    seed = int(rng.integers(0, 1 << 30))
    s.generate(seed, n, 0)
    a = s.run(native.make_opts(rule=rule, max_pivots=budget, loop_mode=native.LOOP_GRAPH), hist_cap=budget)
    torch.cuda.synchronize()
    Ta = T.clone()
    torch.cuda.synchronize()  # the copy runs on torch's stream, the generator below on the solver's
    s.generate(seed, n, 0)
    b = s.run(native.make_opts(rule=rule, max_pivots=budget, loop_mode=native.LOOP_BLOCKED, check_every=K), hist_cap=budget)
    torch.cuda.synchronize()
    # compare the stored columns only (the padding beyond C is scratch)
    Va, Vb = Ta.view(m + 1, ld)[:, :C], T.view(m + 1, ld)[:, :C]
    ok = (a["status"] == b["status"] and a["n_pivots"] == b["n_pivots"] and np.array_equal(a["piv_row"], b["piv_row"])
          and np.array_equal(a["piv_col"], b["piv_col"]) and bool(torch.equal(Va.contiguous().view(torch.int64), Vb.contiguous().view(torch.int64))))
    if not ok:
        bad += 1
        print("MISMATCH", dict(case=case, m=m, n=n, K=K, rule=rule, pad=pad, budget=budget, seed=seed,
                               st=(a["status"], b["status"]), np=(a["n_pivots"], b["n_pivots"])), flush=True)
    del T, Ta
print(f"soak: {N} cases, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
