"""On-GPU probe: how much slower do the look-ahead picks (latency-bound, grid barriers through L2) get while another
stream saturates HBM?  Decides whether overlapping the picks of block b+1 with the flush of block b can pay."""
import os, sys, threading, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native

def tableau(s, n):
    T = torch.empty(n * n, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), n - 1, 1, n, n, n - 1, 2 * n - 2, keep=T)
    s.generate(4, n - 1, 0)
    return T

s1, s2 = native.Solver(0), native.Solver(0)
T1 = tableau(s1, 16384)
for n2 in (4096, 8192):
    T2 = tableau(s2, n2)
    o = native.make_opts(rule=native.RULE_BLAND, max_pivots=4096, loop_mode=native.LOOP_BLOCKED, check_every=32)
    s2.run(o)
    res = {}
    for load in (False, True):
        stop = threading.Event()
        def bg():
            while not stop.is_set():
                s1.time_update(1, 1, native.UPDATE_LDG, 200)
        th = threading.Thread(target=bg)
        if load:
            th.start()
            time.sleep(0.05)
        s2.generate(4, n2 - 1, 0)
        r = s2.run(o)
        res[load] = r["n_pivots"] / (r["device_ms"] * 1e-3)
        if load:
            stop.set()
            th.join()
    print(f"{n2}^2 look-ahead K=32: {res[False]:.0f} pivots/s alone, {res[True]:.0f} pivots/s while another stream streams 16384^2 updates "
          f"({res[False] / res[True]:.2f}x slower)", flush=True)
    del T2
