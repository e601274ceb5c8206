"""One shard of BASELINE config 5 at 8 GPUs (131072 rows x 16384 columns + RHS) driven alone (world = 1) through the
sharded look-ahead / rank-1 drivers: per-kernel times under `ncu --metrics gpu__time_duration.sum`, or wall numbers."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau
R = int(os.environ.get("PROBE_R", "131072"))
ncols = int(os.environ.get("PROBE_C", "16383"))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * K
with torch.cuda.stream(torch.cuda.Stream()):
    eng = CudaShardEngine(R - 1, ncols, 0, ncols, 4, device=0)
    drv = ShardedTableau(eng, 1, 0)
    opts = native.make_opts(rule=native.RULE_BLAND, max_pivots=n)
    drv.run(opts, n, check_every=K, lookahead=K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, got = drv.run(opts, n, check_every=K, lookahead=K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("K", K, "pivots", got, "us/pivot", round(dt / max(got, 1) * 1e6, 1), flush=True)
