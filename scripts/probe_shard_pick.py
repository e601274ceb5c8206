"""One shard of BASELINE config 5 driven alone (world = 1): cost of the pivot DECISION per pivot on 131072-row shards --
the fused peer-memory kernel (k_shard_pick, exchange with itself) against the all-gather chain (price, extract, copy,
winner, ratio) -- next to the update kernel alone.  PROBE_C = stored structural columns (16383: 8 GPUs, 65535: 2 GPUs)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from simplex_solver_b200 import native
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau
R = int(os.environ.get("PROBE_R", "131072"))
n = int(os.environ.get("PROBE_PIVOTS", "16"))
modes = os.environ.get("PROBE_MODES", "p2p,gather").split(",")
for ncols in [int(x) for x in os.environ.get("PROBE_C", "16383,65535").split(",")]:
    with torch.cuda.stream(torch.cuda.Stream()):
        eng = CudaShardEngine(R - 1, ncols, 0, ncols, 4, device=0)
        upd = eng.solver.time_update(1, 1, native.UPDATE_AUTO, 5)
        print(f"C = {ncols + 1}: update kernel alone {upd:.3f} ms", flush=True)
        for mode in modes:
            if mode == "p2p":
                region = torch.zeros(native.Solver.p2p_bytes(R, 1) // 8, dtype=torch.float64, device="cuda:0")
                eng.enable_p2p(1, 0, bases=[region.data_ptr()], region=region)
            else:
                eng.p2p = False
            drv = ShardedTableau(eng, 1, 0)
            opts = native.make_opts(rule=native.RULE_BLAND, max_pivots=n)
            for K in (0, 32):
                nn = n if K == 0 else 2 * K
                o = native.make_opts(rule=native.RULE_BLAND, max_pivots=nn)
                eng.regenerate()
                drv.run(o, nn, check_every=nn if K == 0 else K, lookahead=K)
                eng.regenerate()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _, got = drv.run(o, nn, check_every=nn if K == 0 else K, lookahead=K)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / max(got, 1)
                print(f"  {mode:6s} {'rank-1' if K == 0 else 'look-ahead K=32'}: {ms:.3f} ms/pivot"
                      + (f" (decision {ms - upd:.3f} ms)" if K == 0 else ""), flush=True)
        del drv, eng
        torch.cuda.empty_cache()
