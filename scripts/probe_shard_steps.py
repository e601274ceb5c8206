"""torchrun probe: where does a 16-pivot step of the sharded rank-1 loop spend its time (host timestamps per call and
CUDA events per graph replay), with and without the start-of-run barrier and the history read."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as dist
from simplex_solver_b200 import native
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau

rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank)); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
R = int(os.environ.get("PROBE_R", "131072")); CT = int(os.environ.get("PROBE_CT", "131072")); P = int(os.environ.get("PROBE_PIVOTS", "16"))
with torch.cuda.stream(torch.cuda.Stream()):
    lo, hi = ShardedTableau.columns_of_tableau(CT, world, rank)
    eng = CudaShardEngine(R - 1, CT - 1, lo, hi - lo, 4, device=local)
    eng.enable_p2p(world, rank)
    drv = ShardedTableau(eng, world, rank)
    opts = native.make_opts(rule=native.RULE_BLAND, max_pivots=P)
    for _ in range(2):
        drv.run(opts, P, check_every=P)
    upd = eng.solver.time_update(1, 1, native.UPDATE_AUTO, 5)
    torch.cuda.synchronize(); dist.barrier()
    # (1) the driver as bench.py calls it
    for rep in range(3):
        t0 = time.perf_counter()
        drv.run(opts, P, check_every=P)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.history(P)
        t2 = time.perf_counter()
        print(f"[{rank}] drv.run {1e3 * (t1 - t0):.2f} ms ({1e3 * (t1 - t0) / P:.3f} per pivot; update alone {upd:.3f}), history {1e3 * (t2 - t1):.2f} ms", flush=True)
    # (2) the same, by hand, with timestamps
    key = next(iter(drv._graphs)); g = drv._graphs[key]
    for rep in range(3):
        t0 = time.perf_counter(); eng.reset(P); t1 = time.perf_counter(); dist.barrier(); t2 = time.perf_counter()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); g.replay(); e[1].record(); t3 = time.perf_counter()
        st1 = eng.state(); t4 = time.perf_counter()
        g.replay(); e[2].record(); st2 = eng.state(); t5 = time.perf_counter()
        print(f"[{rank}] reset {1e3 * (t1 - t0):.2f} barrier {1e3 * (t2 - t1):.2f} replay-call {1e3 * (t3 - t2):.2f} state1 {1e3 * (t4 - t3):.2f} "
              f"(graph1 {e[0].elapsed_time(e[1]):.2f} ms = {e[0].elapsed_time(e[1]) / P:.3f}/pivot) replay2+state {1e3 * (t5 - t4):.2f} (graph2 {e[1].elapsed_time(e[2]):.2f} ms) {st1} {st2}", flush=True)
    # (3) back-to-back replays without any host step between them: 4 x P pivots
    eng.reset(4 * P); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"[{rank}] 4 replays back to back: {e0.elapsed_time(e1) / (4 * P):.3f} ms/pivot", flush=True)
dist.barrier()
dist.destroy_process_group()
