"""Multi-GPU parity check of the column-sharded path over NCCL (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/nccl_sharded_check.py

Every rank's shard must hold, bit for bit, the columns of the single-tableau oracle run, and all ranks must log the
same pivot sequence as the oracle.  (The pytest suite covers the same protocol on CPU with gloo and, on one GPU,
with two emulated shards; this script is the real-NCCL leg.)"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import numpy as np
import torch
import torch.distributed as dist

from oracle import oracle as O
from simplex_solver_b200 import native
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    with torch.cuda.stream(torch.cuda.Stream()):
        for rule, lookahead, p2p in ((native.RULE_BLAND, 0, False), (native.RULE_DANTZIG, 0, False),
                                     (native.RULE_BLAND, 8, False), (native.RULE_DANTZIG, 12, False),
                                     (native.RULE_BLAND, 0, True), (native.RULE_DANTZIG, 8, True)):
            # look-ahead cases use shards with an EVEN number of stored columns (no padding element at the end of a row),
            # so an access one column outside a shard would land in real data of the neighbouring row
            m, n_total, seed, budget = 384, (1024 - world if lookahead else 1024), 4, 200
            lo, hi = ShardedTableau.columns_of(n_total, world, rank)
            eng = CudaShardEngine(m, n_total, lo, hi - lo, seed, device=local)
            if p2p:
                eng.enable_p2p(world, rank)
            drv = ShardedTableau(eng, world, rank)
            opts = native.make_opts(rule=rule, max_pivots=budget)
            status, n = drv.run(opts, budget, check_every=48, lookahead=lookahead)
            one = O.OracleTableau.generate(seed, m, n_total)
            ref = one.solve(O.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
            h = eng.history(budget)
            T = eng.tableau()
            rl, cl = eng.labels()
            pos = {int(lab): j for j, lab in enumerate(one.collab[:-1])}
            good = (status == ref["status"] and n == ref["n_pivots"]
                    and np.array_equal(h["piv_row"], ref["piv_row"]) and np.array_equal(h["enter_lab"], ref["enter_lab"])
                    and np.array_equal(h["leave_lab"], ref["leave_lab"]) and np.array_equal(rl, one.rowlab)
                    and all(np.array_equal(T[:, j], one.T[:, pos[int(lab)]]) for j, lab in enumerate(cl[:-1]))
                    and np.array_equal(T[:, -1], one.T[:, -1]))
            print(f"rank {rank} rule {rule} lookahead {lookahead} p2p {p2p}: status {status} pivots {n} bit-exact vs oracle: {good}", flush=True)
            ok = ok and good
            del drv, eng
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("NCCL sharded parity OK")


if __name__ == "__main__":
    main()
