"""Parity cases of the column-sharded path against the oracle, one rank per GPU (torch.distributed must be initialised
with NCCL and the current device set).  Used by bench.py before its timed region at N > 1 (the line's "parity" block),
by tests/nccl_sharded_check.py under torchrun and, through that script, by the `-m gpu` test
tests/test_gpu_multi.py on boxes with >= 2 GPUs.

Every rank's shard must hold, bit for bit, the columns of the single-tableau oracle run, and all ranks must log the
oracle's pivot sequence (row, entering id, leaving id)."""
import numpy as np

from oracle import oracle as O
from simplex_solver_b200 import native
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau

# (rule, look-ahead K, peer-memory exchange, rows m, structural variables n_total)
CASES = (
    (native.RULE_BLAND, 0, False, 384, 1024),
    (native.RULE_DANTZIG, 0, False, 384, 1024),
    (native.RULE_BLAND, 8, False, 384, 1022),     # even stored columns per shard at 2 GPUs: no padding element behind a row
    (native.RULE_DANTZIG, 12, False, 384, 1022),
    (native.RULE_BLAND, 0, True, 384, 1024),
    (native.RULE_DANTZIG, 8, True, 384, 1022),
    (native.RULE_BLAND, 0, True, 301, 1003),      # ragged: odd column counts per shard, rows no multiple of any tile
    (native.RULE_DANTZIG, 32, True, 301, 1003),
    (native.RULE_BLAND, 5, False, 97, 211),
)


def run_cases(rank, world, local, log=None, cases=CASES, seed=4, budget=200):
    """Returns (number of cases, all ok on THIS rank).  The caller reduces `ok` over the ranks."""
    import torch
    ok = True
    with torch.cuda.stream(torch.cuda.Stream()):
        for rule, lookahead, p2p, m, n_total in cases:
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
            if log:
                log(f"rank {rank} rule {rule} lookahead {lookahead} p2p {p2p} {m}x{n_total}: status {status} pivots {n} "
                    f"bit-exact vs oracle: {good}")
            ok = ok and good
            del drv, eng
    return len(cases), ok
