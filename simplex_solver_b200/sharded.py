"""One huge tableau, column-sharded over the GPUs of a node (BASELINE config 5).

Shard g owns a contiguous block of the structural columns of every row (objective row included) plus its own
replica of the right-hand-side column; row labels are replicated.  Per pivot there is ONE exchange:

    candidate  : each shard prices its slice of the objective row and publishes
                 [best reduced cost, variable id, its candidate column (R doubles)]          (device kernels)
    all-gather : G x (R + 2) doubles over NVLink/NVSwitch (NCCL; gloo in the CPU tests)        (the collective)
    pivot      : every shard picks the same global winner from the G headers (total order on
                 (cost, id) or id), runs the ratio test on the winner's column against its RHS
                 replica, and applies the fused rank-1 update to its own columns             (device kernels)

Gathering all G candidate columns instead of "argmin exchange, then broadcast from the owner" keeps the root
of the transfer off the host: nothing in the loop depends on a device-computed value reaching the CPU, so the
whole budget of pivots is enqueued without a single host synchronisation.  The cost is G x 1 MiB per pivot at
R = 131072 -- microseconds on NVSwitch against >= 4 ms of HBM work per pivot.

The engine is pluggable so that the protocol itself is tested on CPU (gloo, world_size 2) with the oracle as
the per-shard engine; the product engine is CudaShardEngine (libb200lp).
"""
from __future__ import annotations

import numpy as np

from . import native
from .batched import shard_range


class CudaShardEngine:
    """A column shard resident in HBM, driven through the C ABI."""

    def __init__(self, m: int, n_total: int, lab0: int, ncols: int, seed: int, device: int = 0):
        import torch
        self.torch = torch
        self.device = device
        self.m, self.n_total, self.lab0, self.ncols = m, n_total, lab0, ncols
        self.R, self.C = m + 1, ncols + 1
        self.ld = (self.C + 15) // 16 * 16
        self.T = torch.empty(self.R * self.ld, dtype=torch.float64, device=f"cuda:{device}")
        self.solver = native.Solver(device)
        self.solver.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.solver.attach(self.T.data_ptr(), m, 1, self.C, self.ld, n_total, n_total + m, keep=self.T)
        self.solver.generate(seed, n_total, lab0)
        self.cand = torch.zeros(self.R + 2, dtype=torch.float64, device=f"cuda:{device}")

    def new_buffer(self, world):
        return self.torch.zeros(world * (self.R + 2), dtype=self.torch.float64, device=f"cuda:{self.device}")

    def reset(self, max_pivots):
        self.solver.shard_reset(max_pivots)

    def candidate(self, opts):
        self.solver.shard_candidate(opts, self.m, self.cand.data_ptr())
        return self.cand

    def pivot(self, opts, gathered, world, rank):
        self.solver.shard_pivot(opts, gathered.data_ptr(), world, rank)

    def state(self):
        return self.solver.shard_state()

    def history(self, cap):
        return self.solver.read_history(cap)

    def solution(self):
        return self.solver.read_solution()

    def tableau(self):
        return self.solver.read_tableau()

    def labels(self):
        return self.solver.get_labels()


class ShardedTableau:
    """Driver of the per-pivot protocol above over torch.distributed (or a single process when world == 1)."""

    def __init__(self, engine, world: int = 1, rank: int = 0, group=None):
        self.engine, self.world, self.rank, self.group = engine, world, rank, group
        self.gathered = engine.new_buffer(world)

    @staticmethod
    def columns_of(n_total: int, world: int, rank: int):
        """Structural columns [lo, hi) owned by `rank`."""
        return shard_range(n_total, world, rank, 2)

    def _all_gather(self, cand):
        if self.world == 1:
            self.gathered.copy_(cand)
            return
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.gathered, cand, group=self.group)

    def run(self, opts, max_pivots: int, check_every: int = 0):
        """Enqueue pivots until optimal / unbounded / max_pivots.  Returns (status, n_pivots)."""
        eng = self.engine
        eng.reset(max_pivots)
        check_every = check_every or max(1, min(max_pivots, 64))
        done_total = 0
        while True:
            for _ in range(check_every):
                self._all_gather(eng.candidate(opts))
                eng.pivot(opts, self.gathered, self.world, self.rank)
            done_total += check_every
            done, status, n = eng.state()
            if done:
                return status, n
            if done_total >= max_pivots + check_every:  # the device sets LIMIT itself; this is a backstop
                return native.STATUS_LIMIT, n
