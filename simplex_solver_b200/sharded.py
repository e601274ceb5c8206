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


def row_stride(C: int) -> int:
    """Row stride (doubles) for a tableau of C stored columns that the library's kernels stream at full speed.  Measured on
    B200 with the pivot-update kernels on 131072-row shards: a stride that is a multiple of 16 doubles but not of 256
    (65552, 32784, 16400) costs 7-8 % of the HBM bandwidth, with both the LDG and the TMA kernel; multiples of 256 doubles
    (2 KB) run at the speed of the power-of-two strides, ragged last strip or not (scripts/probe_shard_shapes.py)."""
    return (C + 255) // 256 * 256 if C >= 4096 else (C + 15) // 16 * 16


class CudaShardEngine:
    """A column shard resident in HBM, driven through the C ABI."""

    def __init__(self, m: int, n_total: int, lab0: int, ncols: int, seed: int, device: int = 0):
        import torch
        self.torch = torch
        self.device = device
        self.m, self.n_total, self.lab0, self.ncols, self.seed = m, n_total, lab0, ncols, seed
        self.R, self.C = m + 1, ncols + 1
        self.ld = row_stride(self.C)
        self.T = torch.empty(self.R * self.ld, dtype=torch.float64, device=f"cuda:{device}")
        self.solver = native.Solver(device)
        self.solver.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.solver.attach(self.T.data_ptr(), m, 1, self.C, self.ld, n_total, n_total + m, keep=self.T)
        self.solver.generate(seed, n_total, lab0)
        self.cand = torch.zeros(self.R + 2, dtype=torch.float64, device=f"cuda:{device}")

    def regenerate(self):
        """Refill the shard with the synthetic LP it was created with (device-side generator)."""
        self.solver.generate(self.seed, self.n_total, self.lab0)

    def initial_labels(self):
        """Row / column labels of the freshly generated shard (host arrays): what travels with a host copy of it."""
        rl = np.concatenate([self.n_total + np.arange(self.m, dtype=np.int32), np.array([-1], dtype=np.int32)])
        cl = np.concatenate([self.lab0 + np.arange(self.ncols, dtype=np.int32), np.array([-1], dtype=np.int32)])
        return rl.astype(np.int32), cl.astype(np.int32)

    def new_buffer(self, world):
        return self.torch.zeros(world * (self.R + 2), dtype=self.torch.float64, device=f"cuda:{self.device}")

    def reset(self, max_pivots):
        self.solver.shard_reset(max_pivots)

    def capture_chunk(self, body):
        """Capture `body` (library kernels + torch.distributed collectives) into a CUDA graph; None if capture fails."""
        torch = self.torch
        outer = torch.cuda.current_stream(self.device).cuda_stream
        g = torch.cuda.CUDAGraph()
        try:
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g):
                self.solver.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
                body()
        except Exception:
            g = None
        finally:
            self.solver.set_stream(outer)
        return g

    def candidate(self, opts, lookahead: bool = False):
        if lookahead:
            self.solver.shard_blk_candidate(opts, self.m, self.cand.data_ptr())
        else:
            self.solver.shard_candidate(opts, self.m, self.cand.data_ptr())
        return self.cand

    def pivot(self, opts, gathered, world, rank, lookahead: bool = False):
        if lookahead:
            self.solver.shard_blk_pivot(opts, gathered.data_ptr(), world, rank)
        else:
            self.solver.shard_pivot(opts, gathered.data_ptr(), world, rank)

    # peer-memory exchange (kernels_shard.cuh: k_shard_pick): candidates are stored straight into every peer's region
    # over NVLink by the kernel that prices, decides and runs the ratio test; no collective call inside the loop
    def enable_p2p(self, world: int, rank: int, group=None, bases=None, region=None):
        """Allocate this shard's exchange region in torch symmetric memory, rendezvous, connect.  `bases`/`region` let a
        single-process test wire two engines by hand."""
        torch = self.torch
        n = native.Solver.p2p_bytes(self.R, world) // 8
        if bases is None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            region = symm.empty(n, dtype=torch.float64, device=f"cuda:{self.device}")
            hdl = symm.rendezvous(region, group if group is not None else dist.group.WORLD)
            bases = [int(p) for p in hdl.buffer_ptrs]
            self._symm = hdl
        self.region = region
        self.solver.p2p_connect(bases, world, rank)
        self.p2p = True
        if hasattr(self, "_symm"):  # every rank has zeroed its flags before anyone pushes
            import torch.distributed as dist
            dist.barrier(group=group)

    def fused(self, opts, lookahead: bool = False):
        """One pivot: the fused kernel prices, exchanges candidates over peer memory, decides and runs the ratio test."""
        self.solver.shard_fused(opts, self.m, lookahead)

    # look-ahead loop (kernels_blocked.cuh): pivots are decided from O(R + C) state and applied K at a time
    def lookahead_begin(self):
        self.solver.shard_blk_begin(self.m)

    def lookahead_flush(self):
        self.solver.shard_blk_flush(self.m)

    def state(self):
        return self.solver.shard_state()

    def epoch(self):
        """Key of captured chunks: moves when the library reallocates a buffer a captured launch has baked in."""
        return self.solver.binding_epoch()

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
        self._graphs = {}  # captured chunks, keyed by the options that are baked into the kernels' arguments

    @staticmethod
    def columns_of(n_total: int, world: int, rank: int):
        """Structural columns [lo, hi) owned by `rank`."""
        return shard_range(n_total, world, rank, 2)

    @staticmethod
    def columns_of_tableau(cols_total: int, world: int, rank: int):
        """Structural columns [lo, hi) of `rank` for ONE tableau of `cols_total` stored columns (RHS included), i.e.
        n_total = cols_total - 1 structural variables whatever the number of GPUs -- the same LP, hence the same pivot
        sequence, at every world size.  Every shard but the last stores exactly cols_total / world columns (its
        structural slice + its RHS replica: row strides stay multiples of 16 doubles for the usual sizes); the last
        one takes the remainder, world - 1 columns more."""
        per = cols_total // world - 1
        lo = rank * per
        hi = (rank + 1) * per if rank < world - 1 else cols_total - 1
        return lo, hi

    def _all_gather(self, cand):
        if self.world == 1:
            self.gathered.copy_(cand)
            return
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.gathered, cand, group=self.group)

    def _expired(self, deadline) -> bool:
        """Has the wall-clock bound passed on ANY rank?  (max-reduction: all ranks must stop at the same chunk)"""
        import time
        late = time.monotonic() >= deadline
        if self.world == 1:
            return late
        import torch
        import torch.distributed as dist
        dev = self.gathered.device if self.gathered is not None else "cpu"
        flag = torch.tensor([1 if late else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        return bool(int(flag.item()))

    def _chunk(self, opts, n, lookahead=0):
        eng = self.engine
        p2p = getattr(eng, "p2p", False)
        for i in range(n):
            if p2p:  # candidates stored into every peer's region over NVLink inside the pick kernel: no collective call
                eng.fused(opts, lookahead > 0)
            else:
                self._all_gather(eng.candidate(opts, lookahead > 0) if lookahead else eng.candidate(opts))
                if lookahead:
                    eng.pivot(opts, self.gathered, self.world, self.rank, True)
                else:
                    eng.pivot(opts, self.gathered, self.world, self.rank)
            if lookahead and ((i + 1) % lookahead == 0 or i + 1 == n):
                eng.lookahead_flush()

    def run(self, opts, max_pivots: int, check_every: int = 0, use_graph: bool = True, lookahead: int = 0):
        """Enqueue pivots until optimal / unbounded / max_pivots / opts.time_limit_s.  Returns (status, n_pivots).

        On GPUs the per-pivot sequence (candidate kernels -> NCCL all-gather -> winner / ratio / update kernels) of a
        whole chunk is captured once into a CUDA graph and replayed, so the host issues one launch per `check_every`
        pivots instead of ~8 calls per pivot; the first chunk runs eagerly (it also warms NCCL up for capture).

        lookahead = K > 0 selects the look-ahead loop: the exchange per pivot is the same, but the tableau is only
        touched once per K pivots (one flush); pivots and tableau stay bit-identical.
        """
        import time
        eng = self.engine
        limit = float(getattr(opts, "time_limit_s", 0.0) or 0.0)
        deadline = time.monotonic() + limit if limit > 0.0 else None
        eng.reset(max_pivots)
        if getattr(eng, "p2p", False) and self.world > 1:
            # the pick kernel gives a missing peer ~4 s before it gives up: start the ranks together
            import torch.distributed as dist
            dist.barrier(group=self.group)
        lookahead = int(max(0, min(lookahead, 32)))
        if lookahead:
            eng.lookahead_begin()
        check_every = check_every or max(1, min(max_pivots, 64))
        if lookahead:
            check_every = max(lookahead, check_every // lookahead * lookahead)  # whole blocks per chunk
        # a captured chunk bakes device pointers and capacities (pivot history arrays sized by max_pivots, look-ahead
        # buffers, ...): it is valid for the binding epoch it was captured in, which is read AFTER reset / begin have
        # made their allocations.  A moved epoch drops every cached chunk (their memory may have been freed).
        epoch = eng.epoch() if hasattr(eng, "epoch") else 0
        if epoch != getattr(self, "_epoch", epoch):
            self._graphs.clear()
        self._epoch = epoch
        key = (opts.rule, opts.update_variant, opts.eps_cost, opts.eps_pivot, check_every, lookahead,
               bool(getattr(eng, "p2p", False)))
        graph = self._graphs.get(key) if use_graph else None
        done_total = 0
        while True:
            if graph is not None:
                graph.replay()
            else:
                self._chunk(opts, check_every, lookahead)
                if use_graph and hasattr(eng, "capture_chunk") and key not in self._graphs:
                    if hasattr(eng, "epoch") and eng.epoch() != epoch:  # the eager chunk allocated: capture next time
                        self._epoch = eng.epoch()
                        self._graphs.clear()
                        epoch = self._epoch
                    self._graphs[key] = eng.capture_chunk(lambda: self._chunk(opts, check_every, lookahead))
                    graph = self._graphs[key]
            done_total += check_every
            done, status, n = eng.state()
            if done:
                return status, n
            if done_total >= max_pivots + check_every:  # the device sets LIMIT itself; this is a backstop
                return native.STATUS_LIMIT, n
            if deadline is not None and self._expired(deadline):
                # the wall-clock bound of the reference's solve (solver_controller.py:76), checked between chunks: every
                # rank leaves at the same chunk (the flag is reduced over the ranks), with a consistent shard
                return native.STATUS_LIMIT, n
