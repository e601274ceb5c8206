"""Batches of independent small LPs (BASELINE config 3): one warp per LP on each GPU, contiguous blocks of the
batch per rank, no collective on the data path -- one gather of (status, z, n_pivots[, x]) at the end.

The reference solves one problem per SolverController.run() (ui_controller.py:194-195); this is the same
solve for B problems of one shape.
"""
from __future__ import annotations

import numpy as np

from . import native


def shard_range(total: int, world: int, rank: int, align: int = 1):
    """Contiguous block [lo, hi) of `total` items owned by `rank`; boundaries are multiples of `align`."""
    units = (total + align - 1) // align
    lo = (units * rank // world) * align
    hi = (units * (rank + 1) // world) * align
    return min(lo, total), min(hi, total)


def solve_batched(A, b, c, ops, rule=native.RULE_DANTZIG, device: int = 0, want_x: bool = True, log_cap: int = 0,
                  engine=None):
    """Host arrays A[B,m,n], b[B,m], c[B,n] (minimisation costs), ops[B,m] -> dict(status, fun, x, n_pivots)."""
    if engine is not None:
        return engine(A, b, c, ops, rule, want_x, log_cap)
    return native.thread_solver(device).solve_batched(A, b, c, ops, native.make_opts(rule=rule), want_x=want_x,
                                                      log_cap=log_cap)


def solve_batched_distributed(make_local, total: int, n: int, rule=native.RULE_DANTZIG, device: int = 0,
                              align: int = 1, engine=None, want_x: bool = False):
    """Every rank solves its block and all ranks receive the full result.

    make_local(lo, hi) -> (A, b, c, ops) for the LPs [lo, hi).  Uses the default torch.distributed group
    (NCCL on GPUs, gloo in the CPU tests); results are gathered with one all_gather of padded blocks.
    """
    import torch
    import torch.distributed as dist

    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    lo, hi = shard_range(total, world, rank, align)
    A, b, c, ops = make_local(lo, hi)
    local = solve_batched(A, b, c, ops, rule=rule, device=device, want_x=want_x, engine=engine)
    if world == 1:
        return local
    width = 3 + (n if want_x else 0)
    per = max(shard_range(total, world, r, align)[1] - shard_range(total, world, r, align)[0] for r in range(world))
    block = np.zeros((per, width))
    k = hi - lo
    block[:k, 0] = local["status"]
    block[:k, 1] = local["fun"]
    block[:k, 2] = local["n_pivots"]
    if want_x:
        block[:k, 3:] = local["x"]
    on_gpu = dist.get_backend() == "nccl"
    t = torch.from_numpy(block)
    if on_gpu:
        t = t.to(f"cuda:{device}")
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), t.reshape(-1).contiguous())
    out = out.cpu().numpy()
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(total, world, r, align)
        parts.append(out[r, : rhi - rlo])
    full = np.concatenate(parts, axis=0)
    res = {"status": full[:, 0].astype(np.int32), "fun": full[:, 1].copy(), "n_pivots": full[:, 2].astype(np.int32),
           "x": full[:, 3:].copy() if want_x else None}
    return res
