"""The N>1 paths on CPU: world_size-2 gloo process groups drive the host-side sharding protocol
(simplex_solver_b200/sharded.py, batched.py) with the oracle standing in for each GPU's kernels.
The column-sharded run must reproduce the single-tableau pivot sequence and solution bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from simplex_solver_b200 import native, workloads as W
from simplex_solver_b200.batched import solve_batched_distributed
from simplex_solver_b200.sharded import ShardedTableau
from tests.oracle_engines import OracleShardEngine, oracle_batched_engine

M, N_TOTAL, SEED, BUDGET = 48, 80, 4, 60


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _sharded_worker(rank, world, port, rule, out_dir):
    _init(rank, world, port)
    lo, hi = ShardedTableau.columns_of(N_TOTAL, world, rank)
    eng = OracleShardEngine(M, N_TOTAL, lo, hi - lo, SEED)
    drv = ShardedTableau(eng, world, rank)
    opts = native.Opts(rule, 0, BUDGET, 1e-9, 1e-9, 1e-7, 0, 1)
    status, n = drv.run(opts, BUDGET, check_every=7)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), status=status, n=n, hist=np.array(eng.hist),
             T=eng.t.T.copy(), rowlab=eng.t.rowlab.copy(), collab=eng.t.collab.copy(), lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("rule", [O.RULE_BLAND, O.RULE_DANTZIG])
def test_column_sharded_matches_single_tableau(tmp_path, rule):
    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), rule, str(tmp_path)), nprocs=world, join=True)
    one = O.OracleTableau.generate(SEED, M, N_TOTAL)
    ref = one.solve(O.make_opts(rule=rule, max_pivots=BUDGET), hist_cap=BUDGET)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for p in parts:
        assert int(p["status"]) == ref["status"] and int(p["n"]) == ref["n_pivots"]
        # same leaving rows, entering and leaving variables on every shard
        np.testing.assert_array_equal(p["hist"][:, 0], ref["piv_row"])
        np.testing.assert_array_equal(p["hist"][:, 1], ref["enter_lab"])
        np.testing.assert_array_equal(p["hist"][:, 2], ref["leave_lab"])
        np.testing.assert_array_equal(p["rowlab"], one.rowlab)
    # every column of the single tableau lives, bit for bit, in the shard that owns its variable
    where = {}
    for r, p in enumerate(parts):
        for j, lab in enumerate(p["collab"][:-1]):
            where[int(lab)] = (r, j)
    assert len(where) == N_TOTAL
    for j, lab in enumerate(one.collab[:-1]):
        r, jj = where[int(lab)]
        np.testing.assert_array_equal(parts[r]["T"][:, jj], one.T[:, j])
    for p in parts:
        np.testing.assert_array_equal(p["T"][:, -1], one.T[:, -1])     # replicated right-hand side


def _batched_worker(rank, world, port, out_dir):
    _init(rank, world, port)
    total = 3000
    res = solve_batched_distributed(lambda lo, hi: W.batched_small_lps(lo, hi - lo), total, 30, align=W.BATCH_BLOCK,
                                    engine=oracle_batched_engine, want_x=True)
    np.savez(os.path.join(out_dir, f"b{rank}.npz"), **{k: v for k, v in res.items() if v is not None})
    dist.destroy_process_group()


def test_batched_blocks_are_gathered_in_order(tmp_path):
    world = 2
    mp.spawn(_batched_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    A, b, c, ops = W.batched_small_lps(0, 3000)
    ref = O.solve_batched(A, b, c, ops, threads=4)
    for r in range(world):
        got = np.load(tmp_path / f"b{r}.npz")
        np.testing.assert_array_equal(got["status"], ref["status"])
        np.testing.assert_array_equal(got["n_pivots"], ref["n_pivots"])
        ok = ref["status"] == 0
        np.testing.assert_array_equal(got["fun"][ok], ref["fun"][ok])
        np.testing.assert_array_equal(got["x"][ok], ref["x"][ok])


def test_single_process_driver_equals_oracle():
    eng = OracleShardEngine(M, N_TOTAL, 0, N_TOTAL, SEED)
    drv = ShardedTableau(eng, 1, 0)
    opts = native.Opts(O.RULE_DANTZIG, 0, 1 << 40, 1e-9, 1e-9, 1e-7, 0, 1)
    status, n = drv.run(opts, 1 << 30, check_every=16)
    one = O.OracleTableau.generate(SEED, M, N_TOTAL)
    ref = one.solve(O.make_opts(rule=O.RULE_DANTZIG))
    assert status == ref["status"] == 0 and n == ref["n_pivots"]
    np.testing.assert_array_equal(eng.t.T, one.T)


def _sharded_limit_worker(rank, world, port, out_dir):
    _init(rank, world, port)
    lo, hi = ShardedTableau.columns_of(N_TOTAL, world, rank)
    eng = OracleShardEngine(M, N_TOTAL, lo, hi - lo, SEED)
    drv = ShardedTableau(eng, world, rank)
    opts = native.make_opts(rule=O.RULE_BLAND, max_pivots=BUDGET, time_limit=1e-7 if rank == 0 else 3600.0)
    status, n = drv.run(opts, BUDGET, check_every=5)
    np.savez(os.path.join(out_dir, f"l{rank}.npz"), status=status, n=n, hist=np.array(eng.hist))
    dist.destroy_process_group()


def test_sharded_driver_time_limit_stops_all_ranks_at_the_same_chunk(tmp_path):
    """solver_controller.py:76 bounds a solve in wall-clock time and :404 maps the limit to "Error".  In the sharded driver
    the bound is checked between chunks and reduced over the ranks: here only rank 0's clock has expired, and both ranks
    stop after the same first chunk with status LIMIT and the oracle's first pivots."""
    world = 2
    mp.spawn(_sharded_limit_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    one = O.OracleTableau.generate(SEED, M, N_TOTAL)
    ref = one.solve(O.make_opts(rule=O.RULE_BLAND, max_pivots=5), hist_cap=5)
    for r in range(world):
        p = np.load(tmp_path / f"l{r}.npz")
        assert int(p["status"]) == native.STATUS_LIMIT and int(p["n"]) == 5
        np.testing.assert_array_equal(p["hist"][:, 0], ref["piv_row"])
