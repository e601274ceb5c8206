"""Memory-safety checks in place of compute-sanitizer (refused on this GPU pool): every caller-owned buffer the kernels
write through raw pointers -- tableau, candidate / gathered buffers, peer exchange regions, snapshot buffer, batched
inputs and outputs -- lives inside a larger tensor whose surroundings hold a NaN bit pattern, and every LIBRARY-owned
buffer lies between guard bands (B200LP_GUARD=1, set in conftest.py for the whole GPU session; b200lp_check_guards).
All loop modes run on ragged shapes (odd column counts, rows that are no multiple of any tile); afterwards the bands
must be untouched.  Results are still compared with the oracle, so a kernel cannot pass by not writing at all."""
import numpy as np
import pytest

from simplex_solver_b200 import native, workloads as W
from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau
from tests.helpers import assert_bit_equal

pytestmark = pytest.mark.gpu
BAND = 4096  # doubles on either side
PATTERN = np.array([0x7FF8DEADBEEF0001], dtype=np.uint64).view(np.float64)[0]


def _banded(n_doubles, torch, dtype=None):
    """A tensor of n_doubles (or elements of dtype) inside a patterned allocation; returns (view, whole, check)."""
    dtype = dtype or torch.float64
    whole = torch.empty(n_doubles + 2 * BAND, dtype=dtype, device="cuda:0")
    if dtype == torch.float64:
        whole.view(torch.int64).fill_(int(np.array([PATTERN]).view(np.int64)[0]))
    else:
        whole.fill_(0x5A)
    ref = whole.clone()

    def check(what):
        torch.cuda.synchronize()
        lo = torch.equal(whole[:BAND].view(torch.uint8), ref[:BAND].view(torch.uint8))
        hi = torch.equal(whole[BAND + n_doubles:].view(torch.uint8), ref[BAND + n_doubles:].view(torch.uint8))
        assert lo and hi, f"{what}: a kernel wrote outside the buffer ({'below' if not lo else 'above'})"
    return whole[BAND:BAND + n_doubles], whole, check


LOOPS = [("launches", dict(loop_mode=native.LOOP_LAUNCHES, update_variant=native.UPDATE_LDG)),
         ("graph-ldg", dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_LDG)),
         ("graph-tma", dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_TMA)),
         ("auto", dict(loop_mode=native.LOOP_AUTO)),
         ("lookahead-5", dict(loop_mode=native.LOOP_BLOCKED, check_every=5)),
         ("lookahead-32", dict(loop_mode=native.LOOP_BLOCKED, check_every=32))]


@pytest.mark.parametrize("shape", [(37, 52), (130, 257), (511, 766), (200, 1301), (1030, 2055)])
def test_tableau_and_workspace_bands_every_loop_mode(oracle, shape):
    """ld == C where C is even, ld = C + 1 where it is odd: no slack between the rows, so a store one column past a row
    lands in the next row's data (caught by the oracle comparison) and one past the last row lands in the band."""
    import torch
    m, n = shape
    C = n + 1
    ld = C + (C & 1)
    budget = 70
    ot = oracle.OracleTableau.generate(11, m, n)
    ref = ot.solve(oracle.make_opts(rule=1, max_pivots=budget), hist_cap=budget)
    for name, kw in LOOPS:
        T, whole, check = _banded((m + 1) * ld, torch)
        s = native.Solver(0)
        s.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=whole)
        s.generate(11, n, 0)
        got = s.run(native.make_opts(rule=native.RULE_BLAND, max_pivots=budget, **kw), hist_cap=budget)
        assert got["n_pivots"] == ref["n_pivots"] and np.array_equal(got["piv_row"], ref["piv_row"]), name
        assert_bit_equal(s.read_tableau(), ot.T, f"{name} {shape}")
        # the single-phase entry points write through the same pointers
        col = s.select_entering(rule=native.RULE_DANTZIG)
        if col >= 0:
            row = s.ratio_test(col)
            if row >= 0:
                s.pivot(row, col, native.UPDATE_TMA)
                s.pivot(row, col, native.UPDATE_LDG)
        check(f"{name} {shape}")
        assert s.check_guards() == 0, f"{name} {shape}: a store left a workspace buffer"
        s.close()


def test_solve_dense_snapshots_and_batched_bands(oracle):
    import ctypes as C
    import torch
    # solve_dense on two-phase LPs (library-owned tableau between guard bands), snapshots in a banded caller buffer
    s = native.Solver(0)
    for k in range(0, 60, 3):
        A, b, c, ops = W.fuzz_lp(k)
        s.build_dense(A, b, c, ops)
        m, n_obj, Cc, _ = s.dims()
        cap = 24
        snaps, whole, check = _banded(cap * (m + n_obj) * Cc, torch)
        s.set_snapshots(snaps.data_ptr(), cap, keep=whole)
        got = s.solve(native.make_opts(), hist_cap=64)
        s.set_snapshots(None, 0)
        ref = oracle.solve_lp(A, b, c, ops, hist_cap=64)
        assert got["status"] == ref["status"] and np.array_equal(got["piv_row"], ref["piv_row"])
        check(f"snapshots fuzz {k}")
        assert s.check_guards() == 0, f"fuzz {k}"
    # batched kernel with device pointers: all four inputs and four outputs banded; ragged batch and shape
    for (B, m, n) in ((777, 20, 30), (1000, 7, 5), (33, 1, 1)):
        A, b, c, ops = W.batched_small_lps(0, B, m, n)
        ins, outs, checks = [], [], []
        for a, dt in ((A, torch.float64), (b, torch.float64), (c, torch.float64), (ops, torch.int8)):
            v, whole, chk = _banded(a.size, torch, dt)
            v.copy_(torch.from_numpy(a.reshape(-1)))
            ins.append((v, whole))
            checks.append(chk)
        for size, dt in ((B, torch.int32), (B, torch.float64), (B * n, torch.float64), (B, torch.int32)):
            v, whole, chk = _banded(size, torch, dt)
            outs.append((v, whole))
            checks.append(chk)
        s.solve_batched_device(B, m, n, ins[0][0].data_ptr(), ins[1][0].data_ptr(), ins[2][0].data_ptr(),
                               ins[3][0].data_ptr(), outs[0][0].data_ptr(), outs[1][0].data_ptr(), outs[2][0].data_ptr(),
                               outs[3][0].data_ptr(), native.make_opts())
        ref = oracle.solve_batched(A, b, c, ops, threads=4)
        assert np.array_equal(outs[0][0].cpu().numpy(), ref["status"])
        assert np.array_equal(outs[3][0].cpu().numpy(), ref["n_pivots"])
        for chk in checks:
            chk(f"batched {B}x{m}x{n}")
        # and the host-pointer path (library-owned staging between guard bands)
        got = s.solve_batched(A, b, c, ops, log_cap=8)
        assert np.array_equal(got["status"], ref["status"])
        assert s.check_guards() == 0, f"batched {B}x{m}x{n}"
    s.close()


@pytest.mark.parametrize("n_total", [157, 160])
def test_shard_exchange_regions_and_candidates_bands(oracle, n_total):
    """Two emulated shards on one GPU: tableaux, candidate / gathered buffers and the peer exchange regions are banded;
    rank-1 and look-ahead loops, all-gather and peer-memory exchange."""
    import torch
    m, seed, budget, world = 95, 4, 60, 2
    one = oracle.OracleTableau.generate(seed, m, n_total)
    for rule in (native.RULE_BLAND, native.RULE_DANTZIG):
        ref = one_ref = None
        t = oracle.OracleTableau.generate(seed, m, n_total)
        ref = t.solve(oracle.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
        opts = native.make_opts(rule=rule, max_pivots=budget)
        for lookahead in (0, 7):
            for p2p in (False, True):
                engs, checks = [], []
                for r in range(world):
                    lo, hi = ShardedTableau.columns_of(n_total, world, r)
                    e = CudaShardEngine.__new__(CudaShardEngine)
                    e.torch, e.device = torch, 0
                    e.m, e.n_total, e.lab0, e.ncols, e.seed = m, n_total, lo, hi - lo, seed
                    e.R, e.C = m + 1, hi - lo + 1
                    e.ld = e.C + (e.C & 1)
                    e.T, e._whole_T, chk = _banded(e.R * e.ld, torch)
                    checks.append(chk)
                    e.solver = native.Solver(0)
                    e.solver.set_stream(torch.cuda.current_stream(0).cuda_stream)
                    e.solver.attach(e.T.data_ptr(), m, 1, e.C, e.ld, n_total, n_total + m, keep=e._whole_T)
                    e.solver.generate(seed, n_total, lo)
                    e.cand, e._whole_c, chk = _banded(e.R + 2, torch)
                    e.cand.zero_()
                    checks.append(chk)
                    engs.append(e)
                gathered, _wg, chk = _banded(world * (m + 3), torch)
                gathered.zero_()
                checks.append(chk)
                if p2p:
                    nreg = native.Solver.p2p_bytes(m + 1, world) // 8
                    regions = []
                    for r in range(world):
                        reg, whole, chk = _banded(nreg, torch)
                        reg.zero_()
                        regions.append((reg, whole))
                        checks.append(chk)
                    for r, e in enumerate(engs):
                        e.enable_p2p(world, r, bases=[t_[0].data_ptr() for t_ in regions], region=regions[r][0])
                for e in engs:
                    e.reset(budget)
                    if lookahead:
                        e.lookahead_begin()
                stride = m + 3
                for it in range(budget + 2):
                    if p2p:  # ONE launch, one cluster per shard: the shards' waits for each other resolve inside the kernel
                        native.Solver.shard_fused_multi([e.solver for e in engs], opts, m, lookahead > 0)
                    else:
                        for r, e in enumerate(engs):
                            gathered[r * stride:(r + 1) * stride].copy_(e.candidate(opts, lookahead > 0))
                        torch.cuda.synchronize()
                        for r, e in enumerate(engs):
                            e.pivot(opts, gathered, world, r, lookahead > 0)
                    for e in engs:
                        if lookahead and (it + 1) % lookahead == 0:
                            e.lookahead_flush()
                    torch.cuda.synchronize()
                if lookahead:
                    for e in engs:
                        e.lookahead_flush()
                what = f"n_total {n_total} rule {rule} lookahead {lookahead} p2p {p2p}"
                for r, e in enumerate(engs):
                    done, status, n = e.state()
                    assert done and status == ref["status"] and n == ref["n_pivots"], what
                    np.testing.assert_array_equal(e.history(budget)["piv_row"], ref["piv_row"])
                    T = e.tableau()
                    _, cl = e.labels()
                    pos = {int(lab): j for j, lab in enumerate(t.collab[:-1])}
                    for j, lab in enumerate(cl[:-1]):
                        assert_bit_equal(T[:, j], t.T[:, pos[int(lab)]], f"{what}: shard {r} column of variable {lab}")
                    assert e.solver.check_guards() == 0, what
                for chk in checks:
                    chk(what)
                for e in engs:
                    e.solver.close()
