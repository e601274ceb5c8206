"""CPU engines for the multi-process (gloo) tests: the oracle plays the part of one GPU's kernels, so that the
host-side sharding protocol of simplex_solver_b200/sharded.py and batched.py is exercised without a GPU."""
import numpy as np
import torch

from oracle import oracle as O


class OracleShardEngine:
    """Same interface as simplex_solver_b200.sharded.CudaShardEngine, on CPU tensors."""

    def __init__(self, m, n_total, lab0, ncols, seed):
        self.t = O.OracleTableau.generate(seed, m, n_total, lab0, ncols)
        self.m, self.R = m, m + 1
        self.cand = torch.zeros(self.R + 2, dtype=torch.float64)
        self.done, self.status, self.n, self.max_pivots = False, 0, 0, 0
        self.local_s = -1
        self.hist = []

    def new_buffer(self, world):
        return torch.zeros(world * (self.R + 2), dtype=torch.float64)

    def reset(self, max_pivots):
        self.done, self.status, self.n, self.max_pivots = False, 0, 0, max_pivots
        self.hist = []

    # the look-ahead loop changes WHEN the tableau is touched, not the protocol: the CPU stand-in ignores it
    def lookahead_begin(self):
        pass

    def lookahead_flush(self):
        pass

    def candidate(self, opts, lookahead=False):
        c = self.cand.numpy()
        c[:] = 0.0
        c[1] = -1.0
        self.local_s = -1
        if self.done:
            return self.cand
        if self.n >= self.max_pivots:
            self.done, self.status = True, O.LIMIT
            return self.cand
        s = self.t.price(rule=opts.rule, eps=opts.eps_cost)
        if s >= 0:
            self.local_s = s
            c[0] = self.t.T[self.m, s]
            c[1] = float(self.t.collab[s])
            c[2:] = self.t.extract_col(s)
        return self.cand

    def pivot(self, opts, gathered, world, rank, lookahead=False):
        if self.done:
            return
        g = gathered.numpy().reshape(world, self.R + 2)
        win = -1
        for k in range(world):
            if g[k, 1] < 0:
                continue
            if win < 0:
                win = k
            elif opts.rule == O.RULE_BLAND:
                if g[k, 1] < g[win, 1]:
                    win = k
            elif g[k, 0] < g[win, 0] or (g[k, 0] == g[win, 0] and g[k, 1] < g[win, 1]):
                win = k
        if win < 0:
            self.done, self.status = True, O.OPT
            return
        col = g[win, 2:].copy()
        r = self.t.ratio(col, opts.eps_pivot)
        if r < 0:
            self.done, self.status = True, O.UNBOUNDED
            return
        enter = int(g[win, 1])
        self.hist.append((r, enter, int(self.t.rowlab[r])))
        self.t.pivot_col(r, col, self.local_s if win == rank else -1, enter)
        self.n += 1

    def state(self):
        return self.done, self.status, self.n


def oracle_batched_engine(A, b, c, ops, rule, want_x, log_cap):
    return O.solve_batched(A, b, c, ops, O.make_opts(rule=rule), log_cap=log_cap, threads=2)
