"""Expected Bland pivot sequences of BASELINE configs 4 and 5, computed by the CPU oracle, for bench.py's parity block.

    python tests/golden/make_pivot_history.py          # writes tests/golden/pivot_history_config{4,5}.npy

Config 5 (131072 x 131072, 137 GB) cannot be held by the oracle, and does not have to be: under Bland's rule the entering
variable is the LOWEST variable id with a negative reduced cost, and a column of the tableau evolves from nothing but
itself, the pivot column and its own entry in the pivot row.  A column slab of the first `ncols` structural variables
(plus the right-hand side) therefore reproduces the full tableau's decisions exactly for as long as every entering
variable has an id below `ncols`: the slab's candidate is then the global minimum id (untouched columns have ids >=
ncols, variables that left the basis have slack ids >= n).  The script checks that condition on every pivot, and checks
the slab method itself against the full-tableau oracle on config 4's first pivots.

Rows of the output (int32): pivot row, entering variable id, leaving variable id -- independent of how the tableau is
stored or sharded, so the same file serves 1, 2, 4 and 8 GPUs and both loops.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as O  # noqa: E402

SEED = 4


def slab_history(m, n_total, ncols, pivots, threads):
    t = O.OracleTableau.generate(SEED, m, n_total, 0, ncols)
    r = t.solve(O.make_opts(rule=O.RULE_BLAND, max_pivots=pivots, threads=threads), hist_cap=pivots)
    k = r["n_pivots"]
    h = np.stack([r["piv_row"][:k], r["enter_lab"][:k], r["leave_lab"][:k]], axis=1).astype(np.int32)
    bad = np.nonzero(h[:, 1] >= ncols)[0]
    valid = int(bad[0]) if len(bad) else k
    if r["status"] != O.LIMIT:  # the slab ran out of candidates: the full tableau would look at higher ids next
        valid = min(valid, k)
    return h[:valid], valid == pivots


def main():
    threads = O.max_threads()
    # the slab method against the full oracle (config 4, first 96 pivots)
    full = O.OracleTableau.generate(SEED, 16383, 16383)
    r = full.solve(O.make_opts(rule=O.RULE_BLAND, max_pivots=96, threads=threads), hist_cap=96)
    want = np.stack([r["piv_row"], r["enter_lab"], r["leave_lab"]], axis=1).astype(np.int32)
    got, ok = slab_history(16383, 16383, 512, 96, threads)
    assert ok and np.array_equal(got, want), "slab history differs from the full-tableau oracle"
    del full
    for name, m, n_total, ncols, pivots in (("config4", 16383, 16383, 4096, 12288), ("config5", 131071, 131071, 1024, 1024)):
        h, ok = slab_history(m, n_total, ncols, pivots, threads)
        print(name, "pivots valid:", len(h), "complete:", ok, "max entering id:", int(h[:, 1].max()))
        np.save(os.path.join(HERE, f"pivot_history_{name}.npy"), h)


if __name__ == "__main__":
    main()
