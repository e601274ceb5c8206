"""The four `simple_simplex` names the reference imports (solver_controller.py:22-27) and calls
(:297-318), re-implemented over the GPU pivot loop:

    tableau = create_tableau(number_of_variables=n, number_of_constraints=m)       # :297
    add_constraint(tableau, "a1,...,an,L|G|E,rhs")                                 # :302-309
    add_objective(tableau, "c1,...,cn,1|0")                                        # :311-316  (1 = maximise)
    result = optimize_json_format(tableau, maximize=bool)                          # :318

`result["pivotSteps"]` is a list of {"step", "pivotRowIndex", "pivotColIndex", "tableau"} -- the only keys the
reference reads (solver_controller.py:332-336): step 0 is the initial tableau with both indices None; step k
(k >= 1) is the tableau AFTER the k-th pivot together with the 0-based (row, column) of that pivot in the
tableau as displayed.

PARITY UNPINNED: simple-simplex==0.0.3 is not under /root/reference and no reference test asserts on a single
tableau cell or pivot index, so the layout below is this package's own definition:
rows F0..F(m-1) = constraints (rows with a negative right-hand side negated), then the objective row, then --
only when the problem needs artificials -- the phase-1 row; columns = structural x1..xn, one slack (<=) or
surplus (>=) column per inequality row, one artificial column per >= / = row, then the right-hand side.  The
objective row holds reduced costs of the minimisation form (for a maximise problem: -c, the textbook max
tableau) and its last cell the running value of the user's objective (negated for a minimise problem).

The GPU works on the condensed (Tucker) tableau: only non-basic columns are stored; the displayed full tableau
is rebuilt on the host by scattering unit vectors into the basic columns -- a pure permutation, no arithmetic.
"""
from __future__ import annotations

import numpy as np

from . import native

_OPS = {"L": native.OP_LE, "G": native.OP_GE, "E": native.OP_EQ}

MAX_RECORDED_STEPS = 4096
MAX_SNAPSHOT_BYTES = 256 << 20


def create_tableau(number_of_variables: int, number_of_constraints: int):
    if number_of_variables < 1:
        raise ValueError("number_of_variables must be >= 1")
    if number_of_constraints < 0:
        raise ValueError("number_of_constraints must be >= 0")
    return {"n": int(number_of_variables), "m": int(number_of_constraints), "rows": [], "objective": None,
            "maximize_flag": None}


def add_constraint(tableau, constraint: str):
    fields = [f.strip() for f in constraint.split(",")]
    n = tableau["n"]
    if len(fields) != n + 2:
        raise ValueError(f"constraint needs {n} coefficients, an operator and a right-hand side: {constraint!r}")
    op = fields[n].upper()
    if op not in _OPS:
        raise ValueError(f"operator must be L, G or E, got {fields[n]!r}")
    if len(tableau["rows"]) >= tableau["m"]:
        raise ValueError("more constraints than declared in create_tableau")
    tableau["rows"].append(([float(v) for v in fields[:n]], _OPS[op], float(fields[n + 1])))


def add_objective(tableau, objective: str):
    fields = [f.strip() for f in objective.split(",")]
    n = tableau["n"]
    if len(fields) != n + 1:
        raise ValueError(f"objective needs {n} coefficients and the 1|0 maximise flag: {objective!r}")
    tableau["objective"] = [float(v) for v in fields[:n]]
    tableau["maximize_flag"] = float(fields[n]) != 0.0


def _column_names(n, m, var_ids):
    names = []
    for v in var_ids:
        if v < n:
            names.append(f"x{v + 1}")
        elif v < n + m:
            names.append(f"s{v - n + 1}")
        else:
            names.append(f"a{v - n - m + 1}")
    return names + ["RHS"]


def expand_full(Tc, rowlab, collab, var_ids, m):
    """Condensed tableau + labels -> displayed full tableau (R x (len(var_ids)+1))."""
    R, C = Tc.shape
    where = {v: k for k, v in enumerate(var_ids)}
    full = np.zeros((R, len(var_ids) + 1))
    for j in range(C - 1):
        full[:, where[int(collab[j])]] = Tc[:, j]
    for i in range(m):
        lab = int(rowlab[i])
        if lab < 0:
            lab = -1 - lab  # artificial left basic in a redundant row
        full[i, where[lab]] = 1.0
    full[:, -1] = Tc[:, C - 1]
    return full


def optimize_json_format(tableau, maximize: bool | None = None, rule: str = "dantzig", device: int = 0):
    import torch

    n, rows = tableau["n"], tableau["rows"]
    if tableau["objective"] is None:
        raise ValueError("add_objective was not called")
    if maximize is None:
        maximize = bool(tableau["maximize_flag"])
    m = len(rows)
    A = np.array([r[0] for r in rows], dtype=np.float64).reshape(m, n)
    ops = np.array([r[1] for r in rows], dtype=np.int8)
    b = np.array([r[2] for r in rows], dtype=np.float64)
    c_user = np.array(tableau["objective"], dtype=np.float64)
    cmin = -c_user if maximize else c_user

    solver = native.thread_solver(device)
    solver.build_dense(A, b, cmin, ops)
    mm, n_obj, C, _ld = solver.dims()
    R = mm + n_obj
    rowlab, collab = solver.get_labels()
    T0 = solver.read_tableau()
    var_ids = sorted(set(int(v) for v in rowlab[:m]) | set(int(v) for v in collab[:C - 1]))

    cap = int(min(MAX_RECORDED_STEPS, max(16, MAX_SNAPSHOT_BYTES // max(8 * R * C, 1))))
    snaps = torch.empty(cap * R * C, dtype=torch.float64, device=f"cuda:{device}")
    solver.set_snapshots(snaps.data_ptr(), cap, keep=snaps)
    try:
        rule_id = native.RULE_BLAND if str(rule).lower() == "bland" else native.RULE_DANTZIG
        res = solver.solve(native.make_opts(rule=rule_id), hist_cap=cap)
    finally:
        solver.set_snapshots(None, 0)
    k = min(res["n_pivots"], cap)
    host = snaps[: k * R * C].cpu().numpy().reshape(k, R, C)

    steps = [{"step": 0, "pivotRowIndex": None, "pivotColIndex": None,
              "tableau": expand_full(T0, rowlab, collab, var_ids, m).tolist(),
              "basis": [int(v) for v in rowlab[:m]]}]
    where = {v: idx for idx, v in enumerate(var_ids)}
    rl, cl = rowlab.copy(), collab.copy()
    for it in range(k):
        r, s = int(res["piv_row"][it]), int(res["piv_col"][it])
        enter = int(res["enter_lab"][it])
        rl[r], cl[s] = cl[s], rl[r]
        steps.append({"step": it + 1, "pivotRowIndex": r, "pivotColIndex": where[enter],
                      "tableau": expand_full(host[it], rl, cl, var_ids, m).tolist(),
                      "basis": [int(v) for v in rl[:m]]})
    if k == res["n_pivots"]:
        # the labels replayed on the host from the pivot history must be the device's own (a row flagged redundant
        # by the drive-out holds -1 - id there): a mismatch means the displayed basic columns would be wrong
        rl_dev, cl_dev = solver.get_labels()
        flagged = rl_dev[:m] < 0
        if not (np.array_equal(np.where(flagged, -1 - rl_dev[:m], rl_dev[:m]), rl[:m])
                and np.array_equal(cl_dev[:C - 1], cl[:C - 1])):
            raise RuntimeError("pivotSteps: label replay disagrees with the device labels")
    z = None
    if res["status"] == native.STATUS_OPTIMAL:
        z = -res["fun"] if maximize else res["fun"]
    return {"pivotSteps": steps, "status": res["status"], "columns": _column_names(n, m, var_ids),
            "optimalValue": z, "solution": res["x"].tolist() if res["status"] == 0 else None,
            "truncated": res["n_pivots"] > cap}
