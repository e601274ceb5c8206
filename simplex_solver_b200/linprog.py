"""`linprog`-shaped entry to the GPU simplex: the seam the reference calls at
/root/reference/app/controllers/solver_controller.py:78-85.

    result = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds,
                     method='highs-ds', options={"presolve": True, "time_limit": 10})

The reference consumes `.success` (:89, :382), `.x[i]` (:388), `.fun` (:393, objective of the MINIMISATION form,
i.e. -z* for a maximise problem), `.status` (:404; 2 = infeasible, 3 = unbounded, 1 = limit) and `.message`
(:398, :411).  Everything else about scipy's OptimizeResult is not part of the contract.

Two quirks of the caller are handled here:
  * an '=' constraint arrives three times -- once in A_eq and as a <= / >= pair in A_ub (:154-161).  The pair is
    implied by the equality and is dropped before a tableau is built (it would only add degenerate rows);
  * a '>=' constraint arrives negated in A_ub with a negative right-hand side (:150-152).  The tableau builder
    flips such rows back (b < 0) and starts them with a surplus + artificial variable (two-phase method).
"""
from __future__ import annotations

import numpy as np

from . import native

_MESSAGES = {
    native.STATUS_OPTIMAL: "Optimization terminated successfully. (B200 tableau simplex: Optimal)",
    native.STATUS_LIMIT: "Iteration or time limit reached. (B200 tableau simplex: pivot budget or time_limit exhausted)",
    native.STATUS_INFEASIBLE: "The problem is infeasible. (B200 tableau simplex: phase 1 optimum > 0)",
    native.STATUS_UNBOUNDED: "The problem is unbounded. (B200 tableau simplex: no leaving row)",
    native.STATUS_NUMERICAL: "Numerical difficulties encountered. (B200 tableau simplex)",
}

_RULES = {"dantzig": native.RULE_DANTZIG, "bland": native.RULE_BLAND,
          native.RULE_DANTZIG: native.RULE_DANTZIG, native.RULE_BLAND: native.RULE_BLAND}


class OptimizeResult(dict):
    """Attribute-style result, like scipy.optimize.OptimizeResult (a dict with attribute access)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    __setattr__ = dict.__setitem__


def _as_2d(A, n):
    if A is None:
        return np.zeros((0, n))
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 1:
        A = A.reshape(1, -1)
    if A.shape[1] != n:
        raise ValueError(f"constraint matrix has {A.shape[1]} columns, expected {n}")
    return A


def _check_bounds(bounds, n):
    """Only x >= 0 is supported -- exactly what the reference passes (solver_controller.py:163)."""
    if bounds is None:
        return
    if isinstance(bounds, tuple) and len(bounds) == 2 and not isinstance(bounds[0], (tuple, list)):
        bounds = [bounds] * n
    for lo, hi in bounds:
        if (lo not in (0, 0.0)) or (hi is not None and np.isfinite(hi)):
            raise NotImplementedError("libb200lp supports bounds (0, None) only, as the reference passes them")


def drop_rows_implied_by_equalities(A_ub, b_ub, A_eq, b_eq):
    """Remove A_ub rows equal to (a, b) or (-a, -b) of some equality row (solver_controller.py:154-161)."""
    if len(b_eq) == 0 or len(b_ub) == 0:
        return A_ub, b_ub
    keys = set()
    for a, b in zip(A_eq, b_eq):
        row = np.append(a, b) + 0.0  # +0.0 normalises -0.0
        keys.add(row.tobytes())
        keys.add((-row + 0.0).tobytes())
    keep = [i for i in range(len(b_ub)) if (np.append(A_ub[i], b_ub[i]) + 0.0).tobytes() not in keys]
    return A_ub[keep], b_ub[keep]


def rows_from_linprog_args(c, A_ub=None, b_ub=None, A_eq=None, b_eq=None):
    """(A_ub, b_ub, A_eq, b_eq) -> one row list (A, b, ops) for the tableau builder."""
    c = np.asarray(c, dtype=np.float64).ravel()
    n = c.size
    A_ub = _as_2d(A_ub, n)
    A_eq = _as_2d(A_eq, n)
    b_ub = np.zeros(0) if b_ub is None else np.asarray(b_ub, dtype=np.float64).ravel()
    b_eq = np.zeros(0) if b_eq is None else np.asarray(b_eq, dtype=np.float64).ravel()
    if len(b_ub) != A_ub.shape[0] or len(b_eq) != A_eq.shape[0]:
        raise ValueError("right-hand side length does not match the constraint matrix")
    A_ub, b_ub = drop_rows_implied_by_equalities(A_ub, b_ub, A_eq, b_eq)
    A = np.vstack([A_ub, A_eq]) if (len(b_ub) + len(b_eq)) else np.zeros((0, n))
    b = np.concatenate([b_ub, b_eq])
    ops = np.concatenate([np.full(len(b_ub), native.OP_LE, dtype=np.int8),
                          np.full(len(b_eq), native.OP_EQ, dtype=np.int8)])
    return c, A, b, ops


def linprog(c, A_ub=None, b_ub=None, A_eq=None, b_eq=None, bounds=None, method="highs-ds", callback=None,
            options=None, x0=None, integrality=None, *, device: int = 0, hist_cap: int = 0):
    """Minimise c'x subject to A_ub x <= b_ub, A_eq x = b_eq, x >= 0 on the GPU (two-phase tableau simplex).

    `method` is accepted and ignored (the reference passes 'highs-ds').  Recognised `options`:
    "rule" ('dantzig' | 'bland', default 'dantzig'), "maxiter" (pivot budget), "eps_cost", "eps_pivot",
    "eps_feas", "update_variant", and "time_limit": seconds of wall clock for the whole call (copies, tableau build
    and both phases); when it expires the solve stops with status 1, which the reference maps to "Error" exactly as
    it does for HiGHS's own limit (solver_controller.py:76, :404).  "presolve" is accepted and has no effect.
    """
    options = dict(options or {})
    c, A, b, ops = rows_from_linprog_args(c, A_ub, b_ub, A_eq, b_eq)
    _check_bounds(bounds, c.size)
    rule = _RULES[options.get("rule", "dantzig")]
    opts = native.make_opts(rule=rule, max_pivots=options.get("maxiter"),
                            eps_cost=options.get("eps_cost", 1e-9), eps_pivot=options.get("eps_pivot", 1e-9),
                            eps_feas=options.get("eps_feas", 1e-7),
                            update_variant=options.get("update_variant", native.UPDATE_AUTO),
                            time_limit=options.get("time_limit"))
    solver = native.thread_solver(device)
    r = solver.solve_dense(A, b, c, ops, opts, hist_cap=hist_cap)
    status = r["status"]
    ok = status == native.STATUS_OPTIMAL
    res = OptimizeResult(
        success=ok, status=status, message=_MESSAGES.get(status, "Unknown status"),
        x=r["x"] if ok else None, fun=r["fun"] if ok else None,
        nit=r["n_pivots"], nit_phase1=r["n_phase1"], device_ms=r["device_ms"],
        kernel_launches=r["kernel_launches"],
    )
    if ok:
        slack_rows = A[ops == native.OP_LE]
        res["slack"] = b[ops == native.OP_LE] - slack_rows @ r["x"] if len(slack_rows) else np.zeros(0)
        eq_rows = A[ops == native.OP_EQ]
        res["con"] = b[ops == native.OP_EQ] - eq_rows @ r["x"] if len(eq_rows) else np.zeros(0)
    for k in ("piv_row", "piv_col", "enter_lab", "leave_lab"):
        if k in r:
            res[k] = r[k]
    return res
