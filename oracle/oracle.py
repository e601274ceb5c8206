"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (simplex_solver_b200/) must never import this module.

The algorithm restated by the C file is described in its header (reference call sites:
/root/reference/app/controllers/solver_controller.py:78-85 and :290-319).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

RULE_DANTZIG, RULE_BLAND = 0, 1
OPT, LIMIT, INFEASIBLE, UNBOUNDED, NUMERICAL = 0, 1, 2, 3, 4
LE, GE, EQ = 0, 1, 2

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i8p = C.POINTER(C.c_int8)


class Tableau(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("n_obj", C.c_int64), ("R", C.c_int64), ("C", C.c_int64), ("ld", C.c_int64),
        ("n_struct", C.c_int64), ("art_base", C.c_int32),
        ("T", _f64p), ("rowlab", _i32p), ("collab", _i32p),
    ]


class Opts(C.Structure):
    _fields_ = [
        ("rule", C.c_int32), ("max_pivots", C.c_int64),
        ("eps_cost", C.c_double), ("eps_pivot", C.c_double), ("eps_feas", C.c_double),
        ("threads", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("fun", C.c_double), ("n_pivots", C.c_int64), ("n_phase1", C.c_int64),
        ("hist_cap", C.c_int64),
        ("piv_row", _i32p), ("piv_col", _i32p), ("enter_lab", _i32p), ("leave_lab", _i32p),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only; no reference sources involved)."""
    src = os.path.join(_HERE, "simplex_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.orc_gen_entry.restype = C.c_double
        L.orc_gen_entry.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.orc_generate.argtypes = [C.POINTER(Tableau), C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.orc_build.argtypes = [C.POINTER(Tableau), _f64p, C.c_int64, _f64p, _f64p, _i8p, C.c_int64, C.c_int64]
        L.orc_free.argtypes = [C.POINTER(Tableau)]
        L.orc_free.restype = None
        L.orc_price.restype = C.c_int64
        L.orc_price.argtypes = [C.POINTER(Tableau), C.c_int64, C.c_int32, C.c_double]
        L.orc_ratio.restype = C.c_int64
        L.orc_ratio.argtypes = [C.POINTER(Tableau), _f64p, C.c_double]
        L.orc_extract_col.restype = None
        L.orc_extract_col.argtypes = [C.POINTER(Tableau), C.c_int64, _f64p]
        L.orc_pivot_col.restype = None
        L.orc_pivot_col.argtypes = [C.POINTER(Tableau), C.c_int64, _f64p, C.c_int64, C.c_int32, C.c_int32]
        L.orc_pivot.restype = None
        L.orc_pivot.argtypes = [C.POINTER(Tableau), C.c_int64, C.c_int64, C.c_int32]
        L.orc_solve.argtypes = [C.POINTER(Tableau), C.POINTER(Opts), C.POINTER(Result)]
        L.orc_read_x.restype = None
        L.orc_read_x.argtypes = [C.POINTER(Tableau), _f64p]
        L.orc_solve_lp.argtypes = [_f64p, C.c_int64, _f64p, _f64p, _i8p, C.c_int64, C.c_int64,
                                   C.POINTER(Opts), C.POINTER(Result), _f64p]
        L.orc_solve_batched.argtypes = [C.c_int64, C.c_int64, C.c_int64, _f64p, _f64p, _f64p, _i8p,
                                        C.POINTER(Opts), _i32p, _f64p, _f64p, _i32p, _i32p, C.c_int64, C.c_int32]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.restype = None
        L.orc_set_threads.argtypes = [C.c_int]
        _i64p = C.POINTER(C.c_int64)
        L.orc_full_dims.argtypes = [_f64p, _i8p, C.c_int64, C.c_int64, _i64p, _i64p]
        L.orc_full_steps.argtypes = [_f64p, C.c_int64, _f64p, _f64p, _i8p, C.c_int64, C.c_int64, C.POINTER(Opts),
                                     C.c_int64, _f64p, _i32p, _i32p, _i32p, _i32p, _i64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def make_opts(rule=RULE_DANTZIG, max_pivots=1 << 40, eps_cost=1e-9, eps_pivot=1e-9, eps_feas=1e-7, threads=1):
    return Opts(rule, max_pivots, eps_cost, eps_pivot, eps_feas, threads)


class OracleTableau:
    """Owning wrapper of an orc_tableau with numpy views on its storage."""

    def __init__(self):
        self.t = Tableau()
        self._alive = False

    @classmethod
    def generate(cls, seed, m, n_total, lab0=0, ncols=None, ld=0):
        self = cls()
        ncols = n_total if ncols is None else ncols
        if lib().orc_generate(C.byref(self.t), seed, m, n_total, lab0, ncols, ld):
            raise MemoryError("orc_generate")
        self._alive = True
        return self

    @classmethod
    def build(cls, A, b, c, ops):
        self = cls()
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        ops = np.ascontiguousarray(ops, dtype=np.int8)
        m, n = (A.shape if A.size else (len(b), len(c)))
        if lib().orc_build(C.byref(self.t), _p(A, _f64p), max(n, 1), _p(b, _f64p), _p(c, _f64p), _p(ops, _i8p), m, n):
            raise MemoryError("orc_build")
        self._alive = True
        return self

    def __del__(self):
        if self._alive and _lib is not None and C is not None:
            try:
                _lib.orc_free(C.byref(self.t))
            except Exception:
                pass
            self._alive = False

    # numpy views (no copy)
    @property
    def T(self):
        t = self.t
        full = np.ctypeslib.as_array(t.T, shape=(t.R, t.ld))
        return full[:, : t.C]

    @property
    def rowlab(self):
        return np.ctypeslib.as_array(self.t.rowlab, shape=(self.t.R,))

    @property
    def collab(self):
        return np.ctypeslib.as_array(self.t.collab, shape=(self.t.C,))

    def price(self, obj_row=None, rule=RULE_DANTZIG, eps=1e-9):
        return int(lib().orc_price(C.byref(self.t), self.t.m if obj_row is None else obj_row, rule, eps))

    def extract_col(self, s):
        col = np.empty(self.t.R, dtype=np.float64)
        lib().orc_extract_col(C.byref(self.t), s, _p(col, _f64p))
        return col

    def ratio(self, col, eps=1e-9):
        col = np.ascontiguousarray(col, dtype=np.float64)
        return int(lib().orc_ratio(C.byref(self.t), _p(col, _f64p), eps))

    def pivot(self, r, s, threads=1):
        lib().orc_pivot(C.byref(self.t), r, s, threads)

    def pivot_col(self, r, col, s_local, enter_lab, threads=1):
        col = np.ascontiguousarray(col, dtype=np.float64)
        lib().orc_pivot_col(C.byref(self.t), r, _p(col, _f64p), s_local, enter_lab, threads)

    def solve(self, opts=None, hist_cap=0):
        opts = opts or make_opts()
        res, keep = _make_result(hist_cap)
        lib().orc_solve(C.byref(self.t), C.byref(opts), C.byref(res))
        return _result_dict(res, keep)

    def read_x(self):
        x = np.empty(self.t.n_struct, dtype=np.float64)
        lib().orc_read_x(C.byref(self.t), _p(x, _f64p))
        return x


def _make_result(hist_cap):
    res = Result()
    keep = None
    if hist_cap > 0:
        keep = [np.full(hist_cap, -1, dtype=np.int32) for _ in range(4)]
        res.hist_cap = hist_cap
        res.piv_row, res.piv_col, res.enter_lab, res.leave_lab = (_p(a, _i32p) for a in keep)
    return res, keep


def _result_dict(res, keep):
    out = {"status": int(res.status), "fun": float(res.fun), "n_pivots": int(res.n_pivots),
           "n_phase1": int(res.n_phase1)}
    if keep is not None:
        k = min(int(res.n_pivots), len(keep[0]))
        out.update(piv_row=keep[0][:k].copy(), piv_col=keep[1][:k].copy(),
                   enter_lab=keep[2][:k].copy(), leave_lab=keep[3][:k].copy())
    return out


def solve_lp(A, b, c, ops, opts=None, hist_cap=0):
    """min c'x s.t. A_i x (ops_i) b_i, x >= 0.  Returns dict(status, fun, x, n_pivots, ...)."""
    opts = opts or make_opts()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    ops = np.ascontiguousarray(ops, dtype=np.int8)
    n = len(c)
    m = len(b)
    x = np.zeros(n, dtype=np.float64)
    res, keep = _make_result(hist_cap)
    rc = lib().orc_solve_lp(_p(A, _f64p), max(n, 1), _p(b, _f64p), _p(c, _f64p), _p(ops, _i8p), m, n,
                            C.byref(opts), C.byref(res), _p(x, _f64p))
    if rc:
        raise MemoryError("orc_solve_lp")
    out = _result_dict(res, keep)
    out["x"] = x
    return out


def solve_batched(A, b, c, ops, opts=None, log_cap=0, threads=1):
    opts = opts or make_opts()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    ops = np.ascontiguousarray(ops, dtype=np.int8)
    B, m, n = A.shape
    status = np.empty(B, dtype=np.int32)
    fun = np.empty(B, dtype=np.float64)
    x = np.empty((B, n), dtype=np.float64)
    npiv = np.empty(B, dtype=np.int32)
    log = np.full((B, max(log_cap, 1), 2), -1, dtype=np.int32) if log_cap > 0 else None
    lib().orc_solve_batched(B, m, n, _p(A, _f64p), _p(b, _f64p), _p(c, _f64p), _p(ops, _i8p), C.byref(opts),
                            _p(status, _i32p), _p(fun, _f64p), _p(x, _f64p), _p(npiv, _i32p),
                            _p(log, _i32p) if log is not None else None, log_cap, threads)
    return {"status": status, "fun": fun, "x": x, "n_pivots": npiv, "piv_log": log}


def full_steps(A, b, c, ops, opts=None, cap=64):
    """Textbook FULL-tableau simplex (orc_full_steps): the displayed tableau of every step, as `pivotSteps` shows it.

    Returns dict(status, n_pivots, n_phase1, var_ids, basis, steps) with steps[k] = (tableau R x W, pivot row,
    pivot column index in the displayed tableau); steps[0] is the initial tableau with (None, None)."""
    opts = opts or make_opts()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    ops = np.ascontiguousarray(ops, dtype=np.int8)
    m, n = len(b), len(c)
    R, W = C.c_int64(), C.c_int64()
    lib().orc_full_dims(_p(b, _f64p), _p(ops, _i8p), m, n, C.byref(R), C.byref(W))
    R, W = int(R.value), int(W.value)
    snaps = np.zeros((cap + 1, R, W), dtype=np.float64)
    pr = np.full(max(cap, 1), -1, dtype=np.int32)
    pc = np.full(max(cap, 1), -1, dtype=np.int32)
    var_ids = np.zeros(max(W - 1, 1), dtype=np.int32)
    basis = np.zeros(max(m, 1), dtype=np.int32)
    out = np.zeros(6, dtype=np.int64)
    rc = lib().orc_full_steps(_p(A, _f64p), max(n, 1), _p(b, _f64p), _p(c, _f64p), _p(ops, _i8p), m, n, C.byref(opts),
                              cap, _p(snaps, _f64p), _p(pr, _i32p), _p(pc, _i32p), _p(var_ids, _i32p),
                              _p(basis, _i32p), out.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc:
        raise MemoryError("orc_full_steps")
    n_piv = int(out[2])
    k = min(n_piv, cap)
    steps = [(snaps[0], None, None)] + [(snaps[i + 1], int(pr[i]), int(pc[i])) for i in range(k)]
    return {"status": int(out[3]), "n_pivots": n_piv, "n_phase1": int(out[4]), "var_ids": var_ids[: W - 1].copy(),
            "basis": basis[:m].copy(), "steps": steps, "truncated": n_piv > cap}


def max_threads():
    return int(lib().orc_max_threads())


def host_cores():
    """Cores this process may run on (its affinity mask) -- not OMP_NUM_THREADS, which launchers such as torchrun set
    to 1 for their workers."""
    import os
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n):
    lib().orc_set_threads(int(n))
