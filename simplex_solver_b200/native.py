"""ctypes binding of libb200lp.so (the C ABI of include/b200lp.h) -- the only compute path of this package.

There is no CPU fallback: if the shared library is missing, or no CUDA device is present, every solver
entry point raises.  PyTorch is used by callers for device memory and streams only; nothing here needs it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200LP_LIB") or os.path.join(_HERE, "libb200lp.so")  # B200LP_LIB: A/B builds of experiments
CSRC = os.path.join(_HERE, "csrc")

RULE_DANTZIG, RULE_BLAND = 0, 1
STATUS_OPTIMAL, STATUS_LIMIT, STATUS_INFEASIBLE, STATUS_UNBOUNDED, STATUS_NUMERICAL = 0, 1, 2, 3, 4
OP_LE, OP_GE, OP_EQ = 0, 1, 2
UPDATE_AUTO, UPDATE_LDG, UPDATE_TMA = 0, 1, 2
LOOP_LAUNCHES, LOOP_GRAPH, LOOP_AUTO, LOOP_BLOCKED = 0, 1, 2, 3

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # nothing is fused except the explicit fma of the arithmetic contract
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]

EXPORTS = [
    "b200lp_version", "b200lp_last_error", "b200lp_default_opts", "b200lp_create", "b200lp_destroy",
    "b200lp_set_stream", "b200lp_synchronize", "b200lp_solve_dense", "b200lp_attach", "b200lp_dims",
    "b200lp_generate", "b200lp_set_labels", "b200lp_get_labels", "b200lp_read_tableau", "b200lp_read_solution",
    "b200lp_run", "b200lp_solve", "b200lp_select_entering", "b200lp_ratio_test", "b200lp_pivot",
    "b200lp_shard_candidate", "b200lp_shard_pivot", "b200lp_shard_state", "b200lp_shard_reset",
    "b200lp_read_history", "b200lp_solve_batched", "b200lp_time_update", "b200lp_build_dense",
    "b200lp_set_snapshots", "b200lp_profile_loop", "b200lp_use_own_stream", "b200lp_shard_blk_begin",
    "b200lp_shard_blk_candidate", "b200lp_shard_blk_pivot", "b200lp_shard_blk_flush", "b200lp_p2p_bytes",
    "b200lp_p2p_connect", "b200lp_shard_fused", "b200lp_shard_fused_multi", "b200lp_check_guards",
    "b200lp_binding_epoch",
]

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i8p = C.POINTER(C.c_int8)


class Opts(C.Structure):
    _fields_ = [
        ("rule", C.c_int32), ("update_variant", C.c_int32), ("max_pivots", C.c_int64),
        ("eps_cost", C.c_double), ("eps_pivot", C.c_double), ("eps_feas", C.c_double),
        ("check_every", C.c_int32), ("loop_mode", C.c_int32), ("time_limit_s", C.c_double),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("n", C.c_int64), ("lda", C.c_int64),
        ("A", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p), ("ops", C.c_void_p),
        ("on_device", C.c_int32), ("reserved", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("reserved", C.c_int32), ("fun", C.c_double),
        ("n_pivots", C.c_int64), ("n_phase1", C.c_int64),
        ("x", C.c_void_p), ("x_len", C.c_int64),
        ("piv_row", C.c_void_p), ("piv_col", C.c_void_p), ("enter_lab", C.c_void_p), ("leave_lab", C.c_void_p),
        ("hist_cap", C.c_int64), ("device_ms", C.c_double), ("kernel_launches", C.c_int64),
    ]


class B200LPError(RuntimeError):
    pass


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile libb200lp.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  Without `force` an existing
    library newer than every source is kept (the GPU box has the prebuilt file and nothing to rebuild it for)."""
    src = os.path.join(CSRC, "b200lp.cu")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "b200lp.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


_lib = None
_lib_lock = threading.Lock()


def lib():
    """The loaded library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise B200LPError(
                        f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(this package has no CPU fallback)")
                L = C.CDLL(LIB_PATH)
                L.b200lp_last_error.restype = C.c_char_p
                L.b200lp_default_opts.restype = None
                L.b200lp_default_opts.argtypes = [C.POINTER(Opts)]
                L.b200lp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
                L.b200lp_destroy.argtypes = [C.c_void_p]
                L.b200lp_set_stream.argtypes = [C.c_void_p, C.c_void_p]
                L.b200lp_synchronize.argtypes = [C.c_void_p]
                L.b200lp_use_own_stream.argtypes = [C.c_void_p]
                L.b200lp_check_guards.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
                L.b200lp_binding_epoch.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
                L.b200lp_solve_dense.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Opts), C.POINTER(Result)]
                L.b200lp_build_dense.argtypes = [C.c_void_p, C.POINTER(Problem)]
                L.b200lp_set_snapshots.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
                L.b200lp_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                            C.c_int64, C.c_int32]
                L.b200lp_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4
                L.b200lp_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64]
                L.b200lp_set_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
                L.b200lp_get_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
                L.b200lp_read_tableau.argtypes = [C.c_void_p, C.c_void_p]
                L.b200lp_read_solution.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
                L.b200lp_run.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.POINTER(Result)]
                L.b200lp_solve.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(Result)]
                L.b200lp_select_entering.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.POINTER(C.c_int64)]
                L.b200lp_ratio_test.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.POINTER(C.c_int64)]
                L.b200lp_pivot.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32]
                L.b200lp_shard_candidate.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.c_void_p]
                L.b200lp_shard_pivot.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_void_p, C.c_int32, C.c_int32]
                L.b200lp_shard_blk_begin.argtypes = [C.c_void_p, C.c_int64]
                L.b200lp_shard_blk_candidate.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.c_void_p]
                L.b200lp_shard_blk_pivot.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_void_p, C.c_int32, C.c_int32]
                L.b200lp_shard_blk_flush.argtypes = [C.c_void_p, C.c_int64]
                L.b200lp_p2p_bytes.restype = C.c_int64
                L.b200lp_p2p_bytes.argtypes = [C.c_int64, C.c_int32]
                L.b200lp_p2p_connect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]
                L.b200lp_shard_fused.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.c_int32]
                L.b200lp_shard_fused_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(Opts), C.c_int64,
                                                       C.c_int32]
                L.b200lp_shard_state.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int64)]
                L.b200lp_shard_reset.argtypes = [C.c_void_p, C.c_int64]
                L.b200lp_read_history.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.POINTER(C.c_int64)]
                L.b200lp_solve_batched.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.POINTER(Opts), C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                                   C.POINTER(C.c_double)]
                L.b200lp_time_update.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                                 C.POINTER(C.c_double)]
                L.b200lp_profile_loop.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.c_int32] + \
                    [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int64)]
                _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise B200LPError(f"libb200lp error {rc}: {lib().b200lp_last_error().decode(errors='replace')}")


def make_opts(rule=RULE_DANTZIG, max_pivots=None, eps_cost=1e-9, eps_pivot=1e-9, eps_feas=1e-7,
              update_variant=UPDATE_AUTO, check_every=0, loop_mode=LOOP_AUTO, use_graph=None,
              time_limit=None) -> Opts:
    o = Opts()
    lib().b200lp_default_opts(C.byref(o))
    o.rule = rule
    if max_pivots is not None:
        o.max_pivots = int(max_pivots)
    o.eps_cost, o.eps_pivot, o.eps_feas = eps_cost, eps_pivot, eps_feas
    o.update_variant = update_variant
    o.check_every = check_every
    if use_graph is not None:  # explicit multi-kernel loop: graph replay or plain launches
        loop_mode = LOOP_GRAPH if use_graph else LOOP_LAUNCHES
    o.loop_mode = loop_mode
    o.time_limit_s = float(time_limit) if time_limit else 0.0  # seconds of wall clock per solve call; 0 = no bound
    return o


def _ptr(a):
    """Host ndarray -> void*; int -> device pointer as is; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def _check_lp_arrays(A, b, c, ops):
    """Host arrays of one LP -> contiguous (A, b, c, ops, m, n); raises ValueError on any shape disagreement (the C ABI
    takes plain pointers and sizes and would read past the end of a short buffer)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    ops = np.ascontiguousarray(ops, dtype=np.int8)
    if b.ndim != 1 or c.ndim != 1:
        raise ValueError(f"b and c must be vectors, got shapes {b.shape} and {c.shape}")
    m, n = len(b), len(c)
    A = np.ascontiguousarray(A, dtype=np.float64)
    if A.size == 0 and m * n == 0:
        A = A.reshape(m, n)
    if A.shape != (m, n):
        raise ValueError(f"A has shape {A.shape}, expected (len(b), len(c)) = {(m, n)}")
    if ops.shape != (m,):
        raise ValueError(f"ops has shape {ops.shape}, expected ({m},)")
    if m and (ops.min() < 0 or ops.max() > 2):
        raise ValueError("ops entries must be 0 ('<='), 1 ('>=') or 2 ('=')")
    return A, b, c, ops, m, n


class Solver:
    """Owning wrapper of one b200lp_solver workspace (one per thread; see `thread_solver`)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().b200lp_create(C.byref(self._h), device))
        self.device = device
        self._keep = None  # keeps an attached torch tensor alive

    def close(self):
        if self._h:
            lib().b200lp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- streams ------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        """Run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream; 0 = legacy default stream)."""
        check(lib().b200lp_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def use_own_stream(self):
        check(lib().b200lp_use_own_stream(self._h))

    def synchronize(self):
        check(lib().b200lp_synchronize(self._h))

    def check_guards(self) -> int:
        """Guard mode (B200LP_GUARD=1): bytes of the bands around the workspace buffers that were overwritten."""
        n = C.c_int64()
        check(lib().b200lp_check_guards(self._h, C.byref(n)))
        return int(n.value)

    # ---- results ------------------------------------------------------------------------------------
    @staticmethod
    def _result(n_x: int, hist_cap: int):
        res = Result()
        keep = {"x": np.zeros(max(n_x, 1), dtype=np.float64)}
        res.x = _ptr(keep["x"])
        res.x_len = n_x
        if hist_cap > 0:
            for k in ("piv_row", "piv_col", "enter_lab", "leave_lab"):
                keep[k] = np.full(hist_cap, -1, dtype=np.int32)
                setattr(res, k, _ptr(keep[k]))
            res.hist_cap = hist_cap
        return res, keep

    @staticmethod
    def _result_dict(res, keep, n_x):
        out = {"status": int(res.status), "fun": float(res.fun), "n_pivots": int(res.n_pivots),
               "n_phase1": int(res.n_phase1), "device_ms": float(res.device_ms),
               "kernel_launches": int(res.kernel_launches), "x": keep["x"][:n_x].copy()}
        if "piv_row" in keep:
            k = min(int(res.n_pivots), len(keep["piv_row"]))
            for name in ("piv_row", "piv_col", "enter_lab", "leave_lab"):
                out[name] = keep[name][:k].copy()
        return out

    # ---- one LP from arrays ---------------------------------------------------------------------------
    def solve_dense(self, A, b, c, ops, opts: Opts | None = None, hist_cap: int = 0, on_device: bool = False,
                    m: int | None = None, n: int | None = None, lda: int | None = None):
        """min c'x s.t. A_i x (ops_i) b_i, x >= 0.  Host ndarrays, or device pointers (ints) with on_device."""
        opts = opts or make_opts()
        ops = np.ascontiguousarray(ops, dtype=np.int8)
        if not on_device:
            A, b, c, ops, m, n = _check_lp_arrays(A, b, c, ops)
            lda = n
        elif m is None or n is None or ops.shape != (m,):
            raise ValueError(f"on_device: m and n must be given and ops must have m = {m} entries, got {ops.shape}")
        p = Problem(m, n, lda if lda else n, _ptr(A), _ptr(b), _ptr(c), _ptr(ops), 1 if on_device else 0, 0)
        res, keep = self._result(n, hist_cap)
        check(lib().b200lp_solve_dense(self._h, C.byref(p), C.byref(opts), C.byref(res)))
        return self._result_dict(res, keep, n)

    def build_dense(self, A, b, c, ops):
        """First half of solve_dense: build the initial tableau on the device (host ndarrays in)."""
        A, b, c, ops, m, n = _check_lp_arrays(A, b, c, ops)
        p = Problem(m, n, n, _ptr(A), _ptr(b), _ptr(c), _ptr(ops), 0, 0)
        check(lib().b200lp_build_dense(self._h, C.byref(p)))
        mm, n_obj, Ccols, ld = self.dims()
        self._dims = (mm, n_obj, Ccols, ld, n)

    def set_snapshots(self, ptr: int | None, cap: int = 0, keep=None):
        check(lib().b200lp_set_snapshots(self._h, C.c_void_p(ptr or 0), cap))
        self._keep_snaps = keep

    # ---- device-resident tableau ----------------------------------------------------------------------
    def attach(self, T_ptr: int, m: int, n_obj: int, Ccols: int, ld: int, n_struct: int, art_base: int, keep=None):
        check(lib().b200lp_attach(self._h, C.c_void_p(T_ptr), m, n_obj, Ccols, ld, n_struct, art_base))
        self._keep = keep
        self._dims = (m, n_obj, Ccols, ld, n_struct)

    def generate(self, seed: int, n_total: int, lab0: int = 0):
        check(lib().b200lp_generate(self._h, seed, n_total, lab0))
        m, n_obj, Ccols, ld, _ = self._dims
        self._dims = (m, n_obj, Ccols, ld, n_total)

    def dims(self):
        v = [C.c_int64() for _ in range(4)]
        check(lib().b200lp_dims(self._h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def set_labels(self, rowlab, collab):
        rowlab = np.ascontiguousarray(rowlab, dtype=np.int32)
        collab = np.ascontiguousarray(collab, dtype=np.int32)
        check(lib().b200lp_set_labels(self._h, _ptr(rowlab), _ptr(collab)))

    def get_labels(self):
        m, n_obj, Ccols, _ = self.dims()
        rl = np.empty(m + n_obj, dtype=np.int32)
        cl = np.empty(Ccols, dtype=np.int32)
        check(lib().b200lp_get_labels(self._h, _ptr(rl), _ptr(cl)))
        return rl, cl

    def read_tableau(self):
        m, n_obj, Ccols, _ = self.dims()
        T = np.empty((m + n_obj, Ccols), dtype=np.float64)
        check(lib().b200lp_read_tableau(self._h, _ptr(T)))
        return T

    def read_solution(self):
        n = self._dims[4]
        x = np.zeros(max(n, 1), dtype=np.float64)
        fun = C.c_double()
        check(lib().b200lp_read_solution(self._h, _ptr(x), C.byref(fun)))
        return x[:n], float(fun.value)

    def run(self, opts: Opts | None = None, obj_row: int | None = None, hist_cap: int = 0):
        opts = opts or make_opts()
        m, n_obj, _, _ = self.dims()
        n = self._dims[4]
        res, keep = self._result(n, hist_cap)
        check(lib().b200lp_run(self._h, C.byref(opts), m if obj_row is None else obj_row, C.byref(res)))
        return self._result_dict(res, keep, n)

    def solve(self, opts: Opts | None = None, hist_cap: int = 0):
        opts = opts or make_opts()
        n = self._dims[4]
        res, keep = self._result(n, hist_cap)
        check(lib().b200lp_solve(self._h, C.byref(opts), C.byref(res)))
        return self._result_dict(res, keep, n)

    # ---- one phase at a time ----------------------------------------------------------------------------
    def select_entering(self, obj_row: int | None = None, rule: int = RULE_DANTZIG, eps_cost: float = 1e-9) -> int:
        m, _, _, _ = self.dims()
        out = C.c_int64()
        check(lib().b200lp_select_entering(self._h, m if obj_row is None else obj_row, rule, eps_cost, C.byref(out)))
        return int(out.value)

    def ratio_test(self, col: int, eps_pivot: float = 1e-9) -> int:
        out = C.c_int64()
        check(lib().b200lp_ratio_test(self._h, col, eps_pivot, C.byref(out)))
        return int(out.value)

    def pivot(self, row: int, col: int, update_variant: int = UPDATE_AUTO):
        check(lib().b200lp_pivot(self._h, row, col, update_variant))

    def time_update(self, row: int, col: int, update_variant: int = UPDATE_AUTO, reps: int = 10) -> float:
        ms = C.c_double()
        check(lib().b200lp_time_update(self._h, row, col, update_variant, reps, C.byref(ms)))
        return float(ms.value)

    def profile_loop(self, opts: Opts, iters: int = 16, obj_row: int | None = None):
        """Average ms of the price / ratio / update kernels over `iters` real loop iterations (CUDA events)."""
        m, _, _, _ = self.dims()
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        check(lib().b200lp_profile_loop(self._h, C.byref(opts), m if obj_row is None else obj_row, iters,
                                        C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        return {"price_ms": a.value, "ratio_ms": b.value, "update_ms": c.value, "pivots": int(n.value)}

    # ---- column shards ----------------------------------------------------------------------------------
    def shard_reset(self, max_pivots: int):
        check(lib().b200lp_shard_reset(self._h, max_pivots))

    def binding_epoch(self) -> int:
        """Changes whenever a caller-captured CUDA graph of this solver's launches may hold stale pointers / capacities."""
        e = C.c_int64()
        check(lib().b200lp_binding_epoch(self._h, C.byref(e)))
        return int(e.value)

    def shard_candidate(self, opts: Opts, obj_row: int, cand_ptr: int):
        check(lib().b200lp_shard_candidate(self._h, C.byref(opts), obj_row, C.c_void_p(cand_ptr)))

    def shard_pivot(self, opts: Opts, gathered_ptr: int, world: int, rank: int):
        check(lib().b200lp_shard_pivot(self._h, C.byref(opts), C.c_void_p(gathered_ptr), world, rank))

    def shard_blk_begin(self, obj_row: int):
        check(lib().b200lp_shard_blk_begin(self._h, obj_row))

    def shard_blk_candidate(self, opts: Opts, obj_row: int, cand_ptr: int):
        check(lib().b200lp_shard_blk_candidate(self._h, C.byref(opts), obj_row, C.c_void_p(cand_ptr)))

    def shard_blk_pivot(self, opts: Opts, gathered_ptr: int, world: int, rank: int):
        check(lib().b200lp_shard_blk_pivot(self._h, C.byref(opts), C.c_void_p(gathered_ptr), world, rank))

    def shard_blk_flush(self, obj_row: int):
        check(lib().b200lp_shard_blk_flush(self._h, obj_row))

    @staticmethod
    def p2p_bytes(R: int, world: int) -> int:
        return int(lib().b200lp_p2p_bytes(R, world))

    def p2p_connect(self, bases, world: int, rank: int):
        arr = (C.c_void_p * world)(*[C.c_void_p(int(b)) for b in bases])
        check(lib().b200lp_p2p_connect(self._h, arr, world, rank))

    def shard_fused(self, opts: Opts, obj_row: int, lookahead: bool = False):
        """One pivot with the peer-memory exchange: fused price / exchange / winner / ratio kernel (+ rank-1 update)."""
        check(lib().b200lp_shard_fused(self._h, C.byref(opts), obj_row, 1 if lookahead else 0))

    @staticmethod
    def shard_fused_multi(solvers, opts: Opts, obj_row: int, lookahead: bool = False):
        """The same for n shards emulated on ONE GPU: one launch, one cluster per shard (tests)."""
        arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
        check(lib().b200lp_shard_fused_multi(arr, len(solvers), C.byref(opts), obj_row, 1 if lookahead else 0))

    def shard_state(self):
        done, status, n = C.c_int32(), C.c_int32(), C.c_int64()
        check(lib().b200lp_shard_state(self._h, C.byref(done), C.byref(status), C.byref(n)))
        return bool(done.value), int(status.value), int(n.value)

    def read_history(self, cap: int):
        arrs = [np.full(cap, -1, dtype=np.int32) for _ in range(4)]
        n = C.c_int64()
        check(lib().b200lp_read_history(self._h, cap, *[_ptr(a) for a in arrs], C.byref(n)))
        k = int(n.value)
        return {"piv_row": arrs[0][:k], "piv_col": arrs[1][:k], "enter_lab": arrs[2][:k], "leave_lab": arrs[3][:k]}

    # ---- batched ------------------------------------------------------------------------------------------
    def solve_batched(self, A, b, c, ops, opts: Opts | None = None, want_x: bool = True, log_cap: int = 0):
        """Host ndarrays A[B,m,n], b[B,m], c[B,n] (minimisation costs), ops[B,m] -> dict of host arrays."""
        opts = opts or make_opts()
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        ops = np.ascontiguousarray(ops, dtype=np.int8)
        if A.ndim != 3:
            raise ValueError(f"A must be [B, m, n], got shape {A.shape}")
        B, m, n = A.shape
        if b.shape != (B, m) or c.shape != (B, n) or ops.shape != (B, m):
            raise ValueError(f"batched LP shapes disagree: A {A.shape}, b {b.shape} (want {(B, m)}), "
                             f"c {c.shape} (want {(B, n)}), ops {ops.shape} (want {(B, m)})")
        status = np.empty(B, dtype=np.int32)
        fun = np.empty(B, dtype=np.float64)
        npiv = np.empty(B, dtype=np.int32)
        x = np.empty((B, n), dtype=np.float64) if want_x else None
        log = np.full((B, log_cap, 2), -1, dtype=np.int32) if log_cap > 0 else None
        ms = C.c_double()
        check(lib().b200lp_solve_batched(self._h, B, m, n, _ptr(A), _ptr(b), _ptr(c), _ptr(ops), C.byref(opts),
                                         _ptr(status), _ptr(fun), _ptr(x), _ptr(npiv), _ptr(log), log_cap, 0,
                                         C.byref(ms)))
        return {"status": status, "fun": fun, "x": x, "n_pivots": npiv, "piv_log": log, "device_ms": float(ms.value)}

    def solve_batched_device(self, B, m, n, A_ptr, b_ptr, c_ptr, ops_ptr, status_ptr, fun_ptr, x_ptr, npiv_ptr,
                             opts: Opts | None = None, log_ptr: int = 0, log_cap: int = 0) -> float:
        """Device pointers in, device pointers out (inputs already resident in HBM).  Returns kernel ms."""
        opts = opts or make_opts()
        ms = C.c_double()
        check(lib().b200lp_solve_batched(self._h, B, m, n, C.c_void_p(A_ptr), C.c_void_p(b_ptr), C.c_void_p(c_ptr),
                                         C.c_void_p(ops_ptr), C.byref(opts), C.c_void_p(status_ptr),
                                         C.c_void_p(fun_ptr), C.c_void_p(x_ptr or 0), C.c_void_p(npiv_ptr),
                                         C.c_void_p(log_ptr or 0), log_cap, 1, C.byref(ms)))
        return float(ms.value)


_tls = threading.local()


def thread_solver(device: int = 0) -> Solver:
    """One workspace per (thread, device): the library is re-entrant across workspaces, not within one."""
    cache = getattr(_tls, "solvers", None)
    if cache is None:
        cache = _tls.solvers = {}
    if device not in cache:
        cache[device] = Solver(device)
    return cache[device]
