"""Synthetic LPs of the BASELINE.json configs and the reference's known-answer problems.

Row operators are encoded as int8: 0 '<=', 1 '>=', 2 '=' (the L/G/E codes of
/root/reference/app/controllers/solver_controller.py:305-306).
"""
from __future__ import annotations

import numpy as np

LE, GE, EQ = 0, 1, 2
OP_CODE = {"<=": LE, ">=": GE, "=": EQ}
OP_TEXT = {LE: "<=", GE: ">=", EQ: "="}

BATCH_BLOCK = 1000  # LPs per RNG block of the batched generator (shards must start on a block boundary)


def dense_feasible_lp(n: int, seed: int = 0, m: int | None = None):
    """BASELINE config 2: max c'x, Ax <= b, x >= 0 with A~U[0,1), b = A x0 + U[0.1,1), c~U[0.1,1).

    Always feasible (x0) and bounded (A > 0).  Returns (A, b, c, ops, maximize=True).
    """
    m = n if m is None else m
    rng = np.random.default_rng(seed)
    A = rng.random((m, n))
    x0 = rng.random(n)
    b = A @ x0 + rng.uniform(0.1, 1.0, m)
    c = rng.uniform(0.1, 1.0, n)
    return A, b, c, np.zeros(m, dtype=np.int8), True


def batched_small_lps(start: int, count: int, m: int = 20, n: int = 30, base_seed: int = 3):
    """BASELINE config 3: `count` independent LPs with global indices [start, start+count).

    Each LP: A~U[-1,1) (m x n); operators i.i.d. {<=:60%, >=:25%, =:15%}; x0~U[0,1);
    b_i = a_i.x0 + s_i (s_i = U[0.1,1) for <=, -U[0.1,1) for >=, 0 for =)  -> feasible by construction;
    objective minimise c~U[0.1,1) (bounded: c > 0, x >= 0).
    LPs with global index = 0 mod 100 get the infeasible pair x1 <= 5, x1 >= 10 in rows 0,1;
    LPs with global index = 1 mod 100 become A <- |A|, all rows >=, maximise c (unbounded along 1).

    Returns A[count,m,n], b[count,m], c[count,n] (costs of the MINIMISATION form), ops[count,m] int8.
    """
    assert start % BATCH_BLOCK == 0, "shards start on a generator block boundary"
    A = np.empty((count, m, n))
    b = np.empty((count, m))
    c = np.empty((count, n))
    ops = np.empty((count, m), dtype=np.int8)
    done = 0
    while done < count:
        blk = (start + done) // BATCH_BLOCK
        k = min(BATCH_BLOCK, count - done)
        rng = np.random.default_rng([base_seed, blk])
        Ab = rng.uniform(-1.0, 1.0, (BATCH_BLOCK, m, n))
        u = rng.random((BATCH_BLOCK, m))
        ob = np.where(u < 0.60, LE, np.where(u < 0.85, GE, EQ)).astype(np.int8)
        x0 = rng.random((BATCH_BLOCK, n))
        s = rng.uniform(0.1, 1.0, (BATCH_BLOCK, m))
        cb = rng.uniform(0.1, 1.0, (BATCH_BLOCK, n))
        gidx = blk * BATCH_BLOCK + np.arange(BATCH_BLOCK)
        unb = gidx % 100 == 1
        Ab[unb] = np.abs(Ab[unb])
        ob[unb] = GE
        cb[unb] = -cb[unb]  # maximise c'x == minimise -c'x
        slack = np.where(ob == LE, s, np.where(ob == GE, -s, 0.0))
        bb = np.einsum("kij,kj->ki", Ab, x0) + slack
        if m >= 2:  # the infeasible pair needs two rows
            inf = gidx % 100 == 0
            Ab[inf, 0, :] = 0.0
            Ab[inf, 0, 0] = 1.0
            bb[inf, 0] = 5.0
            ob[inf, 0] = LE
            Ab[inf, 1, :] = 0.0
            Ab[inf, 1, 0] = 1.0
            bb[inf, 1] = 10.0
            ob[inf, 1] = GE
        A[done:done + k] = Ab[:k]
        b[done:done + k] = bb[:k]
        c[done:done + k] = cb[:k]
        ops[done:done + k] = ob[:k]
        done += k
    return A, b, c, ops


def fuzz_lp(k: int, seed: int = 12345):
    """k-th LP of a ragged stress family (shapes 1..12 x 1..12, mixed operators): rounded/degenerate coefficients
    (k % 4 == 0), dense uniform (1), sparse (2), badly scaled rows 1e-3..1e3 (3); every 5th has a random right-hand
    side (often infeasible); costs in [-1, 2) (often unbounded).  Returns (A, b, c_min, ops)."""
    rng = np.random.default_rng([seed, k])
    mm = int(rng.integers(1, 13))
    nn = int(rng.integers(1, 13))
    kind = k % 4
    if kind == 0:
        A = np.round(rng.uniform(-3, 3, (mm, nn)), 0)
        x0 = np.round(rng.uniform(0, 2, nn), 0)
    elif kind == 1:
        A = rng.uniform(-1, 1, (mm, nn))
        x0 = rng.uniform(0, 1, nn)
    elif kind == 2:
        A = rng.uniform(-2, 2, (mm, nn)) * (rng.random((mm, nn)) < 0.4)
        x0 = rng.uniform(0, 1, nn)
    else:
        A = rng.uniform(-1, 1, (mm, nn)) * 10.0 ** rng.integers(-3, 4, (mm, 1))
        x0 = rng.uniform(0, 1, nn)
    ops = rng.integers(0, 3, mm).astype(np.int8)
    slack = rng.uniform(0, 2, mm) * (rng.random(mm) < 0.7)
    b = A @ x0 + np.where(ops == LE, slack, np.where(ops == GE, -slack, 0.0))
    if k % 5 == 0:
        b = rng.uniform(-3, 3, mm)
    c = rng.uniform(-1, 2, nn)
    if kind == 0:
        c = np.round(c, 0)
    return A + 0.0, b + 0.0, c + 0.0, ops


def lp_to_problem_dict(A, b, c, ops, maximize: bool):
    """Array LP -> the reference's `problema_definicion` wrapper (ui_controller.py:63-67).

    `c` are the coefficients of the objective as the user states it (max or min).  Variable names are
    zero-padded so that the reference's lexicographic `sorted()` (solver_controller.py:46) keeps the order.
    """
    n = len(c)
    width = max(1, len(str(n)))
    names = [f"x{j + 1:0{width}d}" for j in range(n)]
    cons = []
    for i in range(len(b)):
        cons.append({
            "coefficients": {names[j]: float(A[i][j]) for j in range(n)},
            "operator": OP_TEXT[int(ops[i])],
            "rhs": float(b[i]),
        })
    return {"problema_definicion": {
        "funcion_objetivo": {"type": "maximize" if maximize else "minimize",
                             "coefficients": {names[j]: float(c[j]) for j in range(n)}},
        "restricciones": cons,
    }}


def _kat(obj_type, coeffs, rows):
    names = sorted(coeffs)
    return {"problema_definicion": {
        "funcion_objetivo": {"type": obj_type, "coefficients": dict(coeffs)},
        "restricciones": [{"coefficients": {k: float(v) for k, v in zip(names, r[0])}, "operator": r[1],
                           "rhs": float(r[2])} for r in rows],
    }}


def known_answer_problems():
    """The reference's own fixtures in its `problema_definicion` format."""
    return {
        # tests/test_visualization_integration.py:38-48, tests/test_performance_load.py:31-38
        "K1_wyndor": _kat("maximize", {"x1": 3.0, "x2": 5.0},
                          [([1, 0], "<=", 4), ([0, 2], "<=", 12), ([3, 2], "<=", 18)]),
        # tests/test_solver_controller.py:32-40
        "K2_max3": _kat("maximize", {"x1": 15.0, "x2": 18.0},
                        [([4, 2], "<=", 2000), ([2, 6], "<=", 2400), ([20, 28], "<=", 14000)]),
        # tests/test_solver_controller.py:16-24
        "K3_min_ge": _kat("minimize", {"x1": 50.0, "x2": 80.0},
                          [([4, 1], ">=", 4), ([1, 6], ">=", 6), ([4, 6], ">=", 12)]),
        # tests/test_visualization_integration.py:59-68
        "K4_min_ge2": _kat("minimize", {"x1": 2.0, "x2": 3.0}, [([1, 1], ">=", 5), ([2, 1], ">=", 8)]),
        # tests/test_visualization_integration.py:208-217
        "K5_eq": _kat("maximize", {"x1": 1.0, "x2": 1.0}, [([1, 1], "=", 10), ([2, 1], "<=", 15)]),
        # tests/test_visualization_integration.py:244-253
        "K6_infeasible": _kat("maximize", {"x1": 1.0}, [([1], "<=", 5), ([1], ">=", 10)]),
        # tests/test_visualization_integration.py:349-355
        "K7_unbounded": _kat("maximize", {"x1": 1.0, "x2": 1.0}, []),
        # tests/test_visualization_integration.py:379-388
        "K8_zero_cost": _kat("maximize", {"x1": 0.0, "x2": 5.0}, [([1, 0], "<=", 10), ([0, 1], "<=", 5)]),
        # tests/test_visualization_integration.py:422-444
        "K9_three_var": _kat("maximize", {"x1": 1.0, "x2": 2.0, "x3": 3.0},
                             [([1, 1, 1], "<=", 10), ([2, 1, 0], "<=", 15), ([0, 2, 1], "<=", 12)]),
        # docs/technical_documentation.md:307-347
        "K10_doc": _kat("maximize", {"x1": 1.0, "x2": 1.0, "x3": 4.0},
                        [([1, 2, 0], "<=", 10), ([0, 1, 3], "<=", 30)]),
    }


def problem_dict_to_arrays(wrapper):
    """`problema_definicion` wrapper -> (A, b, c_user, ops, maximize, variables) with the reference's
    variable order (sorted keys, solver_controller.py:46) and missing coefficients read as 0 (:131,:142)."""
    d = wrapper["problema_definicion"]
    obj = d["funcion_objetivo"]
    variables = sorted(obj["coefficients"].keys())
    c = np.array([obj["coefficients"].get(v, 0) for v in variables], dtype=np.float64)
    cons = d["restricciones"]
    A = np.array([[k["coefficients"].get(v, 0) for v in variables] for k in cons], dtype=np.float64)
    A = A.reshape(len(cons), len(variables))
    b = np.array([k["rhs"] for k in cons], dtype=np.float64)
    ops = np.array([OP_CODE[k["operator"]] for k in cons], dtype=np.int8)
    return A, b, c, ops, obj["type"] == "maximize", variables
