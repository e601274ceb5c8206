"""Generate tests/golden/reference_golden.json by running the REFERENCE's own solve path.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What runs: app.controllers.solver_controller.SolverController(wrapper).run() from /root/reference,
unmodified, with real scipy HiGHS behind its `linprog` call (solver_controller.py:78-85).  The three
packages that are not installable offline and do not touch status/x*/z* -- gilp, simple_simplex, reportlab --
are stubbed in sys.modules; StorageService.save_solution is redirected to a temp directory.  The scipy
here is newer than the reference's pin (1.12.0): optimal values are version independent, message strings
and non-unique x* are not, so tests only compare x* where the optimum is unique.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)

from simplex_solver_b200 import workloads as W  # noqa: E402


def import_reference_controller():
    for name in ("gilp", "simple_simplex", "reportlab", "reportlab.lib", "reportlab.lib.pagesizes",
                 "reportlab.lib.styles", "reportlab.lib.units", "reportlab.lib.colors", "reportlab.lib.enums",
                 "reportlab.platypus"):
        sys.modules.setdefault(name, types.ModuleType(name))

    def _stub(*a, **k):
        raise RuntimeError("stub")

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    sys.modules["gilp"].LP = _stub
    sys.modules["gilp"].simplex_visual = _stub
    for f in ("create_tableau", "add_constraint", "add_objective", "optimize_json_format"):
        setattr(sys.modules["simple_simplex"], f, _stub)
    for mod in list(sys.modules):
        if mod.startswith("reportlab"):
            sys.modules[mod].__getattr__ = lambda name: _Any()  # type: ignore[attr-defined]
    sys.path.insert(0, REFERENCE)
    import app.controllers.solver_controller as sc  # noqa: E402
    import app.services.storage_service as ss  # noqa: E402
    tmp = tempfile.mkdtemp(prefix="golden_out_")
    ss.OUTPUT_DIR = tmp
    return sc


def run_reference(sc, wrapper):
    seen = {}
    real = sc.linprog

    def spy(*a, **k):
        r = real(*a, **k)
        seen["status"] = int(r.status)
        seen["nit"] = int(getattr(r, "nit", -1))
        return r

    sc.linprog = spy
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            rep = sc.SolverController(wrapper).run()
    finally:
        sc.linprog = real
    sol = rep["solucion_encontrada"]
    out = {"status_text": sol["status"], "scipy_status": seen.get("status"), "nit": seen.get("nit"),
           "z": None if sol["valor_optimo_z"] is None else float(sol["valor_optimo_z"]), "x": None}
    if sol["valores_variables"] is not None:
        out["x"] = [float(sol["valores_variables"][k]) for k in sorted(sol["valores_variables"])]
    return out


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def main():
    sc = import_reference_controller()
    import scipy

    golden = {"generator": "tests/golden/make_golden.py", "scipy_version": scipy.__version__,
              "reference_call": "SolverController.run() -> linprog(method='highs-ds', presolve=True, time_limit=10)",
              "kat": {}, "dense": {}, "batched": {}}

    unique_x = {"K1_wyndor", "K2_max3", "K3_min_ge", "K4_min_ge2", "K9_three_var", "K10_doc"}
    for name, wrapper in W.known_answer_problems().items():
        g = run_reference(sc, wrapper)
        g["x_unique"] = name in unique_x
        g["problem"] = wrapper
        golden["kat"][name] = g
        print(name, g["status_text"], g["scipy_status"], g["z"], g["x"])

    # BASELINE config 2 family: dense random feasible LPs (generator workloads.dense_feasible_lp)
    for n in (16, 64, 128, 256, 512, 1024):
        A, b, c, ops, mx = W.dense_feasible_lp(n, seed=0)
        g = run_reference(sc, W.lp_to_problem_dict(A, b, c, ops, mx))
        g.pop("x")
        g["inputs_sha"] = sha(A, b, c)
        golden["dense"][str(n)] = g
        print("dense", n, g["status_text"], g["z"], g["nit"])

    # BASELINE config 3 family: first 300 LPs of the batched generator (statuses 0/2/3 all occur)
    m, n, count = 20, 30, 300
    A, b, c, ops = W.batched_small_lps(0, count, m, n)
    rows = []
    for k in range(count):
        g = run_reference(sc, W.lp_to_problem_dict(A[k], b[k], c[k], ops[k], False))
        rows.append({"status_text": g["status_text"], "scipy_status": g["scipy_status"], "z": g["z"]})
    golden["batched"] = {"m": m, "n": n, "count": count, "base_seed": 3, "inputs_sha": sha(A, b, c, ops),
                         "results": rows}
    from collections import Counter
    print("batched", Counter(r["status_text"] for r in rows), Counter(r["scipy_status"] for r in rows))

    # a second family with ragged shapes and mixed operators (small, includes degenerate cases)
    mixed = []
    rng = np.random.default_rng(77)
    for k in range(120):
        mm = int(rng.integers(1, 9))
        nn = int(rng.integers(1, 9))
        Am = np.round(rng.uniform(-3, 3, (mm, nn)), 1)
        x0 = np.round(rng.uniform(0, 2, nn), 1)
        opm = rng.integers(0, 3, mm).astype(np.int8)
        slack = np.round(rng.uniform(0, 2, mm), 1)
        bm = Am @ x0 + np.where(opm == 0, slack, np.where(opm == 1, -slack, 0.0))
        if k % 7 == 0:
            bm = np.round(rng.uniform(-3, 3, mm), 1)  # may be infeasible
        cm = np.round(rng.uniform(-1, 2, nn), 1)      # may be unbounded
        mx = bool(k % 2)
        g = run_reference(sc, W.lp_to_problem_dict(Am, bm, cm, opm, mx))
        mixed.append({"A": Am.tolist(), "b": bm.tolist(), "c": cm.tolist(), "ops": opm.tolist(), "maximize": mx,
                      "status_text": g["status_text"], "scipy_status": g["scipy_status"], "z": g["z"]})
    golden["mixed"] = mixed
    print("mixed", Counter(r["status_text"] for r in mixed), Counter(r["scipy_status"] for r in mixed))

    # a ragged stress family regenerated from (seed, k): degenerate, sparse and badly scaled LPs
    fuzz = []
    for k in range(1000):
        Af, bf, cf, of = W.fuzz_lp(k)
        g = run_reference(sc, W.lp_to_problem_dict(Af, bf, cf, of, False))
        fuzz.append([g["scipy_status"], g["z"]])
    golden["fuzz"] = {"seed": 12345, "count": 1000, "inputs_sha": sha(*W.fuzz_lp(7)), "results": fuzz}
    print("fuzz", Counter(r[0] for r in fuzz))

    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote reference_golden.json")


if __name__ == "__main__":
    main()
