import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from simplex_solver_b200 import native, workloads as W
from simplex_solver_b200.linprog import linprog
from simplex_solver_b200.solver_controller import SolverController
wy = W.known_answer_problems()["K1_wyndor"]
A, b, c, ops, mx, _ = W.problem_dict_to_arrays(wy)
def timeit(f, n=200):
    f(); f()
    t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e6
print("linprog(Wyndor) us:", round(timeit(lambda: linprog(-c, A_ub=A, b_ub=b, bounds=[(0, None)] * 2)), 1))
s = native.thread_solver(0)
print("solve_dense(Wyndor) us:", round(timeit(lambda: s.solve_dense(A, b, -c, ops)), 1))
print("solve_batched(B=1) us:", round(timeit(lambda: s.solve_batched(A[None], b[None], (-c)[None], ops[None])), 1))
print("SolverController.run(Wyndor) us:", round(timeit(lambda: SolverController(wy).run(), 50), 1))
k3 = W.known_answer_problems()["K3_min_ge"]
A3, b3, c3, o3, mx3, _ = W.problem_dict_to_arrays(k3)
print("solve_dense(K3 two-phase) us:", round(timeit(lambda: s.solve_dense(A3, b3, c3, o3)), 1))
try:
    from scipy.optimize import linprog as sp
    print("scipy highs-ds (Wyndor) us:", round(timeit(lambda: sp(-c, A_ub=A, b_ub=b, bounds=[(0, None)] * 2, method="highs-ds", options={"presolve": True, "time_limit": 10}), 50), 1))
except Exception as e:
    print("scipy unavailable", e)
