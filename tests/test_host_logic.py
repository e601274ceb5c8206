"""Host-side mirror of the reference interface: argument translation, report format, sharding arithmetic.
No GPU needed (the compute calls are not made here)."""
import numpy as np
import pytest

from simplex_solver_b200 import native, workloads as W
from simplex_solver_b200.batched import shard_range
from simplex_solver_b200.linprog import OptimizeResult, drop_rows_implied_by_equalities, rows_from_linprog_args
from simplex_solver_b200.sharded import ShardedTableau
from simplex_solver_b200.simple_simplex import add_constraint, add_objective, create_tableau, expand_full
from simplex_solver_b200.solver_controller import SolverController, save_solution, status_text

# the reference's own fixtures (tests/test_solver_controller.py:16-40)
OBJ_MIN = {"type": "minimize", "coefficients": {"x1": 50.0, "x2": 80.0}}
CONS_MIN = [
    {"coefficients": {"x1": 4.0, "x2": 1.0}, "operator": ">=", "rhs": 4.0},
    {"coefficients": {"x1": 1.0, "x2": 6.0}, "operator": ">=", "rhs": 6.0},
    {"coefficients": {"x1": 4.0, "x2": 6.0}, "operator": ">=", "rhs": 12.0},
]
OBJ_MAX = {"type": "maximize", "coefficients": {"x1": 15.0, "x2": 18.0}}
CONS_MAX = [
    {"coefficients": {"x1": 4.0, "x2": 2.0}, "operator": "<=", "rhs": 2000.0},
    {"coefficients": {"x1": 2.0, "x2": 6.0}, "operator": "<=", "rhs": 2400.0},
    {"coefficients": {"x1": 20.0, "x2": 28.0}, "operator": "<=", "rhs": 14000.0},
]


def _ctl(obj, cons):
    return SolverController({"problema_definicion": {"funcion_objetivo": obj, "restricciones": cons}})


def test_prepare_model_maximize_matches_reference_expectation():
    # expectations of /root/reference/tests/test_solver_controller.py:66-76
    ctl = _ctl(OBJ_MAX, CONS_MAX)
    c, A_ub, b_ub, A_eq, b_eq, bounds = ctl._prepare_model_for_scipy(OBJ_MAX, CONS_MAX, ctl.variables)
    np.testing.assert_array_equal(c, np.array([-15.0, -18.0]))
    np.testing.assert_array_equal(A_ub, np.array([[4.0, 2.0], [2.0, 6.0], [20.0, 28.0]]))
    np.testing.assert_array_equal(b_ub, np.array([2000.0, 2400.0, 14000.0]))
    assert A_eq is None and b_eq is None
    assert bounds == [(0, None), (0, None)]


def test_prepare_model_minimize_matches_reference_expectation():
    # expectations of /root/reference/tests/test_solver_controller.py:95-103
    ctl = _ctl(OBJ_MIN, CONS_MIN)
    c, A_ub, b_ub, A_eq, b_eq, _ = ctl._prepare_model_for_scipy(OBJ_MIN, CONS_MIN, ctl.variables)
    np.testing.assert_array_equal(c, np.array([50.0, 80.0]))
    np.testing.assert_array_equal(A_ub, np.array([[-4.0, -1.0], [-1.0, -6.0], [-4.0, -6.0]]))
    np.testing.assert_array_equal(b_ub, np.array([-4.0, -6.0, -12.0]))


def test_prepare_model_equality_arrives_three_times_and_is_deduplicated():
    obj = {"type": "maximize", "coefficients": {"x1": 1.0, "x2": 1.0}}
    cons = [{"coefficients": {"x1": 1.0, "x2": 1.0}, "operator": "=", "rhs": 10.0},
            {"coefficients": {"x1": 2.0, "x2": 1.0}, "operator": "<=", "rhs": 15.0}]
    ctl = _ctl(obj, cons)
    c, A_ub, b_ub, A_eq, b_eq, _ = ctl._prepare_model_for_scipy(obj, cons, ctl.variables)
    assert A_ub.shape == (3, 2) and A_eq.shape == (1, 2)       # solver_controller.py:154-161
    cc, A, b, ops = rows_from_linprog_args(c, A_ub, b_ub, A_eq, b_eq)
    assert A.shape == (2, 2)
    np.testing.assert_array_equal(ops, [native.OP_LE, native.OP_EQ])
    np.testing.assert_array_equal(A, [[2.0, 1.0], [1.0, 1.0]])
    np.testing.assert_array_equal(b, [15.0, 10.0])


def test_dedup_keeps_unrelated_rows_and_handles_signed_zero():
    A_eq = np.array([[1.0, 0.0]])
    b_eq = np.array([0.0])
    A_ub = np.array([[1.0, 0.0], [-1.0, -0.0], [1.0, 1.0]])
    b_ub = np.array([0.0, -0.0, 3.0])
    A2, b2 = drop_rows_implied_by_equalities(A_ub, b_ub, A_eq, b_eq)
    np.testing.assert_array_equal(A2, [[1.0, 1.0]])
    np.testing.assert_array_equal(b2, [3.0])


def test_missing_coefficients_read_as_zero_and_variable_order_is_lexicographic():
    obj = {"type": "minimize", "coefficients": {"x1": 1.0, "x10": 3.0, "x2": 2.0}}
    cons = [{"coefficients": {"x2": 5.0}, "operator": "<=", "rhs": 1.0}]
    ctl = _ctl(obj, cons)
    assert ctl.variables == ["x1", "x10", "x2"]                # solver_controller.py:46
    c, A_ub, *_ = ctl._prepare_model_for_scipy(obj, cons, ctl.variables)
    np.testing.assert_array_equal(c, [1.0, 3.0, 2.0])
    np.testing.assert_array_equal(A_ub, [[0.0, 0.0, 5.0]])


def test_empty_wrapper_returns_none_like_the_reference():
    # /root/reference/tests/test_solver_controller.py:225-245
    assert SolverController({}).run() is None
    assert SolverController({"problema_definicion": {}}).run() is None


def test_status_strings():
    assert status_text(OptimizeResult(success=True, status=0)) == "Solucion Factible"
    assert status_text(OptimizeResult(success=False, status=2)) == "Sin Solucion Factible"
    for st in (1, 3, 4):
        assert status_text(OptimizeResult(success=False, status=st)) == "Error"   # solver_controller.py:404


def test_extract_tableaus_format():
    js = {"pivotSteps": [
        {"step": 0, "pivotRowIndex": None, "pivotColIndex": None, "tableau": [[1.0, 2.0], [3.0, 4.123456]]},
        {"step": 1, "pivotRowIndex": 1, "pivotColIndex": 0, "tableau": [[1.0, 2.0], [1.0, 0.33333333]]},
    ]}
    out = SolverController._extract_tableaus_from_simple_simplex(js)
    assert out[0]["title"] == "Iteración 0 (Tabla Inicial)" and out[0]["pivot"] is None
    assert out[1]["title"] == "Iteración 1 (Pivote: Fila 1, Col 0)" and out[1]["pivot"] == (1, 0)
    assert out[0]["table"][0] == ["Base", "C0", "C1"]
    assert out[0]["table"][2] == ["F1", 3.0, 4.1235]            # rounded to 4 dp (solver_controller.py:354)
    assert SolverController._extract_tableaus_from_simple_simplex({}) == []
    html = SolverController._tableau_to_html(out[1]["table"], 1, 0)
    assert html.count("<tr>") == 3 and "0.3333" in html and "background-color:#fff0f0" in html


def test_simple_simplex_string_api():
    t = create_tableau(number_of_variables=2, number_of_constraints=2)
    add_constraint(t, "1.0,0.0,L,4.0")
    add_constraint(t, "3.0,2.0,G,18.0")
    add_objective(t, "3.0,5.0,1")
    assert t["rows"] == [([1.0, 0.0], native.OP_LE, 4.0), ([3.0, 2.0], native.OP_GE, 18.0)]
    assert t["objective"] == [3.0, 5.0] and t["maximize_flag"] is True
    with pytest.raises(ValueError):
        add_constraint(t, "1.0,L,4.0")
    with pytest.raises(ValueError):
        add_constraint(create_tableau(2, 1), "1.0,2.0,X,4.0")


def test_expand_full_scatters_unit_columns():
    # condensed 2 constraint rows + objective; variables: x0, x1 structural, s2, s3 slacks
    Tc = np.array([[2.0, 5.0, 10.0], [3.0, 7.0, 20.0], [-1.0, -4.0, 0.0]])
    rowlab = np.array([2, 0, -1])        # row 0: slack 2 basic, row 1: x0 basic
    collab = np.array([3, 1, -1])        # columns: slack 3, x1, RHS
    full = expand_full(Tc, rowlab, collab, [0, 1, 2, 3], 2)
    np.testing.assert_array_equal(full, [[0, 5, 1, 2, 10], [1, 7, 0, 3, 20], [0, -4, 0, -1, 0]])


def test_save_solution_is_sequential_and_serialises_numpy(tmp_path):
    rep = {"solucion_encontrada": {"valor_optimo_z": np.float64(36.0), "valores_variables": {"x1": np.float64(2.0)}}}
    p1 = save_solution(rep, str(tmp_path))
    p2 = save_solution(rep, str(tmp_path))
    assert p1.endswith("solucion_1.json") and p2.endswith("solucion_2.json")
    import json
    assert json.load(open(p1))["solucion_encontrada"]["valor_optimo_z"] == 36.0


def test_shard_ranges_partition_everything():
    for total in (0, 1, 7, 100000, 131070):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 2, 1000):
                parts = [shard_range(total, world, r, align) for r in range(world)]
                assert parts[0][0] == 0 and parts[-1][1] == total
                for (a, b), (c, d) in zip(parts, parts[1:]):
                    assert b == c and a <= b
                for lo, hi in parts[:-1]:
                    assert lo % align == 0 and hi % align == 0 or hi == total
    assert ShardedTableau.columns_of(131064, 8, 3) == (49148, 65532)
    assert ShardedTableau.columns_of(10, 2, 1) == (4, 10)


def test_workload_generators_shapes_and_special_cases():
    A, b, c, ops = W.batched_small_lps(0, 300)
    assert A.shape == (300, 20, 30) and b.shape == (300, 20) and c.shape == (300, 30) and ops.dtype == np.int8
    assert ops[0, 0] == W.LE and ops[0, 1] == W.GE and b[0, 0] == 5.0 and b[0, 1] == 10.0   # infeasible pattern
    assert (ops[1] == W.GE).all() and (A[1] >= 0).all() and (c[1] < 0).all()                 # unbounded pattern
    A2, b2, c2, ops2 = W.batched_small_lps(1000, 10)
    A3, _, _, _ = W.batched_small_lps(0, 1010)
    np.testing.assert_array_equal(A2, A3[1000:])          # shards regenerate exactly their slice
    wrapper = W.lp_to_problem_dict(A[2], b[2], c[2], ops[2], False)
    Ar, br, cr, opr, mx, names = W.problem_dict_to_arrays(wrapper)
    np.testing.assert_array_equal(Ar, A[2])
    np.testing.assert_array_equal(opr, ops[2])
    assert names == sorted(names) and not mx


def test_linprog_rejects_unsupported_bounds():
    from simplex_solver_b200.linprog import _check_bounds
    _check_bounds([(0, None)] * 3, 3)
    _check_bounds((0, None), 3)
    with pytest.raises(NotImplementedError):
        _check_bounds([(1, None)], 1)
    with pytest.raises(NotImplementedError):
        _check_bounds([(0, 5.0)], 1)


def test_save_solution_fills_gaps_like_the_reference(tmp_path):
    """storage_service.py:35-43 probes 1, 2, ... for the first free name: a deleted solucion_1.json is reused."""
    rep = {"solucion_encontrada": {}}
    paths = [save_solution(rep, str(tmp_path)) for _ in range(3)]
    import os
    os.remove(paths[0])
    assert save_solution(rep, str(tmp_path)).endswith("solucion_1.json")
    assert save_solution(rep, str(tmp_path)).endswith("solucion_4.json")


def test_array_shapes_are_validated_before_the_c_abi():
    """The C ABI takes plain pointers and sizes: a short buffer must be a Python error, never an out-of-bounds read."""
    chk = native._check_lp_arrays
    A, b, c, ops = np.ones((3, 2)), np.ones(3), np.ones(2), np.zeros(3, dtype=np.int8)
    assert chk(A, b, c, ops)[4:] == (3, 2)
    assert chk(np.zeros((0, 2)), np.zeros(0), c, np.zeros(0, dtype=np.int8))[4:] == (0, 2)
    assert chk([], np.zeros(0), c, [])[0].shape == (0, 2)
    for bad in ((np.ones((2, 2)), b, c, ops), (np.ones((3, 3)), b, c, ops), (A, b, c, ops[:2]), (A, b, np.ones((2, 1)), ops),
                (A, b, c, np.array([0, 1, 3], dtype=np.int8)), (np.ones(6), b, c, ops)):
        with pytest.raises(ValueError):
            chk(*bad)
    s = object.__new__(native.Solver)  # no device needed: the checks run before the library is touched
    with pytest.raises(ValueError):
        native.Solver.solve_batched(s, np.ones((4, 3, 2)), np.ones((4, 2)), np.ones((4, 2)), np.zeros((4, 3)), opts=object())
    with pytest.raises(ValueError):
        native.Solver.solve_batched(s, np.ones((3, 2)), np.ones(3), np.ones(2), np.zeros(3), opts=object())


def test_tableau_shards_cover_one_lp_at_every_world_size():
    """bench.py's config 5: ONE LP of cols_total - 1 structural variables at 1/2/4/8 GPUs; every shard but the last stores
    exactly cols_total / world columns (slice + RHS replica), the last one world - 1 more."""
    from simplex_solver_b200.sharded import ShardedTableau
    for cols_total in (131072, 16384, 1024):
        for world in (1, 2, 4, 8):
            spans = [ShardedTableau.columns_of_tableau(cols_total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == cols_total - 1
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo + 1 == cols_total // world for lo, hi in spans[:-1])
            assert spans[-1][1] - spans[-1][0] + 1 == cols_total // world + world - 1


def test_sharded_driver_keys_captured_chunks_by_the_binding_epoch():
    """ShardedTableau replays a captured chunk only while the engine's binding epoch stands: a moved epoch (the library
    reallocated something the chunk bakes in, b200lp_binding_epoch) drops every cached chunk and the next run enqueues
    eagerly and captures again.  Stub engine, no GPU: counts eager chunks, captures and replays."""
    from simplex_solver_b200.sharded import ShardedTableau

    class Graph:
        def __init__(self, eng):
            self.eng = eng

        def replay(self):
            self.eng.replays += 1
            self.eng.n += self.eng.chunk

    class Engine:
        p2p = True  # the fused path: no collective, no torch needed

        def __init__(self):
            self.eager = self.captures = self.replays = 0
            self.n = self.budget = self.chunk = 0
            self._epoch = 7
            self.grow_at = None  # budget from which reset() "reallocates the history"

        def new_buffer(self, world):
            return None

        def reset(self, max_pivots):
            self.n, self.budget = 0, max_pivots
            if self.grow_at is not None and max_pivots >= self.grow_at:
                self._epoch += 1
                self.grow_at = None

        def epoch(self):
            return self._epoch

        def fused(self, opts, lookahead=False):
            self.n += 1

        def capture_chunk(self, body):
            self.captures += 1
            return Graph(self)

        def state(self):
            return self.n >= self.budget, 1, min(self.n, self.budget)

    eng = Engine()
    drv = ShardedTableau(eng, 1, 0)
    opts = native.make_opts(max_pivots=16)
    eng.chunk = 16
    assert drv.run(opts, 16, check_every=16) == (1, 16)
    assert (eng.captures, eng.replays) == (1, 0)           # first run: eager chunk, then the capture
    assert drv.run(opts, 16, check_every=16) == (1, 16)
    assert (eng.captures, eng.replays) == (1, 1)           # same epoch: replay
    eng.grow_at = 128                                      # a bigger budget makes reset() reallocate
    assert drv.run(opts, 128, check_every=16) == (1, 128)
    assert eng.captures == 2, "the chunk must be captured again after the epoch moved"
    replays = eng.replays
    assert drv.run(opts, 128, check_every=16) == (1, 128)
    assert eng.captures == 2 and eng.replays == replays + 8
