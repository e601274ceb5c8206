"""Published known-answer material for the tableau method itself, independent of this repo and of the reference's
(absent) engine: the Wyndor Glass Co. iterations as printed in textbooks (Hillier & Lieberman, "Introduction to Operations
Research", the LP the reference's own tests and load tests use: tests/test_visualization_integration.py:38-48,
tests/test_performance_load.py:31-38) and a two-phase problem with <=, = and >= rows whose optimum is printed there too.
They pin what the golden vectors from the reference's HiGHS call cannot: the displayed tableau of every iteration and the
pivot sequence of the oracle (parity with simple-simplex==0.0.3 itself stays unpinned)."""
import numpy as np

from oracle import oracle as O


def test_wyndor_iterations_match_the_published_tableaux():
    A = np.array([[1.0, 0.0], [0.0, 2.0], [3.0, 2.0]])
    b = np.array([4.0, 12.0, 18.0])
    c = np.array([-3.0, -5.0])  # maximise 3 x1 + 5 x2  ->  minimise -3 x1 - 5 x2 (solver_controller.py:133-134)
    ops = np.zeros(3, dtype=np.int8)
    # columns x1 x2 x3 x4 x5 | rhs; constraint rows first, objective row last
    published = [
        np.array([[1, 0, 1, 0, 0, 4], [0, 2, 0, 1, 0, 12], [3, 2, 0, 0, 1, 18], [-3, -5, 0, 0, 0, 0]], dtype=float),
        np.array([[1, 0, 1, 0, 0, 4], [0, 1, 0, 0.5, 0, 6], [3, 0, 0, -1, 1, 6], [-3, 0, 0, 2.5, 0, 30]], dtype=float),
        np.array([[0, 0, 1, 1 / 3, -1 / 3, 2], [0, 1, 0, 0.5, 0, 6], [1, 0, 0, -1 / 3, 1 / 3, 2], [0, 0, 0, 1.5, 1, 36]]),
    ]
    pivots = [(None, None), (1, 1), (2, 0)]  # x2 enters / row of x4 leaves; x1 enters / row of x5 leaves
    for rule in (O.RULE_DANTZIG,):
        r = O.full_steps(A, b, c, ops, O.make_opts(rule=rule))
        assert r["status"] == O.OPT and r["n_pivots"] == 2
        assert list(r["var_ids"]) == [0, 1, 2, 3, 4]
        assert list(r["basis"]) == [2, 1, 0]  # x3, x2, x1 basic at the optimum
        assert len(r["steps"]) == 3
        for (T, pr, pc), want, (wr, wc) in zip(r["steps"], published, pivots):
            assert (pr, pc) == (wr, wc)
            np.testing.assert_allclose(T, want, rtol=0, atol=1e-14)
    # the condensed tableau takes the same pivots (entering variable ids 1 then 0; rows 1 then 2) and ends at z* = 36
    s = O.solve_lp(A, b, c, ops, O.make_opts(rule=O.RULE_DANTZIG), hist_cap=8)
    assert s["status"] == O.OPT and list(s["piv_row"]) == [1, 2] and list(s["enter_lab"]) == [1, 0]
    assert s["fun"] == -36.0 and list(s["x"]) == [2.0, 6.0]
    # Bland's rule enters x1 first (lowest index) and needs three pivots for the same optimum
    sb = O.solve_lp(A, b, c, ops, O.make_opts(rule=O.RULE_BLAND), hist_cap=8)
    assert sb["status"] == O.OPT and list(sb["enter_lab"])[0] == 0 and sb["fun"] == -36.0 and list(sb["x"]) == [2.0, 6.0]


def test_two_phase_problem_with_all_three_operators_reaches_the_published_optimum():
    """min 0.4 x1 + 0.5 x2  s.t.  0.3 x1 + 0.1 x2 <= 2.7,  0.5 x1 + 0.5 x2 = 6,  0.6 x1 + 0.4 x2 >= 6  (the radiation
    therapy example of the same textbook): x* = (7.5, 4.5), z* = 5.25; phase 1 has to drive two artificials out."""
    A = np.array([[0.3, 0.1], [0.5, 0.5], [0.6, 0.4]])
    b = np.array([2.7, 6.0, 6.0])
    c = np.array([0.4, 0.5])
    ops = np.array([O.LE, O.EQ, O.GE], dtype=np.int8)
    for rule in (O.RULE_DANTZIG, O.RULE_BLAND):
        s = O.solve_lp(A, b, c, ops, O.make_opts(rule=rule), hist_cap=16)
        assert s["status"] == O.OPT and s["n_phase1"] >= 2
        np.testing.assert_allclose(s["x"], [7.5, 4.5], rtol=1e-12)
        assert abs(s["fun"] - 5.25) < 1e-12
        f = O.full_steps(A, b, c, ops, O.make_opts(rule=rule))
        assert f["status"] == O.OPT and f["n_pivots"] == s["n_pivots"]
        # the displayed tableau carries structural, slack / surplus and artificial columns plus the right-hand side
        T_last = f["steps"][-1][0]
        assert T_last.shape[1] == 2 + 2 + 2 + 1
