"""Shared helpers of the parity tests."""
import numpy as np


def to_min_form(c_user, maximize):
    """Costs of the minimisation form handed to the solver (solver_controller.py:133-134)."""
    c = np.asarray(c_user, dtype=np.float64)
    return -c if maximize else c


def z_from_fun(fun, maximize):
    """solver_controller.py:393."""
    return -fun if maximize else fun


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    # +0.0 and -0.0 compare equal here on purpose: the sign of a zero never reaches a decision or a result
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        idx = np.argwhere(~same)[:5]
        raise AssertionError(f"{what}: {np.count_nonzero(~same)} entries differ, first at {idx.tolist()}: "
                             f"{[(a[tuple(i)], b[tuple(i)]) for i in idx]}")


def lp_strings(A, b, c_user, ops, maximize):
    """The strings solver_controller.py:302-315 hands to simple_simplex for this LP (str(float) formatting)."""
    cons = [",".join(str(float(v)) for v in A[i]) + f",{'LGE'[int(ops[i])]},{float(b[i])}" for i in range(len(b))]
    return cons, ",".join(str(float(v)) for v in c_user) + f",{'1' if maximize else '0'}"


def run_pivot_steps(A, b, c_user, ops, maximize, rule="dantzig"):
    """The product's pivotSteps producer driven exactly like the reference drives simple_simplex (:297-318)."""
    from simplex_solver_b200 import simple_simplex as ss
    t = ss.create_tableau(number_of_variables=len(c_user), number_of_constraints=len(b))
    cons, obj = lp_strings(A, b, c_user, ops, maximize)
    for s in cons:
        ss.add_constraint(t, s)
    ss.add_objective(t, obj)
    return ss.optimize_json_format(t, maximize=maximize, rule=rule)


def assert_steps_equal_full_oracle(js, full, what=""):
    """pivotSteps of the product vs the textbook full-tableau oracle (oracle.full_steps): same number of steps, same
    0-based (row, column) of every pivot in the displayed tableau, every displayed cell bit-equal."""
    steps = js["pivotSteps"]
    assert js["status"] == full["status"], f"{what}: status {js['status']} != {full['status']}"
    assert len(steps) == len(full["steps"]), f"{what}: {len(steps)} steps != {len(full['steps'])}"
    assert bool(js["truncated"]) == bool(full["truncated"]), what
    for k, (st, (T, r, c)) in enumerate(zip(steps, full["steps"])):
        assert st["step"] == k
        assert st["pivotRowIndex"] == r and st["pivotColIndex"] == c, f"{what} step {k}: pivot " \
            f"({st['pivotRowIndex']}, {st['pivotColIndex']}) != ({r}, {c})"
        assert_bit_equal(np.array(st["tableau"]), T, f"{what} step {k}")
