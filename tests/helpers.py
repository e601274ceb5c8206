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
