import os
import sys

import pytest

# Guard mode for the whole test session (read by libb200lp when it is first used): every workspace buffer lies between
# two bands of a byte pattern and tests assert the bands stay intact (tests/test_gpu_guard_bands.py).
os.environ.setdefault("B200LP_GUARD", "1")

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def solver():
    from simplex_solver_b200 import native
    s = native.Solver(0)
    yield s
    corrupted = s.check_guards() if os.environ.get("B200LP_GUARD", "0") not in ("", "0") else 0
    s.close()
    assert corrupted == 0, f"{corrupted} guard-band bytes around the shared solver's buffers were overwritten"
