"""Point the UNMODIFIED reference controller at the GPU path.

The reference has no plugin interface; its seam is a set of module-level names in
`app.controllers.solver_controller` that its own tests replace by attribute patching
(/root/reference/tests/test_solver_controller.py:123 patches `...solver_controller.linprog`).  `install()` does
exactly that for the five names on the hot path; `uninstall()` restores them.  See INTEGRATION.md.
"""
from __future__ import annotations

from . import simple_simplex as _ss
from .linprog import linprog as _linprog

_NAMES = {
    "linprog": _linprog,                                   # solver_controller.py:9, called at :78-85
    "create_tableau": _ss.create_tableau,                  # :22-27, called at :297
    "add_constraint": _ss.add_constraint,                  # called at :309
    "add_objective": _ss.add_objective,                    # called at :316
    "optimize_json_format": _ss.optimize_json_format,      # called at :318
}
_saved = {}


def install(module=None):
    """Patch the reference module (default: import app.controllers.solver_controller from sys.path)."""
    if module is None:
        import importlib
        module = importlib.import_module("app.controllers.solver_controller")
    for name, fn in _NAMES.items():
        _saved.setdefault((module, name), getattr(module, name, None))
        setattr(module, name, fn)
    return module


def uninstall(module=None):
    for (mod, name), old in list(_saved.items()):
        if module is None or mod is module:
            if old is None:
                if hasattr(mod, name):
                    delattr(mod, name)
            else:
                setattr(mod, name, old)
            del _saved[(mod, name)]
