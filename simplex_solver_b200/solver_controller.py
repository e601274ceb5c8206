"""Host-side mirror of the reference's solve orchestrator, driving the GPU path.

Same constructor, same `run()` and the same report as
/root/reference/app/controllers/solver_controller.py (`SolverController.__init__` :33-50, `run` :53-120,
`_prepare_model_for_scipy` :122-170, `_extract_tableaus_from_simple_simplex` :322-363, status mapping
:382-414, report keys :417-422), so that its consumers -- templates/solution.html, PdfReportService
(pdf_report_service.py:94-177) and StorageService.save_solution (storage_service.py:106-112) -- are unchanged:

    {"problema_definicion": {...},
     "solucion_encontrada": {"status", "mensaje_solver", "valores_variables", "valor_optimo_z"},
     "visualizacion_gilp_html": str,
     "tablas_intermedias": [{"iteration", "title", "table", "pivot"}, ...]}

What is different underneath: `linprog` is simplex_solver_b200.linprog (GPU two-phase tableau simplex) and the
tableau iterations come from simplex_solver_b200.simple_simplex (same GPU loop with snapshots).  gilp's Plotly
figure (the reference's "Plan A", :208-249) is out of scope; the HTML is always the static "Plan B" tables.

`BulkSolverController` is the non-session entry for problems that the cookie session and the
dict-of-dicts format cannot carry: arrays in, the same `solucion_encontrada` block out.
"""
from __future__ import annotations

import json
import os
import traceback
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import simple_simplex as _ss
from .linprog import linprog

STATUS_OK = "Solucion Factible"
STATUS_INFEASIBLE = "Sin Solucion Factible"
STATUS_ERROR = "Error"

_PIVOT_STYLE = 'style="background-color:#fff0f0; color:#d00; font-weight:bold;"'
_TABLE_OPEN = ('<table class="table table-bordered table-striped" style="border:1px solid #ccc; '
               'justify-content:center; float:none; margin-left:auto; margin-right:auto;">')


def status_text(result) -> str:
    """scipy status integer -> the reference's status strings (solver_controller.py:382, :404)."""
    if result.success:
        return STATUS_OK
    return STATUS_INFEASIBLE if result.status == 2 else STATUS_ERROR


def _json_default(o):
    if isinstance(o, np.generic):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    raise TypeError(f"not JSON serialisable: {type(o)}")


class SolverController:
    """Drop-in for the reference class of the same name.  `output_dir`: where `solucion_N.json` is written; the
    reference always saves into its configured OUTPUT_DIR (storage_service.py:106-112) -- pass that directory for the
    same behaviour; the default None (library use, tests, benchmarks) writes no file."""

    def __init__(self, problem_data_wrapper: dict, output_dir: Optional[str] = None, rule: str = "dantzig",
                 device: int = 0, verbose: bool = False):
        definition = problem_data_wrapper.get("problema_definicion", {})
        self.objective_data = definition.get("funcion_objetivo")
        self.constraints_data = definition.get("restricciones")
        # lexicographic order, as the reference: 'x10' sorts before 'x2' (solver_controller.py:46)
        self.variables = sorted(self.objective_data["coefficients"].keys()) if self.objective_data else []
        self.output_dir = output_dir
        self.rule = rule
        self.device = device
        self.verbose = verbose

    # ---- model translation (solver_controller.py:122-170) ------------------------------------------
    def _prepare_model_for_scipy(self, objective_data: dict, constraints_data: list, variables: list):
        sign = -1 if objective_data["type"] == "maximize" else 1
        c = [sign * objective_data["coefficients"].get(v, 0) for v in variables]
        ub_rows, ub_rhs, eq_rows, eq_rhs = [], [], [], []
        for con in constraints_data:
            row = [con["coefficients"].get(v, 0) for v in variables]
            neg = [-a for a in row]
            op, rhs = con["operator"], con["rhs"]
            if op == "<=":
                ub_rows.append(row)
                ub_rhs.append(rhs)
            elif op == ">=":
                ub_rows.append(neg)
                ub_rhs.append(-rhs)
            elif op == "=":
                eq_rows.append(row)
                eq_rhs.append(rhs)
                ub_rows += [row, neg]
                ub_rhs += [rhs, -rhs]
        as_np = lambda v: np.array(v) if v else None  # noqa: E731
        return np.array(c), as_np(ub_rows), as_np(ub_rhs), as_np(eq_rows), as_np(eq_rhs), [(0, None)] * len(variables)

    # ---- tableau iterations (solver_controller.py:290-363) -------------------------------------------
    def _run_simple_simplex(self) -> dict:
        tab = _ss.create_tableau(number_of_variables=len(self.variables),
                                 number_of_constraints=len(self.constraints_data))
        for con in self.constraints_data:
            coeffs = ",".join(str(con["coefficients"].get(v, 0)) for v in self.variables)
            code = {"<=": "L", ">=": "G"}.get(con["operator"], "E")
            _ss.add_constraint(tab, f"{coeffs},{code},{con['rhs']}")
        is_max = self.objective_data["type"] == "maximize"
        coeffs = ",".join(str(self.objective_data["coefficients"].get(v, 0)) for v in self.variables)
        _ss.add_objective(tab, f"{coeffs},{'1' if is_max else '0'}")
        return _ss.optimize_json_format(tab, maximize=is_max, rule=self.rule, device=self.device)

    @staticmethod
    def _extract_tableaus_from_simple_simplex(simplex_json: dict) -> List[Dict[str, Any]]:
        out = []
        for step in simplex_json.get("pivotSteps", []):
            data = step.get("tableau", [])
            if not data:
                continue
            num = step.get("step", "?")
            prow, pcol = step.get("pivotRowIndex"), step.get("pivotColIndex")
            if num == 0 or prow is None:
                title = "Iteración 0 (Tabla Inicial)"
            else:
                title = f"Iteración {num} (Pivote: Fila {prow}, Col {pcol})"
            table = [["Base"] + [f"C{i}" for i in range(len(data[0]))]]
            for i, row in enumerate(data):
                table.append([f"F{i}"] + [round(v, 4) if isinstance(v, (int, float)) else v for v in row])
            out.append({"iteration": num, "title": title, "table": table,
                        "pivot": (prow, pcol) if prow is not None and pcol is not None else None})
        return out

    @staticmethod
    def _tableau_to_html(table: list, pivot_r, pivot_c) -> str:
        pr = -1 if pivot_r is None else pivot_r
        pc = -1 if pivot_c is None else pivot_c
        parts = [_TABLE_OPEN]
        for ri, row in enumerate(table):
            parts.append("<tr>")
            for ci, cell in enumerate(row):
                tag = "th" if (ci == 0 or ri == 0) else "td"
                style = _PIVOT_STYLE if (ri == pr + 1 and ci == pc + 1) else ""
                text = f"{cell:.4f}" if isinstance(cell, float) else cell
                parts.append(f"<{tag} {style}>{text}</{tag}>")
            parts.append("</tr>")
        parts.append("</table>")
        return "".join(parts)

    def _generate_visualization_html_and_tables(self) -> Tuple[str, List[Dict[str, Any]]]:
        try:
            tables = self._extract_tableaus_from_simple_simplex(self._run_simple_simplex())
            html = []
            for t in tables:
                pr, pc = t["pivot"] if t.get("pivot") else (None, None)
                html.append(f"<h4>{t.get('title', 'Tabla')}</h4>")
                html.append(self._tableau_to_html(t.get("table", []), pr, pc))
            return "<br>".join(html), tables
        except Exception as e:  # the reference swallows Plan-B failures into the HTML (:203-205)
            return f"<p>Error en Plan B: {e}</p>", []

    # ---- report (solver_controller.py:366-432) -----------------------------------------------------------
    def _display_and_save_results(self, result, objective_type: str, html: str, tables: list):
        if result.success:
            values = {name: float(result.x[i]) for i, name in enumerate(self.variables)}
            z = float(-result.fun if objective_type == "maximize" else result.fun)
            found = {"status": STATUS_OK, "mensaje_solver": result.message, "valores_variables": values,
                     "valor_optimo_z": z}
            if self.verbose:
                for name, v in values.items():
                    print(f"   {name} = {v:.4f}")
                print(f"   Z = {z:.4f}")
        else:
            found = {"status": status_text(result), "mensaje_solver": result.message,
                     "valores_variables": None, "valor_optimo_z": None}
        report = {
            "problema_definicion": {"funcion_objetivo": self.objective_data, "restricciones": self.constraints_data},
            "solucion_encontrada": found,
            "visualizacion_gilp_html": html,
            "tablas_intermedias": tables,
        }
        if self.output_dir:
            try:
                save_solution(report, self.output_dir)
            except Exception as e:
                if self.verbose:
                    print(f"Advertencia: No se pudo guardar el reporte de solución: {e}")
        return report

    def run(self):
        """Same contract as the reference: the report dict, or None on missing data / any exception (:65-67, :112-120)."""
        if self.objective_data is None or self.constraints_data is None:
            return None
        try:
            c, A_ub, b_ub, A_eq, b_eq, bounds = self._prepare_model_for_scipy(
                self.objective_data, self.constraints_data, self.variables)
            result = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method="highs-ds",
                             options={"presolve": True, "time_limit": 10, "rule": self.rule}, device=self.device)
            if result.success:
                html, tables = self._generate_visualization_html_and_tables()
            else:
                html = "<p>Visualización no disponible (Problema infactible o no acotado).</p>"
                tables = []
            return self._display_and_save_results(result, self.objective_data["type"], html, tables)
        except Exception:
            if self.verbose:
                traceback.print_exc()
            return None


def save_solution(report: dict, output_dir: str, prefix: str = "solucion_") -> str:
    """`solucion_N.json` with N the FIRST free number counting from 1 -- gaps left by deleted files are filled, as
    StorageService._get_next_filename does (storage_service.py:35-43; save_solution :106-112)."""
    os.makedirs(output_dir, exist_ok=True)
    k = 1
    while os.path.exists(os.path.join(output_dir, f"{prefix}{k}.json")):
        k += 1
    path = os.path.join(output_dir, f"{prefix}{k}.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(report, f, indent=4, ensure_ascii=False, default=_json_default)
    return path


class BulkSolverController:
    """Array entry for LPs that the session/dict format cannot carry.

    `A`, `b`, `c` (objective as the user states it), `ops` (0 '<=', 1 '>=', 2 '=').  `run()` returns the same
    `solucion_encontrada` block as SolverController, with `valores_variables` as an ndarray.
    """

    def __init__(self, A, b, c, ops, maximize: bool, rule: str = "dantzig", device: int = 0):
        self.A, self.b, self.c, self.ops = A, b, np.asarray(c, dtype=np.float64), ops
        self.maximize, self.rule, self.device = maximize, rule, device

    def run(self):
        from . import native
        cmin = -self.c if self.maximize else self.c
        rule = native.RULE_BLAND if self.rule == "bland" else native.RULE_DANTZIG
        r = native.thread_solver(self.device).solve_dense(self.A, self.b, cmin, self.ops, native.make_opts(rule=rule))
        ok = r["status"] == native.STATUS_OPTIMAL
        text = STATUS_OK if ok else (STATUS_INFEASIBLE if r["status"] == 2 else STATUS_ERROR)
        return {"status": text, "mensaje_solver": f"B200 tableau simplex, {r['n_pivots']} pivots",
                "valores_variables": r["x"] if ok else None,
                "valor_optimo_z": (float(-r["fun"]) if self.maximize else float(r["fun"])) if ok else None,
                "n_pivots": r["n_pivots"], "device_ms": r["device_ms"]}
