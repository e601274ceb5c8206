/*
 * simplex_oracle.c -- CPU restatement of the tableau simplex pivot loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file's shared object.  The product path (simplex_solver_b200/) never links, imports or calls it.
 *
 * What it restates.  The reference's /solve path (app/controllers/solver_controller.py:53-120) holds no
 * simplex arithmetic of its own; it delegates to two un-vendored third-party packages:
 *   - scipy==1.12.0  optimize.linprog(method='highs-ds')  -> status, x*, z*   (solver_controller.py:78-85)
 *   - simple-simplex==0.0.3  create_tableau/add_constraint/add_objective/optimize_json_format
 *       -> tableau iterations ("pivotSteps")                                  (solver_controller.py:290-319)
 * simple-simplex's source is not under /root/reference and not installable offline, so the pivot loop below
 * restates the *published* algorithm (primal tableau simplex, textbook two-phase method, Dantzig and Bland
 * entering rules, minimum-ratio leaving rule) under the I/O contract the reference's call sites fix:
 * row operators L/G/E (solver_controller.py:305-306), 0-based pivot (row, col) indices and a step-0 initial
 * tableau (solver_controller.py:332-362), scipy status integers 0/1/2/3 (solver_controller.py:382-414).
 *
 * Pinning.  Status / z* / x* are pinned against golden vectors produced in the build container by the
 * reference's own SolverController.run() with real scipy HiGHS (tests/golden/make_golden.py ->
 * tests/golden/reference_golden.json; checked by tests/test_oracle_golden.py).  The PIVOT SEQUENCE and tableau entries are
 * "parity unpinned": no reference test or fixture holds a single tableau cell or pivot index
 * (SURVEY.md F4), so for those this file is the definition the CUDA path is compared with bit-for-bit.
 * What anchors the iterations outside this repo: the Wyndor LP of the reference's tests reproduces the tableaux printed
 * in textbooks cell for cell, and a two-phase LP with <=, = and >= rows its printed optimum (tests/test_oracle_textbook.py).
 *
 * Arithmetic contract shared with the CUDA kernels (DESIGN.md "Arithmetic contract"):
 *   p = T[r][s]; inv_p = 1/p; col_i = T[i][s];
 *   q_j = T[r][j] / p                          (IEEE division, j != s)
 *   T[i][j] = fma(-col_i, q_j, T[i][j])        (one fused multiply-add, i != r, j != s)
 *   T[i][s] = fma(-col_i, inv_p, 0.0)          (i != r)         T[r][j] = q_j   T[r][s] = inv_p
 * Compile with -ffp-contract=off so that nothing else is fused.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EXPORT __attribute__((visibility("default")))

enum { ORC_RULE_DANTZIG = 0, ORC_RULE_BLAND = 1 };
enum { ORC_OPT = 0, ORC_LIMIT = 1, ORC_INFEASIBLE = 2, ORC_UNBOUNDED = 3, ORC_NUMERICAL = 4 };
enum { ORC_LE = 0, ORC_GE = 1, ORC_EQ = 2 };

/* Condensed (Tucker) tableau.  R = m + n_obj rows, C columns, the last column is the right-hand side.
 * Row m is the objective (z) row, row m+1 (when n_obj == 2) the phase-1 (w) row.
 * rowlab[i] = id of the variable basic in row i (negative: artificial left in a redundant row);
 * collab[j] = id of the non-basic variable that owns column j (-1 for the RHS column).
 * ids: [0,n) structural, n+i logical (slack/surplus) of row i, art_base+i artificial of row i.          */
typedef struct {
    int64_t m, n_obj, R, C, ld;
    int64_t n_struct;
    int32_t art_base;
    double *T;
    int32_t *rowlab;
    int32_t *collab;
} orc_tableau;

typedef struct {
    int32_t rule;
    int64_t max_pivots;
    double eps_cost, eps_pivot, eps_feas;
    int32_t threads; /* OpenMP threads for the rank-1 update; <=1 = serial */
} orc_opts;

typedef struct {
    int32_t status;
    double fun;
    int64_t n_pivots, n_phase1;
    int64_t hist_cap; /* capacity of the four arrays below (may be 0) */
    int32_t *piv_row, *piv_col, *enter_lab, *leave_lab;
} orc_result;

/* ---- counter-based generator shared bit-for-bit with csrc/generate.cuh ------------------------------- */
static inline uint64_t orc_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double orc_u01(uint64_t seed, uint64_t i, uint64_t j) {
    uint64_t h = orc_mix64(orc_mix64(seed ^ (i * 0xD1342543DE82EF95ull)) + j);
    return (double)(h >> 11) * 0x1.0p-53;
}
#define ORC_KEY_RHS 0xFFFFFFFFull

ORC_EXPORT double orc_gen_entry(uint64_t seed, int64_t m, int64_t n_total, int64_t i, int64_t label) {
    /* value of the generated LP "max c'x, Ax <= b" at tableau row i and the column owned by `label`
     * (label < 0: the RHS column).  n_total = number of structural variables of the WHOLE LP.            */
    if (i < m) {
        if (label >= 0) return orc_u01(seed, (uint64_t)i, (uint64_t)label);
        double u = orc_u01(seed, (uint64_t)i, ORC_KEY_RHS);
        double t = 0.9 * u;
        t = t + 0.1;
        return 0.25 * (double)n_total + t;
    }
    if (label >= 0) {
        double u = orc_u01(seed, (uint64_t)m, (uint64_t)label);
        double t = 0.9 * u;
        t = t + 0.1;
        return -t;
    }
    return 0.0;
}

ORC_EXPORT void orc_free(orc_tableau *t) {
    if (!t) return;
    free(t->T);
    free(t->rowlab);
    free(t->collab);
    memset(t, 0, sizeof(*t));
}

static int orc_alloc(orc_tableau *t, int64_t m, int64_t n_obj, int64_t C, int64_t ld) {
    memset(t, 0, sizeof(*t));
    t->m = m;
    t->n_obj = n_obj;
    t->R = m + n_obj;
    t->C = C;
    t->ld = ld < C ? C : ld;
    t->T = (double *)calloc((size_t)(t->R * t->ld), sizeof(double));
    t->rowlab = (int32_t *)calloc((size_t)t->R, sizeof(int32_t));
    t->collab = (int32_t *)calloc((size_t)t->C, sizeof(int32_t));
    if (!t->T || !t->rowlab || !t->collab) {
        orc_free(t);
        return -1;
    }
    return 0;
}

/* Synthetic dense LP of BASELINE configs 4/5: max c'x, Ax <= b, b > 0, as a condensed
 * tableau with ONE objective row; slack basis is feasible so there is no phase 1.  A column shard holds
 * the structural labels [lab0, lab0+ncols) plus its own replica of the RHS as last column.             */
ORC_EXPORT int orc_generate(orc_tableau *t, uint64_t seed, int64_t m, int64_t n_total, int64_t lab0,
                            int64_t ncols, int64_t ld) {
    if (orc_alloc(t, m, 1, ncols + 1, ld)) return -1;
    t->n_struct = n_total;
    t->art_base = (int32_t)(n_total + m);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < t->R; ++i) {
        double *row = t->T + i * t->ld;
        for (int64_t j = 0; j < ncols; ++j) row[j] = orc_gen_entry(seed, m, n_total, i, lab0 + j);
        row[ncols] = orc_gen_entry(seed, m, n_total, i, -1);
        t->rowlab[i] = i < m ? (int32_t)(n_total + i) : -1;
    }
    for (int64_t j = 0; j < ncols; ++j) t->collab[j] = (int32_t)(lab0 + j);
    t->collab[ncols] = -1;
    return 0;
}

/* Initial tableau of "min c'x, A_i x (op_i) b_i, x >= 0".  Rows with b_i < 0 are negated first.         */
ORC_EXPORT int orc_build(orc_tableau *t, const double *A, int64_t lda, const double *b, const double *c,
                         const int8_t *ops, int64_t m, int64_t n) {
    int64_t n_ge = 0, n_art = 0;
    for (int64_t i = 0; i < m; ++i) {
        int op = ops[i];
        if (b[i] < 0.0 && op != ORC_EQ) op = (op == ORC_LE) ? ORC_GE : ORC_LE;
        if (op == ORC_GE) ++n_ge;
        if (op != ORC_LE) ++n_art;
    }
    int64_t n_obj = n_art > 0 ? 2 : 1;
    if (orc_alloc(t, m, n_obj, n + n_ge + 1, 0)) return -1;
    t->n_struct = n;
    t->art_base = (int32_t)(n + m);
    int64_t C = t->C, ld = t->ld, k = 0;
    for (int64_t j = 0; j < n; ++j) t->collab[j] = (int32_t)j;
    t->collab[C - 1] = -1;
    for (int64_t i = 0; i < m; ++i) {
        double *row = t->T + i * ld;
        int op = ops[i];
        int neg = b[i] < 0.0;
        if (neg && op != ORC_EQ) op = (op == ORC_LE) ? ORC_GE : ORC_LE;
        for (int64_t j = 0; j < n; ++j) row[j] = neg ? -A[i * lda + j] : A[i * lda + j];
        row[C - 1] = neg ? -b[i] : b[i];
        if (op == ORC_GE) {
            row[n + k] = -1.0;
            t->collab[n + k] = (int32_t)(n + i);
            ++k;
        }
        t->rowlab[i] = (op == ORC_LE) ? (int32_t)(n + i) : (int32_t)(t->art_base + i);
    }
    double *z = t->T + m * ld;
    for (int64_t j = 0; j < n; ++j) z[j] = c[j];
    t->rowlab[m] = -1;
    if (n_obj == 2) {
        double *w = t->T + (m + 1) * ld;
        t->rowlab[m + 1] = -1;
        for (int64_t j = 0; j < C; ++j) {
            double acc = 0.0;
            for (int64_t i = 0; i < m; ++i)
                if (t->rowlab[i] >= t->art_base) acc = acc + t->T[i * ld + j];
            w[j] = -acc;
        }
    }
    return 0;
}

/* Entering column: Dantzig = most negative reduced cost, ties to the lowest variable id;
 * Bland = lowest variable id with a negative reduced cost.  Artificials never enter.  -1 = optimal.    */
ORC_EXPORT int64_t orc_price(const orc_tableau *t, int64_t obj_row, int32_t rule, double eps_cost) {
    const double *d = t->T + obj_row * t->ld;
    int64_t best = -1;
    for (int64_t j = 0; j < t->C - 1; ++j) {
        int32_t lab = t->collab[j];
        if (lab >= t->art_base) continue;
        if (!(d[j] < -eps_cost)) continue;
        if (best < 0) {
            best = j;
            continue;
        }
        if (rule == ORC_RULE_BLAND) {
            if (lab < t->collab[best]) best = j;
        } else {
            if (d[j] < d[best] || (d[j] == d[best] && lab < t->collab[best])) best = j;
        }
    }
    return best;
}

/* Leaving row over an explicit column copy: min rhs_i / col_i over col_i > eps, ties to the lowest
 * basic-variable id; rows flagged redundant (rowlab < 0) are skipped.  -1 = unbounded direction.        */
ORC_EXPORT int64_t orc_ratio(const orc_tableau *t, const double *col, double eps_pivot) {
    int64_t best = -1;
    double best_ratio = 0.0;
    for (int64_t i = 0; i < t->m; ++i) {
        if (t->rowlab[i] < 0) continue;
        double a = col[i];
        if (!(a > eps_pivot)) continue;
        double ratio = t->T[i * t->ld + t->C - 1] / a;
        if (best < 0 || ratio < best_ratio || (ratio == best_ratio && t->rowlab[i] < t->rowlab[best])) {
            best = i;
            best_ratio = ratio;
        }
    }
    return best;
}

ORC_EXPORT void orc_extract_col(const orc_tableau *t, int64_t s, double *col) {
    for (int64_t i = 0; i < t->R; ++i) col[i] = t->T[i * t->ld + s];
}

/* Fused row-scale + rank-1 update with an explicit pivot column `col` (length R, col[r] = pivot).
 * s_local = position of the entering column in THIS tableau, or -1 when it lives in another shard.     */
ORC_EXPORT void orc_pivot_col(orc_tableau *t, int64_t r, const double *col, int64_t s_local,
                              int32_t enter_lab, int32_t threads) {
    const int64_t R = t->R, C = t->C, ld = t->ld;
    const double p = col[r];
    const double inv_p = 1.0 / p;
    double *q = (double *)malloc((size_t)C * sizeof(double));
    double *rowr = t->T + r * ld;
    for (int64_t j = 0; j < C; ++j) q[j] = rowr[j] / p;
    if (s_local >= 0) q[s_local] = inv_p;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int64_t i = 0; i < R; ++i) {
        if (i == r) continue;
        double *row = t->T + i * ld;
        const double nc = -col[i];
        if (s_local >= 0) row[s_local] = 0.0;
        for (int64_t j = 0; j < C; ++j) row[j] = __builtin_fma(nc, q[j], row[j]);
    }
    for (int64_t j = 0; j < C; ++j) rowr[j] = q[j];
    free(q);
    int32_t leave = t->rowlab[r];
    t->rowlab[r] = enter_lab;
    if (s_local >= 0) t->collab[s_local] = leave;
}

ORC_EXPORT void orc_pivot(orc_tableau *t, int64_t r, int64_t s, int32_t threads) {
    double *col = (double *)malloc((size_t)t->R * sizeof(double));
    orc_extract_col(t, s, col);
    orc_pivot_col(t, r, col, s, t->collab[s], threads);
    free(col);
}

static void orc_log(orc_result *res, int64_t r, int64_t s, int32_t enter, int32_t leave) {
    if (res->n_pivots < res->hist_cap) {
        res->piv_row[res->n_pivots] = (int32_t)r;
        res->piv_col[res->n_pivots] = (int32_t)s;
        res->enter_lab[res->n_pivots] = enter;
        res->leave_lab[res->n_pivots] = leave;
    }
    res->n_pivots++;
}

/* One phase of the loop on objective row obj_row.  Returns ORC_OPT when no entering column is left.    */
static int orc_run_phase(orc_tableau *t, int64_t obj_row, const orc_opts *o, orc_result *res, double *col) {
    for (;;) {
        if (res->n_pivots >= o->max_pivots) return ORC_LIMIT;
        int64_t s = orc_price(t, obj_row, o->rule, o->eps_cost);
        if (s < 0) return ORC_OPT;
        orc_extract_col(t, s, col);
        int64_t r = orc_ratio(t, col, o->eps_pivot);
        if (r < 0) return ORC_UNBOUNDED;
        orc_log(res, r, s, t->collab[s], t->rowlab[r]);
        orc_pivot_col(t, r, col, s, t->collab[s], o->threads);
    }
}

/* After phase 1: pivot artificials that are still basic (at level zero) out of the basis, lowest row
 * first, on the eligible entry of largest magnitude (ties to the lowest variable id); a row with no
 * usable entry is redundant and is flagged (rowlab = -1 - rowlab) so that it never leaves again.       */
static int orc_drive_out(orc_tableau *t, const orc_opts *o, orc_result *res, double *col) {
    for (int64_t i = 0; i < t->m; ++i) {
        if (t->rowlab[i] < t->art_base) continue;
        const double *row = t->T + i * t->ld;
        int64_t best = -1;
        double best_abs = 0.0;
        for (int64_t j = 0; j < t->C - 1; ++j) {
            int32_t lab = t->collab[j];
            if (lab >= t->art_base) continue;
            double a = fabs(row[j]);
            if (!(a > o->eps_pivot)) continue;
            if (best < 0 || a > best_abs || (a == best_abs && lab < t->collab[best])) {
                best = j;
                best_abs = a;
            }
        }
        if (best < 0) {
            t->rowlab[i] = -1 - t->rowlab[i];
            continue;
        }
        if (res->n_pivots >= o->max_pivots) return ORC_LIMIT;
        orc_extract_col(t, best, col);
        orc_log(res, i, best, t->collab[best], t->rowlab[i]);
        orc_pivot_col(t, i, col, best, t->collab[best], o->threads);
    }
    return ORC_OPT;
}

/* Pivot budget.  max_pivots >= 2^40 means "automatic": 200 * (m + C) + 10000 pivots, and -- because Dantzig's
 * rule with lowest-id tie-breaking can cycle on degenerate problems (Beale's example does) -- a phase that
 * exhausts an AUTOMATIC budget under Dantzig continues from the current basis under Bland's rule, which cannot
 * cycle, with one more budget.  An explicit budget is honoured as given (status LIMIT).                       */
#define ORC_AUTO_BUDGET ((int64_t)1 << 40)
static int64_t orc_auto_cap(int64_t m, int64_t C) { return 200 * (m + C) + 10000; }

static int orc_run_phase_fb(orc_tableau *t, int64_t obj_row, orc_opts *o, int is_auto, int64_t cap, orc_result *res,
                            double *col) {
    int st = orc_run_phase(t, obj_row, o, res, col);
    if (st == ORC_LIMIT && is_auto && o->rule == ORC_RULE_DANTZIG) {
        o->rule = ORC_RULE_BLAND;
        o->max_pivots += cap;
        st = orc_run_phase(t, obj_row, o, res, col);
    }
    return st;
}

/* Two-phase driver on a built tableau.  fun = -T[m][C-1] (objective of the minimisation form).         */
ORC_EXPORT int orc_solve(orc_tableau *t, const orc_opts *o_in, orc_result *res) {
    double *col = (double *)malloc((size_t)t->R * sizeof(double));
    orc_opts oo = *o_in;
    orc_opts *o = &oo;
    const int is_auto = o_in->max_pivots >= ORC_AUTO_BUDGET;
    const int64_t cap = is_auto ? orc_auto_cap(t->m, t->C) : o_in->max_pivots;
    o->max_pivots = cap;
    int st = ORC_OPT;
    res->n_pivots = 0;
    res->n_phase1 = 0;
    if (t->n_obj == 2) {
        st = orc_run_phase_fb(t, t->m + 1, o, is_auto, cap, res, col);
        if (st == ORC_UNBOUNDED) st = ORC_NUMERICAL; /* w is bounded below by 0 */
        if (st == ORC_OPT && t->T[(t->m + 1) * t->ld + t->C - 1] < -o->eps_feas) st = ORC_INFEASIBLE;
        if (st == ORC_OPT) st = orc_drive_out(t, o, res, col);
        res->n_phase1 = res->n_pivots;
    }
    if (st == ORC_OPT) st = orc_run_phase_fb(t, t->m, o, is_auto, cap, res, col);
    free(col);
    res->status = st;
    res->fun = -t->T[t->m * t->ld + t->C - 1];
    return st;
}

ORC_EXPORT void orc_read_x(const orc_tableau *t, double *x) {
    for (int64_t j = 0; j < t->n_struct; ++j) x[j] = 0.0;
    for (int64_t i = 0; i < t->m; ++i) {
        int32_t lab = t->rowlab[i];
        if (lab >= 0 && lab < t->n_struct) x[lab] = t->T[i * t->ld + t->C - 1];
    }
}

/* min c'x s.t. rows, x >= 0 from host arrays: build + solve + read x.                                   */
ORC_EXPORT int orc_solve_lp(const double *A, int64_t lda, const double *b, const double *c, const int8_t *ops,
                            int64_t m, int64_t n, const orc_opts *o, orc_result *res, double *x) {
    orc_tableau t;
    if (orc_build(&t, A, lda, b, c, ops, m, n)) return -1;
    orc_solve(&t, o, res);
    if (x) orc_read_x(&t, x);
    orc_free(&t);
    return 0;
}

/* B independent LPs of one shape, packed [B][m][n] / [B][m] / [B][n]; OpenMP over LPs.
 * piv_log (optional) receives up to log_cap (row, col) pairs per LP.                                    */
ORC_EXPORT int orc_solve_batched(int64_t B, int64_t m, int64_t n, const double *A, const double *b,
                                 const double *c, const int8_t *ops, const orc_opts *o, int32_t *status,
                                 double *fun, double *x, int32_t *n_pivots, int32_t *piv_log,
                                 int64_t log_cap, int32_t threads) {
    orc_opts oo = *o;
    oo.threads = 1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 1 ? threads : 1)
#endif
    for (int64_t k = 0; k < B; ++k) {
        orc_result res;
        memset(&res, 0, sizeof(res));
        int32_t *hr = NULL, *hc = NULL, *he = NULL, *hl = NULL;
        if (piv_log && log_cap > 0) {
            hr = (int32_t *)malloc((size_t)log_cap * 4 * sizeof(int32_t));
            hc = hr + log_cap;
            he = hc + log_cap;
            hl = he + log_cap;
            res.hist_cap = log_cap;
            res.piv_row = hr;
            res.piv_col = hc;
            res.enter_lab = he;
            res.leave_lab = hl;
        }
        orc_solve_lp(A + k * m * n, n, b + k * m, c + k * n, ops + k * m, m, n, &oo, &res, x ? x + k * n : NULL);
        status[k] = res.status;
        fun[k] = res.fun;
        n_pivots[k] = (int32_t)res.n_pivots;
        if (hr) {
            for (int64_t e = 0; e < log_cap; ++e) {
                int in = e < res.n_pivots;
                piv_log[(k * log_cap + e) * 2 + 0] = in ? hr[e] : -1;
                piv_log[(k * log_cap + e) * 2 + 1] = in ? hc[e] : -1;
            }
            free(hr);
        }
    }
    return 0;
}

/* ---- textbook FULL-tableau two-phase simplex (pivotSteps checker) ------------------------------------------
 * An independent restatement of the same algorithm on the tableau as a user sees it -- the layout of
 * step["tableau"] consumed at solver_controller.py:332-362 (and printed by pdf_report_service.py:135-177): EVERY
 * variable owns an explicit column, ordered by variable id (structural x1..xn, then the slack / surplus variable of
 * each inequality row, then the artificial of each >= / = row), then the right-hand side; basic columns are carried
 * as unit vectors and updated like any other column.  No label arrays, no condensed storage: "lowest variable id" is
 * simply "lowest column index".  It shares nothing with the condensed code above but the arithmetic contract (one
 * division per pivot-row element, one fma per other element) and the decision rules, so that it can check both the
 * condensed oracle and the GPU's pivotSteps bit for bit:
 *   q_j = T[r][j] / p  gives 1 at the entering column and 1/p at the leaving variable's (unit) column;
 *   fma(-col_i, q_j, T[i][j]) gives exactly 0 at the entering column, leaves other unit columns alone
 *   (q_j = 0) and gives fma(-col_i, 1/p, 0) at the leaving variable's column -- the condensed update rules.
 * snaps receives (cap + 1) tableaux of R x W doubles: step 0 = the initial tableau, step k = after pivot k.
 * out[0..5] = R, W, pivots done, status, pivots in phase 1 (+ drive-out), art_base.                           */
typedef struct {
    int64_t m, R, W, n_real; /* n_real: columns [0, n_real) are not artificial */
    double *T;
    int32_t *basis;          /* column index basic in row i; -1 - index when the row is flagged redundant */
    int32_t *is_basic;       /* per column */
} orc_full;

static int64_t orcf_price(const orc_full *f, int64_t obj_row, int32_t rule, double eps_cost) {
    const double *d = f->T + obj_row * f->W;
    int64_t best = -1;
    for (int64_t j = 0; j < f->n_real; ++j) {
        if (f->is_basic[j]) continue;
        if (!(d[j] < -eps_cost)) continue;
        if (best < 0) best = j;
        else if (rule != ORC_RULE_BLAND && d[j] < d[best]) best = j; /* ties and Bland: the lower index stays */
    }
    return best;
}

static int64_t orcf_ratio(const orc_full *f, int64_t s, double eps_pivot) {
    int64_t best = -1;
    double best_ratio = 0.0;
    for (int64_t i = 0; i < f->m; ++i) {
        if (f->basis[i] < 0) continue;
        double a = f->T[i * f->W + s];
        if (!(a > eps_pivot)) continue;
        double ratio = f->T[i * f->W + f->W - 1] / a;
        if (best < 0 || ratio < best_ratio || (ratio == best_ratio && f->basis[i] < f->basis[best])) {
            best = i;
            best_ratio = ratio;
        }
    }
    return best;
}

static void orcf_pivot(orc_full *f, int64_t r, int64_t s) {
    const int64_t R = f->R, W = f->W;
    double *rowr = f->T + r * W;
    const double p = rowr[s];
    double *col = (double *)malloc((size_t)R * sizeof(double));
    for (int64_t i = 0; i < R; ++i) col[i] = f->T[i * W + s];
    for (int64_t j = 0; j < W; ++j) rowr[j] = rowr[j] / p;
    for (int64_t i = 0; i < R; ++i) {
        if (i == r) continue;
        double *row = f->T + i * W;
        const double nc = -col[i];
        for (int64_t j = 0; j < W; ++j) row[j] = __builtin_fma(nc, rowr[j], row[j]);
    }
    free(col);
    f->is_basic[f->basis[r]] = 0;
    f->is_basic[s] = 1;
    f->basis[r] = (int32_t)s;
}

typedef struct {
    int64_t n_pivots, cap;
    double *snaps;
    int32_t *piv_row, *piv_col;
} orcf_log;

static void orcf_record(const orc_full *f, orcf_log *L, int64_t r, int64_t s) {
    if (L->n_pivots < L->cap) {
        L->piv_row[L->n_pivots] = (int32_t)r;
        L->piv_col[L->n_pivots] = (int32_t)s;
        memcpy(L->snaps + (L->n_pivots + 1) * f->R * f->W, f->T, (size_t)(f->R * f->W) * sizeof(double));
    }
    L->n_pivots++;
}

static int orcf_phase(orc_full *f, int64_t obj_row, int32_t rule, const orc_opts *o, int64_t max_pivots, orcf_log *L) {
    for (;;) {
        if (L->n_pivots >= max_pivots) return ORC_LIMIT;
        int64_t s = orcf_price(f, obj_row, rule, o->eps_cost);
        if (s < 0) return ORC_OPT;
        int64_t r = orcf_ratio(f, s, o->eps_pivot);
        if (r < 0) return ORC_UNBOUNDED;
        orcf_pivot(f, r, s);
        orcf_record(f, L, r, s);
    }
}

ORC_EXPORT int orc_full_dims(const double *b, const int8_t *ops, int64_t m, int64_t n, int64_t *R, int64_t *W) {
    int64_t n_log = 0, n_art = 0;
    for (int64_t i = 0; i < m; ++i) {
        int op = ops[i];
        if (b[i] < 0.0 && op != ORC_EQ) op = (op == ORC_LE) ? ORC_GE : ORC_LE;
        if (op != ORC_EQ) ++n_log;
        if (op != ORC_LE) ++n_art;
    }
    *R = m + (n_art > 0 ? 2 : 1);
    *W = n + n_log + n_art + 1;
    return 0;
}

ORC_EXPORT int orc_full_steps(const double *A, int64_t lda, const double *b, const double *c, const int8_t *ops,
                              int64_t m, int64_t n, const orc_opts *o, int64_t cap, double *snaps, int32_t *piv_row,
                              int32_t *piv_col, int32_t *var_ids, int32_t *basis_out, int64_t *out) {
    orc_full f;
    int64_t R, W;
    orc_full_dims(b, ops, m, n, &R, &W);
    f.m = m;
    f.R = R;
    f.W = W;
    f.T = (double *)calloc((size_t)(R * W), sizeof(double));
    f.basis = (int32_t *)calloc((size_t)(m > 0 ? m : 1), sizeof(int32_t));
    f.is_basic = (int32_t *)calloc((size_t)W, sizeof(int32_t));
    int32_t *logcol = (int32_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int32_t));
    int32_t *artcol = (int32_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int32_t));
    if (!f.T || !f.basis || !f.is_basic || !logcol || !artcol) return -1;
    const int32_t art_base = (int32_t)(n + m);
    /* columns in variable-id order */
    int64_t k = n, n_ge = 0;
    for (int64_t j = 0; j < n; ++j) var_ids[j] = (int32_t)j;
    for (int64_t i = 0; i < m; ++i) {
        int op = ops[i];
        if (b[i] < 0.0 && op != ORC_EQ) op = (op == ORC_LE) ? ORC_GE : ORC_LE;
        logcol[i] = -1;
        if (op != ORC_EQ) {
            logcol[i] = (int32_t)k;
            var_ids[k++] = (int32_t)(n + i);
        }
        if (op == ORC_GE) ++n_ge;
    }
    f.n_real = k;
    for (int64_t i = 0; i < m; ++i) {
        int op = ops[i];
        if (b[i] < 0.0 && op != ORC_EQ) op = (op == ORC_LE) ? ORC_GE : ORC_LE;
        artcol[i] = -1;
        if (op != ORC_LE) {
            artcol[i] = (int32_t)k;
            var_ids[k++] = (int32_t)(art_base + i);
        }
    }
    /* rows */
    for (int64_t i = 0; i < m; ++i) {
        double *row = f.T + i * W;
        int op = ops[i];
        const int neg = b[i] < 0.0;
        if (neg && op != ORC_EQ) op = (op == ORC_LE) ? ORC_GE : ORC_LE;
        for (int64_t j = 0; j < n; ++j) row[j] = neg ? -A[i * lda + j] : A[i * lda + j];
        row[W - 1] = neg ? -b[i] : b[i];
        if (op == ORC_LE) row[logcol[i]] = 1.0;
        if (op == ORC_GE) row[logcol[i]] = -1.0;
        if (op != ORC_LE) row[artcol[i]] = 1.0;
        f.basis[i] = (op == ORC_LE) ? logcol[i] : artcol[i];
        f.is_basic[f.basis[i]] = 1;
    }
    for (int64_t j = 0; j < n; ++j) f.T[m * W + j] = c[j];
    const int two = R == m + 2;
    if (two) {
        /* w = sum of the artificials, expressed in the non-basic variables: minus the sum of the artificial rows */
        double *w = f.T + (m + 1) * W;
        for (int64_t j = 0; j < W; ++j) {
            if (j < W - 1 && f.is_basic[j]) continue;
            double acc = 0.0;
            for (int64_t i = 0; i < m; ++i)
                if (artcol[i] >= 0) acc = acc + f.T[i * W + j];
            w[j] = -acc;
        }
    }
    orcf_log L;
    L.n_pivots = 0;
    L.cap = cap;
    L.snaps = snaps;
    L.piv_row = piv_row;
    L.piv_col = piv_col;
    memcpy(snaps, f.T, (size_t)(R * W) * sizeof(double));

    const int is_auto = o->max_pivots >= ORC_AUTO_BUDGET;
    const int64_t cap_piv = is_auto ? orc_auto_cap(m, n + n_ge + 1) : o->max_pivots;
    int64_t budget = cap_piv;
    int32_t rule = o->rule;
    int st = ORC_OPT;
    int64_t n_phase1 = 0;
    if (two) {
        st = orcf_phase(&f, m + 1, rule, o, budget, &L);
        if (st == ORC_LIMIT && is_auto && rule == ORC_RULE_DANTZIG) {
            rule = ORC_RULE_BLAND;
            budget += cap_piv;
            st = orcf_phase(&f, m + 1, rule, o, budget, &L);
        }
        if (st == ORC_UNBOUNDED) st = ORC_NUMERICAL;
        if (st == ORC_OPT && f.T[(m + 1) * W + W - 1] < -o->eps_feas) st = ORC_INFEASIBLE;
        if (st == ORC_OPT) {
            /* artificials still basic at level zero leave on the largest eligible entry of their row */
            for (int64_t i = 0; i < m && st == ORC_OPT; ++i) {
                if (f.basis[i] < f.n_real) continue; /* also skips nothing negative: flags are set below */
                const double *row = f.T + i * W;
                int64_t best = -1;
                double best_abs = 0.0;
                for (int64_t j = 0; j < f.n_real; ++j) {
                    if (f.is_basic[j]) continue;
                    double a = fabs(row[j]);
                    if (!(a > o->eps_pivot)) continue;
                    if (best < 0 || a > best_abs) {
                        best = j;
                        best_abs = a;
                    }
                }
                if (best < 0) {
                    f.basis[i] = -1 - f.basis[i];
                    continue;
                }
                if (L.n_pivots >= budget) {
                    st = ORC_LIMIT;
                    break;
                }
                orcf_pivot(&f, i, best);
                orcf_record(&f, &L, i, best);
            }
        }
        n_phase1 = L.n_pivots;
    }
    if (st == ORC_OPT) {
        st = orcf_phase(&f, m, rule, o, budget, &L);
        if (st == ORC_LIMIT && is_auto && rule == ORC_RULE_DANTZIG) {
            rule = ORC_RULE_BLAND;
            budget += cap_piv;
            st = orcf_phase(&f, m, rule, o, budget, &L);
        }
    }
    for (int64_t i = 0; i < m; ++i) {
        const int32_t bi = f.basis[i] < 0 ? -1 - f.basis[i] : f.basis[i];
        basis_out[i] = f.basis[i] < 0 ? -1 - var_ids[bi] : var_ids[bi];
    }
    out[0] = R;
    out[1] = W;
    out[2] = L.n_pivots;
    out[3] = st;
    out[4] = n_phase1;
    out[5] = art_base;
    free(f.T);
    free(f.basis);
    free(f.is_basic);
    free(logcol);
    free(artcol);
    return 0;
}

/* Default team size of the parallel loops that take no explicit thread count (the generator).  torchrun exports
 * OMP_NUM_THREADS=1 to its workers; the CPU arm of bench.py sets the team size it reports explicitly. */
ORC_EXPORT void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_EXPORT int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
