/*
 * b200lp.h -- C ABI of libb200lp.so: the simplex tableau pivot loop as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary for the hot path of the-utn-team/simplex-solver's /solve.  The reference has
 * no FFI of its own; its de-facto seams are three groups of module-level names in
 * app/controllers/solver_controller.py that its tests already swap by attribute patching
 * (tests/test_solver_controller.py:123).  Each entry point below names the reference interface it serves:
 *
 *   b200lp_solve_dense   <- scipy.optimize.linprog(...) called at solver_controller.py:78-85
 *                           (status / x* / z*), and the pivot loop of simple_simplex.optimize_json_format
 *                           called at solver_controller.py:318 (pivot history -> "pivotSteps")
 *   b200lp_read_tableau  <- step["tableau"] consumed at solver_controller.py:332-362
 *   b200lp_solve_batched <- the same solve, B independent problems (BASELINE config 3; the reference runs
 *                           one SolverController.run() per problem, ui_controller.py:194-195)
 *   b200lp_attach/_generate/_run, b200lp_shard_* <- the same pivot loop on tableaux that the reference's
 *                           dict-of-dicts input (solver_controller.py:33-50) cannot carry (configs 4, 5)
 *   b200lp_select_entering / b200lp_ratio_test / b200lp_pivot <- the three phases of one iteration, exposed
 *                           one at a time for kernel-level parity tests and ncu.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every function returns 0 on success or a
 * negative B200LP_E_* code; b200lp_last_error() gives the thread-local message.  An LP outcome (infeasible,
 * unbounded, pivot limit) is NOT an error: it is result->status, with scipy.optimize.linprog's integers
 * (solver_controller.py:382-414 maps 0 -> "Solucion Factible", 2 -> "Sin Solucion Factible", rest -> "Error").
 * There is no CPU fallback: without a CUDA device every compute entry point returns B200LP_E_CUDA.
 * A b200lp_solver is a workspace owned by one thread at a time (create one per thread; the Python host
 * keeps them thread-local because the reference's stress tests drive one app from 30 threads,
 * tests/test_performance_load.py:186-199).  Calls release nothing and keep no pointer of the caller's
 * beyond the call, except b200lp_attach, which binds a caller-owned device tableau until the next attach.
 */
#ifndef B200LP_H
#define B200LP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LP_VERSION 200

/* entering-variable rules */
#define B200LP_RULE_DANTZIG 0
#define B200LP_RULE_BLAND 1

/* result->status: scipy.optimize.linprog status integers */
#define B200LP_STATUS_OPTIMAL 0
#define B200LP_STATUS_LIMIT 1
#define B200LP_STATUS_INFEASIBLE 2
#define B200LP_STATUS_UNBOUNDED 3
#define B200LP_STATUS_NUMERICAL 4

/* row operators (the L / G / E codes of solver_controller.py:305-306) */
#define B200LP_OP_LE 0
#define B200LP_OP_GE 1
#define B200LP_OP_EQ 2

/* pivot-update kernel variants */
#define B200LP_UPDATE_AUTO 0
#define B200LP_UPDATE_LDG 1 /* 128-bit vectorised global loads/stores, register resident            */
#define B200LP_UPDATE_TMA 2 /* cp.async.bulk.tensor tiles staged through shared memory (mbarrier ring) */

/* loop drivers */
#define B200LP_LOOP_LAUNCHES 0 /* three plain launches per pivot                                              */
#define B200LP_LOOP_GRAPH 1    /* the iteration replayed as a CUDA graph                                       */
#define B200LP_LOOP_AUTO 2     /* default: tableaux that fit the chip's shared memory (<= ~29 MB) run in ONE  */
                               /* persistent cooperative kernel, SM-sharded and shared-memory resident, with  */
                               /* one grid barrier per pivot; everything larger uses the look-ahead loop      */
                               /* (B200LP_LOOP_BLOCKED, K = 32), which is faster than the rank-1 loop at every */
                               /* size and takes the same pivots bit for bit                                  */

#define B200LP_LOOP_BLOCKED 3  /* look-ahead pivoting: K pivots (check_every, 1..32, default 32) are decided from    */
                               /* O(R + C) state and applied to the tableau in ONE pass -- 2*R*C*8/K bytes of HBM  */
                               /* traffic per pivot, bit-identical pivots and tableau (kernels_blocked.cuh)        */

/* error codes */
#define B200LP_OK 0
#define B200LP_E_INVALID (-1)
#define B200LP_E_CUDA (-2)
#define B200LP_E_NOMEM (-3)
#define B200LP_E_STATE (-4)

typedef struct b200lp_solver b200lp_solver;

typedef struct b200lp_opts {
    int32_t rule;           /* B200LP_RULE_*                                                            */
    int32_t update_variant; /* B200LP_UPDATE_*                                                          */
    int64_t max_pivots;     /* pivot budget of the call (status LIMIT when exhausted); the default      */
                            /* (>= 2^40) means "automatic": 200 * (rows + columns) + 10000              */
    double eps_cost;        /* a reduced cost d_j enters only if d_j < -eps_cost                        */
    double eps_pivot;       /* a column entry takes part in the ratio test only if > eps_pivot          */
    double eps_feas;        /* phase 1 ends infeasible if the sum of artificials exceeds eps_feas       */
    int32_t check_every;    /* pivots enqueued between two host reads of the device status (0: default) */
    int32_t loop_mode;      /* B200LP_LOOP_*: how the device-resident loop is driven                     */
    double time_limit_s;    /* wall-clock bound of one b200lp_solve / _solve_dense / _run call in seconds,  */
                            /* counted from its entry (H2D copies and tableau build included); when it    */
                            /* expires the loop stops at the next pivot (on-chip loop) or block of pivots */
                            /* with status LIMIT and a consistent tableau.  <= 0: no bound.  The reference */
                            /* passes time_limit = 10 and maps the limit status to "Error"                */
                            /* (solver_controller.py:76, :404).  Ignored by b200lp_solve_batched.         */
} b200lp_opts;

/* min c'x  s.t.  A_i x (ops_i) b_i,  x >= 0.   A is m x n row-major with row stride lda. */
typedef struct b200lp_problem {
    int64_t m, n, lda;
    const double *A, *b, *c;
    const int8_t *ops;  /* m entries, B200LP_OP_*; always HOST memory                                  */
    int32_t on_device;  /* 0: A, b, c are host pointers (copied inside the call); 1: device pointers     */
    int32_t reserved;
} b200lp_problem;

typedef struct b200lp_result {
    int32_t status;      /* B200LP_STATUS_*                                                             */
    int32_t reserved;
    double fun;          /* objective of the minimisation form, c'x                                     */
    int64_t n_pivots;    /* pivots performed by the call (both phases)                                  */
    int64_t n_phase1;    /* of which phase 1 + driving artificials out                                  */
    double *x;           /* optional HOST buffer of x_len doubles; receives x*                          */
    int64_t x_len;
    int32_t *piv_row;    /* optional HOST buffers of hist_cap entries: pivot row / column positions in  */
    int32_t *piv_col;    /*   the stored tableau, entering / leaving variable ids, one per pivot        */
    int32_t *enter_lab;
    int32_t *leave_lab;
    int64_t hist_cap;
    double device_ms;        /* GPU time of the pivot loop, CUDA events on the solver's stream          */
    int64_t kernel_launches; /* kernels of this library launched by the call                             */
} b200lp_result;

/* ---- library ---------------------------------------------------------------------------------------- */
int b200lp_version(void);
const char *b200lp_last_error(void);
void b200lp_default_opts(b200lp_opts *opts);

/* ---- workspace -------------------------------------------------------------------------------------- */
int b200lp_create(b200lp_solver **out, int device);
int b200lp_destroy(b200lp_solver *s);
/* run the solver's kernels on a caller stream (cudaStream_t as void*, used as given: 0 is the legacy default
 * stream, on which the loop uses plain launches because CUDA graphs cannot be captured there);
 * b200lp_use_own_stream goes back to the solver's own non-blocking stream (the default after create). */
int b200lp_set_stream(b200lp_solver *s, void *stream);
int b200lp_use_own_stream(b200lp_solver *s);
int b200lp_synchronize(b200lp_solver *s);
/* Guard mode (environment B200LP_GUARD=1 when the library is first used): every buffer of the workspace lies between
 * two 4 KB bands of a byte pattern; this counts the band bytes that were overwritten since the buffers were allocated
 * (0 = no store left its buffer).  Test instrumentation for boxes where compute-sanitizer is not available.      */
int b200lp_check_guards(b200lp_solver *s, int64_t *corrupted_bytes);

/* ---- one LP, reference-facing (linprog seam) --------------------------------------------------------- */
int b200lp_solve_dense(b200lp_solver *s, const b200lp_problem *p, const b200lp_opts *o, b200lp_result *r);
/* the two halves of b200lp_solve_dense: build the initial tableau in the solver's own device buffer (then
 * b200lp_dims / b200lp_get_labels / b200lp_read_tableau describe it), and b200lp_solve below.            */
int b200lp_build_dense(b200lp_solver *s, const b200lp_problem *p);
/* keep a dense R x C copy of the tableau after each of the first `cap` pivots of the following solves in the
 * caller-owned DEVICE buffer snaps_dev (cap * R * C doubles); NULL switches it off.  Feeds the "pivotSteps"
 * of solver_controller.py:332-362 for small LPs.                                                          */
int b200lp_set_snapshots(b200lp_solver *s, double *snaps_dev, int64_t cap);

/* ---- device-resident tableau ------------------------------------------------------------------------ */
/* Stored (condensed) tableau: R = m + n_obj rows, C columns, last column = right-hand side, row stride ld
 * doubles (ld even, ld >= C, base 16-byte aligned).  Row m = objective row, row m+1 = phase-1 row.      */
int b200lp_attach(b200lp_solver *s, double *T_dev, int64_t m, int64_t n_obj, int64_t C, int64_t ld,
                  int64_t n_struct, int32_t art_base);
int b200lp_dims(b200lp_solver *s, int64_t *m, int64_t *n_obj, int64_t *C, int64_t *ld);
/* fill the attached tableau with the synthetic dense LP of BASELINE configs 4/5: structural variables
 * [lab0, lab0 + C-1) of an LP with n_total structural variables, plus the RHS as last column.           */
int b200lp_generate(b200lp_solver *s, uint64_t seed, int64_t n_total, int64_t lab0);
int b200lp_set_labels(b200lp_solver *s, const int32_t *rowlab_host, const int32_t *collab_host);
int b200lp_get_labels(b200lp_solver *s, int32_t *rowlab_host, int32_t *collab_host);
/* copy the stored tableau (R x C, dense, row-major) to a HOST buffer */
int b200lp_read_tableau(b200lp_solver *s, double *T_host);
/* x* (n_struct doubles) and c'x of the current basis, to HOST */
int b200lp_read_solution(b200lp_solver *s, double *x_host, double *fun);
/* the pivot loop on objective row obj_row until optimal / unbounded / o->max_pivots */
int b200lp_run(b200lp_solver *s, const b200lp_opts *o, int64_t obj_row, b200lp_result *r);
/* full two-phase solve of the current tableau (phase 1 only when n_obj == 2) */
int b200lp_solve(b200lp_solver *s, const b200lp_opts *o, b200lp_result *r);

/* ---- one phase at a time ----------------------------------------------------------------------------- */
int b200lp_select_entering(b200lp_solver *s, int64_t obj_row, int32_t rule, double eps_cost, int64_t *col_out);
int b200lp_ratio_test(b200lp_solver *s, int64_t col, double eps_pivot, int64_t *row_out);
int b200lp_pivot(b200lp_solver *s, int64_t row, int64_t col, int32_t update_variant);

/* ---- column-sharded tableau (one shard per GPU; the caller runs the all-gather between the two) ------ */
/* cand_dev: (R + 2) doubles on the device: [best reduced cost, variable id (as double, -1 none), column] */
int b200lp_shard_candidate(b200lp_solver *s, const b200lp_opts *o, int64_t obj_row, double *cand_dev);
/* gathered_dev: world x (R + 2) doubles; picks the global entering column, runs ratio test + update     */
int b200lp_shard_pivot(b200lp_solver *s, const b200lp_opts *o, const double *gathered_dev, int32_t world,
                       int32_t rank);
/* look-ahead (B200LP_LOOP_BLOCKED) form of the same protocol: begin once per run (after shard_reset), then per pivot
 * blk_candidate -> all-gather -> blk_pivot, and blk_flush after every K <= 32 pivots and at the end of the run (the
 * tableau is only up to date after a flush)                                                                  */
int b200lp_shard_blk_begin(b200lp_solver *s, int64_t obj_row);
int b200lp_shard_blk_candidate(b200lp_solver *s, const b200lp_opts *o, int64_t obj_row, double *cand_dev);
int b200lp_shard_blk_pivot(b200lp_solver *s, const b200lp_opts *o, const double *gathered_dev, int32_t world,
                           int32_t rank);
int b200lp_shard_blk_flush(b200lp_solver *s, int64_t obj_row);
/* Peer-memory exchange: instead of handing the candidates to the caller's all-gather, every shard STORES its candidate
 * straight into an exchange region of every peer over NVLink and acquires the generation words of its own region --
 * inside ONE kernel per pivot decision that also prices, picks the winner and runs the ratio test (k_shard_pick).
 * bases[g] = device address, as seen from this process, of rank g's region of b200lp_p2p_bytes(R, world) bytes (e.g.
 * torch symmetric memory buffer_ptrs; 16-byte aligned); all ranks must synchronise once after b200lp_p2p_connect and
 * before the first exchange.  Per pivot: b200lp_shard_fused (lookahead = 0: followed by the rank-1 update; 1: look-ahead
 * loop between b200lp_shard_blk_begin and b200lp_shard_blk_flush).                                                   */
int64_t b200lp_p2p_bytes(int64_t R, int32_t world);
int b200lp_p2p_connect(b200lp_solver *s, void *const *bases, int32_t world, int32_t rank);
int b200lp_shard_fused(b200lp_solver *s, const b200lp_opts *o, int64_t obj_row, int32_t lookahead);
/* n shards of ONE tableau emulated on one GPU (tests): one launch, one thread-block cluster per shard, all resident, so
 * that shards waiting for each other never sit in separate launches; shards[k] must be rank k of n, same device/stream */
int b200lp_shard_fused_multi(b200lp_solver *const *shards, int32_t n, const b200lp_opts *o, int64_t obj_row,
                             int32_t lookahead);
/* synchronise and read the loop state of a sharded run */
int b200lp_shard_state(b200lp_solver *s, int32_t *done, int32_t *status, int64_t *n_pivots);
int b200lp_shard_reset(b200lp_solver *s, int64_t max_pivots);
/* A number that changes whenever anything a CALLER-captured CUDA graph of b200lp_shard_* launches bakes in may have
 * changed (workspace buffers reallocated -- e.g. the pivot history growing with max_pivots --, tableau re-attached, peer
 * regions or snapshots re-bound).  A caller that replays such graphs keys them by it and re-captures when it moves.  */
int b200lp_binding_epoch(b200lp_solver *s, int64_t *epoch);
int b200lp_read_history(b200lp_solver *s, int64_t cap, int32_t *piv_row, int32_t *piv_col, int32_t *enter_lab,
                        int32_t *leave_lab, int64_t *n_out);

/* ---- B independent LPs of one shape ----------------------------------------------------------------- */
/* A[B][m][n], b[B][m], c[B][n] (minimisation costs), ops[B][m]; outputs status[B], fun[B], x[B][n]
 * (may be NULL), n_pivots[B], piv_log[B][log_cap][2] (may be NULL).  on_device: all pointers are device
 * pointers (else host pointers, copied inside the call).                                                */
int b200lp_solve_batched(b200lp_solver *s, int64_t B, int64_t m, int64_t n, const double *A, const double *b,
                         const double *c, const int8_t *ops, const b200lp_opts *o, int32_t *status, double *fun,
                         double *x, int32_t *n_pivots, int32_t *piv_log, int64_t log_cap, int32_t on_device,
                         double *device_ms);
/* device_ms: kernel time with device pointers; with host pointers the span from the first H2D to the last D2H (the
 * batch is cut into 8 chunks on two streams so that copies overlap the kernels).                               */

/* ---- measurement helpers ----------------------------------------------------------------------------- */
/* time `reps` launches of the pivot-update kernel alone on pivot (row, col) of the attached tableau
 * (CUDA events on the solver's stream); the tableau is modified.                                        */
int b200lp_time_update(b200lp_solver *s, int64_t row, int64_t col, int32_t update_variant, int32_t reps,
                       double *ms_per_launch);

/* `iters` loop iterations with plain launches and CUDA events around every kernel: average duration of the
 * price, ratio and update kernels in their real order (what bench.py reports under "roofline").            */
int b200lp_profile_loop(b200lp_solver *s, const b200lp_opts *o, int64_t obj_row, int32_t iters, double *ms_price,
                        double *ms_ratio, double *ms_update, int64_t *pivots_done);

#ifdef __cplusplus
}
#endif
#endif /* B200LP_H */
