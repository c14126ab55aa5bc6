/*
 * pdlp_b200.h -- C ABI of the B200-native restarted-PDHG LP engine.
 *
 * This is the drop-in boundary for torchPDLP's hot path.  torchPDLP has no
 * FFI / plugin registry of its own: its boundary is the
 * Python call surface of PDLP/primal_dual_hybrid_gradient.py,
 * PDLP/primal_dual_hybrid_gradient_step.py, PDLP/helpers.py and
 * PDLP/enhancements.py.  Each entry point below names the reference function
 * (file:line under /root/reference) whose device work it replaces; the ctypes
 * binding a torchPDLP maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - Every function returns 0 on success and a nonzero pdlp_status on error;
 *     pdlp_last_error() returns a thread-local message.  Nothing throws.
 *   - All vector / matrix pointers are DEVICE pointers (cudaMalloc'ed or
 *     tensor.data_ptr()) unless the parameter name starts with `host_`.
 *   - Values are fp64 (PDLP_F64) or fp32 (PDLP_F32); indices are int32.
 *   - Work is enqueued on the `stream` given at pdlp_create (pass torch's
 *     current stream).  Functions that return scalars to the host synchronise
 *     that stream; the others are asynchronous.
 *   - A handle is bound to the CUDA device current at pdlp_create and is not
 *     thread-safe.  The library copies the problem data into its own buffers
 *     (so Ruiz scaling never mutates caller memory) and owns its workspace.
 *   - There is no CPU fallback: with no CUDA device every call fails with
 *     PDLP_ERR_CUDA.
 */
#ifndef PDLP_B200_H_
#define PDLP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdlp_engine* pdlp_handle;

enum pdlp_dtype { PDLP_F64 = 0, PDLP_F32 = 1 };

enum pdlp_status {
  PDLP_OK = 0,
  PDLP_ERR_INVALID = 1, /* bad argument */
  PDLP_ERR_CUDA = 2,    /* CUDA runtime error (message has the cudaError string) */
  PDLP_ERR_STATE = 3,   /* call sequence error (e.g. restart check before any iteration) */
  PDLP_ERR_NOT_CANONICAL = 4 /* pdlp_create_host: unsorted or duplicate column indices inside a row (coalesce first) */
};

/* Restart kinds reported by pdlp_restart_check (reference
 * primal_dual_hybrid_gradient.py:131-141). */
enum pdlp_restart_kind {
  PDLP_RESTART_NONE = 0,
  PDLP_RESTART_SUFFICIENT = 1,
  PDLP_RESTART_NECESSARY = 2,
  PDLP_RESTART_ARTIFICIAL = 3,
  PDLP_RESTART_FORCED = 4 /* inner loop left because the KKT-pass limit was hit (:67) */
};

/* Infeasibility verdicts (reference enhancements.py:141,159). */
enum pdlp_infeas { PDLP_FEASIBLE_SO_FAR = 0, PDLP_DUAL_INFEASIBLE = 1, PDLP_PRIMAL_INFEASIBLE = 2 };

/* Residuals of one iterate: helpers.py:53-96 (compute_residuals_and_duality_gap)
 * plus helpers.py:98-108 (KKT_error).  pr_unscaled / dr_unscaled are the same
 * norms taken on the un-preconditioned problem (reference :157-163); they equal
 * pr / dr when no scaling is active. */
typedef struct pdlp_residuals {
  double pr;          /* ||[ (Kx-q)_eq ; min((Kx-q)_ineq, 0) ]||_2 */
  double dr;          /* ||g - lambda||_2,  g = c - K^T y */
  double gap;         /* adjusted_dual - prim_obj, signed */
  double prim_obj;    /* c^T x */
  double adj_dual;    /* q^T y + l_dual^T max(lambda,0) + u_dual^T min(lambda,0) */
  double kkt;         /* sqrt(w^2 pr^2 + dr^2 / w^2 + gap^2) at the omega in force */
  double pr_unscaled;
  double dr_unscaled;
} pdlp_residuals;

/* Host-visible snapshot of the solver scalars (all kept on the device between
 * calls; this is a copy). */
typedef struct pdlp_state {
  double eta;         /* step size the NEXT iteration will use */
  double omega;       /* primal weight */
  double theta;       /* extrapolation (reference always passes 1.0) */
  double eta_sum;     /* sum of averaging weights since the last restart */
  double last_weight; /* averaging weight of the latest iterate (step.py:110-115) */
  double dx2, dy2, inter; /* ||dx||^2, ||dy||^2, dy^T K dx of the latest step */
  double kkt_first;   /* KKT error at the last restart point */
  double q_norm, c_norm;
  int64_t k;          /* total iterations */
  int32_t t;          /* iterations since the last restart */
  int32_t n_restarts;
  int32_t infeas;     /* pdlp_infeas */
  int32_t infeas_k;   /* iteration at which it was detected */
  int32_t solved;     /* 1 when the last restart point met check_termination */
  int32_t reserved;
} pdlp_state;

/* Result of one restart check. */
typedef struct pdlp_check {
  int32_t kind;        /* pdlp_restart_kind; NONE = keep iterating */
  int32_t use_average; /* 1: restart point is the weighted average (ties -> average, :133) */
  int32_t t;           /* inner iteration count at the check */
  int32_t solved;      /* check_termination (helpers.py:110-128) on the restart point */
  int64_t k;
  double omega;        /* after primal_weight_update when enabled (enhancements.py:73-78) */
  double kkt_current, kkt_average, kkt_previous, kkt_first_new;
  pdlp_residuals point; /* residuals of the chosen restart point (what the reference prints, :167-170) */
} pdlp_check;

const char* pdlp_last_error(void);

/* Number of bytes the engine believes one PDHG iteration must move
 * (SURVEY.md section 8d).  For reporting only. */
int64_t pdlp_bytes_per_iteration(pdlp_handle h);

/* Host-only inspection of the layout the planner would build for a matrix with these row lengths (no device
 * work, usable without a GPU: the `-m "not gpu"` tests check its invariants).  value_bytes = 8 or 4; sms = number
 * of SMs to plan for (148 on B200).  host_row_class (may be NULL) receives per row 0 = lane row (SELL slice),
 * 1 = warp row (32 segments of one row), 2 = CTA row (straight from CSR).                                       */
typedef struct pdlp_plan_info {
  int64_t rows, nnz;
  int64_t lane_rows, warp_rows, cta_rows;
  int64_t slots;        /* padded (index, value) slots of the packed copy; >= nnz - nnz of the CTA rows */
  int64_t packed_bytes; /* headers + row ids + slots */
  int32_t sigma;        /* sorting window chosen by the cost model */
  int32_t grid;         /* CTAs */
  int32_t warps;        /* grid x warps per CTA */
  int32_t pieces;
  int32_t max_piece_bytes;
  int32_t stages;       /* ring stages per warp */
  int64_t max_warp_bytes, mean_warp_bytes; /* balance of the per-warp piece ranges */
  int32_t covered_once; /* 1 iff every row appears in exactly one class and every lane row in exactly one slice lane */
  int32_t ranges_ok;    /* 1 iff the warp ranges are monotone, cover all pieces and start where rows start */
} pdlp_plan_info;
int pdlp_plan_inspect(const int32_t* host_rowptr, int64_t rows, int32_t value_bytes, int32_t sms,
                      pdlp_plan_info* host_out, int8_t* host_row_class);

/* Host-side checker of the run layout (torchpdlp_b200/csrc/run_spmv.cuh; the layout of the two iteration kernels).
 * Plans and packs a host CSR matrix (fp64) with the same planner and per-slot packer the device path uses, walks the
 * chunks with a 32-lane emulation of the kernel's chunk walk and writes out = A x.  mode 0: the run layout (rows
 * longer than long_row take the one-row-per-CTA path).  mode 1: the shared-memory panels of K^T
 * (torchpdlp_b200/csrc/panel_spmv.cuh): panel boundaries, the panel matrix and its per-panel SELL plans are built as
 * on the device path, every CTA's pieces are checked to lie in its own panel and to cover every (panel, row) once,
 * and out = sum over panels of the partial sums (fails when x does not fit PANEL_MAX panels); then `chunks` counts
 * pieces, `stream_entries` padded slots, `chunk_bytes` the largest piece, `max_warp_chunks` the sorting window.
 * No GPU involved; used by tests/test_cpu_run_layout.py.
 * Replaces nothing in the reference: torchPDLP's products are ATen calls
 * (PDLP/primal_dual_hybrid_gradient_step.py:25,33,68,84). */
typedef struct pdlp_run_info {
  int64_t stream_entries; /* nonzeros + empty-row / long-row markers (all panels) */
  int64_t rows_finished;  /* row epilogues the walk ran (mode 1: per panel); must equal rows */
  int32_t grid, warps, chunks, long_rows;
  int32_t panels, panel_cap, chunk_bytes, max_warp_chunks;
} pdlp_run_info;
int pdlp_run_emulate(const int32_t* host_rowptr, const int32_t* host_colidx, const double* host_vals, int64_t rows,
                     int64_t cols, int32_t sms, int32_t long_row, int32_t mode, const double* host_x, double* host_out,
                     pdlp_run_info* host_info);

/* ---- lifetime -------------------------------------------------------------
 * Problem tuple (c, K, q, m_ineq, l, u): reference PDLP/util.py:239-269,
 * consumed at primal_dual_hybrid_gradient.py:7.  K is given as CSR (m x n) AND
 * as CSR of K^T (n x m) with column indices sorted inside each row, so both
 * products are row-parallel and atomic-free.  l/u may hold -inf/+inf.        */
int pdlp_create(pdlp_handle* out, int dtype, int64_t m, int64_t n, int64_t m_ineq, int64_t nnz,
                const int32_t* K_rowptr, const int32_t* K_colidx, const void* K_vals,
                const int32_t* Kt_rowptr, const int32_t* Kt_colidx, const void* Kt_vals,
                const void* c, const void* q, const void* l, const void* u, void* stream);
/* The same engine from a HOST CSR of K alone (pinned memory gives asynchronous copies; pageable works): what the
 * reference spreads over K.to(device) at its call sites (PDLP/main.py -> primal_dual_hybrid_gradient.py:7) and the
 * sparse conversions of PDLP/util.py:239-269.  host_K_rowptr / host_K_colidx hold index_bytes = 4 or 8 byte signed
 * integers (torch's crow_indices / col_indices are int64), host_K_vals values of `dtype`.  The library copies the
 * arrays to the device, narrows the indices to int32, checks the canonical form (columns strictly increasing inside
 * every row: PDLP_ERR_NOT_CANONICAL otherwise), builds the CSR of K^T on the device (column histogram, scan, stable
 * radix sort by column) and plans both layouts on the host while the copies and the transposition run.
 * c, q, l, u may be host or device pointers (unified addressing decides).  Synchronises `stream`.            */
int pdlp_create_host(pdlp_handle* out, int dtype, int64_t m, int64_t n, int64_t m_ineq, int64_t nnz,
                     const void* host_K_rowptr, const void* host_K_colidx, int32_t index_bytes,
                     const void* host_K_vals, const void* c, const void* q, const void* l, const void* u, void* stream);
int pdlp_destroy(pdlp_handle h);
/* Redirect the engine's work to another stream of the same device (e.g. the capture stream of a
 * CUDA graph that also holds the NCCL allreduce of the sharded loop).  The caller orders the streams. */
int pdlp_set_stream(pdlp_handle h, void* stream);

/* ---- iterate / scalars ----------------------------------------------------
 * x_init / y_init of pdlp_algorithm (:31-36).  Also resets the averaging sums,
 * t and the restart snapshot (:58-64).                                        */
int pdlp_set_iterate(pdlp_handle h, const void* x, const void* y);
int pdlp_get_iterate(pdlp_handle h, void* x, void* y);
/* Weighted-average iterate of the current epoch (:118-119). */
int pdlp_get_average(pdlp_handle h, void* x_avg, void* y_avg);
/* eta, omega, theta as at :22-25; adaptive selects step.py:43-115 vs :3-40. */
int pdlp_set_step(pdlp_handle h, double eta, double omega, double theta, int adaptive);
/* tol (:7), primal_update (:150-151), infeasibility detection (:89-101). */
int pdlp_set_options(pdlp_handle h, double tol, int primal_update, int infeasibility_detect, double infeas_tol);
int pdlp_get_state(pdlp_handle h, pdlp_state* host_out);
/* Iteration counters k (total, :76) and t (since restart, :104): the adaptive rule
 * reads k (step.py:102,107), the artificial restart test reads both (:139). */
int pdlp_set_counters(pdlp_handle h, int64_t k, int32_t t);

/* ---- the hot loop -----------------------------------------------------------
 * n_iters PDHG iterations (fixed_one_step_pdhg step.py:3-40 or
 * adaptive_one_step_pdhg step.py:43-115, plus the averaging update :104-112),
 * two kernels per iteration, scalars stay on the device.  With
 * save_previous != 0 the iterate before the LAST step is kept for the restart
 * test (:77-78,:126).  use_graph: 0 = one launch per kernel; 1 = replay a cached
 * CUDA graph of the block; 2 = ONE persistent cooperative kernel for the whole block
 * (grid barriers between the passes, step rule evaluated by every CTA; not with
 * infeasibility detection, which falls back to 1).  All three give bitwise the same
 * iterates.  Async.                                                             */
int pdlp_iterate(pdlp_handle h, int n_iters, int save_previous, int use_graph);

/* Measurement aid for bench.py: run n_iters iterations WITHOUT a graph, bracketing every
 * launch of the two iteration kernels with CUDA events on the engine's stream, and return
 * the average device time per launch (milliseconds) of the primal (K^T y) and dual (K xbar)
 * kernels.  Advances the solver state like pdlp_iterate.  Synchronises. */
int pdlp_time_kernels(pdlp_handle h, int n_iters, float* host_primal_ms, float* host_dual_ms);
/* The same two kernels timed as they run in the loop: n_iters iterations launched back to back with programmatic
 * dependent launch (nothing between the kernels, so the overlap of a kernel's prologue with its predecessor's tail
 * and the warm L2 are those of the block graph); every launch records %globaltimer when its first CTA leaves
 * griddepcontrol.wait and after its last CTA's finalize.  Returns the average of those spans per kernel and the
 * event-timed wall time per iteration (milliseconds).  Default layout (SELL, single pass, no infeasibility
 * detection) only: PDLP_ERR_STATE otherwise.  Advances the solver state.  Synchronises. */
int pdlp_time_kernels_in_loop(pdlp_handle h, int n_iters, float* host_primal_ms, float* host_dual_ms,
                              float* host_iteration_ms);

/* Restart test + restart epilogue (:115-177): averages, the three KKT errors
 * with ONE pass over K and one over K^T, the sufficient / necessary / artificial
 * criteria, candidate selection, primal-weight update, KKT_first, termination
 * test.  force != 0 evaluates and restarts at the current iterate (loop left by
 * the KKT-pass limit, :67 then :148).  Synchronises; fills *host_out.         */
int pdlp_restart_check(pdlp_handle h, int force, pdlp_check* host_out);

/* helpers.py:53-108 for caller-supplied vectors (x: n, y: m) at primal weight
 * omega.  Synchronises. */
int pdlp_kkt(pdlp_handle h, const void* x, const void* y, double omega, pdlp_residuals* host_out);

/* c^T x of the current iterate (what pdlp_algorithm returns as prim_obj when an
 * infeasibility certificate ends the solve, :95).  Synchronises. */
int pdlp_primal_objective(pdlp_handle h, double* host_out);

/* ---- setup-time kernels -----------------------------------------------------
 * Power iteration spectral_norm_estimate_torch (helpers.py:41-51) started from
 * b0 (n values, the caller draws it: reference uses torch.randn).  Synchronises. */
int pdlp_power_iteration(pdlp_handle h, const void* b0, int iters, double* host_sigma);
/* out = K in (transpose == 0, in: n, out: m) or K^T in (transpose != 0). Async. */
int pdlp_spmv(pdlp_handle h, int transpose, const void* in, void* out);
/* Ruiz equilibration, ruiz_precondition (enhancements.py:43-71), in place on the
 * engine's copy of (K, K^T, c, q, l, u).  Synchronises; *host_rounds = rounds run. */
int pdlp_ruiz(pdlp_handle h, int max_iter, double eps, int* host_rounds);
/* Pock-Chambolle (alpha = 1) on top of the current scaling.  No reference
 * implementation exists in the snapshot (main.py:8 imports a missing symbol). */
int pdlp_pock_chambolle(pdlp_handle h);
/* D_row (m), D_col (n): accumulated scalings, ones when none was applied. */
int pdlp_get_scaling(pdlp_handle h, void* D_row, void* D_col);
/* Declare that the problem held by the engine is ALREADY scaled by (D_row, D_col)
 * (data_precond of pdlp_algorithm, :157-163): termination residuals are then taken
 * on the unscaled problem.  Does not touch the data.  Both pointers NULL: forget the scaling (termination on the
 * problem as held: pdlp_algorithm(precondition=False) on an engine that was used with a scaling before). */
int pdlp_set_scaling(pdlp_handle h, const void* D_row, const void* D_col);
/* Current (possibly scaled) problem vectors; any pointer may be NULL. */
int pdlp_get_problem(pdlp_handle h, void* c, void* q, void* l, void* u);
/* Current (possibly scaled) values of K in the caller's CSR order (nnz values). */
int pdlp_get_matrix_values(pdlp_handle h, void* K_vals);
/* Structure of K as the engine holds it (m + 1 row pointers, nnz column indices, int32, device); either may be NULL. */
int pdlp_get_matrix_structure(pdlp_handle h, int32_t* K_rowptr, int32_t* K_colidx);

/* ---- fishnet start-point search (spectral_casting.py), 8 points per call -------------
 * X8: n x 8 and Y8: m x 8 values, point index fastest (row j holds the 8 points' j-th entry,
 * 64-byte aligned).  Fixed-step PDHG on all 8 points with one pass over K^T and one over K:
 * PDHG_step (spectral_casting.py:254-293); Y8 = K X8 (:99); per point c^T x,
 * l_dual^T max(lambda,0), u_dual^T min(lambda,0) of get_best_pts (:217-234) into
 * host_out[3 v + {0,1,2}] (synchronises).  XB8 is an n x 8 scratch buffer.           */
int pdlp_batch_spmv(pdlp_handle h, const void* X8, void* Y8);
int pdlp_batch_step(pdlp_handle h, void* X8, void* XB8, void* Y8, double eta, double omega);
int pdlp_batch_gap(pdlp_handle h, const void* X8, const void* Y8, double* host_out24);

/* ---- row-partitioned engine (multi-GPU) ------------------
 * Each rank builds an ordinary handle on ITS rows of K (m = local rows, m_ineq = local
 * ">=" rows; n, c, l, u replicated) and marks it sharded.  K^T y = sum over ranks of the
 * local partial, so the column side of every pass is split around a host-side allreduce
 * (NCCL through torch.distributed, on the handle's stream):
 *
 *   iteration     pdlp_shard_kty(buf)        buf[0..n) = K_g^T y_g, buf[n..n+4) = partial norms
 *                 allreduce(buf, n + 4)
 *                 pdlp_shard_step(buf, save) step rule of the PREVIOUS iteration (global norms),
 *                                            primal update, local dual kernel
 *   restart test  pdlp_shard_rule_pack / allreduce(4) / pdlp_shard_rule_apply   (pending rule)
 *                 pdlp_shard_check_begin(buf3, have_previous)   buf3[3j+v] = K_g^T Y_g for the 3 iterates
 *                 allreduce(buf3, 3n)
 *                 pdlp_shard_check_mid(buf3, buf9)   column sums (replicated), row sums (partial)
 *                 allreduce(buf9, 9)
 *                 pdlp_shard_check_decide(buf9, force, buf1)   criteria + restart, ||dy_restart||^2 partial
 *                 allreduce(buf1, 1)
 *                 pdlp_shard_check_end(buf1, out)
 * All buffers are device pointers to doubles owned by the caller.                     */
int pdlp_shard_enable(pdlp_handle h, double q_norm_global, double c_norm);
int pdlp_shard_kty(pdlp_handle h, double* buf);
int pdlp_shard_step(pdlp_handle h, const double* buf, int save_previous);
int pdlp_shard_rule_pack(pdlp_handle h, double* buf4);
int pdlp_shard_rule_apply(pdlp_handle h, const double* buf4);
int pdlp_shard_check_begin(pdlp_handle h, double* buf3, int have_previous);
int pdlp_shard_check_mid(pdlp_handle h, const double* buf3, double* buf9);
int pdlp_shard_check_decide(pdlp_handle h, const double* buf9, int force, double* buf1);
int pdlp_shard_check_end(pdlp_handle h, const double* buf1, pdlp_check* host_out);

/* Peer-memory exchange (NVLink / NVSwitch, one process per GPU on one node): replaces the host-issued allreduce of
 * the iteration by kernels that read and write the peers' buffers directly.  Data travels as 16-byte LL words
 * {lo32, seq, hi32, seq} (value + sequence number of the exchange), so there are no flags or fences to wait for.
 *   pdlp_shard_exchange_create   allocates this rank's exchange block and returns its CUDA IPC handle
 *                                (PDLP_IPC_HANDLE_BYTES bytes; all-gather them over any host transport)
 *   pdlp_shard_exchange_connect  maps the peers' blocks; handles = world x PDLP_IPC_HANDLE_BYTES, rank order
 *   pdlp_shard_iterate_p2p       n_iters iterations of three kernels each: the K_g^T y_g SpMV; ONE kernel that
 *                                stores its segments into the owners' memory, sums the owned segment in rank order,
 *                                stores the sums into every peer, evaluates the step rule and does the primal
 *                                update as the sums arrive; the local dual kernel.  Capturable in a CUDA graph.
 *   pdlp_shard_allreduce_p2p     in-place sum of `count` device doubles over the same block (the small collectives
 *                                of the restart test and of the power iteration)
 *   pdlp_shard_exchange_disconnect  unmaps the peers' blocks; call it on every rank, then a host barrier, before
 *                                pdlp_destroy frees the rank's own block
 * Every rank must enqueue the same calls in the same order; a rank that waits > 20 s for a peer stops and the next
 * pdlp_shard_check_end reports the error.                                                             */
#define PDLP_IPC_HANDLE_BYTES 64
int pdlp_shard_exchange_create(pdlp_handle h, int world, int rank, unsigned char* handle_out);
int pdlp_shard_exchange_connect(pdlp_handle h, const unsigned char* handles);
int pdlp_shard_exchange_disconnect(pdlp_handle h);
int pdlp_shard_iterate_p2p(pdlp_handle h, int n_iters, int save_previous);
int pdlp_shard_allreduce_p2p(pdlp_handle h, double* buf, int64_t count);

/* ---- whole solve --------------------------------------------------------------
 * pdlp_algorithm (primal_dual_hybrid_gradient.py:7-181) driven from C++:
 * restart loop, counters k / n / j with the reference's accounting, KKT-pass and
 * wall-clock limits.  The step size / primal weight / iterate must have been
 * set (pdlp_set_step, pdlp_set_iterate).                                       */
typedef struct pdlp_solve_options {
  int64_t max_kkt;       /* :7  default 100000 */
  int32_t restart_period;/* main.py:132 passes 40 */
  int32_t use_graph;
  double time_limit;     /* seconds, :68-74 */
  double time_used;      /* seconds already spent by the caller (:69) */
  int32_t max_trace;     /* capacity of host_trace (records), 0 = none */
  int32_t reserved;
} pdlp_solve_options;

typedef struct pdlp_solve_result {
  int64_t k, j;
  int32_t n;
  int32_t status;        /* 0 Solved, 1 KKT-pass limit, 2 time limit, 3 PRIMAL_INFEASIBLE, 4 DUAL_INFEASIBLE */
  double prim_obj;
  double seconds;        /* wall clock inside pdlp_solve + time_used */
  double device_ms;      /* CUDA-event time of the whole loop */
  int32_t n_trace;       /* records written to host_trace */
  int32_t reserved;
} pdlp_solve_result;

int pdlp_solve(pdlp_handle h, const pdlp_solve_options* opts, pdlp_solve_result* host_out,
               pdlp_check* host_trace);

#ifdef __cplusplus
}
#endif
#endif /* PDLP_B200_H_ */
