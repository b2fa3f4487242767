/*
 * ehm.h -- C-ABI of the B200 explicit-hybrid-MPC partitioning engine (libehm_b200.so).
 *
 * Drop-in boundary for the hot path of dmalyuta/explicit_hybrid_mpc: the per-simplex
 * optimisation oracle (lib/oracle.py), the ECC/L-CSS recursion and task farm
 * (lib/worker.py, lib/scheduler.py), the longest-edge split (lib/tools.py:224-257) and the
 * tree node storage (lib/tree.py).  The reference has no FFI of its own (100% Python over
 * CVXPY->MOSEK and mpi4py); each entry point below names the reference interface it
 * replaces.  Plain C types only: no torch, no C++ in the signatures.  All functions return
 * 0 on success and a negative EHM_E* code on failure; nothing throws.
 *
 * Pointers named d_* are DEVICE pointers (e.g. torch.Tensor.data_ptr() of a CUDA tensor),
 * everything else is a HOST pointer.  `stream` is a cudaStream_t passed as void* (NULL =
 * default stream).  The library never allocates output buffers for the batch calls; the
 * caller owns them.
 */
#ifndef EHM_H
#define EHM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHM_OK            0
#define EHM_EINVAL       -1   /* bad argument / unsupported problem size                     */
#define EHM_ECUDA        -2   /* CUDA runtime error (ehm_last_error() has the text)           */
#define EHM_ENOMEM       -3
#define EHM_EINFEASIBLE  -4   /* "STOP, Theta contains infeasible regions" (worker.py:265-266) */
#define EHM_ECAPACITY    -5   /* node store / frontier capacity exhausted                     */
#define EHM_ENUMERIC     -6   /* midpoint vertex problem failed (worker.py would raise)       */

/* Canonical tolerances; identical, by specification, to oracle/ehm_oracle.h */
#define EHM_TOL_IPM   1e-9
#define EHM_TAU_INT   1e-7
#define EHM_TAU_CONST 1e-7
#define EHM_ZERO_ROW  1e-13
#define EHM_TOL_RD    1e-3   /* guard on the measured dual residual at convergence (ehm_oracle.h) */
#define EHM_MAX_ITER  80

typedef struct ehm_problem ehm_problem;

/* Switched-affine MPC description (host arrays, row-major fp64); see
 * explicit_hybrid_mpc_b200/mpc_library.py:HybridMPCProblem.  Replaces the CVXPY problem
 * templates built by Oracle.__init__ (lib/oracle.py:23-102) from the MPC object
 * (lib/mpc_library.py:474-554). */
typedef struct {
    int n, nu, N, K, p;
    const double *A, *B, *w;        /* (K,n,n) (K,n,nu) (K,n)              */
    const double *C, *D, *e;        /* (K,p,n) (K,p,nu) (K,p)              */
    const double *Q, *P, *R, *ru;   /* (n,n) (n,n) (nu,nu) (K,nu)          */
    const double *theta_scale;      /* (n,) natural magnitude of theta     */
    double eps_a, eps_r;            /* Oracle(mpc,eps_a,eps_r)             */
} ehm_problem_desc;

int ehm_problem_create(const ehm_problem_desc *desc, ehm_problem **out);
int ehm_problem_destroy(ehm_problem *p);
int ehm_set_eps(ehm_problem *p, double eps_a, double eps_r);
int ehm_num_candidates(const ehm_problem *p);     /* K^N one-hot commutations */
const char *ehm_last_error(void);
/* number of hand-written kernels launched by this library since load (bench gpu_launches) */
int64_t ehm_kernel_launches(void);

/* ---- batched oracle sub-problems (one warp per problem) ------------------------------ */
/* P_theta_delta for B (theta, candidate) pairs  [lib/oracle.py:141-173].
 * d_status[b]: 1 feasible & converged (d_J, d_u0 valid), 0 infeasible, -1 numerical failure.
 * d_u0 is the scaled first input uhat_0 (multiply by u_scale on the host). */
int ehm_qp_theta_delta_batch(ehm_problem *p, int B, const double *d_theta, const int *d_cand,
                             double *d_J, double *d_u0, int *d_status, void *stream);

/* Simplex problems over R (B,(n+1),n) for B (simplex, candidate) pairs.
 *  mode EHM_SX_ALIVE : phase-1 only                  -> d_val = sigma*
 *  mode EHM_SX_TQP   : eps-suboptimality epigraph QP -> d_val = t*_c, d_sol = [beta(n), V*]
 *                      (constraints lib/oracle.py:88-97; needs d_Vr (B,n+1))
 *  mode EHM_SX_MINV  : min_{theta in R} V*_c(theta)  -> d_val, d_sol = [beta(n), V*]
 *                      (lib/oracle.py:75-79)
 * with_phase1 != 0 runs the alive test first for TQP/MINV.
 * d_status[b]: 1 ok, 0 not alive / infeasible, -1 numerical failure. */
#define EHM_SX_ALIVE 0
#define EHM_SX_TQP   1
#define EHM_SX_MINV  2
int ehm_simplex_batch(ehm_problem *p, int B, const double *d_R, const double *d_Vr,
                      const int *d_cand, int mode, int with_phase1, double *d_val,
                      double *d_sol, int *d_status, void *stream);

/* ---- the reference's set-valued oracle calls, batched ----------------------------------
 * MOSEK's branch and bound over the binary commutation is replaced by enumeration of all K^N
 * one-hot commutations ("candidates") on the device, followed by device-side selection with the
 * canonical rules of DESIGN.md.  All arrays are DEVICE pointers owned by the caller; scratch is
 * kept (grow-only) inside the problem object.  Candidate index <-> commutation vector:
 * base-K digits, step 0 most significant (mpc_library.HybridMPCProblem.commutation_from_index).
 *
 * Oracle.P_theta  [lib/oracle.py:104-139]: argmin over the commutations of the fixed-delta optimal
 * cost at each of B parameters (ties -> smallest index).  d_cand[b] = -1 / d_status[b] = 0 when no
 * commutation is feasible (the reference's check_feasibility=True answer is d_status).
 * d_J, d_u0 (B,nu scaled first input), d_status may be NULL. */
int ehm_p_theta_batch(ehm_problem *p, int B, const double *d_theta, int *d_cand, double *d_J,
                      double *d_u0, int *d_status, void *stream);
/* Oracle.V_R  [lib/oracle.py:175-218]: for each of B simplices d_R (B,n+1,n) the commutation that is
 * feasible at all n+1 vertices with the smallest sum of vertex costs (-1: none; the reference then
 * returns (None, None)), with its vertex costs d_Jv (B,n+1) and scaled inputs d_Uv (B,n+1,nu). */
int ehm_v_r_batch(ehm_problem *p, int B, const double *d_R, int *d_cand, double *d_Jv, double *d_Uv,
                  int *d_status, void *stream);
/* Oracle.bar_E_delta_R  [lib/oracle.py:285-309]: d_converged[b] = 1 iff no commutation satisfies the
 * two eps-improvement inequalities (:94-95) anywhere in simplex b, i.e. min_c t*_c > 0.
 * d_tmin (optional): that minimum (+inf if no commutation is alive on the simplex). */
int ehm_bar_e_batch(ehm_problem *p, int B, const double *d_R, const double *d_Vr, int *d_converged,
                    double *d_tmin, void *stream);
/* Oracle.bar_D_delta_R + in_variability_ball  [lib/oracle.py:311-414, 220-283]: delta* =
 * argmin t*_c over the commutations that beat delta_ref (t*_c <= 0, t*_c < t*_ref) and are feasible
 * at all vertices; -1 = the reference's (None, None, None, None).  Outputs for delta* != -1:
 * d_theta_star (B,n) minimiser of its epigraph problem, d_Jv (B,n+1) / d_Uv (B,n+1,nu) its vertex
 * costs / scaled inputs, d_var_small (B) the in_variability_ball verdict.  d_status (optional):
 * 1 ok, -1 the min-over-simplex problem of delta_ref failed numerically (answer -1). */
int ehm_bar_d_batch(ehm_problem *p, int B, const double *d_R, const double *d_Vr,
                    const int *d_cand_ref, int *d_cand_star, double *d_theta_star, double *d_Jv,
                    double *d_Uv, int *d_var_small, int *d_status, void *stream);

/* Longest-edge bisection of B simplices  [lib/tools.py:224-257].  d_ij (B,2) int32; n <= EHM_SPLIT_MAXN. */
#define EHM_SPLIT_MAXN 16
int ehm_split_batch(int n, int B, const double *d_R, double *d_S1, double *d_S2, int *d_ij,
                    void *stream);

/* ---- the fused partition loop -------------------------------------------------------- */
/* Replaces scheduler.main()/Worker.spin()/Worker.ecc()/Worker.lcss() and build_tree()
 * (lib/scheduler.py:364-691, lib/worker.py:164-461): a device-resident LIFO frontier of
 * open simplices is popped `wave` nodes at a time; every node is run through the oracle
 * kernels, closed or split, and its children are pushed back.  Results land in a flat node
 * store (the device analogue of lib/tree.py) which is copied to the caller's HOST buffers. */
typedef struct {
    int wave;              /* nodes popped per wave (0 = default)                          */
    int64_t max_nodes;     /* node-store capacity (0 = default)                            */
    int64_t max_frontier;  /* frontier capacity in entries (0 = default)                   */
    int64_t node_budget;   /* stop (cleanly) once this many nodes exist; 0 = run to the end */
    int verbose;
} ehm_partition_opts;

typedef struct {
    int64_t nnodes, nleaves, nsolves, niters, nfail, nwaves, nlaunches;
    int64_t open_left;     /* open work left: frontier entries + root simplices still queued */
    int64_t frontier;      /* of these, the entries ehm_partition_export can hand over      */
    int max_depth;
    double device_ms;      /* CUDA-event time of the whole loop                            */
    /* per pipeline stage, summed over waves (CUDA events on the launching stream):
     * 0 pop, 1 lcss_ref, 2 lcss_items, 3 lcss_eval, 4 lcss_select, 5 lcss_points,
     * 6 lcss_final, 7 the ECC phase */
    double stage_ms[8];
    double stage_flops[8]; /* algorithmic fp64 flops of the solve stages (1,3,5,7)         */
    double stage_bytes[8]; /* algorithmic HBM bytes of the streaming stages (0,2,6)        */
} ehm_partition_stats;

typedef struct ehm_partition ehm_partition;

/* Resumable form: create the node store / frontier with the root simplices (host
 * (nroots,(n+1),n); nroots may be 0 for a rank that only receives work), then advance by at
 * most max_waves waves (<= 0: until the frontier is empty or node_budget is hit).  stats are
 * cumulative since creation; stats.open_left > 0 means there is work left. */
int ehm_partition_create(ehm_problem *p, int nroots, const double *roots,
                         const ehm_partition_opts *opts, ehm_partition **out);
/* optional, right after create: global ids of the root simplices (default 0..nroots-1); they
 * are inherited by all descendants and travel with exported records */
int ehm_partition_set_root_ids(ehm_partition *t, int nroots, const int32_t *ids);
int ehm_partition_advance(ehm_partition *t, int64_t max_waves, ehm_partition_stats *stats);
/* the same with a time budget: returns after the first wave that ends more than budget_ms after
 * the call (<= 0: no budget).  The multi-GPU driver uses it so that all ranks reach the
 * depth exchange at about the same time (the reference's scheduler polls its workers every
 * loop turn, lib/scheduler.py:425-441); the tree does not depend on where rounds end. */
int ehm_partition_advance_for(ehm_partition *t, int64_t max_waves, double budget_ms,
                              ehm_partition_stats *stats);
/* Frontier exchange between ranks (replaces the worker->scheduler->worker task hand-over,
 * lib/worker.py:232-236, lib/scheduler.py:515-523).  export pops the k OLDEST entries into
 * caller-owned DEVICE buffers: d_meta int32[k][6] {node, cand, depth, root_id, skip, 0},
 * d_geom f64[k][(n+1)n+(n+1)+(n+1)nu] (vertices, vertex costs, vertex inputs), d_lb
 * f32[k][ncand]; the exported nodes are marked closed=2.  import appends k received records as
 * new nodes (origin = (d_origin_rank[e], meta node id)) and pushes them on the frontier. */
int ehm_partition_export(ehm_partition *t, int k, int32_t *d_meta, double *d_geom, float *d_lb);
int ehm_partition_import(ehm_partition *t, int k, const int32_t *d_meta, const double *d_geom,
                         const float *d_lb, const int32_t *d_origin_rank);
/* origin int32[nnodes][2] (rank, node) of imported nodes, -1 for local ones (HOST buffer) */
int ehm_partition_fetch_origin(ehm_partition *t, int32_t *origin);
/* One-shot form = create + advance(0).
 * roots: host (nroots,(n+1),n).  Runs until the frontier is empty (or node_budget). */
int ehm_partition_run(ehm_problem *p, int nroots, const double *roots,
                      const ehm_partition_opts *opts, ehm_partition **out,
                      ehm_partition_stats *stats);
/* copy the node store to HOST buffers sized from stats.nnodes:
 * left/right/root_id/depth/cand int32[nnodes], closed uint8[nnodes],
 * vertices f64[nnodes,(n+1),n], vcost f64[nnodes,n+1], vinput f64[nnodes,(n+1),nu] (scaled) */
int ehm_partition_fetch(ehm_partition *t, int32_t *left, int32_t *right, int32_t *root_id,
                        int32_t *depth, int32_t *cand, uint8_t *closed, double *vertices,
                        double *vcost, double *vinput);
int ehm_partition_free(ehm_partition *t);

/* ---- measurement helpers (bench.py) --------------------------------------------------- */
/* counters since problem creation */
int ehm_read_counters(ehm_problem *p, int64_t *iters, int64_t *solves, int64_t *fails);
/* {pipeline stage, candidate, node id, iterations} of the most recent numerical failure */
int ehm_last_failure(ehm_problem *p, int32_t *info4);
/* debugging aid: record the per-iteration trace {it, rp, gap, obj, rho*rd0, |lam|, rows, nv} of
 * the solves that run while enabled (last one wins); fetch copies 128 x 8 doubles */
int ehm_debug_trace(int on);
int ehm_debug_fetch(double *out);
/* dependent-free DFMA loop on every SM: the fp64 FMA roofline denominator, in TFLOP/s */
int ehm_fp64_peak(double *tflops);
/* STREAM-style device copy of `bytes` bytes: GB/s (read+write counted) */
int ehm_hbm_peak(int64_t bytes, double *gbs);

/* ---- online evaluation of the explicit control law --------------------------------------
 * ExplicitMPC (lib/mpc_library.py:662-792): compute_simplex_basis_inverse (:684-712) =
 * ehm_locator_create (Minv of every node, on the device); get_containing_cell +
 * __call__ (:738-792) = ehm_locator_eval for a batch of B parameters.  The tree is the FULL
 * tree (triangulation spine included) as flat HOST arrays: left/right child ids (-1 = leaf),
 * vertices (nnodes,(n+1),n), vertex inputs (nnodes,(n+1),nu); `top` = id of the tree root.
 * d_theta (B,n), d_u (B,nu), d_leaf (B, optional) are DEVICE pointers. */
typedef struct ehm_locator ehm_locator;
int ehm_locator_create(int n, int nu, int64_t nnodes, const int32_t *left, const int32_t *right,
                       const double *vertices, const double *vinput, int32_t top, ehm_locator **out);
int ehm_locator_eval(ehm_locator *l, int64_t B, const double *d_theta, double *d_u, int32_t *d_leaf,
                     void *stream);
int ehm_locator_destroy(ehm_locator *l);

#ifdef __cplusplus
}
#endif
#endif
