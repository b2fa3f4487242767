/*
 * ehm_oracle.h -- CPU restatement (plain C, fp64) of the partitioning hot path of
 * dmalyuta/explicit_hybrid_mpc.  TEST INFRASTRUCTURE ONLY: the product (the CUDA engine in
 * explicit_hybrid_mpc_b200/csrc) never links, imports or calls anything in oracle/.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PINNING.  The reference delegates its arithmetic to CVXPY 1.0.21 -> MOSEK 9.0.87
 * (requirements.txt:1-6, lib/global_vars.py:25), neither of which exists in this container,
 * and ships no tests or golden vectors; three of its oracle calls are feasibility MICPs
 * (objective 0, lib/oracle.py:46,201,347) whose returned commutation is solver-internal.
 * PARITY WITH MOSEK'S OWN CHOICES IS THEREFORE UNPINNED and this restatement fixes a
 * documented canonical selection rule (DESIGN.md "Canonical oracle semantics").  Everything
 * else is pinned against the reference's own code, imported from /root/reference and run
 * here (generating scripts and fixtures under tests/golden/):
 *   - problem construction (lib/mpc_library.py:410-554) -> fixed-delta optima, MICP / V_R
 *     feasibility in reference form (SciPy HiGHS / SLSQP): pendulum_golden.json
 *   - the eps-suboptimality problems of bar_E / bar_D / in_variability_ball in reference form
 *     (lib/oracle.py:70-102): epigraph_golden.json
 *   - Worker.ecc / Worker.lcss / tools.split_along_longest_edge / tree.py run unmodified over
 *     the six oracle calls: worker_tree_golden*.npz, reproduced node for node
 *   - tools.delaunay / join_triangulation / simplex_volume, examples.create_oracle's eps_a:
 *     tools_golden.npz
 *
 * What is restated (reference file:line):
 *   lib/oracle.py:104-139   P_theta          -> ehmo_p_theta
 *   lib/oracle.py:141-173   P_theta_delta    -> ehmo_p_theta_delta
 *   lib/oracle.py:175-218   V_R              -> ehmo_v_r
 *   lib/oracle.py:220-283   in_variability_ball (inside ehmo_bar_d)
 *   lib/oracle.py:285-309   bar_E_delta_R    -> ehmo_bar_e
 *   lib/oracle.py:311-414   bar_D_delta_R    -> ehmo_bar_d
 *   lib/tools.py:224-257    split_along_longest_edge -> ehmo_split
 *   lib/tools.py:134-150    simplex_volume   -> ehmo_simplex_volume
 *   lib/worker.py:241-417   Worker.ecc / Worker.lcss recursion -> ehmo_partition
 */
#ifndef EHM_ORACLE_H
#define EHM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ehmo_problem ehmo_problem;

/* Canonical tolerances (shared, by specification, with the CUDA engine). */
#define EHMO_TOL_IPM   1e-9   /* primal residual and duality gap s'lam, relative              */
#define EHMO_TAU_INT   1e-7   /* a candidate is alive iff some point has all normalised
                                 slacks >= TAU_INT (phase-1 optimum sigma* <= -TAU_INT)      */
#define EHMO_TAU_CONST 1e-7   /* constant rows (no variable coefficient) may be violated by
                                 at most this much (closed feasibility, slab boundaries)      */
#define EHMO_ZERO_ROW  1e-13  /* |coef| below this (relative to the row's natural scale)
                                 makes the row a constant row                                 */
#define EHMO_TAU_DROP  1e-3   /* row pre-solve margin (normalised rows)                     */
#define EHMO_TOL_RD    1e-3   /* guard on the MEASURED dual residual |Hz+f+G'lam| at convergence, relative to
                                 1+|f|+|lam| (it carries the round-off of the normal equations)     */
#define EHMO_MAX_ITER  80

/* Switched-affine MPC description; array shapes as in
 * explicit_hybrid_mpc_b200/mpc_library.py:HybridMPCProblem (row-major, fp64). */
ehmo_problem *ehmo_create(int n, int nu, int N, int K, int p,
                          const double *A, const double *B, const double *w,
                          const double *C, const double *D, const double *e,
                          const double *Q, const double *P, const double *R,
                          const double *ru, const double *theta_scale,
                          double eps_a, double eps_r);
void ehmo_destroy(ehmo_problem *pb);
void ehmo_set_eps(ehmo_problem *pb, double eps_a, double eps_r);
void ehmo_set_presolve(int on); /* 1 (default): drop rows that cannot bind in the input box */
void ehmo_set_planes(int on);     /* 1 (default): tangent-plane bounds in the partition (result-invariant) */
long ehmo_plane_skips(void);
void ehmo_set_early_exit(int on); /* 1 (default): stop a phase-1 LP at the first certified interior point */
void ehmo_stats(long *out6);      /* phase-1 alive / its, phase-1 dead / its, QP solves / its */
long ehmo_total_iters(void);
int ehmo_num_candidates(const ehmo_problem *pb);
long ehmo_solve_count(const ehmo_problem *pb); /* convex sub-problems solved so far */

/* P_theta_delta: returns 1 if feasible (J, u0[nu] filled), 0 if infeasible, <0 on failure. */
int ehmo_p_theta_delta(ehmo_problem *pb, const double *theta, int cand, double *J, double *u0);
/* P_theta: argmin over all commutations. returns best candidate or -1. */
int ehmo_p_theta(ehmo_problem *pb, const double *theta, double *J, double *u0);
/* V_R: R is (n+1) x n row-major. returns candidate or -1; Jv[n+1], Uv[(n+1)*nu]. */
int ehmo_v_r(ehmo_problem *pb, const double *R, double *Jv, double *Uv);
/* One L-CSS node evaluation (bar_E + bar_D + in_variability_ball, see ehm_oracle.c). */
typedef struct {
    int converged;  /* bar_E: 1 = eps-suboptimal leaf                               */
    int cstar;      /* bar_D: delta* candidate or -1                                */
    int var_small;  /* in_variability_ball                                          */
    double t_ref;   /* t* of delta_ref                                              */
    int n_eval;     /* candidates whose epigraph problem was evaluated              */
    int n_fail;     /* numerical failures (should be 0)                             */
    int n_plane;    /* candidates skipped by their tangent-plane bound              */
} ehmo_lcss_result;
int ehmo_lcss_eval(ehmo_problem *pb, const double *R, const double *Vr, int cand_ref,
                   double *lb, int prune, double *theta_star, double *Jv, double *Uv,
                   ehmo_lcss_result *res);
void ehmo_lb_rebase(int n, int ncand, const double *Vr_old, const double *Vr_new, double *lb);
/* Stand-alone reference-shaped calls (brute force over all candidates).
 * bar_E: returns 1 if eps-suboptimal; tstar (optional, ncand) receives t*_c (NaN = dead). */
int ehmo_bar_e(ehmo_problem *pb, const double *R, const double *Vr, double *tstar);
/* bar_D: returns delta* candidate or -1. theta_star[n], Jv[n+1], Uv[(n+1)*nu], var_small. */
int ehmo_bar_d(ehmo_problem *pb, const double *R, const double *Vr, int cand_ref,
               double *theta_star, double *Jv, double *Uv, int *var_small);

/* min over the simplex of V*_c; returns 0 on success */
int ehmo_min_over_simplex(ehmo_problem *pb, const double *R, int cand, double *val);

/* geometry */
void ehmo_split(int n, const double *R, double *S1, double *S2, int *ij);
double ehmo_simplex_volume(int n, const double *R);

/* Whole-tree partition: nthreads workers over one shared LIFO task queue (every node visit is
 * a task, as in the reference's scheduler/worker farm).
 * roots: nroots x (n+1) x n.  Output node arrays are allocated by the library and returned
 * through ehmo_tree; free with ehmo_tree_free. */
typedef struct {
    long nnodes;
    int n, nu;
    int *left, *right;       /* child node index or -1                                   */
    int *root_id;            /* root simplex this node descends from                     */
    int *depth;              /* depth below its root                                     */
    int *cand;               /* commutation candidate index or -1 (ecc interior nodes)   */
    unsigned char *closed;   /* 1 = eps-suboptimal leaf                                  */
    double *vertices;        /* nnodes x (n+1) x n                                       */
    double *vcost;           /* nnodes x (n+1)                                           */
    double *vinput;          /* nnodes x (n+1) x nu                                      */
    long nleaves, nsolves, nfail;
    int max_depth;
    int error;               /* 1: Theta contains infeasible regions (worker.py:265-266) */
} ehmo_tree;

/* prune: 1 = skip candidates by inherited bounds (same tree, fewer solves), 0 = brute force */
ehmo_tree *ehmo_partition(ehmo_problem *pb, int nroots, const double *roots, int nthreads,
                          long max_nodes, int prune);
void ehmo_tree_free(ehmo_tree *t);

#ifdef __cplusplus
}
#endif
#endif
