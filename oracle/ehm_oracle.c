/*
 * ehm_oracle.c -- CPU restatement (plain C, fp64) of the explicit hybrid MPC partitioning
 * oracle + ECC/L-CSS recursion.  See ehm_oracle.h for the scope, the list of restated
 * reference functions (file:line) and the "parity unpinned" statement.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product.
 *
 * Algorithm in one paragraph.  The reference's MICPs range over binary commutations delta with
 * sum_i z[k][i]==1 (lib/mpc_library.py:526-527), i.e. over K^N one-hot mode sequences
 * ("candidates").  For a fixed candidate the active big-M pair (:530-534) is an equality, so
 * the states are affine in (theta,u) and every oracle problem condenses to a small dense
 * convex QP/LP in u (and, for the simplex problems of lib/oracle.py:70-102, in the
 * barycentric coordinates beta of theta and one epigraph variable).  MOSEK's branch & bound
 * is replaced by explicit enumeration with the canonical selection rules written next to each
 * function below.  Each convex sub-problem is solved by a dense Mehrotra predictor-corrector
 * interior-point method after an always-feasible "phase-1" LP decided feasibility.
 */
#include "ehm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define MAXNV 40

static long g_iters = 0;
static double g_rdmax_seen = 0; /* diagnostic: largest scaled measured dual residual at convergence */
double ehmo_rdmax_seen(void) { return g_rdmax_seen; }
/* Counters are accumulated per thread and flushed once per task / API call: a shared counter
 * bumped in every IPM iteration turns into a cache line bouncing between all cores (it capped
 * the 128-thread host of the GPU box at the throughput of 8 cores). */
static __thread long t_iters = 0, t_solves = 0, t_stat[6] = {0, 0, 0, 0, 0, 0};
static int env_flag(const char *name, int *cache)
{
    if (*cache < 0) *cache = getenv(name) != NULL;
    return *cache;
}
static int dbg_debug = -1, dbg_failtrace = -1, dbg_trace = -1;
long ehmo_total_iters(void) { return g_iters + t_iters; }
/* statistics of the decision solves (debugging aid, not thread-exact): phase-1 solves that
 * ended alive / dead and their iterations, QP solves and their iterations */
static long g_stat[6];
void ehmo_stats(long *out) { for (int i = 0; i < 6; ++i) out[i] = g_stat[i] + t_stat[i]; }
/* Early exit of the phase-1 LP (0 disables): the elastic LP starts strictly feasible, so every
 * iterate (z_k, sigma_k) is a feasible point of it; as soon as z_k leaves all (normalised)
 * slacks of the ORIGINAL rows >= TAU_INT the answer "alive" is certified (sigma* <= -TAU_INT)
 * and the remaining iterations cannot change it. */
/* optional output of the last qp_solve of this thread: multipliers of the rows (tangent planes) */
static __thread double *t_lam_out = NULL;
static int g_planes = 2; /* 0 off, 1 inherited planes, 2 local planes (what the CUDA engine does) */
static __thread double *t_planes = NULL;
static long g_plane_skips = 0;
long ehmo_plane_skips(void) { return g_plane_skips; }
void ehmo_set_planes(int on) { g_planes = on; }
static int g_early = 1;
void ehmo_set_early_exit(int on) { g_early = on; }
static inline double fmax2(double a, double b) { return a > b ? a : b; }

struct ehmo_problem {
    int n, nu, N, K, p;
    int U, m, ncand, Y;          /* U = N*nu inputs, m = N*p rows, Y = U+n+1 */
    double *A, *B, *w, *C, *D, *e, *Q, *P, *R, *ru, *tscale;
    double eps_a, eps_r;
    double *cG, *cS, *cg, *cM, *crho; /* condensed data of every candidate */
    long nsolves;
};

/* ------------------------------------------------------------------------------------ */
/* small dense helpers                                                                  */
/* ------------------------------------------------------------------------------------ */
static double *dalloc(size_t k) { return (double *)calloc(k ? k : 1, sizeof(double)); }

static double *dcopy(const double *src, size_t k)
{
    double *d = dalloc(k);
    memcpy(d, src, k * sizeof(double));
    return d;
}

/* in-place lower Cholesky of the nv x nv SPD matrix K (row-major).  A pivot that collapses
 * below 1e-13 of its own original diagonal entry (rank-deficient normal matrix close to an LP
 * vertex) is replaced by a huge one, which freezes that direction for this Newton step
 * (the classical LIPSOL safeguard).  returns 0 on success, 1 on NaN/Inf. */
static int cholesky(int nv, double *K)
{
    double maxd = 0;
    for (int j = 0; j < nv; ++j) maxd = fmax2(maxd, K[j * nv + j]);
    if (!isfinite(maxd) || !(maxd > 0.0)) return 1;
    for (int j = 0; j < nv; ++j) {
        const double d0 = K[j * nv + j];
        double d = d0;
        for (int k = 0; k < j; ++k) d -= K[j * nv + k] * K[j * nv + k];
        if (!isfinite(d)) return 1;
        if (d <= 1e-13 * d0 || d0 <= 0.0) d = 1e100;
        d = sqrt(d);
        K[j * nv + j] = d;
        for (int i = j + 1; i < nv; ++i) {
            double v = K[i * nv + j];
            for (int k = 0; k < j; ++k) v -= K[i * nv + k] * K[j * nv + k];
            K[i * nv + j] = v / d;
        }
    }
    return 0;
}

static void chol_solve(int nv, const double *L, double *b)
{
    for (int i = 0; i < nv; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= L[i * nv + k] * b[k];
        b[i] = v / L[i * nv + i];
    }
    for (int i = nv - 1; i >= 0; --i) {
        double v = b[i];
        for (int k = i + 1; k < nv; ++k) v -= L[k * nv + i] * b[k];
        b[i] = v / L[i * nv + i];
    }
}

/* ------------------------------------------------------------------------------------ */
/* Dense Mehrotra predictor-corrector for  min 0.5 z'Hz + f'z  s.t.  G z <= h           */
/* (what CVXPY->MOSEK does for the reference at every .solve(), lib/oracle.py:131-378)   */
/* z0 != NULL gives the primal start (slacks max(h-Gz0,1)).  returns 0 converged,        */
/* 1 iteration limit, 2 factorisation breakdown.                                         */
/* ------------------------------------------------------------------------------------ */
static int qp_solve_x(int nv, int m, const double *H, const double *f, const double *G,
                      const double *h, const double *z0, double *z, double *obj_out,
                      int *iters_out, int early);
static int qp_solve(int nv, int m, const double *H, const double *f, const double *G,
                    const double *h, const double *z0, double *z, double *obj_out,
                    int *iters_out)
{
    return qp_solve_x(nv, m, H, f, G, h, z0, z, obj_out, iters_out, 0);
}

/* early != 0: the problem is a phase-1 LP whose last variable is sigma with coefficient -1 in
 * every row; stop as soon as the current z certifies sigma* <= -TAU_INT. */
static int qp_solve_x(int nv, int m, const double *H, const double *f, const double *G,
                      const double *h, const double *z0, double *z, double *obj_out,
                      int *iters_out, int early)
{
    double *s = dalloc(m), *lam = dalloc(m), *rp = dalloc(m), *ds = dalloc(m),
           *dl = dalloc(m), *rc = dalloc(m), *wv = dalloc(m);
    double K[MAXNV * MAXNV], rd[MAXNV], dz[MAXNV], rhs[MAXNV];
    int status = 1, it = 0;
    double rho = 1.0, rd0 = 0.0; /* exact-arithmetic dual residual = rho * rd0 (see below) */
    double fmax = 0, obj = 0;
    for (int j = 0; j < nv; ++j) fmax = fmax2(fmax, fabs(f[j]));
    for (int j = 0; j < nv; ++j) z[j] = z0 ? z0[j] : 0.0;
    for (int r = 0; r < m; ++r) {
        double gz = 0;
        for (int j = 0; j < nv; ++j) gz += G[r * nv + j] * z[j];
        s[r] = fmax2(h[r] - gz, 1.0);
        lam[r] = 1.0 / s[r]; /* centred start: s.lam = 1 on every row */
    }
    for (it = 0; it <= EHMO_MAX_ITER; ++it) {
        double rpmax = 0, rdmax = 0, mu = 0;
        obj = 0;
        for (int j = 0; j < nv; ++j) {
            double v = f[j], hz = 0;
            if (H)
                for (int k = 0; k < nv; ++k) hz += H[j * nv + k] * z[k];
            obj += z[j] * (f[j] + 0.5 * hz);
            rd[j] = v + hz;
        }
        int cert = early;
        for (int r = 0; r < m; ++r) {
            double gz = 0;
            for (int j = 0; j < nv; ++j) {
                gz += G[r * nv + j] * z[j];
                rd[j] += G[r * nv + j] * lam[r];
            }
            /* slack of the original row (without the sigma column) at z_k */
            if (cert && !(h[r] - gz - z[nv - 1] >= EHMO_TAU_INT)) cert = 0;
            rp[r] = gz + s[r] - h[r];
            rpmax = fmax2(rpmax, fabs(rp[r]));
            mu += s[r] * lam[r];
        }
        for (int j = 0; j < nv; ++j) rdmax = fmax2(rdmax, fabs(rd[j]));
        double zmax = 0, lmax = 0;
        for (int j = 0; j < nv; ++j) zmax = fmax2(zmax, fabs(z[j]));
        for (int r = 0; r < m; ++r) lmax = fmax2(lmax, lam[r]);
        if (!isfinite(mu) || !isfinite(rdmax) || !isfinite(rpmax)) { status = 2; break; }
        if (env_flag("EHMO_DEBUG", &dbg_debug)) fprintf(stderr, "it %d rp %.3e rd %.3e mu %.3e obj %.10g lmax %.3e zmax %.3e\n", it, rpmax, rdmax, mu, obj, lmax, zmax);
        /* residuals relative to the magnitude of the terms they are made of (rows are
         * normalised, |G| <= 1): |Gz| <= |z|, |G'lam| <= |lam| */
        if (it == 0) rd0 = rdmax;
        if (cert) {
            obj = -EHMO_TAU_INT; /* certified: sigma* <= -TAU_INT */
            status = 0;
            break;
        }
        /* Every step scales the dual residual by exactly (1 - alpha (1 - sigma)); the value
         * measured above additionally carries the round-off of the normal equations (up to
         * ~1e-4 relative once lam/s spans 20+ orders), so convergence is tested on the
         * propagated value rho*rd0; the measured one is held to the loose guard TOL_RD only (a
         * point whose measured residual is larger than that is not accepted as converged). */
        if (rdmax > g_rdmax_seen * (1 + fmax + lmax) && rpmax <= EHMO_TOL_IPM * (1 + zmax) &&
            rho * rd0 <= EHMO_TOL_IPM * (1 + fmax + lmax) && mu <= EHMO_TOL_IPM * (1 + fabs(obj)))
            g_rdmax_seen = rdmax / (1 + fmax + lmax); /* diagnostic only (racy across threads) */
        if (rpmax <= EHMO_TOL_IPM * (1 + zmax) && rho * rd0 <= EHMO_TOL_IPM * (1 + fmax + lmax) &&
            mu <= EHMO_TOL_IPM * (1 + fabs(obj)) && rdmax <= EHMO_TOL_RD * (1 + fmax + lmax)) {
            status = 0;
            break;
        }
        if (it == EHMO_MAX_ITER) break;
        mu /= m;
        /* K = H + G' diag(lam/s) G */
        for (int j = 0; j < nv * nv; ++j) K[j] = H ? H[j] : 0.0;
        for (int r = 0; r < m; ++r) {
            double wr = lam[r] / s[r];
            wv[r] = wr;
            const double *g = G + r * nv;
            for (int j = 0; j < nv; ++j) {
                double gj = wr * g[j];
                if (gj == 0.0) continue;
                for (int k = 0; k <= j; ++k) K[j * nv + k] += gj * g[k];
            }
        }
        for (int j = 0; j < nv; ++j)
            for (int k = j + 1; k < nv; ++k) K[j * nv + k] = K[k * nv + j];
        if (cholesky(nv, K)) { status = 2; break; }
        double alpha = 1.0, resc = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            /* pass 0: affine predictor (rc = s.lam); pass 1: centring + second-order corrector
             * with sigma = (1-alpha_aff)^3 and the residuals scaled by (1-sigma), so that
             * infeasibility and complementarity shrink at the same rate (the CVXOPT/ECOS form
             * of Mehrotra's method; the plain form cycles on thin feasible sets). */
            if (pass == 0) {
                for (int r = 0; r < m; ++r) rc[r] = s[r] * lam[r];
            } else {
                double sig = (1.0 - alpha) * (1.0 - alpha) * (1.0 - alpha);
                resc = 1.0 - sig;
                for (int r = 0; r < m; ++r) rc[r] = s[r] * lam[r] + ds[r] * dl[r] - sig * mu;
            }
            for (int j = 0; j < nv; ++j) rhs[j] = -resc * rd[j];
            for (int r = 0; r < m; ++r) {
                double t = (-rc[r] + lam[r] * resc * rp[r]) / s[r];
                const double *g = G + r * nv;
                for (int j = 0; j < nv; ++j) rhs[j] -= g[j] * t;
            }
            chol_solve(nv, K, rhs);
            memcpy(dz, rhs, nv * sizeof(double));
            double amax = 1e300;
            for (int r = 0; r < m; ++r) {
                double gdz = 0;
                const double *g = G + r * nv;
                for (int j = 0; j < nv; ++j) gdz += g[j] * dz[j];
                ds[r] = -resc * rp[r] - gdz;
                dl[r] = (-rc[r] - lam[r] * ds[r]) / s[r];
                if (ds[r] < 0) amax = fmin(amax, -s[r] / ds[r]);
                if (dl[r] < 0) amax = fmin(amax, -lam[r] / dl[r]);
            }
            alpha = (pass == 0) ? fmin(1.0, amax) : fmin(1.0, 0.99 * amax);
        }
        {
            static double gam = -1;
            if (gam < 0) gam = getenv("EHMO_GAMMA") ? atof(getenv("EHMO_GAMMA")) : 0.0;
            if (gam > 0) {
                for (int tries = 0; tries < 20; ++tries) {
                    double mun = 0, mn = 1e300;
                    for (int r = 0; r < m; ++r) {
                        double pr = (s[r] + alpha * ds[r]) * (lam[r] + alpha * dl[r]);
                        mun += pr;
                        mn = fmin(mn, pr);
                    }
                    if (mn >= gam * mun / m) break;
                    alpha *= 0.8;
                }
            }
        }
        t_iters++;
        rho *= 1.0 - alpha * resc;
        if (env_flag("EHMO_DEBUG", &dbg_debug)) {
            double mn = 1e300, mx = 0; int imn = -1;
            for (int r = 0; r < m; ++r) { double pr = s[r] * lam[r]; if (pr < mn) { mn = pr; imn = r; } if (pr > mx) mx = pr; }
            fprintf(stderr, "    alpha %.3e  min s.lam %.3e (row %d s %.3e lam %.3e) max %.3e mu %.3e\n", alpha, mn, imn, s[imn], lam[imn], mx, mu);
        }
        for (int j = 0; j < nv; ++j) z[j] += alpha * dz[j];
        for (int r = 0; r < m; ++r) {
            s[r] += alpha * ds[r];
            lam[r] += alpha * dl[r];
        }
    }
    *obj_out = obj;
    if (t_lam_out) memcpy(t_lam_out, lam, (size_t)m * sizeof(double));
    if (status != 0 && env_flag("EHMO_FAILTRACE", &dbg_failtrace)) {
        double rpm = 0, rdm = 0, mu = 0;
        for (int r = 0; r < m; ++r) { rpm = fmax2(rpm, fabs(rp[r])); mu += s[r] * lam[r]; }
        for (int j = 0; j < nv; ++j) rdm = fmax2(rdm, fabs(rd[j]));
        fprintf(stderr, "FAIL status %d nv %d m %d it %d rp %.3e rd %.3e mu %.3e obj %.6g hmax %.3e fmax %.3e\n", status, nv, m, it, rpm, rdm, mu / m, obj, 0.0, fmax);
    }
    if (iters_out) *iters_out = it;
    free(s); free(lam); free(rp); free(ds); free(dl); free(rc); free(wv);
    return status;
}

/* ------------------------------------------------------------------------------------ */
/* Row preparation: drop constant rows (checked, closed tolerance), normalise the rest.   */
/* rho[r] is the row's natural scale in (u,theta) space (used only to make the            */
/* constant-row test scale-free).  returns number of kept rows, or -1 if a constant row   */
/* is violated (problem infeasible).                                                      */
/* ------------------------------------------------------------------------------------ */
static int g_presolve = 1;
void ehmo_set_presolve(int on) { g_presolve = on; }

/* nbox leading variables live in [-1,1] (the scaled inputs: the input-bound rows of the MPC are
 * part of every problem), the next nsx are barycentric coordinates of a simplex.  Pre-solve:
 * a normalised row a'z <= h whose left-hand side cannot exceed h - TAU_DROP anywhere in that
 * box is dropped -- exact for every decision (see DESIGN.md "Row pre-solve"); it removes the
 * ~75% of the reference's big-M rows (lib/mpc_library.py:529-534) that can never bind. */
static __thread int t_mstage = 0, t_stage_kept = 0;
static int prep_rows(int nv, int m, const double *G, const double *h, const double *rho,
                     double *Go, double *ho, int nbox, int nsx)
{
    int mo = 0;
    t_stage_kept = 0;
    for (int r = 0; r < m; ++r) {
        if (r == t_mstage) t_stage_kept = mo;
        double nrm = 0;
        for (int j = 0; j < nv; ++j) nrm = fmax2(nrm, fabs(G[r * nv + j]));
        double sc = rho[r] > 0 ? rho[r] : 1.0;
        if (nrm <= EHMO_ZERO_ROW * sc) {
            if (h[r] < -EHMO_TAU_CONST * sc) return -1;
            continue;
        }
        double inv = 1.0 / nrm;
        if (g_presolve) {
            double sup = 0, bmax = 0;
            int other = 0;
            for (int j = 0; j < nv; ++j) {
                double a = G[r * nv + j] * inv;
                if (j < nbox) sup += fabs(a);
                else if (j < nbox + nsx) bmax = fmax2(bmax, a);
                else if (a != 0.0) other = 1;
            }
            if (!other && sup + bmax <= h[r] * inv - EHMO_TAU_DROP) continue;
        }
        for (int j = 0; j < nv; ++j) Go[mo * nv + j] = G[r * nv + j] * inv;
        ho[mo] = h[r] * inv;
        ++mo;
    }
    if (t_mstage >= m) t_stage_kept = mo;
    return mo;
}

/* phase 1:  min sigma  s.t.  G z - sigma <= h  (rows already normalised).
 * returns 1 alive (sigma* <= -TAU_INT), 0 not alive, <0 solver failure. */
static int phase1(ehmo_problem *pb, int nv, int m, const double *G, const double *h,
                  double *sigma_out)
{
    (void)pb;
    int na = nv + 1;
    double *Ga = dalloc((size_t)m * na);
    double f[MAXNV], z0[MAXNV], z[MAXNV], obj;
    double s0 = 0;
    for (int r = 0; r < m; ++r) {
        memcpy(Ga + (size_t)r * na, G + (size_t)r * nv, nv * sizeof(double));
        Ga[(size_t)r * na + nv] = -1.0;
        s0 = fmax2(s0, -h[r]);
    }
    for (int j = 0; j < nv; ++j) { f[j] = 0; z0[j] = 0; }
    f[nv] = 1.0;
    z0[nv] = s0 + 1.0; /* strictly feasible start: slacks h+sigma >= 1 */
    int its = 0;
    int st = qp_solve_x(na, m, NULL, f, Ga, h, z0, z, &obj, &its, sigma_out ? 0 : g_early);
    t_solves++;
    free(Ga);
    if (sigma_out) *sigma_out = obj;
    if (st != 0) return -1;
    const int alive = obj <= -EHMO_TAU_INT ? 1 : 0;
    t_stat[alive ? 0 : 2]++; t_stat[alive ? 1 : 3] += its;
    return alive;
}

/* ------------------------------------------------------------------------------------ */
/* Condensation of one candidate (mode sequence)                                          */
/* ------------------------------------------------------------------------------------ */
static void seq_of(const ehmo_problem *pb, int c, int *seq)
{
    for (int k = pb->N - 1; k >= 0; --k) { seq[k] = c % pb->K; c /= pb->K; }
}

static void condense(ehmo_problem *pb, int c)
{
    const int n = pb->n, nu = pb->nu, N = pb->N, p = pb->p, U = pb->U, Y = pb->Y, m = pb->m;
    int seq[64];
    seq_of(pb, c, seq);
    double *G = pb->cG + (size_t)c * m * U, *S = pb->cS + (size_t)c * m * n,
           *g = pb->cg + (size_t)c * m, *M = pb->cM + (size_t)c * Y * Y,
           *rho = pb->crho + (size_t)c * m;
    /* T (n x Y): x_k = T [u; theta; 1] */
    double *T = dalloc((size_t)n * Y), *Tn = dalloc((size_t)n * Y), *WT = dalloc((size_t)n * Y);
    for (int i = 0; i < n; ++i) T[i * Y + U + i] = 1.0;
    memset(M, 0, sizeof(double) * Y * Y);
    for (int k = 0; k < N; ++k) {
        const int i = seq[k];
        const double *Ci = pb->C + (size_t)i * p * n, *Di = pb->D + (size_t)i * p * nu,
                     *ei = pb->e + (size_t)i * p, *Ai = pb->A + (size_t)i * n * n,
                     *Bi = pb->B + (size_t)i * n * nu, *wi = pb->w + (size_t)i * n;
        for (int r = 0; r < p; ++r) {
            const int row = k * p + r;
            double *Gr = G + (size_t)row * U, *Sr = S + (size_t)row * n;
            for (int j = 0; j < U; ++j) {
                double v = 0;
                for (int a = 0; a < n; ++a) v += Ci[r * n + a] * T[a * Y + j];
                Gr[j] = v;
            }
            for (int j = 0; j < nu; ++j) Gr[k * nu + j] += Di[r * nu + j];
            double gc = ei[r], nat = 0;
            for (int j = 0; j < n; ++j) {
                double v = 0;
                for (int a = 0; a < n; ++a) v += Ci[r * n + a] * T[a * Y + U + j];
                Sr[j] = -v;
                nat = fmax2(nat, fabs(v) * pb->tscale[j]);
            }
            for (int a = 0; a < n; ++a) gc -= Ci[r * n + a] * T[a * Y + Y - 1];
            g[row] = gc;
            for (int j = 0; j < U; ++j) nat = fmax2(nat, fabs(Gr[j]));
            rho[row] = nat;
        }
        /* input cost u_k' R u_k + ru' u_k */
        for (int a = 0; a < nu; ++a) {
            for (int b = 0; b < nu; ++b) M[(k * nu + a) * Y + k * nu + b] += pb->R[a * nu + b];
            M[(k * nu + a) * Y + Y - 1] += 0.5 * pb->ru[i * nu + a];
            M[(Y - 1) * Y + k * nu + a] += 0.5 * pb->ru[i * nu + a];
        }
        /* x_{k+1} = A x_k + B u_k + w */
        for (int a = 0; a < n; ++a)
            for (int j = 0; j < Y; ++j) {
                double v = 0;
                for (int b = 0; b < n; ++b) v += Ai[a * n + b] * T[b * Y + j];
                Tn[a * Y + j] = v;
            }
        for (int a = 0; a < n; ++a) {
            for (int j = 0; j < nu; ++j) Tn[a * Y + k * nu + j] += Bi[a * nu + j];
            Tn[a * Y + Y - 1] += wi[a];
        }
        memcpy(T, Tn, sizeof(double) * n * Y);
        /* state cost on x_{k+1}: Q for k+1<N, P for k+1==N (mpc_library.py:515-517) */
        const double *W = (k + 1 < N) ? pb->Q : pb->P;
        for (int a = 0; a < n; ++a)
            for (int j = 0; j < Y; ++j) {
                double v = 0;
                for (int b = 0; b < n; ++b) v += W[a * n + b] * T[b * Y + j];
                WT[a * Y + j] = v;
            }
        for (int i2 = 0; i2 < Y; ++i2)
            for (int j = 0; j < Y; ++j) {
                double v = 0;
                for (int a = 0; a < n; ++a) v += T[a * Y + i2] * WT[a * Y + j];
                M[i2 * Y + j] += v;
            }
    }
    free(T); free(Tn); free(WT);
}

ehmo_problem *ehmo_create(int n, int nu, int N, int K, int p, const double *A, const double *B,
                          const double *w, const double *C, const double *D, const double *e,
                          const double *Q, const double *P, const double *R, const double *ru,
                          const double *theta_scale, double eps_a, double eps_r)
{
    ehmo_problem *pb = (ehmo_problem *)calloc(1, sizeof(*pb));
    pb->n = n; pb->nu = nu; pb->N = N; pb->K = K; pb->p = p;
    pb->U = N * nu; pb->m = N * p; pb->Y = pb->U + n + 1;
    if (pb->U + n + 2 > MAXNV || N > 60) { free(pb); return NULL; }
    long nc = 1;
    for (int k = 0; k < N; ++k) nc *= K;
    pb->ncand = (int)nc;
    pb->A = dcopy(A, (size_t)K * n * n); pb->B = dcopy(B, (size_t)K * n * nu);
    pb->w = dcopy(w, (size_t)K * n); pb->C = dcopy(C, (size_t)K * p * n);
    pb->D = dcopy(D, (size_t)K * p * nu); pb->e = dcopy(e, (size_t)K * p);
    pb->Q = dcopy(Q, (size_t)n * n); pb->P = dcopy(P, (size_t)n * n);
    pb->R = dcopy(R, (size_t)nu * nu); pb->ru = dcopy(ru, (size_t)K * nu);
    pb->tscale = dcopy(theta_scale, n);
    pb->eps_a = eps_a; pb->eps_r = eps_r;
    pb->cG = dalloc((size_t)nc * pb->m * pb->U); pb->cS = dalloc((size_t)nc * pb->m * n);
    pb->cg = dalloc((size_t)nc * pb->m); pb->cM = dalloc((size_t)nc * pb->Y * pb->Y);
    pb->crho = dalloc((size_t)nc * pb->m);
    for (int c = 0; c < pb->ncand; ++c) condense(pb, c);
    return pb;
}

void ehmo_destroy(ehmo_problem *pb)
{
    if (!pb) return;
    free(pb->A); free(pb->B); free(pb->w); free(pb->C); free(pb->D); free(pb->e);
    free(pb->Q); free(pb->P); free(pb->R); free(pb->ru); free(pb->tscale);
    free(pb->cG); free(pb->cS); free(pb->cg); free(pb->cM); free(pb->crho);
    free(pb);
}

void ehmo_set_eps(ehmo_problem *pb, double eps_a, double eps_r)
{
    pb->eps_a = eps_a;
    pb->eps_r = eps_r;
}
int ehmo_num_candidates(const ehmo_problem *pb) { return pb->ncand; }
static void flush_thread_counters(ehmo_problem *pb)
{
    if (t_solves) { __sync_fetch_and_add(&pb->nsolves, t_solves); t_solves = 0; }
    if (t_iters) { __sync_fetch_and_add(&g_iters, t_iters); t_iters = 0; }
    for (int i = 0; i < 6; ++i)
        if (t_stat[i]) { __sync_fetch_and_add(&g_stat[i], t_stat[i]); t_stat[i] = 0; }
}
long ehmo_solve_count(const ehmo_problem *pb)
{
    flush_thread_counters((ehmo_problem *)pb);
    return pb->nsolves;
}

/* ------------------------------------------------------------------------------------ */
/* P_theta_delta (lib/oracle.py:141-173): fixed-commutation convex QP at one theta.       */
/* ------------------------------------------------------------------------------------ */
int ehmo_p_theta_delta(ehmo_problem *pb, const double *theta, int c, double *J, double *u0)
{
    const int n = pb->n, U = pb->U, m = pb->m, Y = pb->Y;
    if (c < 0 || c >= pb->ncand) return 0; /* not one-hot: sum(z[k])==1 violated */
    const double *G = pb->cG + (size_t)c * m * U, *S = pb->cS + (size_t)c * m * n,
                 *g = pb->cg + (size_t)c * m, *M = pb->cM + (size_t)c * Y * Y,
                 *rho = pb->crho + (size_t)c * m;
    double *h = dalloc(m), *Gn = dalloc((size_t)m * U), *hn = dalloc(m);
    int ret = 0;
    for (int r = 0; r < m; ++r) {
        double v = g[r];
        for (int j = 0; j < n; ++j) v += S[r * n + j] * theta[j];
        h[r] = v;
    }
    int mo = prep_rows(U, m, G, h, rho, Gn, hn, U, 0);
    if (mo >= 0) {
        int alive = phase1(pb, U, mo, Gn, hn, NULL);
        if (alive < 0) ret = -1;
        else if (alive == 1) {
            double H[MAXNV * MAXNV], f[MAXNV], z[MAXNV], obj, cst;
            for (int i = 0; i < U; ++i) {
                for (int j = 0; j < U; ++j) H[i * U + j] = 2.0 * M[i * Y + j];
                double v = M[i * Y + Y - 1];
                for (int j = 0; j < n; ++j) v += M[i * Y + U + j] * theta[j];
                f[i] = 2.0 * v;
            }
            cst = M[(Y - 1) * Y + Y - 1];
            for (int i = 0; i < n; ++i) {
                cst += 2.0 * M[(U + i) * Y + Y - 1] * theta[i];
                for (int j = 0; j < n; ++j) cst += theta[i] * M[(U + i) * Y + U + j] * theta[j];
            }
            double cs = 1.0;
            for (int i = 0; i < U; ++i) cs = fmax2(cs, fmax2(fabs(H[i * U + i]), fabs(f[i])));
            for (int i = 0; i < U; ++i) {
                f[i] /= cs;
                for (int j = 0; j < U; ++j) H[i * U + j] /= cs;
            }
            int st = qp_solve(U, mo, H, f, Gn, hn, NULL, z, &obj, NULL);
            obj *= cs;
            t_solves++;
            if (st == 0) {
                ret = 1;
                if (J) *J = obj + cst;
                if (u0)
                    for (int j = 0; j < pb->nu; ++j) u0[j] = z[j];
            } else ret = -1;
        }
    }
    free(h); free(Gn); free(hn);
    return ret;
}

/* P_theta (lib/oracle.py:104-139): MICP = min over commutations.
 * Canonical rule: argmin_c J_c(theta), ties -> smallest candidate index. */
int ehmo_p_theta(ehmo_problem *pb, const double *theta, double *J, double *u0)
{
    int best = -1;
    double Jb = 0, ub[MAXNV], Jc, uc[MAXNV];
    for (int c = 0; c < pb->ncand; ++c) {
        if (ehmo_p_theta_delta(pb, theta, c, &Jc, uc) == 1 && (best < 0 || Jc < Jb)) {
            best = c; Jb = Jc;
            memcpy(ub, uc, sizeof(double) * pb->nu);
        }
    }
    if (best >= 0) {
        if (J) *J = Jb;
        if (u0) memcpy(u0, ub, sizeof(double) * pb->nu);
    }
    return best;
}

/* V_R (lib/oracle.py:175-218): a commutation feasible at every vertex of R, plus the vertex
 * inputs/costs under it.  The reference takes whichever delta MOSEK returns for objective 0
 * and blacklists it if a vertex QP then fails (:198-218).
 * Canonical rule: among candidates whose P_theta_delta is feasible and converges at all n+1
 * vertices, the one minimising sum_i J_c(v_i); ties -> smallest candidate index. */
int ehmo_v_r(ehmo_problem *pb, const double *R, double *Jv, double *Uv)
{
    const int n = pb->n, nu = pb->nu;
    int best = -1;
    double sb = 0, Jc[MAXNV], Uc[MAXNV * MAXNV];
    for (int c = 0; c < pb->ncand; ++c) {
        int ok = 1;
        double sum = 0;
        for (int i = 0; i <= n && ok; ++i) {
            ok = ehmo_p_theta_delta(pb, R + i * n, c, &Jc[i], Uc + i * nu) == 1;
            sum += Jc[i];
        }
        if (ok && (best < 0 || sum < sb)) {
            best = c; sb = sum;
            memcpy(Jv, Jc, sizeof(double) * (n + 1));
            memcpy(Uv, Uc, sizeof(double) * (n + 1) * nu);
        }
    }
    return best;
}

/* ------------------------------------------------------------------------------------ */
/* Simplex problems (theta = v0 + E beta, beta >= 0, sum beta <= 1; lib/oracle.py:70-74)  */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int nv, m;           /* nv = U+n (+1 with tau) */
    double *G, *h, *rho; /* un-normalised rows */
    double H[MAXNV * MAXNV], f[MAXNV], cst;
    double cs; /* objective/tau scale: reported value = cs * obj + cst */
} simplex_qp;

/* build rows + cost of candidate c over simplex R in variables [u; beta] (with_tau=0) or
 * [u; beta; tau] with the two epigraph rows of the eps-suboptimality test (with_tau=1). */
static void build_simplex_qp(const ehmo_problem *pb, int c, const double *R, const double *Vr,
                             int with_tau, simplex_qp *q)
{
    const int n = pb->n, U = pb->U, m = pb->m, Y = pb->Y;
    const int nb = U + n, nv = nb + (with_tau ? 1 : 0);
    const int mt = m + n + 1 + (with_tau ? 2 : 0);
    const double *G = pb->cG + (size_t)c * m * U, *S = pb->cS + (size_t)c * m * n,
                 *g = pb->cg + (size_t)c * m, *M = pb->cM + (size_t)c * Y * Y,
                 *rho = pb->crho + (size_t)c * m;
    double E[MAXNV * MAXNV]; /* n x n, column i = v_{i+1} - v_0 */
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) E[j * n + i] = R[(i + 1) * n + j] - R[j];
    q->nv = nv; q->m = mt; q->cs = 1.0;
    q->G = dalloc((size_t)mt * nv); q->h = dalloc(mt); q->rho = dalloc(mt);
    for (int r = 0; r < m; ++r) {
        double *row = q->G + (size_t)r * nv;
        for (int j = 0; j < U; ++j) row[j] = G[r * U + j];
        double hv = g[r];
        for (int i = 0; i < n; ++i) {
            double v = 0;
            for (int j = 0; j < n; ++j) v += S[r * n + j] * E[j * n + i];
            row[U + i] = -v;
        }
        for (int j = 0; j < n; ++j) hv += S[r * n + j] * R[j];
        q->h[r] = hv;
        q->rho[r] = rho[r];
    }
    for (int i = 0; i < n; ++i) { /* beta_i >= 0 */
        q->G[(size_t)(m + i) * nv + U + i] = -1.0;
        q->h[m + i] = 0.0; q->rho[m + i] = 1.0;
    }
    for (int i = 0; i < n; ++i) q->G[(size_t)(m + n) * nv + U + i] = 1.0; /* sum beta <= 1 */
    q->h[m + n] = 1.0; q->rho[m + n] = 1.0;
    if (with_tau) {
        /* bar_V(beta) = Vr[0] + sum_i beta_i (Vr[i+1]-Vr[0])   (lib/oracle.py:88-90)
         * bar_V - V >= eps_a        <=>  V + tau <= 0 with tau >= eps_a - bar_V
         * bar_V - V >= eps_r V      <=>  V + tau <= 0 with tau >= -bar_V/(1+eps_r)
         * (lib/oracle.py:94-95) */
        double k2 = 1.0 / (1.0 + pb->eps_r);
        double *r1 = q->G + (size_t)(m + n + 1) * nv, *r2 = q->G + (size_t)(m + n + 2) * nv;
        double cs = 1.0; /* cost scale: tau = cs * tau', objective divided by cs */
        for (int i = 0; i <= n; ++i) cs = fmax2(cs, fabs(Vr[i]));
        q->cs = cs;
        double n1 = cs, n2 = cs;
        for (int i = 0; i < n; ++i) {
            double dc = Vr[i + 1] - Vr[0];
            r1[U + i] = -dc; r2[U + i] = -dc * k2;
            n1 = fmax2(n1, fabs(dc)); n2 = fmax2(n2, fabs(dc * k2));
        }
        r1[nb] = -cs; r2[nb] = -cs;
        q->h[m + n + 1] = Vr[0] - pb->eps_a; q->rho[m + n + 1] = n1;
        q->h[m + n + 2] = Vr[0] * k2; q->rho[m + n + 2] = n2;
    }
    /* cost: y = [u; theta; 1] = Tm [u; beta; 1] */
    double *Tm = dalloc((size_t)Y * Y), *MT = dalloc((size_t)Y * Y), *Mp = dalloc((size_t)Y * Y);
    for (int i = 0; i < U; ++i) Tm[i * Y + i] = 1.0;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) Tm[(U + j) * Y + U + i] = E[j * n + i];
        Tm[(U + j) * Y + Y - 1] = R[j];
    }
    Tm[(Y - 1) * Y + Y - 1] = 1.0;
    for (int i = 0; i < Y; ++i)
        for (int j = 0; j < Y; ++j) {
            double v = 0;
            for (int k = 0; k < Y; ++k) v += M[i * Y + k] * Tm[k * Y + j];
            MT[i * Y + j] = v;
        }
    for (int i = 0; i < Y; ++i)
        for (int j = 0; j < Y; ++j) {
            double v = 0;
            for (int k = 0; k < Y; ++k) v += Tm[k * Y + i] * MT[k * Y + j];
            Mp[i * Y + j] = v;
        }
    memset(q->H, 0, sizeof(q->H));
    if (!with_tau) { /* scale the plain QP by its own Hessian/gradient magnitude */
        double cs = 1.0;
        for (int i = 0; i < nb; ++i) cs = fmax2(cs, fmax2(2.0 * fabs(Mp[i * Y + i]), 2.0 * fabs(Mp[i * Y + Y - 1])));
        q->cs = cs;
    }
    for (int i = 0; i < nb; ++i) {
        for (int j = 0; j < nb; ++j) q->H[i * nv + j] = 2.0 * Mp[i * Y + j] / q->cs;
        q->f[i] = 2.0 * Mp[i * Y + Y - 1] / q->cs;
    }
    if (with_tau) q->f[nb] = 1.0;
    q->cst = Mp[(Y - 1) * Y + Y - 1];
    free(Tm); free(MT); free(Mp);
}

static void free_simplex_qp(simplex_qp *q) { free(q->G); free(q->h); free(q->rho); }

/* Is candidate c alive on R?  (phase-1 over [u; beta]).  1 alive, 0 dead, <0 failure */
static int alive_on_simplex(ehmo_problem *pb, int c, const double *R)
{
    simplex_qp q;
    build_simplex_qp(pb, c, R, NULL, 0, &q);
    double *Gn = dalloc((size_t)q.m * q.nv), *hn = dalloc(q.m);
    int mo = prep_rows(q.nv, q.m, q.G, q.h, q.rho, Gn, hn, pb->U, pb->n);
    int ret = (mo < 0) ? 0 : phase1(pb, q.nv, mo, Gn, hn, NULL);
    free(Gn); free(hn); free_simplex_qp(&q);
    return ret;
}

/* Solve the simplex QP of candidate c: with_tau=1 -> t*_c (eps-suboptimality epigraph
 * problem), with_tau=0 -> min_{theta in R} V*_c(theta) (lib/oracle.py:75-79).
 * returns 0 ok; zsol (nv) optional. */
static int solve_simplex_qp_pl(ehmo_problem *pb, int c, const double *R, const double *Vr,
                               int with_tau, double *val, double *zsol, double *plane);
static int solve_simplex_qp(ehmo_problem *pb, int c, const double *R, const double *Vr,
                            int with_tau, double *val, double *zsol)
{
    return solve_simplex_qp_pl(pb, c, R, Vr, with_tau, val, zsol, NULL);
}

/* plane (n+1 values, optional, with_tau only): values at the vertices of R of the tangent plane
 * of the candidate's value function W_c(theta) = min_u V_c(u,theta) at the minimiser theta* of
 * the epigraph problem: W_c >= plane everywhere (convexity), so for ANY vertex costs Vr' on a
 * sub-simplex  t*_c >= min_i (plane_i - Vr'_i) + eps_a  and  >= min_i (plane_i - Vr'_i/(1+eps_r)).
 * Gradient from the KKT multipliers of the stage rows: dW/dbeta = dV/dbeta + sum_r lam_r a_r,beta. */
static int solve_simplex_qp_pl(ehmo_problem *pb, int c, const double *R, const double *Vr,
                               int with_tau, double *val, double *zsol, double *plane)
{
    simplex_qp q;
    build_simplex_qp(pb, c, R, Vr, with_tau, &q);
    double *Gn = dalloc((size_t)q.m * q.nv), *hn = dalloc(q.m);
    double z[MAXNV], obj = 0;
    t_mstage = pb->m;
    int mo = prep_rows(q.nv, q.m, q.G, q.h, q.rho, Gn, hn, pb->U, pb->n), st = 2;
    const int nstage = t_stage_kept;
    t_mstage = 0;
    double *lam = (plane && mo > 0) ? dalloc(mo) : NULL;
    if (mo >= 0) {
        int its = 0;
        t_lam_out = lam;
        st = qp_solve(q.nv, mo, q.H, q.f, Gn, hn, NULL, z, &obj, &its);
        t_lam_out = NULL;
        t_stat[4]++; t_stat[5] += its;
        t_solves++;
    }
    if (st == 0 && plane && with_tau) {
        const int n = pb->n, U = pb->U, nv = q.nv;
        double g[MAXNV], Vst = q.cs * obj + q.cst - q.cs * z[nv - 1], sg = 0, p0;
        for (int i = 0; i < n; ++i) {
            double v = q.f[U + i];
            for (int k = 0; k < nv; ++k) v += q.H[(U + i) * nv + k] * z[k];
            for (int r = 0; r < nstage; ++r) v += lam[r] * Gn[(size_t)r * nv + U + i];
            g[i] = q.cs * v;
            sg += fabs(g[i]);
        }
        p0 = Vst;
        for (int i = 0; i < n; ++i) p0 -= g[i] * z[U + i];
        const double mg = 1e-6 * (1.0 + fabs(Vst) + sg); /* safety margin of the certificate */
        plane[0] = p0 - mg;
        for (int i = 0; i < n; ++i) plane[i + 1] = p0 + g[i] - mg;
    }
    free(lam);
    if (st == 0) {
        *val = q.cs * obj + q.cst;
        if (zsol) {
            memcpy(zsol, z, sizeof(double) * q.nv);
            if (with_tau) zsol[q.nv - 1] *= q.cs;
        }
    } else if (env_flag("EHMO_FAILTRACE", &dbg_failtrace)) {
        fprintf(stderr, "SIMPLEXFAIL c %d tau %d R", c, with_tau);
        for (int i = 0; i < (pb->n + 1) * pb->n; ++i) fprintf(stderr, " %.17g", R[i]);
        fprintf(stderr, " Vr");
        if (Vr) for (int i = 0; i <= pb->n; ++i) fprintf(stderr, " %.17g", Vr[i]);
        fprintf(stderr, "\n");
    }
    free(Gn); free(hn); free_simplex_qp(&q);
    return st;
}

/* One L-CSS node evaluation = bar_E_delta_R (lib/oracle.py:285-309) + bar_D_delta_R
 * (:311-414) + in_variability_ball (:220-283), evaluated together because all three range
 * over the same per-candidate epigraph problem
 *
 *     t*_c = min_{u,beta,tau}  V_c(u,theta(beta)) + tau
 *            s.t. rows of c,  beta in simplex,  tau >= eps_a - bar_V(beta),
 *                 tau >= -bar_V(beta)/(1+eps_r)
 *
 * so that  t*_c <= 0  <=>  candidate c satisfies the two bar_E inequalities (:94-95)
 * somewhere in R.
 *
 * Canonical semantics:
 *   converged  <=>  no alive candidate (delta_ref included) has t*_c <= 0        [bar_E]
 *   delta*      =   argmin t*_c over {delta_ref} U {c alive, t*_c <= 0, P_theta_delta feasible
 *                   and convergent at all n+1 vertices}; ties -> delta_ref first, then the
 *                   smallest index.  If the argmin is delta_ref itself the result is "none",
 *                   exactly the reference's rule for MOSEK returning delta_ref (:384-394).
 *   theta*, V*  =   minimiser / V-part of delta*'s epigraph problem
 *   var_small   =   max(Vr) - min_R V*_{delta_ref} < max(eps_a, eps_r V*)     (:282-283)
 *
 * lb[] (optional, one double per candidate, in/out) carries certified lower bounds of t*_c
 * down a lineage: +inf = dead on R (phase-1 found no interior point; inherited by every
 * descendant), finite = t*_c of an ancestor simplex (valid because bar_V of a child is
 * <= bar_V of its parent on the child), -inf = unknown.  With prune!=0 a candidate whose
 * bound already exceeds the incumbent is skipped -- this never changes the result. */
int ehmo_lcss_eval(ehmo_problem *pb, const double *R, const double *Vr, int cref, double *lb,
                   int prune, double *theta_star, double *Jv, double *Uv,
                   ehmo_lcss_result *res)
{
    const int n = pb->n, nu = pb->nu, U = pb->U, nc = pb->ncand;
    double *ts = dalloc(nc);
    double tref = INFINITY, zref[MAXNV];
    double *pl = (lb && g_planes) ? t_planes : NULL; /* nc x (n+1), see solve_simplex_qp_pl */
    const double k2 = 1.0 / (1.0 + pb->eps_r);
    res->converged = 0; res->cstar = -1; res->var_small = 0; res->n_eval = 0; res->n_fail = 0;
    res->n_plane = 0;
    /* delta_ref is feasible at every vertex by construction: no phase-1 */
    if (cref >= 0) {
        if (solve_simplex_qp(pb, cref, R, Vr, 1, &tref, zref) != 0) { tref = INFINITY; res->n_fail++; }
        res->n_eval++;
    }
    res->t_ref = tref;
    const double thr = tref <= 0.0 ? tref : 0.0; /* incumbent: only t*_c below this matter */
    int any_improving = tref <= 0.0;
    for (int c = 0; c < nc; ++c) {
        ts[c] = NAN;
        if (c == cref) continue;
        if (lb && lb[c] == INFINITY) continue;                 /* dead */
        if (lb && prune && lb[c] > thr) continue;              /* cannot be improving/argmin */
        if (pl && prune) { /* tangent-plane bound under the CURRENT vertex costs */
            const double *P = pl + (size_t)c * (n + 1);
            double a = INFINITY, bq = INFINITY;
            for (int i = 0; i <= n; ++i) {
                a = fmin(a, P[i] - Vr[i]);
                bq = fmin(bq, P[i] - k2 * Vr[i]);
            }
            const double lbp = fmax2(a + pb->eps_a, bq);
            if (lbp > thr) {
                if (lbp > lb[c]) lb[c] = lbp;
                res->n_plane++;
                continue;
            }
        }
        int al = alive_on_simplex(pb, c, R);
        res->n_eval++;
        if (al != 1) {
            /* a numerical failure is NOT a certificate: the inherited bound stays (the reference
             * blacklists a delta inside one bar_D call only, lib/oracle.py:398-414), so the
             * descendants evaluate the candidate again; only a certified "dead" is inherited */
            if (al < 0) res->n_fail++;
            else if (lb) lb[c] = INFINITY;
            continue;
        }
        double t;
        if (solve_simplex_qp_pl(pb, c, R, Vr, 1, &t, NULL, pl ? pl + (size_t)c * (n + 1) : NULL) != 0) {
            res->n_fail++;
            if (pl) pl[(size_t)c * (n + 1)] = -INFINITY;
            continue;
        }
        ts[c] = t;
        if (lb) lb[c] = t;
        if (t <= 0.0) any_improving = 1;
    }
    res->converged = !any_improving;
    /* delta*: ascending t*, strictly better than delta_ref, eligible at all vertices */
    int best = -1;
    unsigned char *tried = (unsigned char *)calloc(nc, 1);
    while (!res->converged) {
        int c = -1;
        for (int k = 0; k < nc; ++k)
            if (k != cref && !tried[k] && ts[k] <= 0.0 && ts[k] < tref && (c < 0 || ts[k] < ts[c]))
                c = k;
        if (c < 0) break;
        tried[c] = 1;
        int ok = 1;
        for (int i = 0; i <= n && ok; ++i)
            ok = ehmo_p_theta_delta(pb, R + i * n, c, &Jv[i], Uv + i * nu) == 1;
        if (ok) { best = c; break; }
    }
    free(tried);
    if (best >= 0) {
        double minV, t, z[MAXNV];
        if (solve_simplex_qp(pb, cref, R, NULL, 0, &minV, NULL) != 0 ||
            solve_simplex_qp(pb, best, R, Vr, 1, &t, z) != 0) {
            best = -1;
            res->n_fail++;
        } else {
            double Vstar = t - z[U + n], maxV = Vr[0];
            for (int i = 1; i <= n; ++i) maxV = fmax2(maxV, Vr[i]);
            for (int j = 0; j < n; ++j) {
                double v = R[j];
                for (int i = 0; i < n; ++i) v += (R[(i + 1) * n + j] - R[j]) * z[U + i];
                theta_star[j] = v;
            }
            res->var_small = (maxV - minV) < fmax2(pb->eps_a, pb->eps_r * Vstar);
        }
    }
    res->cstar = best;
    free(ts);
    return 0;
}

/* When a lineage changes its commutation (vertex costs Vr_old -> Vr_new on the same
 * simplex) the inherited bounds stay valid after lowering them by
 * d = max(0, max_i Vr_new[i]-Vr_old[i]):  bar_V_new <= bar_V_old + d on R. */
void ehmo_lb_rebase(int n, int ncand, const double *Vr_old, const double *Vr_new, double *lb)
{
    double d = 0;
    for (int i = 0; i <= n; ++i) d = fmax2(d, Vr_new[i] - Vr_old[i]);
    if (d > 0) {
        d += 1e-9 * (1.0 + d);
        for (int c = 0; c < ncand; ++c)
            if (isfinite(lb[c])) lb[c] -= d;
    }
}

/* bar_E_delta_R as a stand-alone call (brute force over all candidates).  The reference
 * signature has no delta_ref; every candidate goes through phase-1. */
int ehmo_bar_e(ehmo_problem *pb, const double *R, const double *Vr, double *tstar)
{
    int subopt = 1;
    for (int c = 0; c < pb->ncand; ++c) {
        double t = NAN;
        if (alive_on_simplex(pb, c, R) == 1 && solve_simplex_qp(pb, c, R, Vr, 1, &t, NULL) == 0) {
            if (t <= 0.0) subopt = 0;
        } else t = NAN;
        if (tstar) tstar[c] = t;
    }
    return subopt;
}

/* min_{theta in R} V*_c(theta)  (min_over_simplex_for_this_delta, lib/oracle.py:75-79,276) */
int ehmo_min_over_simplex(ehmo_problem *pb, const double *R, int c, double *val)
{
    return solve_simplex_qp(pb, c, R, NULL, 0, val, NULL);
}

/* bar_D_delta_R as a stand-alone call (brute force). */
int ehmo_bar_d(ehmo_problem *pb, const double *R, const double *Vr, int cref,
               double *theta_star, double *Jv, double *Uv, int *var_small)
{
    ehmo_lcss_result res;
    ehmo_lcss_eval(pb, R, Vr, cref, NULL, 0, theta_star, Jv, Uv, &res);
    *var_small = res.var_small;
    return res.cstar;
}

/* ------------------------------------------------------------------------------------ */
/* geometry (lib/tools.py:224-257, :134-150)                                              */
/* ------------------------------------------------------------------------------------ */
/* Longest edge = first maximum, in itertools.combinations order, of the Euclidean length;
 * compared through the squared length accumulated left-to-right without FMA contraction so
 * that CPU and GPU take bit-identical decisions (DESIGN.md "bit-exact geometry"). */
void ehmo_split(int n, const double *R, double *S1, double *S2, int *ij)
{
    /* lib/tools.py:245-250: argmax (first maximum) over itertools.combinations of
     * la.norm(R[i]-R[j]).  numpy's norm of a 1-D vector is sqrt(x.dot(x)); for these short
     * vectors the BLAS dot is a left-to-right chain of fused multiply-adds starting from 0
     * (verified against numpy 2.3 / OpenBLAS 0.3.30 on 20000 vectors, tests/test_oracle_golden.py),
     * so that is the arithmetic restated here -- ties between edges of equal length are
     * decided by exactly these roundings. */
    int bi = 0, bj = 1;
    double best = -1.0;
    for (int i = 0; i <= n; ++i)
        for (int j = i + 1; j <= n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) {
                volatile double d = R[i * n + k] - R[j * n + k];
                acc = fma(d, d, acc);
            }
            const double len = sqrt(acc);
            if (len > best) { best = len; bi = i; bj = j; }
        }
    memcpy(S1, R, sizeof(double) * (n + 1) * n);
    memcpy(S2, R, sizeof(double) * (n + 1) * n);
    for (int k = 0; k < n; ++k) {
        double mid = (R[bi * n + k] + R[bj * n + k]) / 2.0;
        S1[bi * n + k] = mid;
        S2[bj * n + k] = mid;
    }
    ij[0] = bi; ij[1] = bj;
}

double ehmo_simplex_volume(int n, const double *R)
{
    double M[MAXNV * MAXNV], det = 1.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) M[j * n + i] = R[(i + 1) * n + j] - R[j];
    for (int c = 0; c < n; ++c) { /* partial-pivot LU */
        int pv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(M[r * n + c]) > fabs(M[pv * n + c])) pv = r;
        if (M[pv * n + c] == 0.0) return 0.0;
        if (pv != c) {
            for (int j = 0; j < n; ++j) { double t = M[c * n + j]; M[c * n + j] = M[pv * n + j]; M[pv * n + j] = t; }
            det = -det;
        }
        det *= M[c * n + c];
        for (int r = c + 1; r < n; ++r) {
            double l = M[r * n + c] / M[c * n + c];
            for (int j = c; j < n; ++j) M[r * n + j] -= l * M[c * n + j];
        }
    }
    for (int k = 2; k <= n; ++k) det /= k;
    return fabs(det);
}

/* ------------------------------------------------------------------------------------ */
/* ECC / L-CSS (lib/worker.py:241-417) driven by a shared LIFO task queue                  */
/* (lib/scheduler.py:364-599: tasks popped from the tail; every split pushes both children, */
/* i.e. the reference's "offload the child" path of worker.py:187-239 taken always).        */
/* The tree does not depend on who processes a node, nor on the order.                      */
/* ------------------------------------------------------------------------------------ */
typedef struct task {
    long id;
    double *lb;   /* NULL: ECC task; else L-CSS with these inherited bounds (owned) */
    struct task *next;
} task;

typedef struct {
    ehmo_problem *pb;
    ehmo_tree t;
    long cap;
    int prune;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    task *top;
    int busy, nthreads, stop;
} builder;

static long new_node(builder *b, const double *R, int root_id, int depth)
{
    ehmo_tree *t = &b->t;
    const int n = t->n, nu = t->nu;
    long id = __sync_fetch_and_add(&t->nnodes, 1);
    if (id >= b->cap) { t->error = 3; __sync_fetch_and_sub(&t->nnodes, 1); return -1; }
    t->left[id] = t->right[id] = -1;
    t->root_id[id] = root_id; t->depth[id] = depth; t->cand[id] = -1; t->closed[id] = 0;
    memcpy(t->vertices + id * (n + 1) * n, R, sizeof(double) * (n + 1) * n);
    memset(t->vcost + id * (n + 1), 0, sizeof(double) * (n + 1));
    memset(t->vinput + id * (n + 1) * nu, 0, sizeof(double) * (n + 1) * nu);
    int md;
    while (depth > (md = t->max_depth) && !__sync_bool_compare_and_swap(&t->max_depth, md, depth)) {}
    return id;
}

static void push_task(builder *b, long id, double *lb)
{
    task *k = (task *)malloc(sizeof(task));
    k->id = id; k->lb = lb;
    pthread_mutex_lock(&b->mu);
    k->next = b->top;
    b->top = k;
    pthread_cond_signal(&b->cv);
    pthread_mutex_unlock(&b->mu);
}

/* Worker.ecc (lib/worker.py:241-291), one node */
static void ecc_step(builder *b, long id)
{
    ehmo_problem *pb = b->pb;
    ehmo_tree *t = &b->t;
    const int n = t->n, nu = t->nu;
    double R[MAXNV * MAXNV], Jv[MAXNV], Uv[MAXNV * MAXNV];
    memcpy(R, t->vertices + id * (n + 1) * n, sizeof(double) * (n + 1) * n);
    int c = ehmo_v_r(pb, R, Jv, Uv);
    if (c < 0) {
        /* worker.py:264-266: the barycentre must admit some commutation */
        double cR[MAXNV], J;
        for (int j = 0; j < n; ++j) {
            cR[j] = 0;
            for (int i = 0; i <= n; ++i) cR[j] += R[i * n + j];
            cR[j] /= (n + 1);
        }
        int any = 0;
        for (int k = 0; k < pb->ncand && !any; ++k)
            any = ehmo_p_theta_delta(pb, cR, k, &J, NULL) == 1;
        if (!any) { t->error = 1; return; }
        double S1[MAXNV * MAXNV], S2[MAXNV * MAXNV];
        int ij[2];
        ehmo_split(n, R, S1, S2, ij);
        long l = new_node(b, S1, t->root_id[id], t->depth[id] + 1);
        long r = new_node(b, S2, t->root_id[id], t->depth[id] + 1);
        if (l < 0 || r < 0) return;
        t->left[id] = (int)l; t->right[id] = (int)r;
        push_task(b, l, NULL);
        push_task(b, r, NULL);
    } else {
        t->cand[id] = c;
        memcpy(t->vcost + id * (n + 1), Jv, sizeof(double) * (n + 1));
        memcpy(t->vinput + id * (n + 1) * nu, Uv, sizeof(double) * (n + 1) * nu);
        /* bounds, then one tangent plane (n+1 vertex values) per candidate; -inf = none yet */
        double *lb = dalloc((size_t)pb->ncand * (n + 2));
        for (size_t k = pb->ncand; k < (size_t)pb->ncand * (n + 2); ++k) lb[k] = -INFINITY;
        for (int k = 0; k < pb->ncand; ++k) lb[k] = -INFINITY;
        push_task(b, id, lb);
    }
}

/* Worker.lcss (lib/worker.py:293-417), one node; lb is consumed */
static void lcss_step(builder *b, long id, double *lb)
{
    ehmo_problem *pb = b->pb;
    ehmo_tree *t = &b->t;
    const int n = t->n, nu = t->nu, nc = pb->ncand;
    double R[MAXNV * MAXNV], Vr[MAXNV], Ur[MAXNV * MAXNV];
    memcpy(R, t->vertices + id * (n + 1) * n, sizeof(double) * (n + 1) * n);
    memcpy(Vr, t->vcost + id * (n + 1), sizeof(double) * (n + 1));
    memcpy(Ur, t->vinput + id * (n + 1) * nu, sizeof(double) * (n + 1) * nu);
    int cref = t->cand[id];
    double th[MAXNV], Jn[MAXNV], Un[MAXNV * MAXNV];
    ehmo_lcss_result res;
    t_planes = lb + nc;
    ehmo_lcss_eval(pb, R, Vr, cref, lb, b->prune, th, Jn, Un, &res);
    t_planes = NULL;
    __sync_fetch_and_add(&t->nfail, res.n_fail);
    __sync_fetch_and_add(&g_plane_skips, res.n_plane);
    if (env_flag("EHMO_TRACE", &dbg_trace))
        fprintf(stderr, "node %ld depth %d cref %d evaluated %d tref %.4g conv %d cstar %d vs %d\n",
                id, t->depth[id], cref, res.n_eval, res.t_ref, res.converged, res.cstar, res.var_small);
    if (res.converged) { /* worker.py:368-375 close leaf */
        t->closed[id] = 1;
        __sync_fetch_and_add(&t->nleaves, 1);
        free(lb);
        return;
    }
    int cnew = cref, have_mid = 0;
    double Jmid = 0, Umid[MAXNV];
    if (res.cstar >= 0 && !res.var_small) {
        /* delta* is adopted only if its own midpoint problem goes through; a numerical failure falls
         * back to delta_ref for this visit (the reference's "blacklist delta* and retry",
         * lib/oracle.py:398-414) -- same rule as k_lcss_final of the CUDA engine */
        double T1[MAXNV * MAXNV], T2[MAXNV * MAXNV];
        int tij[2];
        ehmo_split(n, R, T1, T2, tij);
        if (ehmo_p_theta_delta(pb, T1 + tij[0] * n, res.cstar, &Jmid, Umid) != 1) {
            __sync_fetch_and_add(&t->nfail, 1);
            res.cstar = -1;
        } else {
            have_mid = 1;
        }
    }
    if (res.cstar >= 0) { /* worker.py:389-395 */
        cnew = res.cstar;
        ehmo_lb_rebase(n, nc, Vr, Jn, lb);
        lb[cref] = -INFINITY; /* the old commutation becomes an ordinary candidate */
        memcpy(Vr, Jn, sizeof(double) * (n + 1));
        memcpy(Ur, Un, sizeof(double) * (n + 1) * nu);
    }
    if (res.cstar >= 0 && res.var_small) { /* worker.py:396-401: swap, same node again */
        t->cand[id] = cnew;
        memcpy(t->vcost + id * (n + 1), Vr, sizeof(double) * (n + 1));
        memcpy(t->vinput + id * (n + 1) * nu, Ur, sizeof(double) * (n + 1) * nu);
        push_task(b, id, lb);
        return;
    }
    /* worker.py:403-417: split, re-solve only the midpoint vertex */
    double S1[MAXNV * MAXNV], S2[MAXNV * MAXNV], Jm, Um[MAXNV];
    int ij[2];
    ehmo_split(n, R, S1, S2, ij);
    if (have_mid) {
        Jm = Jmid;
        memcpy(Um, Umid, sizeof(double) * nu);
    } else if (ehmo_p_theta_delta(pb, S1 + ij[0] * n, cnew, &Jm, Um) != 1) {
        t->error = 2; /* cannot happen for a convex feasible set; the reference would raise */
        free(lb);
        return;
    }
    long l = new_node(b, S1, t->root_id[id], t->depth[id] + 1);
    long r = new_node(b, S2, t->root_id[id], t->depth[id] + 1);
    if (l < 0 || r < 0) { free(lb); return; }
    t->left[id] = (int)l; t->right[id] = (int)r;
    for (int side = 0; side < 2; ++side) {
        long ch = side ? r : l;
        t->cand[ch] = cnew;
        memcpy(t->vcost + ch * (n + 1), Vr, sizeof(double) * (n + 1));
        memcpy(t->vinput + ch * (n + 1) * nu, Ur, sizeof(double) * (n + 1) * nu);
        t->vcost[ch * (n + 1) + ij[side]] = Jm;
        memcpy(t->vinput + (ch * (n + 1) + ij[side]) * nu, Um, sizeof(double) * nu);
    }
    /* tangent planes follow the geometry: the new vertex is the midpoint of vertices ij[0], ij[1] */
    double *lbl = dcopy(lb, (size_t)nc * (n + 2));
    for (int c = 0; c < nc; ++c) {
        double *Pl = lbl + nc + (size_t)c * (n + 1), *Pr = lb + nc + (size_t)c * (n + 1);
        const double mid = 0.5 * (Pr[ij[0]] + Pr[ij[1]]);
        Pl[ij[0]] = mid;
        Pr[ij[1]] = mid;
    }
    if (g_planes == 2) {
        /* "local" variant (what the CUDA engine does): the planes of THIS node's evaluation
         * tighten the children's bounds under the children's own vertex costs, then are dropped */
        const double k2 = 1.0 / (1.0 + pb->eps_r);
        for (int side = 0; side < 2; ++side) {
            double *lc = side ? lb : lbl;
            const double *Vc = t->vcost + (side ? r : l) * (n + 1);
            for (int c = 0; c < nc; ++c) {
                double *P = lc + nc + (size_t)c * (n + 1);
                if (P[0] > -INFINITY && lc[c] != INFINITY) {
                    double a = INFINITY, bq = INFINITY;
                    for (int i = 0; i <= n; ++i) { a = fmin(a, P[i] - Vc[i]); bq = fmin(bq, P[i] - k2 * Vc[i]); }
                    const double lbp = fmax2(a + pb->eps_a, bq);
                    if (lbp > lc[c]) lc[c] = lbp;
                }
                for (int i = 0; i <= n; ++i) P[i] = -INFINITY;
            }
        }
    }
    push_task(b, l, lbl);
    push_task(b, r, lb);
}

static void *thread_main(void *p)
{
    builder *b = (builder *)p;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        while (!b->top && b->busy > 0 && !b->stop) pthread_cond_wait(&b->cv, &b->mu);
        if (b->stop || (!b->top && b->busy == 0)) {
            pthread_cond_broadcast(&b->cv);
            pthread_mutex_unlock(&b->mu);
            break;
        }
        task *k = b->top;
        b->top = k->next;
        b->busy++;
        pthread_mutex_unlock(&b->mu);
        if (!b->t.error) {
            if (k->lb) lcss_step(b, k->id, k->lb);
            else ecc_step(b, k->id);
        } else free(k->lb);
        free(k);
        flush_thread_counters(b->pb);
        pthread_mutex_lock(&b->mu);
        b->busy--;
        if (b->t.error) b->stop = 1;
        if ((!b->top && b->busy == 0) || b->stop) pthread_cond_broadcast(&b->cv);
        pthread_mutex_unlock(&b->mu);
    }
    return NULL;
}

ehmo_tree *ehmo_partition(ehmo_problem *pb, int nroots, const double *roots, int nthreads,
                          long max_nodes, int prune)
{
    if (nthreads < 1) nthreads = 1;
    const int n = pb->n, nu = pb->nu;
    builder b;
    memset(&b, 0, sizeof b);
    b.pb = pb; b.prune = prune; b.nthreads = nthreads;
    b.cap = max_nodes > 0 ? max_nodes : (1L << 21);
    pthread_mutex_init(&b.mu, NULL);
    pthread_cond_init(&b.cv, NULL);
    ehmo_tree *t = &b.t;
    t->n = n; t->nu = nu;
    t->left = (int *)malloc(b.cap * sizeof(int)); t->right = (int *)malloc(b.cap * sizeof(int));
    t->root_id = (int *)malloc(b.cap * sizeof(int)); t->depth = (int *)malloc(b.cap * sizeof(int));
    t->cand = (int *)malloc(b.cap * sizeof(int)); t->closed = (unsigned char *)malloc(b.cap);
    t->vertices = (double *)malloc(b.cap * (n + 1) * n * sizeof(double));
    t->vcost = (double *)malloc(b.cap * (n + 1) * sizeof(double));
    t->vinput = (double *)malloc(b.cap * (n + 1) * nu * sizeof(double));
    long solves0 = pb->nsolves;
    for (int r = nroots - 1; r >= 0; --r) { /* reverse: root 0 is popped first */
        long id = new_node(&b, roots + (size_t)r * (n + 1) * n, r, 0);
        push_task(&b, id, NULL);
    }
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, thread_main, &b);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
    while (b.top) { task *k = b.top; b.top = k->next; free(k->lb); free(k); }
    pthread_mutex_destroy(&b.mu);
    pthread_cond_destroy(&b.cv);
    ehmo_tree *out = (ehmo_tree *)malloc(sizeof(ehmo_tree));
    *out = *t;
    out->nsolves = pb->nsolves - solves0;
    return out;
}

void ehmo_tree_free(ehmo_tree *t)
{
    if (!t) return;
    free(t->left); free(t->right); free(t->root_id); free(t->depth); free(t->cand);
    free(t->closed); free(t->vertices); free(t->vcost); free(t->vinput);
    free(t);
}
