// ehm_solve.cuh -- per-warp assembly + solution of the oracle sub-problems.
//
// Sub-problems (reference lib/oracle.py; algebra as in oracle/ehm_oracle.c, written
// independently here for the warp-parallel layout):
//   point   : P_theta_delta (:141-173)            variables u            (U)
//   alive   : phase-1 over a simplex              variables [u;beta;sig] (U+NX+1)
//   tqp     : eps-suboptimality epigraph problem  variables [u;beta;tau] (U+NX+1)  (:88-97)
//   minv    : min over simplex of V*_c            variables [u;beta]     (U+NX)    (:75-79)
//
// Every problem class runs through ONE instantiation of ipm_solve (point problems NV=U+1,
// simplex problems NV=U+NX+1): the phase-1 variable sigma, the epigraph variable tau and --
// where a problem has one variable less -- an inert dummy column (zero coefficients, zero
// cost; its Cholesky pivot is frozen by the guard) share the last column.
//
// Row pre-solve.  The condensed problems carry every big-M row of the reference
// (lib/mpc_library.py:529-534): 210 rows for the pendulum, of which ~160 can never bind inside
// the box |uhat| <= 1 the input-bound rows impose.  A normalised row a'z <= h is dropped when
//      sup_{|u|<=1, beta in simplex} a'z  =  |a_u|_1 + max(0, max_i a_beta_i)  <=  h - TAU_DROP,
// which is exact for every decision the engine takes (a point that satisfies the kept rows,
// which include the box rows, with margin TAU_INT satisfies the dropped ones with margin
// TAU_DROP > TAU_INT; optima are unchanged because the feasible set is).  The surviving rows
// (34-61 for the pendulum) are compacted with warp ballots, so the solver works on <= 96 rows
// (MRC = 3 per lane) instead of 224: 3-4x less work per iteration and 26 KB instead of 31 KB of
// shared memory per problem (8 resident one-warp CTAs per SM at 255 registers).
//
// Candidate tables live in global memory (L2 resident: 60 MB for the pendulum), column-major
// with the row index fastest so that a warp's load of one column is one coalesced 256 B
// segment per 32 rows.
#pragma once
#include "ehm_ipm.cuh"

#define EHM_TAU_DROP 1e-3

namespace ehm {

struct DevTables {
    int n, nu, N, K, p;
    int U, m, MP, Y, ncand;  // U=N*nu, m=N*p, MP = rows padded to 32, Y=U+n+1
    const double *G;         // [ncand][U][MP]
    const double *S;         // [ncand][n][MP]
    const double *g;         // [ncand][MP]
    const double *rho;       // [ncand][MP]
    const double *M;         // [ncand][Y*Y]
    double eps_a, eps_r;
};

// Shared-memory workspace of one warp.  MRI = input row slots per lane (ceil(rows/32) of the
// un-presolved problem), MRC = row capacity per lane after the pre-solve.
template <int U, int NX, int MRI, int MRC>
struct WarpWs {
    static constexpr int NVS = U + NX + 1;
    static constexpr int NVP = U + 1;
    using WS = IpmWs<NVS, MRC>;
    using WP = IpmWs<NVP, MRC>;
    static constexpr int MRP = 32 * MRC;
    static constexpr int Y = U + NX + 1;
    // assembly scratch behind the solver block (cost transform, simplex geometry)
    static constexpr int oM = WS::kDoubles;         // candidate M, later Mp (Y*Y)
    static constexpr int oT = oM + Y * Y;           // T (Y*Y)
    static constexpr int oMT = oT + Y * Y;          // M*T (Y*Y)
    static constexpr int oR = oMT + Y * Y;          // simplex vertices (NX+1)*NX
    static constexpr int oE = oR + (NX + 1) * NX;   // E NX*NX
    static constexpr int oV = oE + NX * NX;         // Vr (NX+1), padded 16
    static constexpr int kDoubles = oV + 16;
    static_assert(WP::kDoubles <= WS::kDoubles, "point block must fit the simplex block");
};

struct SolveOut {
    int status;   // 1 ok / feasible, 0 infeasible / not alive, -1 numerical failure
    double val;   // J, sigma*, t* or min V
    int iters;    // IPM iterations spent
    int nsolves;  // IPM solves spent
    double flops; // algorithmic flops spent (see ipm_solve)
};

// bit j set <=> row lane+32*j < nrows
template <int MRC>
__device__ __forceinline__ unsigned act_of(int nrows, int lane)
{
    unsigned a = 0;
#pragma unroll
    for (int j = 0; j < MRC; ++j)
        if (lane + 32 * j < nrows) a |= 1u << j;
    return a;
}

template <int NV, int MRC>
__device__ __forceinline__ int run_ipm(double *ws, unsigned act, int lane, SolveOut &out, double &obj,
                                       bool early = false)
{
    using W = IpmWs<NV, MRC>;
    const int st = ipm_solve<NV, MRC>(ws, act, lane, early);
    obj = ws[W::oOut + 0];
    out.iters += (int)ws[W::oOut + 1];
    out.flops += ws[W::oOut + 2];
    out.nsolves++;
    __syncwarp();
    return st;
}

// zero rows [from, 32*MRC) of an NV-column tile (G = 0, h = 1: the ipm_solve contract for rows
// that do not take part)
template <int NV, int MRC>
__device__ __forceinline__ void clear_rows(double *Gs, double *hs, int from, int lane)
{
    constexpr int MRP = 32 * MRC;
#pragma unroll
    for (int j = 0; j < MRC; ++j) {
        const int r = lane + 32 * j;
        if (r >= from) {
#pragma unroll
            for (int k = 0; k < NV; ++k) Gs[k * MRP + r] = 0.0;
            hs[r] = 1.0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// P_theta_delta at one theta for candidate c.   u0 (nu values) written to u0_out.
// ---------------------------------------------------------------------------------------
template <int U, int NX, int MRI, int MRC>
__device__ SolveOut solve_point(const DevTables &tb, int c, const double *__restrict__ theta,
                                double *ws, double *u0_out, int lane)
{
    using W = typename WarpWs<U, NX, MRI, MRC>::WP;
    constexpr int MRP = W::MRP, NV = U + 1;
    double *Gs = ws + W::oG, *hs = ws + W::oH, *Hs = ws + W::oHs, *fs = ws + W::oF, *zs = ws + W::oZ;
    SolveOut out{0, 0.0, 0, 0, 0.0};
    const int MP = tb.MP, m = tb.m, Y = tb.Y;
    const double *Gc = tb.G + (size_t)c * U * MP, *Sc = tb.S + (size_t)c * NX * MP,
                 *gc = tb.g + (size_t)c * MP, *rc = tb.rho + (size_t)c * MP;
    double th[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) th[i] = theta[i];
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    bool viol = false;
    double s0 = 0.0;
#pragma unroll 1
    for (int j = 0; j < MRI; ++j) {
        const int r = lane + 32 * j;
        bool keep = false;
        double g[U], hv = 0.0;
#pragma unroll
        for (int k = 0; k < U; ++k) g[k] = 0.0;
        if (r < m) {
            double nrm = 0.0, l1 = 0.0;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                g[k] = Gc[k * MP + r];
                nrm = fmax(nrm, fabs(g[k]));
                l1 += fabs(g[k]);
            }
            hv = gc[r];
#pragma unroll
            for (int i = 0; i < NX; ++i) hv = fma(Sc[i * MP + r], th[i], hv);
            const double rho = rc[r];
            const double sc = rho > 0.0 ? rho : 1.0;
            if (nrm <= EHM_ZERO_ROW * sc) {
                if (hv < -EHM_TAU_CONST * sc) viol = true;
            } else {
                const double inv = 1.0 / nrm;
                hv *= inv;
#pragma unroll
                for (int k = 0; k < U; ++k) g[k] *= inv;
                keep = !(l1 * inv <= hv - EHM_TAU_DROP);   // pre-solve: cannot bind in |u| <= 1
            }
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        if (keep) {
            const int pos = cnt + __popc(bal & lt);
            if (pos < MRP) {
#pragma unroll
                for (int k = 0; k < U; ++k) Gs[k * MRP + pos] = g[k];
                Gs[U * MRP + pos] = -1.0;
                hs[pos] = hv;
            }
            s0 = fmax(s0, -hv);
        }
        cnt += __popc(bal);
    }
    if (__any_sync(kFull, viol)) return out;
    if (cnt > MRP) {  // capacity of the compacted tile exceeded (never for the shipped examples)
        out.status = -1;
        return out;
    }
    clear_rows<NV, MRC>(Gs, hs, cnt, lane);
    const unsigned act = act_of<MRC>(cnt, lane);
    s0 = warp_max(s0);
    // ---- phase 1: min sigma s.t. G u - sigma <= h --------------------------------------
    for (int e = lane; e < NV * NV; e += 32) Hs[e] = 0.0;
    if (lane < 16) {
        fs[lane] = (lane == U) ? 1.0 : 0.0;
        zs[lane] = (lane == U) ? s0 + 1.0 : 0.0;
    }
    __syncwarp();
    double obj;
    int st = run_ipm<NV, MRC>(ws, act, lane, out, obj, true);
    if (st != 0) {
        out.status = -1;
        return out;
    }
    if (!(obj <= -EHM_TAU_INT)) return out;
    // ---- phase 2: the QP (last column becomes an inert dummy) ------------------------------
    const double *Mc = tb.M + (size_t)c * Y * Y;
    double cs = 1.0, cst = 0.0;
    {
        double fv = 0.0, hd = 0.0;
        if (lane < U) {
            fv = Mc[lane * Y + Y - 1];
#pragma unroll
            for (int i = 0; i < NX; ++i) fv = fma(Mc[lane * Y + U + i], th[i], fv);
            fv *= 2.0;
            hd = 2.0 * Mc[lane * Y + lane];
        }
        cs = fmax(1.0, warp_max(fmax(fabs(fv), fabs(hd))));
        if (lane < 16) {
            fs[lane] = (lane < U) ? fv / cs : 0.0;
            zs[lane] = 0.0;
        }
        for (int e = lane; e < NV * NV; e += 32) {
            const int a = e / NV, b = e % NV;
            Hs[e] = (a < U && b < U) ? 2.0 * Mc[a * Y + b] / cs : 0.0;
        }
#pragma unroll
        for (int j = 0; j < MRC; ++j) Gs[U * MRP + lane + 32 * j] = 0.0;
        cst = Mc[(Y - 1) * Y + Y - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            cst = fma(2.0 * Mc[(U + i) * Y + Y - 1], th[i], cst);
#pragma unroll
            for (int k = 0; k < NX; ++k) cst = fma(th[i] * Mc[(U + i) * Y + U + k], th[k], cst);
        }
        __syncwarp();
    }
    st = run_ipm<NV, MRC>(ws, act, lane, out, obj);
    if (st != 0) {
        out.status = -1;
        return out;
    }
    out.status = 1;
    out.val = cs * obj + cst;
    if (u0_out != nullptr && lane < tb.nu && lane < U) u0_out[lane] = zs[lane];
    __syncwarp();
    return out;
}

// ---------------------------------------------------------------------------------------
// Simplex problems.  Rv: (NX+1)*NX vertices (global), Vr: NX+1 vertex costs (global, TQP).
// mode: EHM_SX_ALIVE / EHM_SX_TQP / EHM_SX_MINV;  phase1: run the alive test first.
// sol (optional, global, NX+1 doubles): [beta*, V*] written for TQP / MINV.
// Row order of the compacted tile: surviving stage rows, the NX+1 simplex rows, then (TQP)
// the two epigraph rows -- so phase 1 simply runs on the first total-2 rows.
// plane (optional, global, NX+1 floats, TQP): values at the vertices of R of the tangent plane
// of the candidate's value function W_c(theta) = min_u V_c(u, theta) at the minimiser theta* of
// the epigraph problem, lowered by a safety margin and rounded down.  W_c >= plane everywhere
// (convexity), hence for ANY vertex costs Vr' on a sub-simplex with vertex values plane'_i
//     t*_c >= min_i (plane'_i - Vr'_i) + eps_a   and   t*_c >= min_i (plane'_i - Vr'_i/(1+eps_r)).
// The gradient comes from the KKT multipliers of the stage rows (oracle/ehm_oracle.c:
// solve_simplex_qp_pl):  dW/dbeta = dV/dbeta + sum_r lam_r a_(r,beta).
// ---------------------------------------------------------------------------------------
template <int U, int NX, int MRI, int MRC>
__device__ SolveOut solve_simplex(const DevTables &tb, int c, const double *__restrict__ Rv,
                                  const double *__restrict__ Vr, int mode, bool phase1,
                                  double *ws, double *sol, int lane, float *plane = nullptr)
{
    using WW = WarpWs<U, NX, MRI, MRC>;
    using W = typename WW::WS;
    constexpr int MRP = W::MRP, NB = U + NX, NVS = U + NX + 1, Y = WW::Y;
    double *Gs = ws + W::oG, *hs = ws + W::oH, *Hs = ws + W::oHs, *fs = ws + W::oF, *zs = ws + W::oZ,
           *Ms = ws + WW::oM, *Ts = ws + WW::oT, *MTs = ws + WW::oMT, *Rs = ws + WW::oR,
           *Es = ws + WW::oE, *Vs = ws + WW::oV;
    SolveOut out{0, 0.0, 0, 0, 0.0};
    const int MP = tb.MP, m = tb.m;
    const double *Gc = tb.G + (size_t)c * U * MP, *Sc = tb.S + (size_t)c * NX * MP,
                 *gc = tb.g + (size_t)c * MP, *rc = tb.rho + (size_t)c * MP;
    // simplex geometry: E[j][i] = v_{i+1}[j] - v_0[j]
    for (int e = lane; e < (NX + 1) * NX; e += 32) Rs[e] = Rv[e];
    if (mode == EHM_SX_TQP && lane <= NX) Vs[lane] = Vr[lane];
    __syncwarp();
    for (int e = lane; e < NX * NX; e += 32) {
        const int j = e / NX, i = e % NX;
        Es[e] = Rs[(i + 1) * NX + j] - Rs[j];
    }
    __syncwarp();
    double cs = 1.0;  // objective / tau scale
    if (mode == EHM_SX_TQP) {
#pragma unroll
        for (int i = 0; i <= NX; ++i) cs = fmax(cs, fabs(Vs[i]));
    }
    const bool tau = (mode == EHM_SX_TQP);
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    bool viol = false;
    double s0 = 0.0;
    // ---- stage rows: normalise, pre-solve, compact ------------------------------------------
#pragma unroll 1
    for (int j = 0; j < MRI; ++j) {
        const int r = lane + 32 * j;
        bool keep = false;
        double co[NB], hv = 0.0;
#pragma unroll
        for (int k = 0; k < NB; ++k) co[k] = 0.0;
        if (r < m) {
            double nrm = 0.0, l1 = 0.0, bmax = 0.0;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                co[k] = Gc[k * MP + r];
                nrm = fmax(nrm, fabs(co[k]));
                l1 += fabs(co[k]);
            }
            double sr[NX];
            hv = gc[r];
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) {
                sr[jj] = Sc[jj * MP + r];
                hv = fma(sr[jj], Rs[jj], hv);
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double v = 0.0;
#pragma unroll
                for (int jj = 0; jj < NX; ++jj) v = fma(sr[jj], Es[jj * NX + i], v);
                co[U + i] = -v;
                nrm = fmax(nrm, fabs(v));
                bmax = fmax(bmax, -v);
            }
            const double rho = rc[r];
            const double sc = rho > 0.0 ? rho : 1.0;
            if (nrm <= EHM_ZERO_ROW * sc) {
                if (hv < -EHM_TAU_CONST * sc) viol = true;
            } else {
                const double inv = 1.0 / nrm;
                hv *= inv;
#pragma unroll
                for (int k = 0; k < NB; ++k) co[k] *= inv;
                // sup over |u| <= 1 and beta in the simplex of the row's left-hand side
                keep = !((l1 + bmax) * inv <= hv - EHM_TAU_DROP);
            }
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        if (keep) {
            const int pos = cnt + __popc(bal & lt);
            if (pos < MRP) {
#pragma unroll
                for (int k = 0; k < NB; ++k) Gs[k * MRP + pos] = co[k];
                Gs[NB * MRP + pos] = 0.0;
                hs[pos] = hv;
            }
            s0 = fmax(s0, -hv);
        }
        cnt += __popc(bal);
    }
    if (__any_sync(kFull, viol)) return out;
    // ---- simplex rows (-beta_i <= 0, sum beta <= 1) and the two epigraph rows ---------------
    const int nextra = NX + 1 + (tau ? 2 : 0);
    if (cnt + nextra > MRP) {
        out.status = -1;
        return out;
    }
    if (lane < nextra) {
        const int pos = cnt + lane;
        double co[NB], last = 0.0, hv = 0.0;
#pragma unroll
        for (int k = 0; k < NB; ++k) co[k] = 0.0;
        if (lane < NX) {
#pragma unroll
            for (int k = 0; k < NX; ++k) co[U + k] = (k == lane) ? -1.0 : 0.0;
        } else if (lane == NX) {
#pragma unroll
            for (int k = 0; k < NX; ++k) co[U + k] = 1.0;
            hv = 1.0;
        } else {
            // bar_V(beta) = Vr0 + dc'beta;  tau' scaled by cs   (oracle.py:88-95)
            const double k2 = (lane == NX + 1) ? 1.0 : 1.0 / (1.0 + tb.eps_r);
            double n1 = cs;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double dc = (Vs[i + 1] - Vs[0]) * k2;
                co[U + i] = -dc;
                n1 = fmax(n1, fabs(dc));
            }
            const double inv = 1.0 / n1;
#pragma unroll
            for (int i = 0; i < NX; ++i) co[U + i] *= inv;
            last = -cs * inv;
            hv = ((lane == NX + 1) ? Vs[0] - tb.eps_a : Vs[0] * k2) * inv;
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) Gs[k * MRP + pos] = co[k];
        Gs[NB * MRP + pos] = last;
        hs[pos] = hv;
        if (lane <= NX) s0 = fmax(s0, -hv);
    }
    const int total = cnt + nextra, n1rows = cnt + NX + 1;
    clear_rows<NVS, MRC>(Gs, hs, total, lane);
    s0 = warp_max(s0);
    __syncwarp();
    double obj;
    if (phase1 || mode == EHM_SX_ALIVE) {
        // sigma column: -1 on the phase-1 rows, restored to 0 afterwards (the epigraph rows keep
        // their tau coefficient: they sit behind n1rows and are masked out of this solve)
        const unsigned act1 = act_of<MRC>(n1rows, lane);
#pragma unroll
        for (int j = 0; j < MRC; ++j)
            if (act1 >> j & 1u) Gs[NB * MRP + lane + 32 * j] = -1.0;
        for (int e = lane; e < NVS * NVS; e += 32) Hs[e] = 0.0;
        if (lane < 16) {
            fs[lane] = (lane == NB) ? 1.0 : 0.0;
            zs[lane] = (lane == NB) ? s0 + 1.0 : 0.0;
        }
        __syncwarp();
        const int st = run_ipm<NVS, MRC>(ws, act1, lane, out, obj, true);
        if (st != 0) {
            out.status = -1;
            return out;
        }
        if (mode == EHM_SX_ALIVE) {
            out.val = obj;
            out.status = (obj <= -EHM_TAU_INT) ? 1 : 0;
            return out;
        }
        if (!(obj <= -EHM_TAU_INT)) return out;
#pragma unroll
        for (int j = 0; j < MRC; ++j)
            if (act1 >> j & 1u) Gs[NB * MRP + lane + 32 * j] = 0.0;
        __syncwarp();
    }
    // ---- cost in [u; beta; 1]: Mp = T' M T,  [u;theta;1] = T [u;beta;1] ----------------
    const double *Mc = tb.M + (size_t)c * Y * Y;
    for (int e = lane; e < Y * Y; e += 32) {
        Ms[e] = Mc[e];
        const int a = e / Y, b = e % Y;
        double t = 0.0;
        if (a < U) t = (a == b) ? 1.0 : 0.0;
        else if (a < U + NX) {
            const int jj = a - U;
            if (b >= U && b < U + NX) t = Es[jj * NX + (b - U)];
            else if (b == Y - 1) t = Rs[jj];
        } else t = (b == Y - 1) ? 1.0 : 0.0;
        Ts[e] = t;
    }
    __syncwarp();
    for (int e = lane; e < Y * Y; e += 32) {
        const int a = e / Y, b = e % Y;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < Y; ++k) v = fma(Ms[a * Y + k], Ts[k * Y + b], v);
        MTs[e] = v;
    }
    __syncwarp();
    for (int e = lane; e < Y * Y; e += 32) {  // Mp into Ms
        const int a = e / Y, b = e % Y;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < Y; ++k) v = fma(Ts[k * Y + a], MTs[k * Y + b], v);
        Ms[e] = v;
    }
    __syncwarp();
    const double cst = Ms[(Y - 1) * Y + Y - 1];
    if (!tau) {
        double v = 0.0;
        if (lane < NB) v = fmax(2.0 * fabs(Ms[lane * Y + lane]), 2.0 * fabs(Ms[lane * Y + Y - 1]));
        cs = fmax(1.0, warp_max(v));
    }
    for (int e = lane; e < NVS * NVS; e += 32) {
        const int a = e / NVS, b = e % NVS;
        Hs[e] = (a < NB && b < NB) ? 2.0 * Ms[a * Y + b] / cs : 0.0;
    }
    if (lane < 16) {
        fs[lane] = (lane < NB) ? 2.0 * Ms[lane * Y + Y - 1] / cs : ((lane == NB && tau) ? 1.0 : 0.0);
        zs[lane] = 0.0;
    }
    __syncwarp();
    const unsigned act = act_of<MRC>(total, lane);
    const int st = run_ipm<NVS, MRC>(ws, act, lane, out, obj);
    if (st != 0) {
        out.status = -1;
        return out;
    }
    out.status = 1;
    out.val = cs * obj + cst;
    if (plane != nullptr && tau) {
        const double *lv = ws + W::oL;
        double gb[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) gb[i] = 0.0;
#pragma unroll 1
        for (int j = 0; j < MRC; ++j) {
            const int r = lane + 32 * j;
            if (r < cnt) {
                const double l = lv[r];
#pragma unroll
                for (int i = 0; i < NX; ++i) gb[i] = fma(l, Gs[(U + i) * MRP + r], gb[i]);
            }
        }
        double sg = 0.0, p0 = out.val - zs[NB] * cs;
        const double vst = p0;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = warp_sum(gb[i]) + fs[U + i];
#pragma unroll
            for (int k = 0; k < NVS; ++k) v = fma(Hs[(U + i) * NVS + k], zs[k], v);
            gb[i] = cs * v;
            sg += fabs(gb[i]);
            p0 = fma(-gb[i], zs[U + i], p0);
        }
        const double mg = 1e-6 * (1.0 + fabs(vst) + sg);
        if (lane == 0) plane[0] = __double2float_rd(p0 - mg);
        if (lane >= 1 && lane <= NX) {
            double gl = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) gl = (lane == i + 1) ? gb[i] : gl;
            plane[lane] = __double2float_rd(p0 + gl - mg);
        }
    }
    if (sol != nullptr) {
        if (lane < NX) sol[lane] = zs[U + lane];
        if (lane == NX) sol[NX] = tau ? out.val - zs[NB] * cs : out.val;
    }
    __syncwarp();
    return out;
}

}  // namespace ehm
