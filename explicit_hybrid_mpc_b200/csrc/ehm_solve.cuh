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
// cost; its Cholesky pivot is frozen by the guard) share the last column.  That keeps the hot
// code of a kernel inside the instruction cache.
//
// Candidate tables live in global memory (L2 resident: 60 MB for the pendulum), column-major
// with the row index fastest so that a warp's load of one column is one coalesced 256 B
// segment per 32 rows.
#pragma once
#include "ehm_ipm.cuh"

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

// Shared-memory workspace of one warp: the solver block of the widest problem; the assembly
// scratch (cost transform, simplex geometry) aliases the per-row state vectors, which are
// only live inside ipm_solve.
template <int U, int NX, int MR>
struct WarpWs {
    static constexpr int NVS = U + NX + 1;
    static constexpr int NVP = U + 1;
    using WS = IpmWs<NVS, MR>;
    using WP = IpmWs<NVP, MR>;
    static constexpr int MRP = 32 * MR;
    static constexpr int Y = U + NX + 1;
    static constexpr int kDoubles = WS::kDoubles;
    // scratch inside the simplex block's s/lam/a/b region (4*MRP doubles)
    static constexpr int oM = WS::oS;               // candidate M, later Mp (Y*Y)
    static constexpr int oT = oM + Y * Y;           // T (Y*Y)
    static constexpr int oMT = oT + Y * Y;          // M*T (Y*Y)
    static constexpr int oR = oMT + Y * Y;          // simplex vertices (NX+1)*NX
    static constexpr int oE = oR + (NX + 1) * NX;   // E NX*NX
    static constexpr int oV = oE + NX * NX;         // Vr (NX+1), padded 16
    static_assert(oV + 16 <= WS::oHs, "assembly scratch must fit the row-state region");
    static_assert(WP::kDoubles <= WS::kDoubles, "point block must fit the simplex block");
};

struct SolveOut {
    int status;   // 1 ok / feasible, 0 infeasible / not alive, -1 numerical failure
    double val;   // J, sigma*, t* or min V
    int iters;    // IPM iterations spent
    int nsolves;  // IPM solves spent
    double flops; // algorithmic flops spent (see ipm_solve)
};

template <int NV, int MR>
__device__ __forceinline__ int run_ipm(double *ws, unsigned act, int lane, SolveOut &out, double &obj)
{
    using W = IpmWs<NV, MR>;
    const int st = ipm_solve<NV, MR>(ws, act, lane);
    obj = ws[W::oOut + 0];
    out.iters += (int)ws[W::oOut + 1];
    out.flops += ws[W::oOut + 2];
    out.nsolves++;
    __syncwarp();
    return st;
}

// ---------------------------------------------------------------------------------------
// P_theta_delta at one theta for candidate c.   u0 (nu values) written to u0_out by lane 0.
// ---------------------------------------------------------------------------------------
template <int U, int NX, int MR>
__device__ SolveOut solve_point(const DevTables &tb, int c, const double *__restrict__ theta,
                                double *ws, double *u0_out, int lane)
{
    using W = typename WarpWs<U, NX, MR>::WP;
    constexpr int MRP = W::MRP, NV = U + 1;
    double *Gs = ws + W::oG, *hs = ws + W::oH, *Hs = ws + W::oHs, *fs = ws + W::oF, *zs = ws + W::oZ;
    SolveOut out{0, 0.0, 0, 0, 0.0};
    const int MP = tb.MP, m = tb.m, Y = tb.Y;
    const double *Gc = tb.G + (size_t)c * U * MP, *Sc = tb.S + (size_t)c * NX * MP,
                 *gc = tb.g + (size_t)c * MP, *rc = tb.rho + (size_t)c * MP;
    double th[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) th[i] = theta[i];
    unsigned act = 0;
    bool viol = false;
    double s0 = 0.0;
#pragma unroll 1
    for (int j = 0; j < MR; ++j) {
        const int r = lane + 32 * j;
        if (r < m) {
            double g[U], nrm = 0.0;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                g[k] = Gc[k * MP + r];
                nrm = fmax(nrm, fabs(g[k]));
            }
            double hv = gc[r];
#pragma unroll
            for (int i = 0; i < NX; ++i) hv = fma(Sc[i * MP + r], th[i], hv);
            const double rho = rc[r];
            const double sc = rho > 0.0 ? rho : 1.0;
            if (nrm <= EHM_ZERO_ROW * sc) {
                if (hv < -EHM_TAU_CONST * sc) viol = true;
            } else {
                const double inv = 1.0 / nrm;
#pragma unroll
                for (int k = 0; k < U; ++k) Gs[k * MRP + r] = g[k] * inv;
                Gs[U * MRP + r] = -1.0;
                hs[r] = hv * inv;
                s0 = fmax(s0, -hv * inv);
                act |= 1u << j;
            }
        }
        if (!(act >> j & 1u)) {  // rows that do not take part: G = 0, h = 1 (ipm_solve contract)
#pragma unroll
            for (int k = 0; k < NV; ++k) Gs[k * MRP + r] = 0.0;
            hs[r] = 1.0;
        }
    }
    if (__any_sync(kFull, viol)) return out;
    s0 = warp_max(s0);
    // ---- phase 1: min sigma s.t. G u - sigma <= h --------------------------------------
    for (int e = lane; e < NV * NV; e += 32) Hs[e] = 0.0;
    if (lane < 16) {
        fs[lane] = (lane == U) ? 1.0 : 0.0;
        zs[lane] = (lane == U) ? s0 + 1.0 : 0.0;
    }
    __syncwarp();
    double obj;
    int st = run_ipm<NV, MR>(ws, act, lane, out, obj);
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
#pragma unroll 1
        for (int j = 0; j < MR; ++j)
            if (act >> j & 1u) Gs[U * MRP + lane + 32 * j] = 0.0;
        cst = Mc[(Y - 1) * Y + Y - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            cst = fma(2.0 * Mc[(U + i) * Y + Y - 1], th[i], cst);
#pragma unroll
            for (int k = 0; k < NX; ++k) cst = fma(th[i] * Mc[(U + i) * Y + U + k], th[k], cst);
        }
        __syncwarp();
    }
    st = run_ipm<NV, MR>(ws, act, lane, out, obj);
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
// ---------------------------------------------------------------------------------------
template <int U, int NX, int MR>
__device__ SolveOut solve_simplex(const DevTables &tb, int c, const double *__restrict__ Rv,
                                  const double *__restrict__ Vr, int mode, bool phase1,
                                  double *ws, double *sol, int lane)
{
    using WW = WarpWs<U, NX, MR>;
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
    unsigned act = 0, tau_bits = 0;
    bool viol = false;
    double s0 = 0.0;
#pragma unroll 1
    for (int j = 0; j < MR; ++j) {
        const int r = lane + 32 * j;
        double co[NB], hv = 0.0, rho = 1.0, last = 0.0;
        bool have = false;
#pragma unroll
        for (int k = 0; k < NB; ++k) co[k] = 0.0;
        if (r < m) {
#pragma unroll
            for (int k = 0; k < U; ++k) co[k] = Gc[k * MP + r];
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
            }
            rho = rc[r];
            have = true;
        } else if (r < m + NX) {  // -beta_i <= 0
            const int i = r - m;
#pragma unroll
            for (int k = 0; k < NX; ++k) co[U + k] = (k == i) ? -1.0 : 0.0;
            have = true;
        } else if (r == m + NX) {  // sum beta <= 1
#pragma unroll
            for (int k = 0; k < NX; ++k) co[U + k] = 1.0;
            hv = 1.0;
            have = true;
        } else if (tau && r <= m + NX + 2) {
            // bar_V(beta) = Vr0 + dc'beta;  tau' scaled by cs   (oracle.py:88-95)
            const double k2 = (r == m + NX + 1) ? 1.0 : 1.0 / (1.0 + tb.eps_r);
            double n1 = cs;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const double dc = (Vs[i + 1] - Vs[0]) * k2;
                co[U + i] = -dc;
                n1 = fmax(n1, fabs(dc));
            }
            last = -cs;
            hv = (r == m + NX + 1) ? Vs[0] - tb.eps_a : Vs[0] * k2;
            rho = n1;
            have = true;
            tau_bits |= 1u << j;
        }
        if (have) {
            double nrm = fabs(last);
#pragma unroll
            for (int k = 0; k < NB; ++k) nrm = fmax(nrm, fabs(co[k]));
            const double sc = rho > 0.0 ? rho : 1.0;
            if (nrm <= EHM_ZERO_ROW * sc) {
                if (hv < -EHM_TAU_CONST * sc) viol = true;
            } else {
                const double inv = 1.0 / nrm;
#pragma unroll
                for (int k = 0; k < NB; ++k) Gs[k * MRP + r] = co[k] * inv;
                Gs[NB * MRP + r] = last * inv;  // tau column
                hs[r] = hv * inv;
                // rows carrying tau do not exist in the phase-1 problem
                if (!(tau_bits >> j & 1u)) s0 = fmax(s0, -hv * inv);
                act |= 1u << j;
            }
        }
        if (!(act >> j & 1u)) {  // rows that do not take part: G = 0, h = 1 (ipm_solve contract)
#pragma unroll
            for (int k = 0; k < NVS; ++k) Gs[k * MRP + r] = 0.0;
            hs[r] = 1.0;
        }
    }
    if (__any_sync(kFull, viol)) return out;
    s0 = warp_max(s0);
    __syncwarp();
    double obj;
    if (phase1 || mode == EHM_SX_ALIVE) {
        // sigma column: -1 on every phase-1 row; the tau column is restored afterwards from
        // registers (two rows at most)
        const unsigned act1 = act & ~tau_bits;
        double keep0 = 0.0;  // tau-column value of this lane's non-tau rows is 0
        (void)keep0;
#pragma unroll 1
        for (int j = 0; j < MR; ++j)
            if (act1 >> j & 1u) Gs[NB * MRP + lane + 32 * j] = -1.0;
        for (int e = lane; e < NVS * NVS; e += 32) Hs[e] = 0.0;
        if (lane < 16) {
            fs[lane] = (lane == NB) ? 1.0 : 0.0;
            zs[lane] = (lane == NB) ? s0 + 1.0 : 0.0;
        }
        __syncwarp();
        const int st = run_ipm<NVS, MR>(ws, act1, lane, out, obj);
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
#pragma unroll 1
        for (int j = 0; j < MR; ++j)
            if (act1 >> j & 1u) Gs[NB * MRP + lane + 32 * j] = 0.0;
        // the row-state region was clobbered by the solver: reload the geometry scratch
        for (int e = lane; e < (NX + 1) * NX; e += 32) Rs[e] = Rv[e];
        __syncwarp();
        for (int e = lane; e < NX * NX; e += 32) {
            const int j = e / NX, i = e % NX;
            Es[e] = Rs[(i + 1) * NX + j] - Rs[j];
        }
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
    const int st = run_ipm<NVS, MR>(ws, act, lane, out, obj);
    if (st != 0) {
        out.status = -1;
        return out;
    }
    out.status = 1;
    out.val = cs * obj + cst;
    if (sol != nullptr) {
        if (lane < NX) sol[lane] = zs[U + lane];
        if (lane == NX) sol[NX] = tau ? out.val - zs[NB] * cs : out.val;
    }
    __syncwarp();
    return out;
}

}  // namespace ehm
