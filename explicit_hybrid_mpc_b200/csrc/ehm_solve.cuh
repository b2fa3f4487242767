// ehm_solve.cuh -- per-warp assembly + solution of the oracle sub-problems.
//
// Sub-problems (reference lib/oracle.py; algebra as in oracle/ehm_oracle.c, written
// independently here for the warp-parallel layout):
//   point   : P_theta_delta (:141-173)            variables u            (U)
//   alive   : phase-1 over a simplex              variables [u;beta;sig] (U+NX+1)
//   tqp     : eps-suboptimality epigraph problem  variables [u;beta;tau] (U+NX+1)  (:88-97)
//   minv    : min over simplex of V*_c            variables [u;beta]     (U+NX)    (:75-79)
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

// Shared-memory workspace of one warp (all sizes in doubles, compile-time).
template <int U, int NX, int MR>
struct WarpWs {
    static constexpr int NVS = U + NX + 1;
    static constexpr int MRP = 32 * MR;
    static constexpr int Y = U + NX + 1;
    static constexpr int oG = 0;
    static constexpr int oH = oG + NVS * MRP;    // hs
    static constexpr int oHs = oH + MRP;         // Hessian NVS*NVS
    static constexpr int oF = oHs + NVS * NVS;   // f (padded 16)
    static constexpr int oRed = oF + 16;         // 64
    static constexpr int oM = oRed + 64;         // candidate M (Y*Y)
    static constexpr int oT = oM + Y * Y;        // T (Y*Y)
    static constexpr int oMT = oT + Y * Y;       // M*T (Y*Y)
    static constexpr int oR = oMT + Y * Y;       // simplex vertices (NX+1)*NX
    static constexpr int oE = oR + (NX + 1) * NX;  // E NX*NX
    static constexpr int oV = oE + NX * NX;      // Vr (NX+1), padded 16
    static constexpr int kDoubles = oV + 16;
};

struct SolveOut {
    int status;   // 1 ok / feasible, 0 infeasible / not alive, -1 numerical failure
    double val;   // J, sigma*, t* or min V
    int iters;    // IPM iterations spent
    int nsolves;  // IPM solves spent
    double flops; // algorithmic flops spent (see ipm_solve)
};

// ---------------------------------------------------------------------------------------
// P_theta_delta at one theta for candidate c.   u0 (nu values) written to u0_out by lane 0.
// skip_phase1: the caller knows the point is strictly feasible (midpoints, vertices).
// ---------------------------------------------------------------------------------------
template <int U, int NX, int MR>
__device__ SolveOut solve_point(const DevTables &tb, int c, const double *__restrict__ theta,
                                double *ws, double *u0_out, int lane)
{
    using W = WarpWs<U, NX, MR>;
    constexpr int MRP = W::MRP;
    double *Gs = ws + W::oG, *hs = ws + W::oH, *Hs = ws + W::oHs, *fs = ws + W::oF,
           *red = ws + W::oRed;
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
#pragma unroll
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
    }
    if (__any_sync(kFull, viol)) return out;
    s0 = warp_max(s0);
    __syncwarp();
    // ---- phase 1: min sigma s.t. G u - sigma <= h --------------------------------------
    {
        if (lane < 16) fs[lane] = (lane == U) ? 1.0 : 0.0;
        __syncwarp();
        double z[U + 1];
#pragma unroll
        for (int k = 0; k < U; ++k) z[k] = 0.0;
        z[U] = s0 + 1.0;
        double obj;
        int it;
        const int st = ipm_solve<U + 1, MR, false>(Gs, hs, act, Hs, fs, red, z, obj, it, lane, &out.flops);
        out.iters += it;
        out.nsolves++;
        if (st != 0) {
            out.status = -1;
            return out;
        }
        if (!(obj <= -EHM_TAU_INT)) return out;
    }
    // ---- phase 2: the QP ------------------------------------------------------------------
    const double *Mc = tb.M + (size_t)c * Y * Y;
    double cs = 1.0, cst = 0.0;
    {
        // H = 2 M_uu, f = 2 (M_u,theta theta + M_u,1)
        double fv = 0.0, hd = 0.0;
        if (lane < U) {
            fv = Mc[lane * Y + Y - 1];
#pragma unroll
            for (int i = 0; i < NX; ++i) fv = fma(Mc[lane * Y + U + i], th[i], fv);
            fv *= 2.0;
            hd = 2.0 * Mc[lane * Y + lane];
        }
        cs = fmax(1.0, warp_max(fmax(fabs(fv), fabs(hd))));
        if (lane < U) fs[lane] = fv / cs;
        for (int e = lane; e < U * U; e += 32) Hs[e] = 2.0 * Mc[(e / U) * Y + (e % U)] / cs;
        cst = Mc[(Y - 1) * Y + Y - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            cst = fma(2.0 * Mc[(U + i) * Y + Y - 1], th[i], cst);
#pragma unroll
            for (int k = 0; k < NX; ++k) cst = fma(th[i] * Mc[(U + i) * Y + U + k], th[k], cst);
        }
        __syncwarp();
    }
    double z[U];
#pragma unroll
    for (int k = 0; k < U; ++k) z[k] = 0.0;
    double obj;
    int it;
    const int st = ipm_solve<U, MR, true>(Gs, hs, act, Hs, fs, red, z, obj, it, lane, &out.flops);
    out.iters += it;
    out.nsolves++;
    if (st != 0) {
        out.status = -1;
        return out;
    }
    out.status = 1;
    out.val = cs * obj + cst;
    if (u0_out != nullptr && lane == 0) {
#pragma unroll
        for (int k = 0; k < U; ++k)
            if (k < tb.nu) u0_out[k] = z[k];
    }
    return out;
}

// ---------------------------------------------------------------------------------------
// Simplex problems.  Rv: (NX+1)*NX vertices (global), Vr: NX+1 vertex costs (global, TQP).
// mode: EHM_SX_ALIVE / EHM_SX_TQP / EHM_SX_MINV;  phase1: run the alive test first.
// sol (optional, global, NX+1 doubles): [beta*, V*] written by lane 0 for TQP / MINV.
// ---------------------------------------------------------------------------------------
template <int U, int NX, int MR>
__device__ SolveOut solve_simplex(const DevTables &tb, int c, const double *__restrict__ Rv,
                                  const double *__restrict__ Vr, int mode, bool phase1,
                                  double *ws, double *sol, int lane)
{
    using W = WarpWs<U, NX, MR>;
    constexpr int MRP = W::MRP, NB = U + NX, NVS = U + NX + 1, Y = W::Y;
    double *Gs = ws + W::oG, *hs = ws + W::oH, *Hs = ws + W::oHs, *fs = ws + W::oF,
           *red = ws + W::oRed, *Ms = ws + W::oM, *Ts = ws + W::oT, *MTs = ws + W::oMT,
           *Rs = ws + W::oR, *Es = ws + W::oE, *Vs = ws + W::oV;
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
    unsigned act = 0;
    bool viol = false;
    double s0 = 0.0;
#pragma unroll
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
                Gs[NB * MRP + r] = last * inv;  // tau column (overwritten with -1 for sigma)
                hs[r] = hv * inv;
                // rows carrying tau do not exist in the phase-1 problem
                if (!(tau && r > m + NX)) s0 = fmax(s0, -hv * inv);
                act |= 1u << j;
            }
        }
    }
    if (__any_sync(kFull, viol)) return out;
    __syncwarp();
    // bits of the two tau rows in this lane's mask
    unsigned tau_bits = 0;
#pragma unroll
    for (int j = 0; j < MR; ++j) {
        const int r = lane + 32 * j;
        if (tau && r > m + NX && r <= m + NX + 2) tau_bits |= 1u << j;
    }
    if (phase1 || mode == EHM_SX_ALIVE) {
        // sigma column: -1 on every phase-1 row
        const unsigned act1 = act & ~tau_bits;
        double keep[MR];
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            const int r = lane + 32 * j;
            keep[j] = Gs[NB * MRP + r];
            if (act1 >> j & 1u) Gs[NB * MRP + r] = -1.0;
        }
        if (lane < 16) fs[lane] = (lane == NB) ? 1.0 : 0.0;
        s0 = warp_max(s0);
        __syncwarp();
        double z[NVS];
#pragma unroll
        for (int k = 0; k < NB; ++k) z[k] = 0.0;
        z[NB] = s0 + 1.0;
        double obj;
        int it;
        const int st = ipm_solve<NVS, MR, false>(Gs, hs, act1, Hs, fs, red, z, obj, it, lane, &out.flops);
        out.iters += it;
        out.nsolves++;
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
        for (int j = 0; j < MR; ++j) {
            const int r = lane + 32 * j;
            if (act1 >> j & 1u) Gs[NB * MRP + r] = keep[j];
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
    // Mp into Ms
    for (int e = lane; e < Y * Y; e += 32) {
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
    const int nv = tau ? NVS : NB;
    for (int e = lane; e < nv * nv; e += 32) {
        const int a = e / nv, b = e % nv;
        Hs[e] = (a < NB && b < NB) ? 2.0 * Ms[a * Y + b] / cs : 0.0;
    }
    if (lane < 16) fs[lane] = (lane < NB) ? 2.0 * Ms[lane * Y + Y - 1] / cs : (lane == NB ? 1.0 : 0.0);
    __syncwarp();
    double obj;
    int it, st;
    double zsol[NVS];
    if (tau) {
        double z[NVS];
#pragma unroll
        for (int k = 0; k < NVS; ++k) z[k] = 0.0;
        st = ipm_solve<NVS, MR, true>(Gs, hs, act, Hs, fs, red, z, obj, it, lane, &out.flops);
#pragma unroll
        for (int k = 0; k < NVS; ++k) zsol[k] = z[k];
    } else {
        double z[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) z[k] = 0.0;
        st = ipm_solve<NB, MR, true>(Gs, hs, act, Hs, fs, red, z, obj, it, lane, &out.flops);
#pragma unroll
        for (int k = 0; k < NB; ++k) zsol[k] = z[k];
        zsol[NB] = 0.0;
    }
    out.iters += it;
    out.nsolves++;
    if (st != 0) {
        out.status = -1;
        return out;
    }
    out.status = 1;
    out.val = cs * obj + cst;
    if (sol != nullptr && lane == 0) {
#pragma unroll
        for (int i = 0; i < NX; ++i) sol[i] = zsol[U + i];
        sol[NX] = tau ? out.val - zsol[NB] * cs : out.val;
    }
    return out;
}

}  // namespace ehm
