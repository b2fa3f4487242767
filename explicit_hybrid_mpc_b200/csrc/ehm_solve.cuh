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
// cost; its Cholesky pivot is frozen by the guard; it rides for free in the 8-wide DMMA blocks)
// share the last column.
//
// Row pre-solve.  The condensed problems carry every big-M row of the reference
// (lib/mpc_library.py:529-534): 210 rows for the pendulum, of which ~160 can never bind inside
// the box |uhat| <= 1 the input-bound rows impose.  A normalised row a'z <= h is dropped when
//      sup_{|u|<=1, beta in simplex} a'z  =  |a_u|_1 + max(0, max_i a_beta_i)  <=  h - TAU_DROP,
// which is exact for every decision the engine takes (a point that satisfies the kept rows,
// which include the box rows, with margin TAU_INT satisfies the dropped ones with margin
// TAU_DROP > TAU_INT; optima are unchanged because the feasible set is).  The surviving rows
// (34-61 for the pendulum) are compacted with warp ballots into the row-major tile of
// ehm_ipm.cuh (<= CAP rows).
//
// Candidate tables live in global memory (L2 resident: 60 MB for the pendulum), column-major
// with the row index fastest so that a warp's load of one column is one coalesced 256 B
// segment per 32 rows.  The candidate's cost matrix M (one contiguous block) is staged by TMA
// (cp.async.bulk + mbarrier, SASS UBLKCP) behind the row assembly and the phase-1 solve when its size
// is a multiple of 16 bytes; measured effect: none (profiles/r2_tma_before_after.txt).
#pragma once
#include "ehm_ipm.cuh"

#define EHM_TAU_DROP 1e-3

namespace ehm {

struct DevTables {
    int n, nu, N, K, p;
    int U, m, MP, Y, ncand;  // U=N*nu, m=N*p, MP = rows padded to 32, Y=U+n+1
    int is_lp;               // cost has no quadratic part (Q = P = R = 0)
    const double *G;         // [ncand][U][MP]
    const double *S;         // [ncand][n][MP]
    const double *g;         // [ncand][MP]
    const double *rho;       // [ncand][MP]
    const double *M;         // [ncand][Y*Y]
    double eps_a, eps_r;
};

// Shared-memory workspace of one warp: the solver block of the simplex problems (the point
// problems reuse it with their own, smaller layout) + the simplex geometry.
template <int U, int NX, int CAP>
struct WarpWs {
    static constexpr int NVS = U + NX + 1;
    static constexpr int NVP = U + 1;
    using WS = IpmWs<NVS, CAP>;
    using WP = IpmWs<NVP, CAP>;
    static constexpr int Y = U + NX + 1;
    static constexpr int oR = WS::kDoubles;          // simplex vertices (NX+1)*NX
    static constexpr int oE = oR + (NX + 1) * NX;    // E NX*NX
    static constexpr int oV = oE + NX * NX;          // Vr (NX+1)
    static constexpr int oW1 = oV + NX + 1;          // M_thth E (NX*NX)
    static constexpr int oW2 = oW1 + NX * NX;        // M_thth v0 + m_th (NX), theta scratch (NX)
    // TMA staging of the candidate's cost matrix M (Y*Y doubles, one contiguous block in the
    // table): cp.async.bulk needs 16-byte sizes and addresses, i.e. an even Y
    static constexpr bool kTma = (Y * Y * 8) % 16 == 0;
    static constexpr int oMst = (oW2 + 2 * NX + 1) & ~1;            // staged M (16-byte aligned)
    static constexpr int oBar = oMst + (kTma ? Y * Y : 0);         // mbarrier (8 bytes)
    static constexpr int kDoubles = (oBar + (kTma ? 1 : 0) + 1) & ~1;
    static_assert(WP::kDoubles <= WS::kDoubles, "point block must fit the simplex block");
};

// ---- TMA (bulk async copy) of one contiguous global block into shared memory ------------------
// One elected lane arms the warp's mbarrier with the byte count and issues cp.async.bulk; the
// copy runs behind the row assembly and the phase-1 solve, every lane waits on the barrier's
// phase before the first use.  SASS: UBLKCP + SYNCS.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_bar_init(void *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// once per warp and kernel: arm the warp's mbarrier; returns the initial phase parity
template <int U, int NX, int CAP>
__device__ __forceinline__ unsigned warp_ws_init(double *ws, int lane);

struct SolveOut {
    int status;   // 1 ok / feasible, 0 infeasible / not alive, -1 numerical failure
    double val;   // J, sigma*, t* or min V
    int iters;    // IPM iterations spent
    int nsolves;  // IPM solves spent
    double flops; // algorithmic flops spent (see ipm_solve)
};

template <int U, int NX, int CAP>
__device__ __forceinline__ unsigned warp_ws_init(double *ws, int lane)
{
    using WW = WarpWs<U, NX, CAP>;
    if constexpr (WW::kTma) {
        if (lane == 0) tma_bar_init(ws + WW::oBar);
        __syncwarp();
    }
    return 0u;
}

template <int NV, int CAP>
__device__ __forceinline__ int run_ipm(double *ws, int m, int lane, bool hasH, SolveOut &out, double &obj,
                                       bool early = false)
{
    using W = IpmWs<NV, CAP>;
    const int st = ipm_solve<NV, CAP>(ws, m, lane, hasH, early);
    obj = ws[W::oOut + 0];
    out.iters += (int)ws[W::oOut + 1];
    out.flops += ws[W::oOut + 2];
    out.nsolves++;
    __syncwarp();
    return st;
}

// zero the coefficients of rows [from, roundup4(from)) (the ipm_solve contract: finite data in
// the partial last DMMA group)
template <int NV, int CAP>
__device__ __forceinline__ void clear_tail(double *T, int from, int lane)
{
    using W = IpmWs<NV, CAP>;
    const int to = (from + 3) & ~3;
    for (int e = lane; e < (to - from) * W::TW; e += 32) T[from * W::TW + e] = 0.0;
}

// ---------------------------------------------------------------------------------------
// P_theta_delta at one theta for candidate c.   u0 (nu values) written to u0_out.
// ---------------------------------------------------------------------------------------
template <int U, int NX, int CAP>
__device__ SolveOut solve_point(const DevTables &tb, int c, const double *__restrict__ theta,
                                double *ws, double *u0_out, int lane)
{
    using W = typename WarpWs<U, NX, CAP>::WP;
    constexpr int TW = W::TW, NV = U + 1;
    double *T = ws + W::oT, *Hs = ws + W::oHs, *fs = ws + W::oF, *zs = ws + W::oZ;
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
    for (int r0 = 0; r0 < m; r0 += 32) {
        const int r = r0 + lane;
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
            if (pos < CAP) {
#pragma unroll
                for (int k = 0; k < U; ++k) T[pos * TW + k] = g[k];
                T[pos * TW + U] = -1.0;
                T[pos * TW + W::cH] = hv;
            }
            s0 = fmax(s0, -hv);
        }
        cnt += __popc(bal);
    }
    if (__any_sync(kFull, viol)) return out;
    if (cnt > CAP) {  // capacity of the compacted tile exceeded (never for the shipped examples)
        out.status = -1;
        return out;
    }
    clear_tail<NV, CAP>(T, cnt, lane);
    s0 = warp_max(s0);
    // ---- phase 1: min sigma s.t. G u - sigma <= h --------------------------------------
    if (lane < W::NVP) {
        fs[lane] = (lane == U) ? 1.0 : 0.0;
        zs[lane] = (lane == U) ? s0 + 1.0 : 0.0;
    }
    __syncwarp();
    double obj;
    int st = run_ipm<NV, CAP>(ws, cnt, lane, false, out, obj, true);
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
        if (lane < W::NVP) {
            fs[lane] = (lane < U) ? fv / cs : 0.0;
            zs[lane] = 0.0;
        }
        if constexpr (W::kHasH) {
            for (int e = lane; e < NV * NV; e += 32) {
                const int a = e / NV, b = e % NV;
                Hs[a * W::HS + b] = (a < U && b < U) ? 2.0 * Mc[a * Y + b] / cs : 0.0;
            }
        }
        for (int r = lane; r < cnt; r += 32) T[r * TW + U] = 0.0;
        cst = Mc[(Y - 1) * Y + Y - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            cst = fma(2.0 * Mc[(U + i) * Y + Y - 1], th[i], cst);
#pragma unroll
            for (int k = 0; k < NX; ++k) cst = fma(th[i] * Mc[(U + i) * Y + U + k], th[k], cst);
        }
        __syncwarp();
    }
    st = run_ipm<NV, CAP>(ws, cnt, lane, tb.is_lp == 0, out, obj);
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
template <int U, int NX, int CAP>
__device__ SolveOut solve_simplex(const DevTables &tb, int c, const double *__restrict__ Rv,
                                  const double *__restrict__ Vr, int mode, bool phase1,
                                  double *ws, double *sol, int lane, unsigned &tma_parity, float *plane = nullptr)
{
    using WW = WarpWs<U, NX, CAP>;
    using W = typename WW::WS;
    constexpr int TW = W::TW, NB = U + NX, NVS = U + NX + 1, Y = WW::Y, HS = W::HS;
    double *T = ws + W::oT, *Hs = ws + W::oHs, *fs = ws + W::oF, *zs = ws + W::oZ, *Rs = ws + WW::oR,
           *Es = ws + WW::oE, *Vs = ws + WW::oV, *W1 = ws + WW::oW1, *W2 = ws + WW::oW2;
    SolveOut out{0, 0.0, 0, 0, 0.0};
    const int MP = tb.MP, m = tb.m;
    const double *Gc = tb.G + (size_t)c * U * MP, *Sc = tb.S + (size_t)c * NX * MP,
                 *gc = tb.g + (size_t)c * MP, *rc = tb.rho + (size_t)c * MP;
    // the candidate's cost matrix travels by TMA behind the row assembly and the phase-1 solve
    const double *Mg = tb.M + (size_t)c * Y * Y;
    bool tma_pending = false;
    if constexpr (WW::kTma) {
        if (mode != EHM_SX_ALIVE) {
            if (lane == 0) tma_load_1d(ws + WW::oMst, Mg, Y * Y * 8, ws + WW::oBar);
            tma_pending = true;
        }
    }
    // every exit drains the copy (a CTA must not retire, nor the barrier be re-armed, with a bulk
    // copy in flight); tma_parity = phase parity of the next completion of this warp's barrier
    auto leave = [&](const SolveOut &o) {
        if constexpr (WW::kTma) {
            if (tma_pending) {
                tma_wait(ws + WW::oBar, tma_parity);
                tma_parity ^= 1u;
            }
        }
        __syncwarp();
        return o;
    };
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
    for (int r0 = 0; r0 < m; r0 += 32) {
        const int r = r0 + lane;
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
            if (pos < CAP) {
#pragma unroll
                for (int k = 0; k < NB; ++k) T[pos * TW + k] = co[k];
                T[pos * TW + NB] = 0.0;
                T[pos * TW + W::cH] = hv;
            }
            s0 = fmax(s0, -hv);
        }
        cnt += __popc(bal);
    }
    if (__any_sync(kFull, viol)) return leave(out);
    // ---- simplex rows (-beta_i <= 0, sum beta <= 1) and the two epigraph rows ---------------
    const int nextra = NX + 1 + (tau ? 2 : 0);
    if (cnt + nextra > CAP) {
        out.status = -1;
        return leave(out);
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
        for (int k = 0; k < NB; ++k) T[pos * TW + k] = co[k];
        T[pos * TW + NB] = last;
        T[pos * TW + W::cH] = hv;
        if (lane <= NX) s0 = fmax(s0, -hv);
    }
    const int total = cnt + nextra, n1rows = cnt + NX + 1;
    clear_tail<NVS, CAP>(T, total, lane);
    s0 = warp_max(s0);
    __syncwarp();
    double obj;
    if (phase1 || mode == EHM_SX_ALIVE) {
        // sigma column: -1 on the phase-1 rows, restored to 0 afterwards (the epigraph rows keep
        // their tau coefficient: they sit behind n1rows and take no part in this solve)
        for (int r = lane; r < n1rows; r += 32) T[r * TW + NB] = -1.0;
        if (lane < W::NVP) {
            fs[lane] = (lane == NB) ? 1.0 : 0.0;
            zs[lane] = (lane == NB) ? s0 + 1.0 : 0.0;
        }
        __syncwarp();
        const int st = run_ipm<NVS, CAP>(ws, n1rows, lane, false, out, obj, true);
        if (st != 0) {
            out.status = -1;
            return leave(out);
        }
        if (mode == EHM_SX_ALIVE) {
            out.val = obj;
            out.status = (obj <= -EHM_TAU_INT) ? 1 : 0;
            return leave(out);
        }
        if (!(obj <= -EHM_TAU_INT)) return leave(out);
        for (int r = lane; r < n1rows; r += 32) T[r * TW + NB] = 0.0;
        __syncwarp();
    }
    // ---- cost in [u; beta; 1]:  V = x'Mx, x = [u; theta; 1], theta = v0 + E beta ---------------
    //   H_uu = 2 Muu, H_ub = 2 Mut E, H_bb = 2 E' Mtt E, f_u = 2 (Mut v0 + mu),
    //   f_b = 2 E' (Mtt v0 + mt), const = v0' Mtt v0 + 2 mt' v0 + m0
    const double *Mc = Mg;
    if constexpr (WW::kTma) {
        tma_wait(ws + WW::oBar, tma_parity);
        tma_parity ^= 1u;
        tma_pending = false;
        Mc = ws + WW::oMst;
    }
    for (int e = lane; e < NX * NX + NX; e += 32) {
        if (e < NX * NX) {
            const int a = e / NX, i = e % NX;   // W1[a][i] = sum_j Mtt[a][j] E[j][i]
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) v = fma(Mc[(U + a) * Y + U + j], Es[j * NX + i], v);
            W1[e] = v;
        } else {
            const int a = e - NX * NX;          // W2[a] = sum_j Mtt[a][j] v0[j] + mt[a]
            double v = Mc[(U + a) * Y + Y - 1];
#pragma unroll
            for (int j = 0; j < NX; ++j) v = fma(Mc[(U + a) * Y + U + j], Rs[j], v);
            W2[a] = v;
        }
    }
    __syncwarp();
    double cst = Mc[(Y - 1) * Y + Y - 1];
#pragma unroll
    for (int j = 0; j < NX; ++j) cst = fma(Rs[j], W2[j] + Mc[(U + j) * Y + Y - 1], cst);
    // unscaled H (NB x NB) and f (NB) entries, one per lane and round
    double fmine = 0.0;
    if (lane < NB) {
        if (lane < U) {
            fmine = Mc[lane * Y + Y - 1];
#pragma unroll
            for (int j = 0; j < NX; ++j) fmine = fma(Mc[lane * Y + U + j], Rs[j], fmine);
        } else {
#pragma unroll
            for (int j = 0; j < NX; ++j) fmine = fma(Es[j * NX + lane - U], W2[j], fmine);
        }
        fmine *= 2.0;
    }
    if constexpr (W::kHasH) {
        for (int e = lane; e < NVS * NVS; e += 32) {
            const int a = e / NVS, b = e % NVS;
            double v = 0.0;
            if (a < NB && b < NB) {
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                if (hi < U) v = Mc[a * Y + b];
                else if (lo < U) {   // u-beta block: sum_j Mut[lo][j] E[j][hi-U]
#pragma unroll
                    for (int j = 0; j < NX; ++j) v = fma(Mc[lo * Y + U + j], Es[j * NX + hi - U], v);
                } else {             // beta-beta block: sum_j E[j][a-U] W1[j][b-U]
#pragma unroll
                    for (int j = 0; j < NX; ++j) v = fma(Es[j * NX + a - U], W1[j * NX + b - U], v);
                }
                v *= 2.0;
            }
            Hs[a * HS + b] = v;
        }
        __syncwarp();
    }
    if (!tau) {
        double v = 0.0;
        if (lane < NB) v = fmax(W::kHasH ? fabs(Hs[lane * HS + lane]) : 0.0, fabs(fmine));
        cs = fmax(1.0, warp_max(v));
    }
    {
        const double ics = 1.0 / cs;
        if constexpr (W::kHasH)
            for (int e = lane; e < NVS * NVS; e += 32) Hs[(e / NVS) * HS + e % NVS] *= ics;
        if (lane < W::NVP) {
            fs[lane] = (lane < NB) ? fmine * ics : ((lane == NB && tau) ? 1.0 : 0.0);
            zs[lane] = 0.0;
        }
    }
    __syncwarp();
    const int st = run_ipm<NVS, CAP>(ws, total, lane, tb.is_lp == 0, out, obj);
    if (st != 0) {
        out.status = -1;
        return leave(out);
    }
    out.status = 1;
    out.val = cs * obj + cst;
    if (plane != nullptr && tau) {
        double gb[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) gb[i] = 0.0;
        for (int r = lane; r < cnt; r += 32) {
            const double l = T[r * TW + W::cW];
#pragma unroll
            for (int i = 0; i < NX; ++i) gb[i] = fma(l, T[r * TW + U + i], gb[i]);
        }
        double sg = 0.0, p0 = out.val - zs[NB] * cs;
        const double vst = p0;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = warp_sum(gb[i]) + fs[U + i];
            if constexpr (W::kHasH)
                if (tb.is_lp == 0)
                    for (int k = 0; k < NVS; ++k) v = fma(Hs[(U + i) * HS + k], zs[k], v);
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
