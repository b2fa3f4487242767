// ehm_ipm.cuh -- warp-level dense primal-dual interior-point solver (sm_100a, fp64).
//
// One warp solves one small convex QP/LP
//        min 0.5 z'Hz + f'z   s.t.  G z <= h            (NV variables, m <= CAP rows)
// held in shared memory + registers.  It replaces one `.solve()` of the reference
// (CVXPY -> MOSEK, lib/oracle.py:131,134,166,169,203,276,305,350).
//
// Mapping to the hardware (round 2, "v11"; history and ncu findings in profiles/README.md):
//   * one warp = one problem; 4 independent warps per CTA, 4 CTAs per SM (16 problems per SM,
//     <= 128 registers, <= 14 KB of shared memory per problem).
//   * the row tile is ROW-major with an even row stride TW = 2*odd: a lane reads its own rows
//     with conflict-free 128-bit loads; every row carries four extra columns behind its NV
//     coefficients: a = (lam/s)*rp, w = lam/s, v (corrector term) and h.
//   * per-row state (s, lam, h, rp, 1/s, ds*dl) lives in REGISTERS: lane L owns rows L, L+32, ...
//   * the normal matrix K = H + G' diag(w) G is built on the fp64 tensor path:
//     mma.sync.m8n8k4.f64 (DMMA) over groups of four rows, A = [w*G | a]' and B = G taken from
//     the same tile loads, accumulators initialised with H.  The extra A row a gives the
//     predictor's right-hand side G'(w*rp) in the same instructions.  Measured on B200
//     (scripts/ubench/dmma_bench.cu): DMMA and DFMA share one pipe (37.2 vs 34.1 TFLOP/s, no
//     overlap), so DMMA buys issue slots (1 instruction per 256 FMA) and registers (2
//     accumulators per 8x8 block instead of NV(NV+1)/2 per lane), not pipe time.
//   * Cholesky and the triangular solves are cooperative: lane i owns row i of the factor (in
//     shared memory, odd stride), pivots / solution components travel by shuffle.  The round-1
//     version factorised redundantly in every lane's registers (32x the work, 112 registers).
//   * G'v for the corrector is a lane-serial pass (lane k sums column k over the rows).
//
// The numerical method is, by specification, the one of oracle/ehm_oracle.c:qp_solve_x
// (Mehrotra predictor-corrector in the CVXOPT/ECOS form: sigma=(1-alpha_aff)^3, residuals
// scaled by 1-sigma, centred start s.lam=1, pivot guard, convergence on primal residual,
// gap s'lam and the propagated dual residual rho*rd0, the measured dual residual as a guard).
// Only the summation order differs.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/ehm.h"

#ifndef EHM_SMALL_NV
#define EHM_SMALL_NV 25   // systems up to this size: fully unrolled Cholesky / solves, own row in registers (NV = 25: -16 % vs rolled)
#endif

namespace ehm {

constexpr unsigned kFull = 0xffffffffu;

// debugging aid: per-iteration trace of the LAST solve that ran with tracing enabled
__device__ double g_trace[8 * 128];
__device__ int g_trace_on = 0;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
// max / min over the warp of NON-NEGATIVE doubles (NaN propagates as the largest pattern): the
// IEEE order of non-negative doubles is the unsigned order of their bit patterns, so two
// 32-bit REDUX instructions replace a 5-level tree of 64-bit shuffles (~70 SASS instructions).
__device__ __forceinline__ double warp_max_nonneg(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(kFull, hi);
    const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double warp_min_nonneg(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(kFull, hi);
    const unsigned ml = __reduce_min_sync(kFull, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}

// 1/x and 1/sqrt(x) for finite positive x: hardware seed (MUFU.RCP64H / RSQ64H, ~2^-23) + two
// Newton steps (error below 2 ulp), branch free -- the IEEE-rounded library versions carry a slow
// path that costs a divergence frame (BSSY/BSYNC) at every call site.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-(x * y), y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-(x * y), y, 1.0);
    return fma(0.5 * y, e, y);
}

// D(8x8) += A(8x4) * B(4x8), fp64 tensor path.  Fragments: a = A[lane/4][lane%4],
// b = B[lane%4][lane/4], c0/c1 = C[lane/4][2*(lane%4) + 0/1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Sum of one double per lane over the warp, result in every lane, with two DMMAs: the lane's
// value is its B fragment B[lane%4][lane/4], A = ones, so D[.][j] = sum of lanes 4j..4j+3 and the
// lane ends up with two of those; a second pass over their sum adds the four pairs.  6 SASS
// instructions instead of the ~60 of a 5-level 64-bit shuffle tree (see chol_small).
__device__ __forceinline__ double warp_sum_mma(double v)
{
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    dmma884(c0, c1, 1.0, v);
    dmma884(d0, d1, 1.0, c0 + c1);
    return d0;
}

constexpr int tile_width(int nv)
{
    int t = ((nv + 1) & ~1) + 4;
    while (t % 2 != 0 || (t / 2) % 2 == 0) ++t;   // even, half of it odd: conflict-free 128-bit row reads
    return t;
}

// Shapes compiled for LP costs only keep no Hessian in shared memory (5 KB of the 39 KB a problem
// of the synthetic n=8 shape needs: 6 instead of 5 resident problems per SM).  Keyed on the
// (variables, row capacity) pair of the solver block; ehm_problem_create refuses a QP cost there.
template <int NV, int CAP>
struct IpmHasH { static constexpr bool value = true; };
template <> struct IpmHasH<25, 92> { static constexpr bool value = false; };
template <> struct IpmHasH<17, 92> { static constexpr bool value = false; };

// Per-warp shared-memory block of the solver (offsets in doubles from the block base).
template <int NV, int CAP>
struct IpmWs {
    static constexpr bool kHasH = IpmHasH<NV, CAP>::value;
    static_assert(CAP % 4 == 0, "row capacity must be a multiple of the DMMA k-step");
    static constexpr int TW = tile_width(NV);    // row stride of the tile
    static constexpr int cA = (NV + 1) & ~1;     // column a = w * rp (even: a and w are stored as one double2)
    static constexpr int cW = cA + 1;            // column w = lam / s (on exit: lam)
    static constexpr int cV = cA + 2;            // column v: corrector term / scratch vector
    static constexpr int cH = cA + 3;            // column h (right-hand side, never overwritten)
    static constexpr int NBA = (cA + 1 + 7) / 8; // 8-row blocks of [vars ; (pad) ; a]
    static constexpr int NBB = (NV + 7) / 8;     // 8-column blocks of the variables
    static constexpr int LS = 8 * NBB + 1;       // row stride of K / L (odd: lane i reads row i)
    static constexpr int NVP = 8 * NBB;
    static constexpr int HS = NV | 1;            // row stride of H (odd)
    static constexpr int MR = (CAP + 31) / 32;   // row slots per lane
    static constexpr int oT = 0;                               // tile [CAP][TW] (+ over-read pad)
    static constexpr int oK = oT + CAP * TW + 8 * NBA;         // K ext [8*NBA][LS]: rows < NV = K / L, row cA = G'(w rp)
    static constexpr int oHs = oK + 8 * NBA * LS + 1;          // H [NV][HS]
    static constexpr int oF = (oHs + (kHasH ? NV * HS : 0) + 1) & ~1;   // f        (NVP)
    static constexpr int oZ = oF + NVP;                        // z in/out (NVP)
    static constexpr int oDz = oZ + NVP;                       // dz       (NVP)
    static constexpr int oDiag = oDz + NVP;                    // original diagonal of K (NVP)
    static constexpr int oRed = oDiag + NVP;                   // partial sums of gt_vec (32)
    static constexpr int oOut = oRed + 32;                     // [0] objective [1] iterations [2] flops [3] flags
    static constexpr int kDoubles = oOut + 8;
};

// acc_k = sum_{r < m} T[r][k] * T[r][col] for k < NV, returned in lane k.  The lanes form
// 32/NV groups of NV lanes that take the rows round-robin; the partial sums meet by shuffle.
template <int NV, int CAP>
__device__ __forceinline__ double gt_vec(const double *__restrict__ T, double *red, int col, int m, int lane)
{
    using W = IpmWs<NV, CAP>;
    constexpr int NG = 32 / NV;
    const int g = lane / NV, k = lane - g * NV;
    double acc = 0.0, acc2 = 0.0;
    if (g < NG) {
        const double *p = T + g * W::TW;
        int r = g;
#pragma unroll 2
        for (; r + NG < m; r += 2 * NG, p += 2 * NG * W::TW) {
            acc = fma(p[k], p[col], acc);
            acc2 = fma(p[NG * W::TW + k], p[NG * W::TW + col], acc2);
        }
        if (r < m) acc = fma(p[k], p[col], acc);
    }
    acc += acc2;
    if constexpr (NG > 1) {   // partial sums of the groups meet in shared memory (no shuffles, see chol_small)
        if (g > 0 && g < NG) red[(g - 1) * NV + k] = acc;
        __syncwarp();
        if (lane < NV) {
#pragma unroll
            for (int q = 0; q < NG - 1; ++q) acc += red[q * NV + lane];
        }
        __syncwarp();
    }
    return acc;
}

// ---- cooperative Cholesky / triangular solves: lane i owns row i of the factor ----------------
// Small systems (NV <= 12): every loop is unrolled at compile time, the lane's own row lives in
// registers, row j reaches the other lanes by broadcast loads from shared memory and the pivot by
// one shuffle.  `diag` holds the original diagonal of K (the guard compares against it).
template <int NV, int LS>
__device__ __forceinline__ bool chol_small(double *Ks, const double *diag, int lane, double (&Lr)[NV], double &myinv)
{
    // No shuffles: on sm_100a every __shfl_sync / __reduce_*_sync is wrapped in a WARPSYNC ...
    // ENDCOLLECTIVE frame (~12 SASS instructions per 64-bit value).  The pivot of column j is
    // recomputed by every lane from the broadcast loads of row j it needs anyway.
    const int li = lane < NV ? lane : 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) Lr[k] = Ks[li * LS + k];
    double chk = 0.0;
    myinv = 1.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const double d0 = diag[j];
        double v = Lr[j], d = d0, v2 = 0.0, d2 = 0.0;
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const double ljk = Ks[j * LS + k];
            if (k & 1) {
                v2 = fma(-Lr[k], ljk, v2);
                d2 = fma(-ljk, ljk, d2);
            } else {
                v = fma(-Lr[k], ljk, v);
                d = fma(-ljk, ljk, d);
            }
        }
        v += v2;
        d += d2;
        chk += d;                                          // NaN / Inf of any pivot surfaces here
        d = (d <= 1e-13 * d0 || d0 <= 0.0) ? 1e100 : d;   // pivot guard (oracle: cholesky)
        const double iv = fast_rsqrt(d);
        if (lane == j) {
            v = d;
            myinv = iv;
        }
        Lr[j] = v * iv;
        if (lane >= j && lane < NV) Ks[lane * LS + j] = Lr[j];
        __syncwarp();
    }
    return !isfinite(chk);
}

// b_i in lane i on entry, x_i = (K^-1 b)_i in lane i on return (lanes >= NV: 0).  The solution
// components travel through the scratch vector sb (shared memory, NV doubles); a variant in which
// every lane runs the whole substitution redundantly (no round trips) needs 40 more registers
// and spilled at the 128 of this kernel: 1.6x slower.
template <int NV, int LS>
__device__ __forceinline__ double trisolve_small(const double *Ks, double *sb, int lane, const double (&Lr)[NV],
                                                 double myinv, double b)
{
    const double *col = Ks + (lane < NV ? lane : 0);
    if (lane >= NV) b = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {   // forward  L y = b: lanes i > k update b_i with y_k
        if (lane == k) sb[k] = b * myinv;
        __syncwarp();
        const double yk = sb[k];
        if (lane > k) b = fma(-Lr[k], yk, b);
    }
    b *= myinv;                       // lane k was last touched at step k-1: this is y_k
    __syncwarp();
#pragma unroll
    for (int i = NV - 1; i >= 0; --i) {   // backward  L' x = y: lanes k < i subtract L[i][k] x_i
        if (lane == i) sb[i] = b * myinv;
        __syncwarp();
        const double xi = sb[i];
        if (lane < i) b = fma(-col[i * LS], xi, b);
    }
    return lane < NV ? sb[lane < NV ? lane : 0] : 0.0;
}

// Larger systems: rolled loops, the factor stays in shared memory.
template <int NV, int LS>
__device__ __forceinline__ bool chol_rolled(double *Ks, const double *diag, int lane, double &myinv)
{
    const double *rowi = Ks + (lane < NV ? lane : 0) * LS;
    bool bad = false;
    myinv = 1.0;
#pragma unroll 1
    for (int j = 0; j < NV; ++j) {
        const double *rowj = Ks + j * LS;
        double v = rowi[j], v2 = 0.0;
        int k = 0;
#pragma unroll 4
        for (; k + 1 < j; k += 2) {
            v = fma(-rowi[k], rowj[k], v);
            v2 = fma(-rowi[k + 1], rowj[k + 1], v2);
        }
        if (k < j) v = fma(-rowi[k], rowj[k], v);
        v += v2;
        double d = __shfl_sync(kFull, v, j);
        const double d0 = diag[j];
        bad = bad || !isfinite(d);
        d = (d <= 1e-13 * d0 || d0 <= 0.0) ? 1e100 : d;
        const double iv = fast_rsqrt(d);
        if (lane >= j && lane < NV) Ks[lane * LS + j] = (lane == j) ? d * iv : v * iv;
        myinv = (lane == j) ? iv : myinv;
        __syncwarp();
    }
    return bad;
}

template <int NV, int LS>
__device__ __forceinline__ double trisolve_rolled(const double *Ks, int lane, double myinv, double b)
{
    const double *rowi = Ks + (lane < NV ? lane : 0) * LS;
    const double *col = Ks + (lane < NV ? lane : 0);
    if (lane >= NV) b = 0.0;
#pragma unroll 4
    for (int k = 0; k < NV; ++k) {
        const double yk = __shfl_sync(kFull, b * myinv, k);
        if (lane > k) b = fma(-rowi[k], yk, b);
    }
    b *= myinv;
#pragma unroll 4
    for (int i = NV - 1; i >= 0; --i) {
        const double xi = __shfl_sync(kFull, b * myinv, i);
        if (lane < i) b = fma(-col[i * LS], xi, b);
    }
    b *= myinv;
    return lane < NV ? b : 0.0;
}

// Solve  min 0.5 z'Hz + f'z  s.t.  G z <= h  with one warp.  `w` = base of an IpmWs block: the
// tile rows [0, m) (coefficients in columns [0, NV), h in column cH), H (if hasH), f and the
// primal start z must be filled in.  Rows [m, roundup4(m)) may hold anything finite.
// On return z holds the solution, column cW of the tile the multipliers, out[0..2] the
// objective / iteration count / algorithmic flops.
// returns 0 converged, 1 iteration limit, 2 numerical breakdown.
//
// early (warp-uniform): the problem is a phase-1 LP  min sigma : G z - sigma <= h  (sigma = last
// variable, started strictly feasible, so every iterate is a feasible point of it).  The solve
// stops as soon as z_k leaves every original row a slack >= TAU_INT: "alive" is then certified
// (sigma* <= -TAU_INT) and later iterations cannot change the answer (oracle/ehm_oracle.c:
// qp_solve_x).  The objective returned in that case is the bound -TAU_INT.
template <int NV, int CAP>
__device__ __noinline__ int ipm_solve(double *w, int m, int lane, bool hasH_, bool early = false)
{
    using W = IpmWs<NV, CAP>;
    const bool hasH = W::kHasH && hasH_;
    constexpr int TW = W::TW, MR = W::MR, LS = W::LS, NBA = W::NBA, NBB = W::NBB;
    constexpr int NVE = NV & ~1;   // even part of the columns (128-bit loads), NV - NVE = 0/1 tail
    double *T = w + W::oT, *Ks = w + W::oK, *zs = w + W::oZ, *dzs = w + W::oDz;
    const double *__restrict__ Hs = w + W::oHs;
    const double *__restrict__ fs = w + W::oF;
    // per-row state in registers
    double s[MR], lam[MR], rp[MR], rs[MR], dsdl[MR];
    const double fk = lane < NV ? fs[lane] : 0.0;
    const double fmx = warp_max_nonneg(fabs(fk));
    double zk = lane < NV ? zs[lane] : 0.0;   // lane k owns z_k
    const double mact = (double)m;
    const int ngroups = (m + 3) >> 2;
    // ---- centred start -------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < MR; ++j) {
        const int r = lane + 32 * j;
        s[j] = 1.0; lam[j] = 0.0; rp[j] = 0.0; rs[j] = 1.0; dsdl[j] = 0.0;
        if (r < m) {
            double gz = 0.0;
#pragma unroll
            for (int k = 0; k < NVE; k += 2) {
                const double2 g = *reinterpret_cast<const double2 *>(T + r * TW + k);
                const double2 zz = *reinterpret_cast<const double2 *>(zs + k);
                gz = fma(g.x, zz.x, fma(g.y, zz.y, gz));
            }
            if (NV & 1) gz = fma(T[r * TW + NV - 1], zs[NV - 1], gz);
            s[j] = fmax(T[r * TW + W::cH] - gz, 1.0);
            lam[j] = 1.0 / s[j];
            T[r * TW + W::cV] = lam[j];
        }
    }
    __syncwarp();
    // initial dual residual rd0 = |H z0 + f + G' lam0|
    double rd0;
    {
        double v = gt_vec<NV, CAP>(T, w + W::oRed, W::cV, m, lane);
        if (lane < NV) {
            v += fk;
            if (hasH)
                for (int q = 0; q < NV; ++q) v = fma(Hs[lane * W::HS + q], zs[q], v);
        } else v = 0.0;
        rd0 = warp_max_nonneg(fabs(v));
    }
    int status = 1, it = 0;
    double rho = 1.0, obj = 0.0;
#pragma unroll 1
    for (it = 0; it <= EHM_MAX_ITER; ++it) {
        // ---- pass A: primal residual, gap, 1/s, w, a -----------------------------------------
        double rpm = 0.0, lmx = 0.0, mu = 0.0;
        bool cert = early;
        const double cthr = EHM_TAU_INT + zs[NV - 1];
        {
            double gz[MR];
#pragma unroll
            for (int j = 0; j < MR; ++j) gz[j] = 0.0;
#pragma unroll
            for (int k = 0; k < NVE; k += 2) {
                const double2 zz = *reinterpret_cast<const double2 *>(zs + k);
#pragma unroll
                for (int j = 0; j < MR; ++j) {
                    const int r = lane + 32 * j;
                    if (32 * j < m) {
                        const double2 g = *reinterpret_cast<const double2 *>(T + (r < CAP ? r : 0) * TW + k);
                        gz[j] = fma(g.x, zz.x, fma(g.y, zz.y, gz[j]));
                    }
                }
            }
            if (NV & 1) {
                const double zl = zs[NV - 1];
#pragma unroll
                for (int j = 0; j < MR; ++j) {
                    const int r = lane + 32 * j;
                    if (32 * j < m) gz[j] = fma(T[(r < CAP ? r : 0) * TW + NV - 1], zl, gz[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < MR; ++j) {
                const int r = lane + 32 * j;
                if (32 * j < m && r < 4 * ngroups) {
                    double a = 0.0, wt = 0.0;
                    if (r < m) {
                        rp[j] = gz[j] + s[j] - T[r * TW + W::cH];
                        rs[j] = fast_rcp(s[j]);
                        wt = lam[j] * rs[j];
                        a = wt * rp[j];
                        rpm = fmax(rpm, fabs(rp[j]));
                        // slack of the original row (without the sigma column) at z_k: h - Gz - sigma_k
                        cert = cert && (s[j] - rp[j] >= cthr);
                        mu = fma(s[j], lam[j], mu);
                        lmx = fmax(lmx, lam[j]);
                    }
                    *reinterpret_cast<double2 *>(T + r * TW + W::cA) = make_double2(a, wt);
                }
            }
        }
        if (early && __all_sync(kFull, cert)) {
            obj = -EHM_TAU_INT;
            status = 0;
            break;
        }
        rpm = warp_max_nonneg(rpm);
        lmx = warp_max_nonneg(lmx);
        mu = warp_sum_mma(mu);
        // (H z + f)_k in lane k, objective, |z|
        double hzf = fk;
        if (hasH && lane < NV) {
#pragma unroll 4
            for (int q = 0; q < NV; ++q) hzf = fma(Hs[lane * W::HS + q], zs[q], hzf);
        }
        const double zm = warp_max_nonneg(lane < NV ? fabs(zk) : 0.0);
        if (!isfinite(mu) || !isfinite(rpm) || !isfinite(zm)) {
            status = 2;
            break;
        }
        // the objective is only needed once the residual tests pass (and on exit)
        bool conv = rpm <= EHM_TOL_IPM * (1.0 + zm) && rho * rd0 <= EHM_TOL_IPM * (1.0 + fmx + lmx);
        if (conv || it == EHM_MAX_ITER || g_trace_on) {
            obj = warp_sum_mma(lane < NV ? zk * 0.5 * (hzf + fk) : 0.0);
            conv = conv && mu <= EHM_TOL_IPM * (1.0 + fabs(obj));
        }
        if (g_trace_on && lane == 0 && it < 120) {
            g_trace[8 * it + 0] = it; g_trace[8 * it + 1] = rpm; g_trace[8 * it + 2] = mu;
            g_trace[8 * it + 3] = obj; g_trace[8 * it + 4] = rho * rd0; g_trace[8 * it + 5] = lmx;
            g_trace[8 * it + 6] = mact; g_trace[8 * it + 7] = NV;
        }
        if (conv) {
            // guard on the MEASURED dual residual |Hz + f + G'lam| (it carries the round-off of the
            // normal equations, so it cannot be held to TOL_IPM; see oracle/ehm_oracle.c)
#pragma unroll
            for (int j = 0; j < MR; ++j) {
                const int r = lane + 32 * j;
                if (r < m) T[r * TW + W::cV] = lam[j];
            }
            __syncwarp();
            double v = gt_vec<NV, CAP>(T, w + W::oRed, W::cV, m, lane);
            v = lane < NV ? fabs(v + hzf) : 0.0;
            const double rdm = warp_max_nonneg(v);
            if (rdm <= EHM_TOL_RD * (1.0 + fmx + lmx)) {
                status = 0;
                break;
            }
        }
        if (it == EHM_MAX_ITER) break;
        const double mu_mean = mu / mact;
        __syncwarp();
        // ---- K ext = [H + G'WG ; (G'(w rp))'] on the fp64 tensor path ---------------------------
        {
            double c[NBA][NBB][2];
            const int cr = lane >> 2, cc = 2 * (lane & 3);
#pragma unroll
            for (int a = 0; a < NBA; ++a)
#pragma unroll
                for (int b = 0; b < NBB; ++b) {
                    c[a][b][0] = c[a][b][1] = 0.0;
                    if (hasH && b <= a) {
                        const int i = 8 * a + cr, q = 8 * b + cc;
                        if (i < NV && q < NV) c[a][b][0] = Hs[i * W::HS + q];
                        if (i < NV && q + 1 < NV) c[a][b][1] = Hs[i * W::HS + q + 1];
                    }
                }
            const double *tp = T + (lane & 3) * TW + (lane >> 2);
#pragma unroll 2
            for (int g = 0; g < ngroups; ++g, tp += 4 * TW) {
                double x[NBA];
#pragma unroll
                for (int a = 0; a < NBA; ++a) x[a] = tp[8 * a];
                const double wr = T[(4 * g + (lane & 3)) * TW + W::cW];
                double aw[NBA];
#pragma unroll
                for (int a = 0; a < NBA; ++a) aw[a] = (8 * a + (lane >> 2) < NV) ? x[a] * wr : x[a];
#pragma unroll
                for (int a = 0; a < NBA; ++a)
#pragma unroll
                    for (int b = 0; b < NBB; ++b)
                        if (b <= a) dmma884(c[a][b][0], c[a][b][1], aw[a], x[b]);
            }
#pragma unroll
            for (int a = 0; a < NBA; ++a)
#pragma unroll
                for (int b = 0; b < NBB; ++b)
                    if (b <= a) {
                        double *q = Ks + (8 * a + cr) * LS + 8 * b + cc;
                        q[0] = c[a][b][0];
                        q[1] = c[a][b][1];
                        if (a == b) {   // the diagonal, kept aside for the pivot guard
                            if (cr == cc) w[W::oDiag + 8 * a + cr] = c[a][b][0];
                            if (cr == cc + 1) w[W::oDiag + 8 * a + cr] = c[a][b][1];
                        }
                    }
        }
        __syncwarp();
        const double ak = lane < NV ? Ks[W::cA * LS + lane] : 0.0;   // (G'(w rp))_k
        // ---- Cholesky, lane i = row i, left-looking, in place ------------------------------------
        constexpr bool kSmall = NV <= EHM_SMALL_NV;
        double Lr[kSmall ? NV : 1];
        double myinv;
        bool bad;
        if constexpr (kSmall) bad = chol_small<NV, LS>(Ks, w + W::oDiag, lane, Lr, myinv);
        else bad = chol_rolled<NV, LS>(Ks, w + W::oDiag, lane, myinv);
        if (bad) {
            status = 2;
            break;
        }
        // ---- predictor (trip 0) and corrector (trip 1) --------------------------------------------
        double sig = 0.0, cres = 1.0, alpha = 1.0, dzk = 0.0;
#pragma unroll 1
        for (int trip = 0; trip < 2; ++trip) {
            double b = 0.0;
            if (trip == 0) {
                b = -(hzf + ak);
            } else {
                // v = sigma lam + (ds.dl - sigma mu) / s;  rhs = -c (Hz + f + a) + G'v
                const double smu = sig * mu_mean;
#pragma unroll
                for (int j = 0; j < MR; ++j) {
                    const int r = lane + 32 * j;
                    if (r < m) T[r * TW + W::cV] = fma(sig, lam[j], (dsdl[j] - smu) * rs[j]);
                }
                __syncwarp();
                const double c1 = gt_vec<NV, CAP>(T, w + W::oRed, W::cV, m, lane);
                b = fma(-cres, hzf + ak, c1);
            }
            if (lane >= NV) b = 0.0;
                        if constexpr (kSmall) b = trisolve_small<NV, LS>(Ks, dzs, lane, Lr, myinv, b);
            else b = trisolve_rolled<NV, LS>(Ks, lane, myinv, b);
            dzk = b;
            if (lane < W::NVP) dzs[lane] = b;
            __syncwarp();
            // step to the boundary: min over rows of s/(-ds) and lam/(-dl).  The ratios are formed
            // in fp32 (the step only has to stay inside the cone: it is scaled by 0.99 below, which
            // dwarfs the 1e-7 of fp32, and the predictor's length merely sets sigma), which lets one
            // 32-bit REDUX find the minimum.
            float amin = 3.0e38f;
            {
                double gd[MR];
#pragma unroll
                for (int j = 0; j < MR; ++j) gd[j] = 0.0;
#pragma unroll
                for (int k = 0; k < NVE; k += 2) {
                    const double2 dd = *reinterpret_cast<const double2 *>(dzs + k);
#pragma unroll
                    for (int j = 0; j < MR; ++j) {
                        const int r = lane + 32 * j;
                        if (32 * j < m) {
                            const double2 g = *reinterpret_cast<const double2 *>(T + (r < CAP ? r : 0) * TW + k);
                            gd[j] = fma(g.x, dd.x, fma(g.y, dd.y, gd[j]));
                        }
                    }
                }
                if (NV & 1) {
                    const double dl_ = dzs[NV - 1];
#pragma unroll
                    for (int j = 0; j < MR; ++j) {
                        const int r = lane + 32 * j;
                        if (32 * j < m) gd[j] = fma(T[(r < CAP ? r : 0) * TW + NV - 1], dl_, gd[j]);
                    }
                }
                const double smu = sig * mu_mean;
#pragma unroll
                for (int j = 0; j < MR; ++j) {
                    const int r = lane + 32 * j;
                    if (r < m) {
                        const double rc = trip ? fma(s[j], lam[j], dsdl[j] - smu) : s[j] * lam[j];
                        const double ds = -cres * rp[j] - gd[j];
                        const double dl = (-rc - lam[j] * ds) * rs[j];
                        // s / (-ds) where ds < 0 (else +inf), fp32 with the hardware reciprocal (1 ulp)
                        const float qs = __double2float_rz(s[j]) * __frcp_rn(__double2float_ru(fmax(-ds, 0.0)));
                        const float ql = __double2float_rz(lam[j]) * __frcp_rn(__double2float_ru(fmax(-dl, 0.0)));
                        amin = fminf(amin, fminf(qs, ql));
                        if (trip) {
                            rp[j] = ds;     // rp / dsdl are dead from here on: they carry the step
                            dsdl[j] = dl;
                        } else {
                            dsdl[j] = ds * dl;
                        }
                    }
                }
            }
            const double amax = (double)__uint_as_float(__reduce_min_sync(kFull, __float_as_uint(amin)));
            if (trip == 0) {
                const double a_aff = fmin(1.0, amax);
                sig = (1.0 - a_aff) * (1.0 - a_aff) * (1.0 - a_aff);
                cres = 1.0 - sig;
            } else {
                alpha = fmin(1.0, 0.99 * amax);
            }
        }
        // ---- step ----------------------------------------------------------------------------
        zk = fma(alpha, dzk, zk);
        __syncwarp();
        if (lane < NV) zs[lane] = zk;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            const int r = lane + 32 * j;
            if (r < m) {
                s[j] = fma(alpha, rp[j], s[j]);
                lam[j] = fma(alpha, dsdl[j], lam[j]);
            }
        }
        rho *= 1.0 - alpha * cres;
        __syncwarp();
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < MR; ++j) {
        const int r = lane + 32 * j;
        if (r < m) T[r * TW + W::cW] = lam[j];
    }
    if (lane == 0) {
        w[W::oOut + 0] = obj;
        w[W::oOut + 1] = (double)it;
        // algorithmic flops (2 per FMA): K build m*NV(NV+1), five mat-vec passes 10*m*NV,
        // Cholesky NV^3/3, two triangular solve pairs 4*NV^2 -- per iteration
        w[W::oOut + 2] = (double)it * (mact * (NV * (NV + 1) + 10.0 * NV) + NV * NV * NV / 3.0 + 4.0 * NV * NV);
    }
    __syncwarp();
    return status;
}

}  // namespace ehm
