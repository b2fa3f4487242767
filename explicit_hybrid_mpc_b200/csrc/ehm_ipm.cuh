// ehm_ipm.cuh -- warp-level dense primal-dual interior-point solver (sm_100a, fp64).
//
// One warp solves one small convex QP/LP
//        min 0.5 z'Hz + f'z   s.t.  G z <= h            (NV <= 10 variables, <= 32*MR rows)
// held entirely in shared memory + registers.  It replaces one `.solve()` of the
// reference (CVXPY -> MOSEK, lib/oracle.py:131,134,166,169,203,276,305,350).
//
// Mapping to the hardware (round-1 final; history and ncu findings in profiles/README.md):
//   * one warp = one problem, one-warp CTAs, 8 per SM at 255 registers.  Lane L owns rows L,
//     L+32, L+64 of the (pre-solved, compacted) 96-row tile; G is column-major in shared memory
//     (32 lanes read 32 consecutive doubles: conflict free); the per-row state (slack s,
//     multiplier lam, 1/s / ds, rc / dl, primal residual) also lives in shared memory so that
//     the row loops stay ROLLED (a fully unrolled first version was 25k SASS instructions and
//     stalled on instruction fetch) and handle two rows per trip.
//   * the normal matrix K = H + G' diag(lam/s) G is accumulated per lane in NV(NV+1)/2
//     registers and summed over the warp through a 32x33 shared-memory transposition (two
//     rounds of 32 values; the gap sum s'lam rides along, H is added by the summing lane).
//   * max / min reductions use two 32-bit REDUX on the bit patterns of non-negative doubles.
//   * the Cholesky factorisation is computed redundantly by every lane in registers (no
//     broadcasts), the packed factor is left in shared memory for the triangular solves.
//   * predictor and corrector share one code path (a 2-trip loop).
//
// The numerical method is, by specification, the one of oracle/ehm_oracle.c:qp_solve
// (Mehrotra predictor-corrector in the CVXOPT/ECOS form: sigma=(1-alpha_aff)^3, residuals
// scaled by 1-sigma, centred start s.lam=1, pivot guard, convergence on primal residual,
// gap s'lam and the propagated dual residual rho*rd0).  Only the summation order differs.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/ehm.h"

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

// Packed lower-triangular index.
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// Per-warp shared-memory block of the solver (offsets in doubles from the block base).
template <int NV, int MR>
struct IpmWs {
    static constexpr int MRP = 32 * MR;
    static constexpr int oG = 0;                 // G, column-major [NV][MRP]
    static constexpr int oH = oG + NV * MRP;     // h
    static constexpr int oS = oH + MRP;          // slack
    static constexpr int oL = oS + MRP;          // multiplier
    static constexpr int oA = oL + MRP;          // rp / ds
    static constexpr int oB = oA + MRP;          // ds.dl / rc / dl
    static constexpr int oR = oB + MRP;          // primal residual G z + s - h of the iteration
    static constexpr int oHs = oR + MRP;         // Hessian NV*NV row-major (zeros for an LP)
    static constexpr int oF = oHs + NV * NV;     // f (16)
    static constexpr int oZ = oF + 16;           // z in/out (16)
    static constexpr int oRed = oZ + 16;         // reduction scratch / packed K / packed L (64)
    static constexpr int oOut = oRed + 64;       // [0] objective [1] iterations [2] flops (8)
    static constexpr int oLs = oOut + 8;         // packed Cholesky factor / working matrix (64)
    static constexpr int oD0 = oLs + 64;         // original diagonal of K (16)
    static constexpr int oHp = oD0 + 16;         // packed lower triangle of the Hessian (64)
    static constexpr int oScr = oHp + 64;        // 32 x 33 transposition scratch of the reductions
    static constexpr int kDoubles = oScr + 32 * 33;
};

// Warp sums through shared memory.  Up to 32 per-lane values at a time are transposed through
// a 32x33 scratch: every lane stores its partial of value i at scr[i*33 + lane] (unit stride:
// conflict free), then lane L adds the 32 partials of value L (addresses L*33 + j: 16 distinct
// 8-byte banks per half-warp: conflict free) in a depth-5 tree.  No shuffles: a 64-bit shuffle
// costs ~12 SASS instructions here and the earlier 8-value version of this routine (11 dependent
// calls per iteration) was 22 % of all stall samples (profiles/README.md).
constexpr int kScrStride = 33;

// Sum CNT per-lane partials v[base .. base+CNT) over the warp into out[base .. base+CNT).
template <int CNT, int P, bool ADD = false>
__device__ __forceinline__ void reduce_round(const double (&v)[P], int base, double *out, double *scr, int lane,
                                             const double *add = nullptr)
{
#pragma unroll
    for (int i = 0; i < CNT; ++i) scr[i * kScrStride + lane] = v[base + i];
    __syncwarp();
    if (lane < CNT) {
        const double *q = scr + lane * kScrStride;
        double t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = q[j] + q[j + 16];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += t[j + 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] += t[j + 4];
        const double sum = (t[0] + t[2]) + (t[1] + t[3]);
        out[base + lane] = ADD ? sum + add[base + lane] : sum;
    }
    __syncwarp();
}

// Sum the first CNT of P per-lane partials over the warp into out[0..CNT) (shared memory).
// With ADD the constant vector add[0..CNT) is added to the sums (the Hessian joins the normal
// matrix here, two additions per lane instead of one per entry and lane later).
template <int CNT, int P, bool ADD = false>
__device__ __forceinline__ void reduce_all(const double (&v)[P], double *out, double *scr, int lane,
                                           const double *add = nullptr)
{
    constexpr int R0 = CNT < 32 ? CNT : 32;
    reduce_round<R0, P, ADD>(v, 0, out, scr, lane, add);
    if constexpr (CNT > 32) reduce_round<CNT - 32, P, ADD>(v, 32, out, scr, lane, add);
}

// Solve  min 0.5 z'Hz + f'z  s.t.  G z <= h  with one warp.  `w` = base of an IpmWs block:
// G, h, Hs, f and the primal start z must be filled in; act: bit j set <=> row lane+32*j takes
// part (normalised, non-constant row); rows that do not take part must hold G = 0, h = 1.
// On return z holds the solution, out[0..2] the objective / iteration count / algorithmic
// flops.  returns 0 converged, 1 iteration limit, 2 numerical breakdown.
//
// Row passes handle TWO rows per trip of a rolled loop (independent dependency chains: the
// solver is latency bound with 7 resident warps per SM), are branch free (rows that do not
// take part carry s = 1, lam = 0 and a zero mask) and contain no division: 1/s is formed once
// per row and iteration, step-length ratios are compared by cross-multiplication.
//
// early (warp-uniform): the problem is a phase-1 LP  min sigma : G z - sigma <= h  (sigma = last
// variable, started strictly feasible, so every iterate is a feasible point of it).  The solve
// stops as soon as z_k leaves every original row a slack >= TAU_INT: "alive" is then certified
// (sigma* <= -TAU_INT) and later iterations cannot change the answer (oracle/ehm_oracle.c:
// qp_solve_x).  The objective returned in that case is the bound -TAU_INT.
template <int NV, int MR>
__device__ __noinline__ int ipm_solve(double *w, unsigned act, int lane, bool early = false)
{
    using W = IpmWs<NV, MR>;
    constexpr int MRP = W::MRP;
    constexpr int NK = NV * (NV + 1) / 2;
    constexpr int PK = 8 * ((NK + 8) / 8);   // accumulators + the gap sum, padded to 8-value chunks
    constexpr int PV = 8 * ((NV + 7) / 8);
    const double *__restrict__ Gs = w + W::oG;
    const double *__restrict__ hs = w + W::oH;
    double *sv = w + W::oS, *lv = w + W::oL, *av = w + W::oA, *bv = w + W::oB, *rv = w + W::oR;
    const double *__restrict__ Hs = w + W::oHs;
    double *red = w + W::oRed, *Ls = w + W::oLs, *scr = w + W::oScr, *Hp = w + W::oHp;
    double z[NV], f[NV];
    double fmx = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        f[k] = w[W::oF + k];
        z[k] = w[W::oZ + k];
        fmx = fmax(fmx, fabs(f[k]));
    }
    const double mact = warp_sum((double)__popc(act));
    for (int idx = lane; idx < NK; idx += 32) {   // packed lower triangle of H
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= idx) ++i;
        Hp[idx] = Hs[i * NV + idx - i * (i + 1) / 2];
    }
    if (lane == 0) Hp[NK] = 0.0;   // slot of the gap sum s'lam in the reduction
    // row pairs in which no lane has a participating row are skipped (warp-uniform)
    unsigned pairs = 0;
#pragma unroll
    for (int j = 0; j < MR; j += 2)
        if (__any_sync(kFull, (act >> j & 3u) != 0u)) pairs |= 1u << j;
    // ---- centred start + initial dual residual rd0 = |H z0 + f + G' lam0| -----------------
    double rd0;
    {
        double gl[PV];
#pragma unroll
        for (int k = 0; k < PV; ++k) gl[k] = 0.0;
#pragma unroll 1
        for (int j = 0; j < MR; ++j) {
            const int r = lane + 32 * j;
            if (!(act >> j & 1u)) {
                sv[r] = 1.0;
                lv[r] = 0.0;
                av[r] = 1.0;
                bv[r] = 0.0;
                continue;
            }
            double g[NV], gz = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                g[k] = Gs[k * MRP + r];
                gz = fma(g[k], z[k], gz);
            }
            const double s = fmax(hs[r] - gz, 1.0);
            const double l = 1.0 / s;
            sv[r] = s;
            lv[r] = l;
            av[r] = l;
            bv[r] = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) gl[k] = fma(g[k], l, gl[k]);
        }
        reduce_all<NV>(gl, red, scr, lane);
        double m = 0.0;
        if (lane < NV) {
            double hz = w[W::oF + lane] + red[lane];
#pragma unroll
            for (int l = 0; l < NV; ++l) hz = fma(Hs[lane * NV + l], z[l], hz);
            m = fabs(hz);
        }
        rd0 = warp_max(m);
        __syncwarp();
    }
    // masks of the two rows handled per loop trip
    int status = 1, it = 0;
    double rho = 1.0, obj = 0.0;
#pragma unroll 1
    for (it = 0; it <= EHM_MAX_ITER; ++it) {
        // ---- pass A: primal residual, gap, 1/s, normal matrix -------------------------------
        double acc[PK];
#pragma unroll
        for (int k = 0; k < PK; ++k) acc[k] = 0.0;
        double rpm = 0.0, lmx = 0.0;
        double &mu_acc = acc[NK];
        bool cert = early;
        const double cthr = EHM_TAU_INT + z[NV - 1];
#pragma unroll 1
        for (int j = 0; j < MR; j += 2) {
            if (!(pairs >> j & 1u)) continue;
            const bool on1 = (j + 1 < MR);
            const int r0 = lane + 32 * j, r1 = on1 ? r0 + 32 : r0;
            // rows that do not take part in THIS solve (e.g. the epigraph rows during phase 1)
            // may hold data: mask them (their s = 1, lam = 0 already keep K and the gap clean)
            const double m0 = (act >> j & 1u) ? 1.0 : 0.0;
            const double m1 = (on1 && (act >> (j + 1) & 1u)) ? 1.0 : 0.0;
            double g0[NV], g1[NV], gz0 = 0.0, gz1 = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                g0[k] = Gs[k * MRP + r0];
                g1[k] = Gs[k * MRP + r1];
                gz0 = fma(g0[k], z[k], gz0);
                gz1 = fma(g1[k], z[k], gz1);
            }
            const double s0 = sv[r0], s1 = sv[r1], l0 = lv[r0] * m0, l1 = lv[r1] * m1;
            const double rs0 = 1.0 / s0, rs1 = 1.0 / s1;
            av[r0] = rs0;
            if (on1) av[r1] = rs1;
            const double rp0 = gz0 + s0 - hs[r0], rp1 = gz1 + s1 - hs[r1];
            rv[r0] = rp0;
            if (on1) rv[r1] = rp1;
            rpm = fmax(rpm, fmax(m0 * fabs(rp0), m1 * fabs(rp1)));
            // slack of the original row (without the sigma column) at z_k: h - Gz - sigma_k
            cert = cert && (m0 == 0.0 || s0 - rp0 >= cthr) && (m1 == 0.0 || s1 - rp1 >= cthr);
            mu_acc = fma(s0, l0, fma(s1, l1, mu_acc));
            lmx = fmax(lmx, fmax(l0, l1));
            const double wt0 = l0 * rs0, wt1 = l1 * rs1;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const double gw0 = wt0 * g0[k], gw1 = wt1 * g1[k];
#pragma unroll
                for (int q = 0; q <= k; ++q)
                    acc[tri(k, q)] = fma(gw0, g0[q], fma(gw1, g1[q], acc[tri(k, q)]));
            }
        }
        if (early && __all_sync(kFull, cert)) {
            obj = -EHM_TAU_INT;
            status = 0;
            break;
        }
        rpm = warp_max_nonneg(rpm);
        lmx = warp_max_nonneg(lmx);
        // (H z + f): lane k forms row k, exchanged through the scratch line
        double hzf[NV], zm;
        {
            if (lane < NV) {
                double v = w[W::oF + lane];
#pragma unroll
                for (int q = 0; q < NV; ++q) v = fma(Hs[lane * NV + q], z[q], v);
                red[lane] = v;
            }
            __syncwarp();
            zm = 0.0;
            obj = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                hzf[k] = red[k];
                obj = fma(z[k], 0.5 * (hzf[k] + f[k]), obj);
                zm = fmax(zm, fabs(z[k]));
            }
            __syncwarp();
        }
        // K (+H) and the gap sum -> red; Cholesky input stays there until it is loaded below
        reduce_all<NK + 1, PK, true>(acc, red, scr, lane, Hp);
        const double mu = red[NK];
        {
            if (!isfinite(mu) || !isfinite(rpm) || !isfinite(obj)) {
                status = 2;
                break;
            }
            if (rpm <= EHM_TOL_IPM * (1.0 + zm) && rho * rd0 <= EHM_TOL_IPM * (1.0 + fmx + lmx) &&
                mu <= EHM_TOL_IPM * (1.0 + fabs(obj))) {
                status = 0;
                break;
            }
        }
        if (g_trace_on && lane == 0 && it < 120) {
            g_trace[8 * it + 0] = it; g_trace[8 * it + 1] = rpm; g_trace[8 * it + 2] = mu;
            g_trace[8 * it + 3] = obj; g_trace[8 * it + 4] = rho * rd0; g_trace[8 * it + 5] = lmx;
            g_trace[8 * it + 6] = mact; g_trace[8 * it + 7] = NV;
        }
        if (it == EHM_MAX_ITER) break;
        const double mu_mean = mu / mact;
        // ---- K -> packed scratch -> Cholesky, redundantly in every lane's registers (SIMT makes
        //      the redundancy free; the lane-parallel shuffle version cost 1.5k instructions:
        //      a 64-bit shuffle is ~13 SASS instructions) -> packed factor in shared memory -----
        double inv[NV];
        {
            double Lk[NK];
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int q = 0; q <= i; ++q) Lk[tri(i, q)] = red[tri(i, q)];
            __syncwarp();
            bool bad = false;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const double d0 = Lk[tri(j, j)];
                double d = d0;
#pragma unroll
                for (int k = 0; k < j; ++k) d = fma(-Lk[tri(j, k)], Lk[tri(j, k)], d);
                if (!isfinite(d)) bad = true;
                if (d <= 1e-13 * d0 || d0 <= 0.0) d = 1e100;   // pivot guard (oracle: cholesky)
                const double iv = rsqrt(d);
                inv[j] = iv;
                Lk[tri(j, j)] = d * iv;
#pragma unroll
                for (int i = j + 1; i < NV; ++i) {
                    double v = Lk[tri(i, j)];
#pragma unroll
                    for (int k = 0; k < j; ++k) v = fma(-Lk[tri(i, k)], Lk[tri(j, k)], v);
                    Lk[tri(i, j)] = v * iv;
                }
            }
            if (bad) {
                status = 2;
                break;
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NK; ++k) Ls[k] = Lk[k];
            }
            __syncwarp();
        }
        // ---- predictor (trip 0) and corrector (trip 1) share one code path -------------------
        double sig = 0.0, cres = 1.0, alpha = 1.0;
        double dz[NV];
#pragma unroll 1
        for (int trip = 0; trip < 2; ++trip) {
            double rh[PV];
#pragma unroll
            for (int k = 0; k < PV; ++k) rh[k] = 0.0;
#pragma unroll 1
            for (int j = 0; j < MR; j += 2) {
                if (!(pairs >> j & 1u)) continue;
                const bool on1 = (j + 1 < MR);
                const int r0 = lane + 32 * j, r1 = on1 ? r0 + 32 : r0;
                const double m0 = (act >> j & 1u) ? 1.0 : 0.0;
                const double m1 = (on1 && (act >> (j + 1) & 1u)) ? 1.0 : 0.0;
                double g0[NV], g1[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    g0[k] = Gs[k * MRP + r0];
                    g1[k] = Gs[k * MRP + r1];
                }
                const double s0 = sv[r0], s1 = sv[r1], l0 = lv[r0], l1 = lv[r1];
                const double rp0 = rv[r0], rp1 = rv[r1];
                // rc: s.lam (predictor) or s.lam + ds_aff.dl_aff - sigma mu (corrector)
                const double rc0 = m0 * (s0 * l0 + (trip ? bv[r0] - sig * mu_mean : 0.0));
                const double rc1 = m1 * (s1 * l1 + (trip ? bv[r1] - sig * mu_mean : 0.0));
                if (trip) {
                    bv[r0] = rc0;
                    if (on1) bv[r1] = rc1;
                }
                // rhs = -c (Hz+f) - G'(c lam + t),  t = (lam c rp - rc)/s
                const double q0 = m0 * fma(cres, l0, (l0 * cres * rp0 - rc0) * av[r0]);
                const double q1 = m1 * fma(cres, l1, (l1 * cres * rp1 - rc1) * av[r1]);
#pragma unroll
                for (int k = 0; k < NV; ++k) rh[k] = fma(g0[k], q0, fma(g1[k], q1, rh[k]));
            }
            reduce_all<NV>(rh, red, scr, lane);
#pragma unroll
            for (int k = 0; k < NV; ++k) dz[k] = -cres * hzf[k] - red[k];
            __syncwarp();
            // column-oriented triangular solves with the replicated factor (chain length 2/step)
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                dz[k] *= inv[k];
#pragma unroll
                for (int i = k + 1; i < NV; ++i) dz[i] = fma(-Ls[tri(i, k)], dz[k], dz[i]);
            }
#pragma unroll
            for (int k = NV - 1; k >= 0; --k) {
                dz[k] *= inv[k];
#pragma unroll
                for (int i = 0; i < k; ++i) dz[i] = fma(-Ls[tri(k, i)], dz[k], dz[i]);
            }
            // step to the boundary: min over rows of s/(-ds) and lam/(-dl), as fractions
            double an = 1.0, ad = 1e-300;   // num/den = "infinite" step
            an = 1e300;
            ad = 1.0;
#pragma unroll 1
            for (int j = 0; j < MR; j += 2) {
                if (!(pairs >> j & 1u)) continue;
                const bool on1 = (j + 1 < MR);
                const int r0 = lane + 32 * j, r1 = on1 ? r0 + 32 : r0;
                const double m1 = (on1 && (act >> (j + 1) & 1u)) ? 1.0 : 0.0;
                const double m0 = (act >> j & 1u) ? 1.0 : 0.0;
                double gd0 = 0.0, gd1 = 0.0;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    gd0 = fma(Gs[k * MRP + r0], dz[k], gd0);
                    gd1 = fma(Gs[k * MRP + r1], dz[k], gd1);
                }
                const double s0 = sv[r0], s1 = sv[r1], l0 = lv[r0], l1 = lv[r1];
                const double rp0 = rv[r0], rp1 = rv[r1];
                const double rc0 = trip ? bv[r0] : s0 * l0, rc1 = trip ? bv[r1] : s1 * l1;
                const double ds0 = m0 * (-cres * rp0 - gd0), ds1 = m1 * (-cres * rp1 - gd1);
                const double dl0 = m0 * (-rc0 - l0 * ds0) * av[r0], dl1 = m1 * (-rc1 - l1 * ds1) * av[r1];
                if (ds0 < 0.0 && s0 * ad < an * (-ds0)) { an = s0; ad = -ds0; }
                if (dl0 < 0.0 && l0 * ad < an * (-dl0)) { an = l0; ad = -dl0; }
                if (ds1 < 0.0 && s1 * ad < an * (-ds1)) { an = s1; ad = -ds1; }
                if (dl1 < 0.0 && l1 * ad < an * (-dl1)) { an = l1; ad = -dl1; }
                if (trip) {
                    av[r0] = ds0;
                    bv[r0] = dl0;
                    if (on1) {
                        av[r1] = ds1;
                        bv[r1] = dl1;
                    }
                } else {
                    bv[r0] = ds0 * dl0;
                    if (on1) bv[r1] = ds1 * dl1;
                }
            }
            const double amax = warp_min_nonneg(an / ad);
            if (trip == 0) {
                const double a_aff = fmin(1.0, amax);
                sig = (1.0 - a_aff) * (1.0 - a_aff) * (1.0 - a_aff);
                cres = 1.0 - sig;
            } else {
                alpha = fmin(1.0, 0.99 * amax);
            }
        }
        // ---- step ----------------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < NV; ++k) z[k] = fma(alpha, dz[k], z[k]);
#pragma unroll 1
        for (int j = 0; j < MR; ++j) {
            if (!(act >> j & 1u)) continue;
            const int r = lane + 32 * j;
            sv[r] = fma(alpha, av[r], sv[r]);
            lv[r] = fma(alpha, bv[r], lv[r]);
        }
        rho *= 1.0 - alpha * cres;
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) w[W::oZ + k] = z[k];
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
