// ehm_ipm.cuh -- warp-level dense primal-dual interior-point solver (sm_100a, fp64).
//
// One warp solves one small convex QP/LP
//        min 0.5 z'Hz + f'z   s.t.  G z <= h            (NV <= 10 variables, <= 32*MR rows)
// held entirely in shared memory + registers.  It replaces one `.solve()` of the
// reference (CVXPY -> MOSEK, lib/oracle.py:131,134,166,169,203,276,305,350).
//
// Mapping to the hardware (version 2 -- sized for the instruction cache):
//   * lane L owns rows L, L+32, ...; G is column-major in shared memory (32 lanes read 32
//     consecutive doubles: conflict free); the per-row state (slack s, multiplier lam and the
//     two step vectors) also lives in shared memory so that the row loops can stay ROLLED.
//     The first version kept that state in registers, which forced a 7x unroll of every
//     row pass: 25k SASS instructions, and ncu showed the warps stalled on instruction
//     fetch (stall_no_instruction 5.5 per issue, profiles/README.md).  The hot loop is now
//     ~1.5k instructions and fits the 32 KB L1.5 instruction cache.
//   * the normal matrix K = H + G' diag(lam/s) G is accumulated per lane in NV(NV+1)/2
//     registers (static indices) and combined with a recursive-halving reduce-scatter
//     (62 shuffles instead of 5*55), exchanged through a 64-double scratch line.
//   * Cholesky is lane-parallel: lane i holds row i of K in registers, column j is formed
//     with j shuffle broadcasts; the factor is then replicated into every lane's registers so
//     that the two triangular solves of the predictor / corrector are shuffle-free.
//   * predictor and corrector share one code path (a 2-trip loop), step lengths are
//     warp-shuffle min reductions.
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
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

__host__ __device__ constexpr int pow2_ge(int x) { return x <= 1 ? 1 : 2 * pow2_ge((x + 1) / 2); }

// Recursive-halving reduce-scatter of P per-lane partials (P a power of two).  After the
// call lane holds LF = max(P/32,1) fully reduced values v[0..LF) whose global indices start
// at `base`.
template <int L, int B>
struct ReduceScatter {
    template <int P>
    static __device__ __forceinline__ void run(double (&v)[P], int lane, int &base)
    {
        if constexpr (B >= 1) {
            if constexpr (L > 1) {
                constexpr int H = L / 2;
                const bool up = (lane & B) != 0;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const double send = up ? v[i] : v[i + H];
                    const double keep = up ? v[i + H] : v[i];
                    v[i] = keep + __shfl_xor_sync(kFull, send, B);
                }
                if (up) base += H;
                ReduceScatter<H, B / 2>::run(v, lane, base);
            } else {
                v[0] += __shfl_xor_sync(kFull, v[0], B);
                ReduceScatter<1, B / 2>::run(v, lane, base);
            }
        }
    }
};

// Sum P per-lane partials over the warp and leave the first `cnt` sums in out[0..cnt)
// (shared memory), visible to every lane on return.
template <int P>
__device__ __forceinline__ void warp_reduce_to_smem(double (&v)[P], double *out, int cnt, int lane)
{
    int base = 0;
    ReduceScatter<P, 16>::run(v, lane, base);
    constexpr int LF = (P >= 32) ? P / 32 : 1;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < LF; ++i)
        if (base + i < cnt) out[base + i] = v[i];
    __syncwarp();
}

// Packed lower-triangular index.
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// Guarded pivot of the Cholesky factorisation (oracle/ehm_oracle.c:cholesky): returns
// (sqrt(d), 1/sqrt(d)); a pivot that collapsed below 1e-13 of its own original diagonal
// freezes the direction.
__device__ __noinline__ double2 chol_pivot(double d, double d0)
{
    if (d <= 1e-13 * d0 || d0 <= 0.0) d = 1e100;
    const double sd = sqrt(d);
    return make_double2(sd, 1.0 / sd);
}

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
    static constexpr int oHs = oB + MRP;         // Hessian NV*NV row-major (zeros for an LP)
    static constexpr int oF = oHs + NV * NV;     // f (16)
    static constexpr int oZ = oF + 16;           // z in/out (16)
    static constexpr int oRed = oZ + 16;         // reduction scratch / packed K / packed L (64)
    static constexpr int oOut = oRed + 64;       // [0] objective [1] iterations [2] flops (8)
    static constexpr int kDoubles = oOut + 8;
};

// Solve  min 0.5 z'Hz + f'z  s.t.  G z <= h  with one warp.  `w` = base of an IpmWs block:
// G, h, Hs, f and the primal start z must be filled in; act: bit j set <=> row lane+32*j takes
// part (normalised, non-constant row).  On return z holds the solution, out[0..2] the
// objective / iteration count / algorithmic flops.  returns 0 converged, 1 iteration limit,
// 2 numerical breakdown.
template <int NV, int MR>
__device__ __noinline__ int ipm_solve(double *w, unsigned act, int lane)
{
    using W = IpmWs<NV, MR>;
    constexpr int MRP = W::MRP;
    constexpr int NK = NV * (NV + 1) / 2;
    constexpr int PK = pow2_ge(NK);
    constexpr int PV = pow2_ge(NV);
    const double *__restrict__ Gs = w + W::oG;
    const double *__restrict__ hs = w + W::oH;
    double *sv = w + W::oS, *lv = w + W::oL, *av = w + W::oA, *bv = w + W::oB;
    const double *__restrict__ Hs = w + W::oHs;
    double *red = w + W::oRed;
    double z[NV], f[NV];
    double fmx = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        f[k] = w[W::oF + k];
        z[k] = w[W::oZ + k];
        fmx = fmax(fmx, fabs(f[k]));
    }
    const double mact = warp_sum((double)__popc(act));
    // ---- centred start + initial dual residual rd0 = |H z0 + f + G' lam0| -----------------
    double rd0;
    {
        double gl[PV];
#pragma unroll
        for (int k = 0; k < PV; ++k) gl[k] = 0.0;
#pragma unroll 1
        for (int j = 0; j < MR; ++j) {
            if (!(act >> j & 1u)) continue;
            const int r = lane + 32 * j;
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
#pragma unroll
            for (int k = 0; k < NV; ++k) gl[k] = fma(g[k], l, gl[k]);
        }
        warp_reduce_to_smem<PV>(gl, red, NV, lane);
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double hz = f[k] + red[k];
#pragma unroll
            for (int l = 0; l < NV; ++l) hz = fma(Hs[k * NV + l], z[l], hz);
            m = fmax(m, fabs(hz));
        }
        rd0 = m;
        __syncwarp();
    }
    int status = 1, it = 0;
    double rho = 1.0, obj = 0.0;
#pragma unroll 1
    for (it = 0; it <= EHM_MAX_ITER; ++it) {
        // ---- pass A: primal residual, gap, normal matrix ------------------------------------
        double acc[PK];
#pragma unroll
        for (int k = 0; k < PK; ++k) acc[k] = 0.0;
        double mu = 0.0, rpm = 0.0, lmx = 0.0;
#pragma unroll 1
        for (int j = 0; j < MR; ++j) {
            if (!(act >> j & 1u)) continue;
            const int r = lane + 32 * j;
            double g[NV], gz = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                g[k] = Gs[k * MRP + r];
                gz = fma(g[k], z[k], gz);
            }
            const double s = sv[r], l = lv[r];
            const double rp = gz + s - hs[r];
            av[r] = rp;
            rpm = fmax(rpm, fabs(rp));
            mu = fma(s, l, mu);
            lmx = fmax(lmx, l);
            const double wt = l / s;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const double gw = wt * g[k];
#pragma unroll
                for (int q = 0; q <= k; ++q) acc[tri(k, q)] = fma(gw, g[q], acc[tri(k, q)]);
            }
        }
        mu = warp_sum(mu);
        rpm = warp_max(rpm);
        lmx = warp_max(lmx);
        // (H z + f): lane k forms row k, exchanged through the scratch line
        double hzf[NV];
        {
            if (lane < NV) {
                double v = w[W::oF + lane];
#pragma unroll
                for (int q = 0; q < NV; ++q) v = fma(Hs[lane * NV + q], z[q], v);
                red[lane] = v;
            }
            __syncwarp();
            double zm = 0.0;
            obj = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                hzf[k] = red[k];
                obj = fma(z[k], 0.5 * (hzf[k] + f[k]), obj);
                zm = fmax(zm, fabs(z[k]));
            }
            __syncwarp();
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
        if (it == EHM_MAX_ITER) break;
        const double mu_mean = mu / mact;
        // ---- K -> packed scratch -> lane-parallel Cholesky -> replicated factor --------------
        warp_reduce_to_smem<PK>(acc, red, NK, lane);
        double L[NK], inv[NV];
        {
            double Lr[NV], d0 = 1.0;
            const int li = lane < NV ? lane : 0;
#pragma unroll
            for (int k = 0; k < NV; ++k)
                Lr[k] = (lane < NV && k <= lane) ? red[tri(li, k)] + Hs[li * NV + k] : 0.0;
            if (lane < NV) d0 = red[tri(li, li)] + Hs[li * NV + li];
            __syncwarp();
            bool bad = false;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                double v = Lr[j];
#pragma unroll
                for (int k = 0; k < j; ++k) v = fma(-Lr[k], __shfl_sync(kFull, Lr[k], j), v);
                const double d = __shfl_sync(kFull, v, j);
                const double dd0 = __shfl_sync(kFull, d0, j);
                if (!isfinite(d)) bad = true;
                const double2 pv = chol_pivot(d, dd0);
                inv[j] = pv.y;
                Lr[j] = (lane == j) ? pv.x : ((lane > j && lane < NV) ? v * pv.y : 0.0);
            }
            if (bad) {
                status = 2;
                break;
            }
            if (lane < NV) {
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    if (k <= lane) red[tri(li, k)] = Lr[k];
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NK; ++k) L[k] = red[k];
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
            for (int j = 0; j < MR; ++j) {
                if (!(act >> j & 1u)) continue;
                const int r = lane + 32 * j;
                const double s = sv[r], l = lv[r], rp = av[r];
                // rc: s.lam (predictor) or s.lam + ds_aff.dl_aff - sigma mu (corrector)
                double rc = s * l;
                if (trip) {
                    rc = rc + bv[r] - sig * mu_mean;
                    bv[r] = rc;
                }
                // rhs = -c (Hz+f) - G'(c lam + t),  t = (lam c rp - rc)/s
                const double q = fma(cres, l, (l * cres * rp - rc) / s);
#pragma unroll
                for (int k = 0; k < NV; ++k) rh[k] = fma(Gs[k * MRP + r], q, rh[k]);
            }
            warp_reduce_to_smem<PV>(rh, red, NV, lane);
#pragma unroll
            for (int k = 0; k < NV; ++k) dz[k] = -cres * hzf[k] - red[k];
            __syncwarp();
            // triangular solves with the replicated factor
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                double v = dz[i];
#pragma unroll
                for (int k = 0; k < i; ++k) v = fma(-L[tri(i, k)], dz[k], v);
                dz[i] = v * inv[i];
            }
#pragma unroll
            for (int i = NV - 1; i >= 0; --i) {
                double v = dz[i];
#pragma unroll
                for (int k = i + 1; k < NV; ++k) v = fma(-L[tri(k, i)], dz[k], v);
                dz[i] = v * inv[i];
            }
            double amax = 1e300;
#pragma unroll 1
            for (int j = 0; j < MR; ++j) {
                if (!(act >> j & 1u)) continue;
                const int r = lane + 32 * j;
                double gdz = 0.0;
#pragma unroll
                for (int k = 0; k < NV; ++k) gdz = fma(Gs[k * MRP + r], dz[k], gdz);
                const double s = sv[r], l = lv[r];
                const double rc = trip ? bv[r] : s * l;
                const double ds = -cres * av[r] - gdz;
                const double dl = (-rc - l * ds) / s;
                if (ds < 0.0) amax = fmin(amax, -s / ds);
                if (dl < 0.0) amax = fmin(amax, -l / dl);
                if (trip) {
                    av[r] = ds;
                    bv[r] = dl;
                } else {
                    bv[r] = ds * dl;
                }
            }
            amax = warp_min(amax);
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
    if (lane < NV) {
        // static indexing of z by the owning lane
#pragma unroll
        for (int k = 0; k < NV; ++k)
            if (lane == k) w[W::oZ + k] = z[k];
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
