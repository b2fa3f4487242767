// ehm_ipm.cuh -- warp-level dense primal-dual interior-point solver (sm_100a, fp64).
//
// One warp solves one small convex QP/LP
//        min 0.5 z'Hz + f'z   s.t.  G z <= h            (NV <= 10 variables, <= 32*MR rows)
// held entirely in shared memory + registers.  It replaces one `.solve()` of the
// reference (CVXPY -> MOSEK, lib/oracle.py:131,134,166,169,203,276,305,350).
//
// Mapping to the hardware
//   * lane L owns rows L, L+32, ... (MR of them): slack s, multiplier lam and the step
//     vectors live in registers; G is stored column-major in shared memory so that the 32
//     lanes read 32 consecutive doubles (conflict free, 256 B per LDS.64 wavefront pair).
//   * the normal matrix K = H + G' diag(lam/s) G is accumulated per lane in NV(NV+1)/2
//     registers and combined with a recursive-halving reduce-scatter (31 or 62 shuffles
//     instead of 5*NK), the packed result is exchanged through a 64-double scratch line.
//   * every lane then factors K redundantly in registers (fully unrolled Cholesky, static
//     indexing) -- SIMT makes the redundancy free and avoids any further exchange for the
//     two triangular solves of the predictor and the corrector.
//   * step length = warp-shuffle min reduction.
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

// In-register Cholesky of the packed SPD matrix (lower).  inv[j] = 1/L_jj.
// Pivot guard as in oracle/ehm_oracle.c:cholesky.  Returns false on NaN/Inf.
template <int NV>
__device__ __forceinline__ bool chol_packed(double (&L)[NV * (NV + 1) / 2], double (&inv)[NV])
{
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const double d0 = L[tri(j, j)];
        double d = d0;
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-L[tri(j, k)], L[tri(j, k)], d);
        if (!isfinite(d)) ok = false;
        if (d <= 1e-13 * d0 || d0 <= 0.0) d = 1e100;
        d = sqrt(d);
        const double id = 1.0 / d;
        L[tri(j, j)] = d;
        inv[j] = id;
#pragma unroll
        for (int i = j + 1; i < NV; ++i) {
            double v = L[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-L[tri(i, k)], L[tri(j, k)], v);
            L[tri(i, j)] = v * id;
        }
    }
    return ok;
}

template <int NV>
__device__ __forceinline__ void chol_solve(const double (&L)[NV * (NV + 1) / 2],
                                           const double (&inv)[NV], double (&b)[NV])
{
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v = fma(-L[tri(i, k)], b[k], v);
        b[i] = v * inv[i];
    }
#pragma unroll
    for (int i = NV - 1; i >= 0; --i) {
        double v = b[i];
#pragma unroll
        for (int k = i + 1; k < NV; ++k) v = fma(-L[tri(k, i)], b[k], v);
        b[i] = v * inv[i];
    }
}

struct IpmStats {
    int iters;
};

// Solve  min 0.5 z'Hz + f'z  s.t.  G z <= h  with one warp.
//   Gs   : shared, column-major, column k at Gs + k*MRP, MRP = 32*MR rows (pad rows unused)
//   hs   : shared, MRP
//   act  : bit j set <=> row lane+32*j takes part (normalised, non-constant row)
//   Hs   : shared, NV*NV row-major (ignored when !HASH), fs: shared NV
//   red  : shared scratch, >= 64 doubles, private to the warp
//   z    : in = primal start, out = solution (replicated in every lane)
// returns 0 converged, 1 iteration limit, 2 numerical breakdown.
template <int NV, int MR, bool HASH>
__device__ __noinline__ int ipm_solve(const double *__restrict__ Gs, const double *__restrict__ hs,
                                      unsigned act, const double *__restrict__ Hs,
                                      const double *__restrict__ fs, double *red,
                                      double (&z)[NV], double &obj_out, int &iters_out, int lane,
                                      double *flops_out = nullptr)
{
    constexpr int MRP = 32 * MR;
    constexpr int NK = NV * (NV + 1) / 2;
    constexpr int PK = pow2_ge(NK);
    constexpr int PV = pow2_ge(NV);
    double s[MR], lam[MR], a[MR], b[MR];
    double f[NV];
    double fmx = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        f[k] = fs[k];
        fmx = fmax(fmx, fabs(f[k]));
    }
    const double mact = warp_sum((double)__popc(act));
#pragma unroll
    for (int j = 0; j < MR; ++j) {
        const int r = lane + 32 * j;
        if (act >> j & 1u) {
            double gz = 0.0;
#pragma unroll
            for (int k = 0; k < NV; ++k) gz = fma(Gs[k * MRP + r], z[k], gz);
            s[j] = fmax(hs[r] - gz, 1.0);
            lam[j] = 1.0 / s[j];
        } else {
            s[j] = 1.0;
            lam[j] = 0.0;
        }
        a[j] = 0.0;
        b[j] = 0.0;
    }
    int status = 1, it = 0;
    double rho = 1.0, rd0 = 0.0, obj = 0.0;
    for (it = 0; it <= EHM_MAX_ITER; ++it) {
        // ---- residuals + normal matrix in one pass over G -------------------------------
        double gl[PV], acc[PK];
#pragma unroll
        for (int k = 0; k < PV; ++k) gl[k] = 0.0;
#pragma unroll
        for (int k = 0; k < PK; ++k) acc[k] = 0.0;
        double mu = 0.0, rpm = 0.0, lmx = 0.0;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            if (act >> j & 1u) {
                const int r = lane + 32 * j;
                double g[NV], gz = 0.0;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    g[k] = Gs[k * MRP + r];
                    gz = fma(g[k], z[k], gz);
                }
                const double rp = gz + s[j] - hs[r];
                a[j] = rp;
                rpm = fmax(rpm, fabs(rp));
                mu = fma(s[j], lam[j], mu);
                lmx = fmax(lmx, lam[j]);
                const double w = lam[j] / s[j];
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    gl[k] = fma(g[k], lam[j], gl[k]);
                    const double gw = w * g[k];
#pragma unroll
                    for (int l = 0; l <= k; ++l) acc[tri(k, l)] = fma(gw, g[l], acc[tri(k, l)]);
                }
            }
        }
        mu = warp_sum(mu);
        rpm = warp_max(rpm);
        lmx = warp_max(lmx);
        warp_reduce_to_smem<PV>(gl, red, NV, lane);
        double rd[NV], rdm = 0.0, zm = 0.0;
        obj = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double hz = 0.0;
            if (HASH) {
#pragma unroll
                for (int l = 0; l < NV; ++l) hz = fma(Hs[k * NV + l], z[l], hz);
            }
            obj += z[k] * (f[k] + 0.5 * hz);
            rd[k] = f[k] + hz + red[k];
            rdm = fmax(rdm, fabs(rd[k]));
            zm = fmax(zm, fabs(z[k]));
        }
        __syncwarp();
        if (it == 0) rd0 = rdm;
        if (!isfinite(mu) || !isfinite(rdm) || !isfinite(rpm)) {
            status = 2;
            break;
        }
        if (rpm <= EHM_TOL_IPM * (1.0 + zm) && rho * rd0 <= EHM_TOL_IPM * (1.0 + fmx + lmx) &&
            rdm <= EHM_TOL_DUAL * (1.0 + fmx + lmx) && mu <= EHM_TOL_IPM * (1.0 + fabs(obj))) {
            status = 0;
            break;
        }
        if (it == EHM_MAX_ITER) break;
        const double mu_mean = mu / mact;
        // ---- factor K ---------------------------------------------------------------------
        warp_reduce_to_smem<PK>(acc, red, NK, lane);
        double L[NK], inv[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int l = 0; l <= k; ++l)
                L[tri(k, l)] = red[tri(k, l)] + (HASH ? Hs[k * NV + l] : 0.0);
        __syncwarp();
        if (!chol_packed<NV>(L, inv)) {
            status = 2;
            break;
        }
        // ---- predictor ---------------------------------------------------------------------
        double rh[PV];
#pragma unroll
        for (int k = 0; k < PV; ++k) rh[k] = 0.0;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            if (act >> j & 1u) {
                const int r = lane + 32 * j;
                const double t = fma(lam[j], a[j] / s[j], -lam[j]); // (-s.lam + lam rp)/s
#pragma unroll
                for (int k = 0; k < NV; ++k) rh[k] = fma(Gs[k * MRP + r], t, rh[k]);
            }
        }
        warp_reduce_to_smem<PV>(rh, red, NV, lane);
        double dz[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) dz[k] = -rd[k] - red[k];
        __syncwarp();
        chol_solve<NV>(L, inv, dz);
        double amax = 1e300;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            if (act >> j & 1u) {
                const int r = lane + 32 * j;
                double gdz = 0.0;
#pragma unroll
                for (int k = 0; k < NV; ++k) gdz = fma(Gs[k * MRP + r], dz[k], gdz);
                const double ds = -a[j] - gdz;
                const double dl = -lam[j] - lam[j] * ds / s[j];
                if (ds < 0.0) amax = fmin(amax, -s[j] / ds);
                if (dl < 0.0) amax = fmin(amax, -lam[j] / dl);
                b[j] = ds * dl;
            }
        }
        const double a_aff = fmin(1.0, warp_min(amax));
        const double sig = (1.0 - a_aff) * (1.0 - a_aff) * (1.0 - a_aff);
        const double resc = 1.0 - sig;
        // ---- corrector ---------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < PV; ++k) rh[k] = 0.0;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            if (act >> j & 1u) {
                const int r = lane + 32 * j;
                const double rc = fma(s[j], lam[j], b[j]) - sig * mu_mean;
                b[j] = rc;
                const double t = (lam[j] * resc * a[j] - rc) / s[j];
#pragma unroll
                for (int k = 0; k < NV; ++k) rh[k] = fma(Gs[k * MRP + r], t, rh[k]);
            }
        }
        warp_reduce_to_smem<PV>(rh, red, NV, lane);
#pragma unroll
        for (int k = 0; k < NV; ++k) dz[k] = -resc * rd[k] - red[k];
        __syncwarp();
        chol_solve<NV>(L, inv, dz);
        amax = 1e300;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            if (act >> j & 1u) {
                const int r = lane + 32 * j;
                double gdz = 0.0;
#pragma unroll
                for (int k = 0; k < NV; ++k) gdz = fma(Gs[k * MRP + r], dz[k], gdz);
                const double ds = -resc * a[j] - gdz;
                const double dl = (-b[j] - lam[j] * ds) / s[j];
                if (ds < 0.0) amax = fmin(amax, -s[j] / ds);
                if (dl < 0.0) amax = fmin(amax, -lam[j] / dl);
                a[j] = ds;
                b[j] = dl;
            }
        }
        const double alpha = fmin(1.0, 0.99 * warp_min(amax));
#pragma unroll
        for (int k = 0; k < NV; ++k) z[k] = fma(alpha, dz[k], z[k]);
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            if (act >> j & 1u) {
                s[j] = fma(alpha, a[j], s[j]);
                lam[j] = fma(alpha, b[j], lam[j]);
            }
        }
        rho *= 1.0 - alpha * resc;
    }
    obj_out = obj;
    iters_out = it;
    // algorithmic flops (2 per FMA): K build m*NV(NV+1), six mat-vecs 12*m*NV, Cholesky NV^3/3,
    // two triangular solve pairs 4*NV^2 -- per iteration
    if (flops_out)
        *flops_out += (double)it * (mact * (NV * (NV + 1) + 12.0 * NV) + NV * NV * NV / 3.0 + 4.0 * NV * NV);
    return status;
}

}  // namespace ehm
