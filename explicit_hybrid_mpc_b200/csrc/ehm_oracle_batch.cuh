// ehm_oracle_batch.cuh -- the reference's four set-valued oracle calls as batched C-ABI entry
// points (included by ehm_lib.cu after the partition engine, whose stage kernels they reuse):
//
//   ehm_p_theta_batch   Oracle.P_theta        lib/oracle.py:104-139
//   ehm_v_r_batch       Oracle.V_R            lib/oracle.py:175-218
//   ehm_bar_e_batch     Oracle.bar_E_delta_R  lib/oracle.py:285-309
//   ehm_bar_d_batch     Oracle.bar_D_delta_R  lib/oracle.py:311-414 (+ in_variability_ball :220-283)
//
// Every call enumerates all K^N commutations on the device (MOSEK's branch and bound in the
// reference) and applies the canonical selection rules of DESIGN.md with device-side argmin /
// ranking kernels: no torch reductions, no host round trip per candidate.  The caller's arrays are
// viewed as a node store (entry e = node e), so the kernels of the fused partition loop
// (k_ecc_points / k_ecc_pick, k_lcss_ref / items / eval / select / points) run unchanged.
#pragma once

namespace ehm {

struct OracleScratch {   // grow-only device scratch of the batched calls, owned by the problem
    int cap = 0;
    WaveBuf wv{};
    EccBuf eb{};
    Ctl *dctl = nullptr;
    int *more = nullptr;   // [1] entries that need another vetting round
    int *neg1 = nullptr;   // [cap] constant -1: the commutation of entries that have none (bar_E)
    std::vector<void *> allocs;
    ~OracleScratch()
    {
        for (void *p : allocs) cudaFree(p);
    }
};

__global__ void k_batch_init(WaveBuf wv, EccBuf eb, int B, int ncand, Ctl *ctl, int *more, int *neg1)
{
    const size_t total = (size_t)B * ncand;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        wv.lb[i] = -INFINITY;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int e = t; e < B; e += gridDim.x * blockDim.x) {
        wv.node[e] = e;
        wv.skip[e] = 0;
        wv.tref[e] = INFINITY;
        eb.queue[e] = e;
        neg1[e] = -1;
    }
    if (t == 0) {
        ctl->nitems = 0;
        ctl->npitems = 0;
        ctl->need_centroid = 0;
        ctl->error = 0;
        for (int k = 0; k < 8; ++k) ctl->cur[k] = 0ull;
        *more = 0;
    }
}

__global__ void k_batch_round(Ctl *ctl, int *more)
{
    if (threadIdx.x == 0) {
        ctl->npitems = 0;
        for (int k = 0; k < 8; ++k) ctl->cur[k] = 0ull;
        *more = 0;
    }
}

// P_theta / V_R results from the ECC buffers (one thread per entry)
template <int NX>
__global__ void k_batch_pick_out(DevTables tb, EccBuf eb, int B, int nslots, int *cand, double *J, double *U,
                                 int *status)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    const int c = eb.pick[e], nu = tb.nu;
    cand[e] = c;
    if (status) status[e] = c >= 0 ? 1 : 0;
    for (int i = 0; i < nslots; ++i) {
        const size_t src = c >= 0 ? ((size_t)e * tb.ncand + c) : 0;
        if (J) J[(size_t)e * nslots + i] = c >= 0 ? eb.J[src * (NX + 2) + i] : INFINITY;
        if (U)
            for (int q = 0; q < nu; ++q)
                U[((size_t)e * nslots + i) * nu + q] = c >= 0 ? eb.U[(src * (NX + 1) + i) * nu + q] : 0.0;
    }
}

// bar_E: converged flags (+ optional minimum of t* over the candidates, +inf if none is alive)
__global__ void k_batch_bare_out(WaveBuf wv, int B, int ncand, int *converged, double *tmin)
{
    const int lane = threadIdx.x & 31;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= B) return;
    double best = INFINITY;
    for (int c = lane; c < ncand; c += 32) best = fmin(best, wv.tval[(size_t)e * ncand + c]);
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(kFull, best, o));
    if (lane == 0) {
        converged[e] = wv.conv[e];
        if (tmin) tmin[e] = best;
    }
}

// bar_D: first vetted candidate that is feasible at all vertices, or another round, or "none"
template <int NX>
__global__ void k_batch_bard_out(DevTables tb, NodeStore ns, WaveBuf wv, int B, int *cstar, double *theta_star,
                                 double *Jv, double *Uv, int *var_small, int *status, int *more)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    const int ncand = tb.ncand, nu = tb.nu;
    int out = -1, vs = 0, st = 1;
    const int nsel = wv.nsel[e];
    int kstar = -1;
    if (!wv.conv[e]) {
        for (int k = 0; k < nsel && kstar < 0; ++k) {
            bool ok = true;
            for (int i = 0; i <= NX; ++i) ok = ok && wv.ptS[((size_t)e * (kSelT + 1) + k) * (NX + 2) + i] == 1;
            if (ok) kstar = k;
        }
        if (kstar < 0 && nsel == kSelT && wv.nimp[e] > wv.skip[e] + kSelT) {
            wv.skip[e] += kSelT;   // more improving candidates than vetted so far: another round
            atomicAdd(more, 1);
            return;
        }
    }
    if (kstar >= 0) {
        if (wv.minv_st[e] != 1) {
            st = -1;   // min_R V*_ref failed numerically (the reference would blacklist and retry)
        } else {
            out = wv.sel[e * kSelT + kstar];
            const int item = wv.titem[(size_t)e * ncand + out];
            const double *sol = wv.it_sol + (size_t)item * (NX + 1);
            const double *R = ns.verts + (size_t)e * (NX + 1) * NX, *Vr = ns.vcost + (size_t)e * (NX + 1);
            double maxV = Vr[0];
            for (int i = 1; i <= NX; ++i) maxV = fmax(maxV, Vr[i]);
            vs = (maxV - wv.minv[e]) < fmax(tb.eps_a, tb.eps_r * sol[NX]) ? 1 : 0;   // oracle.py:282-283
            for (int j = 0; j < NX; ++j) {
                double v = R[j];
                for (int i = 0; i < NX; ++i) v = fma(R[(i + 1) * NX + j] - R[j], sol[i], v);
                theta_star[(size_t)e * NX + j] = v;
            }
            const size_t ps = ((size_t)e * (kSelT + 1) + kstar) * (NX + 2);
            for (int i = 0; i <= NX; ++i) {
                Jv[(size_t)e * (NX + 1) + i] = wv.ptJ[ps + i];
                for (int q = 0; q < nu; ++q) Uv[((size_t)e * (NX + 1) + i) * nu + q] = wv.ptU[(ps + i) * nu + q];
            }
        }
    }
    cstar[e] = out;
    var_small[e] = vs;
    if (status) status[e] = st;
}

}  // namespace ehm
