// ehm_partition.cuh -- the device-resident partition loop (included by ehm_lib.cu).
//
// Replaces, for the hot path, the reference's MPI task farm and recursion:
//   Scheduler.task_queue / work_dispatcher (lib/scheduler.py:364-599)  -> LIFO frontier in HBM
//   Worker.ecc  (lib/worker.py:241-291)                                -> ECC waves
//   Worker.lcss (lib/worker.py:293-417)                                -> L-CSS waves
//   tools.split_along_longest_edge (lib/tools.py:224-257)              -> k_lcss_select / finalize
//   tree.Tree / NodeData (lib/tree.py)                                 -> flat node store
//
// One wave pops up to W frontier entries and pushes them through a fixed pipeline of
// persistent kernels whose work lists are built on the device (dynamic cursors, no host
// round trip between stages; the host reads one small control block per wave):
//
//   k_pop          move the popped entries (node id + per-candidate bounds) to wave buffers
//   k_lcss_ref     t*_ref of every entry's own commutation            (1 warp / entry)
//   k_lcss_items   candidates whose inherited bound does not exclude them -> item list
//   k_lcss_eval    phase-1 + epigraph QP of every item                (1 warp / item)
//   k_lcss_select  convergence test, ranked improving candidates, split geometry,
//                  list of vertex / midpoint / variability problems
//   k_lcss_points  those problems                                     (1 warp / problem)
//   k_lcss_final   delta*, var_small, close / swap / split, push children
//
// Canonical semantics: DESIGN.md "Canonical oracle semantics" (identical to
// oracle/ehm_oracle.c:ehmo_lcss_eval; the inherited bounds only skip work).
#pragma once
#include <cfloat>

namespace ehm {

constexpr int kSelT = 4;       // improving candidates vetted per visit (rest: revisit)
constexpr int kEccWave = 16;   // ECC nodes per ECC wave

struct Ctl {
    unsigned long long cur[8];  // dynamic work cursors of the stage kernels
    int ftop;                   // frontier stack top
    int nnodes;
    int nleaves;
    int nitems;                 // L-CSS evaluation items of this wave
    int npitems;                // point / minv items of this wave
    int error;                  // EHM_E* (0 = ok)
    int max_depth;
    int necc_next;              // ECC queue tail
    int need_centroid;          // ECC entries without a common commutation
    int pad;
};

struct NodeStore {
    int cap;
    double *verts, *vcost, *vinput;
    int *cand, *left, *right, *root_id, *depth;
    unsigned char *closed;   // 0 open / interior, 1 eps-suboptimal leaf, 2 exported to another rank
    int *origin;             // [cap][2] (rank, node) this node continues, -1 = local
};

struct Frontier {
    int cap;
    int *node;
    int *skip;     // rank offset for the delta* search (see k_lcss_select)
    float *lb;     // [cap][ncand]
};

struct PItem {  // point / minv problem of the L-CSS pipeline
    int e;      // wave entry
    int k;      // candidate slot 0..kSelT-1, kSelT = delta_ref
    int i;      // vertex 0..n, n+1 = midpoint, -1 = min over simplex
    int c;      // candidate
};

struct WaveBuf {
    int W;
    int *node, *skip;
    float *lb;          // [W][ncand]
    double *tref;       // [W]
    double *tval;       // [W][ncand]   t*_c of this visit (+inf = not improving / skipped)
    int *titem;         // [W][ncand]   item index of (e,c) or -1
    int2 *items;        // [W*ncand]
    double *it_sol;     // [W*ncand][n+1]
    float *it_pl;       // [W*ncand][n+1]  tangent plane of the item's value function (ehm_solve.cuh)
    int *sel;           // [W][kSelT]
    int *nsel, *nimp, *conv, *ij;
    double *vmid;       // [W][n]
    PItem *pitems;
    double *ptJ;        // [W][kSelT+1][n+2]
    double *ptU;        // [W][kSelT+1][n+2][nu]
    int *ptS;
    double *minv;
    int *minv_st;
};

struct EccBuf {
    int *queue;         // node ids awaiting ECC (cap = node store cap)
    double *J;          // [kEccWave][ncand][n+2]   (slot n+1 = barycentre)
    double *U;          // [kEccWave][ncand][n+1][nu]
    int *S;             // [kEccWave][ncand][n+2]
    int *pick;          // [kEccWave]
};

__device__ __forceinline__ long long grab(unsigned long long *cur, int lane)
{
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(cur, 1ull);
    return (long long)__shfl_sync(kFull, v, 0);
}

__device__ __forceinline__ float f32_down(double t) { return __double2float_rd(t); }

// ------------------------------------------------------------------------------------------
// geometry helpers (single thread)
// ------------------------------------------------------------------------------------------
template <int NX>
__device__ void longest_edge(const double *R, int &bi, int &bj)
{
    double best = -1.0;
    bi = 0;
    bj = 1;
    for (int i = 0; i <= NX; ++i)
        for (int j = i + 1; j <= NX; ++j) {
            double acc = 0.0;
            for (int k = 0; k < NX; ++k) {
                const double d = __dsub_rn(R[i * NX + k], R[j * NX + k]);
                acc = __fma_rn(d, d, acc);   // the reference's la.norm arithmetic, see k_split_batch
            }
            const double len = __dsqrt_rn(acc);
            if (len > best) {
                best = len;
                bi = i;
                bj = j;
            }
        }
}

// ------------------------------------------------------------------------------------------
// k_pop: move the top W entries of the frontier into the wave buffers
// ------------------------------------------------------------------------------------------
__global__ void k_pop(Frontier fr, WaveBuf wv, int base, int W, int ncand, Ctl *ctl)
{
    const size_t total = (size_t)W * ncand;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
        wv.lb[i] = fr.lb[(size_t)base * ncand + i];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int q = t; q < W; q += gridDim.x * blockDim.x) {
        wv.node[q] = fr.node[base + q];
        wv.skip[q] = fr.skip[base + q];
    }
    if (t == 0) {
        ctl->ftop = base;
        ctl->nitems = 0;
        ctl->npitems = 0;
        for (int k = 0; k < 8; ++k) ctl->cur[k] = 0ull;
    }
}

// ------------------------------------------------------------------------------------------
// k_lcss_ref: t*_ref (no phase 1: delta_ref is feasible at every vertex)
// ------------------------------------------------------------------------------------------
template <class SH>
__global__ void __launch_bounds__(32 * SH::WPC, SH::CPS)
k_lcss_ref(DevTables tb, NodeStore ns, WaveBuf wv, int W, Ctl *ctl, DevCounters *cnt)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int U = SH::U, NX = SH::NX, CAP = SH::CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ws = smem + (size_t)warp * WarpWs<U, NX, CAP>::kDoubles;
    unsigned tma_parity = warp_ws_init<U, NX, CAP>(ws, lane);
    for (;;) {
        const long long e = grab(&ctl->cur[0], lane);
        if (e >= W) break;
        const int nd = wv.node[e];
        const SolveOut o = solve_simplex<U, NX, CAP>(tb, ns.cand[nd], ns.verts + (size_t)nd * (NX + 1) * NX,
                                                    ns.vcost + (size_t)nd * (NX + 1), EHM_SX_TQP, false,
                                                    ws, nullptr, lane, tma_parity);
        account(cnt, o, lane, 1, ns.cand[nd], nd);
        if (lane == 0) wv.tref[e] = (o.status == 1) ? o.val : INFINITY;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// k_lcss_items: candidates not excluded by their inherited bound
// ------------------------------------------------------------------------------------------
__global__ void k_lcss_items(NodeStore ns, WaveBuf wv, int W, int ncand, Ctl *ctl)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = gw; e < W; e += nw) {
        const double tref = wv.tref[e];
        const double thr = tref <= 0.0 ? tref : 0.0;
        const int cref = ns.cand[wv.node[e]];
        for (int c0 = 0; c0 < ncand; c0 += 32) {
            const int c = c0 + lane;
            bool take = false;
            if (c < ncand) {
                const float lb = wv.lb[(size_t)e * ncand + c];
                take = (c != cref) && !(lb == INFINITY) && !((double)lb > thr);
                wv.tval[(size_t)e * ncand + c] = INFINITY;
                wv.titem[(size_t)e * ncand + c] = -1;
            }
            const unsigned m = __ballot_sync(kFull, take);
            if (m) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&ctl->nitems, __popc(m));
                pos = __shfl_sync(kFull, pos, 0);
                if (take) {
                    const int idx = pos + __popc(m & ((1u << lane) - 1u));
                    wv.items[idx] = make_int2(e, c);
                    wv.titem[(size_t)e * ncand + c] = idx;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_lcss_eval: alive test + epigraph QP per item
// ------------------------------------------------------------------------------------------
template <class SH>
__global__ void __launch_bounds__(32 * SH::WPC, SH::CPS)
k_lcss_eval(DevTables tb, NodeStore ns, WaveBuf wv, Ctl *ctl, DevCounters *cnt)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int U = SH::U, NX = SH::NX, CAP = SH::CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ws = smem + (size_t)warp * WarpWs<U, NX, CAP>::kDoubles;
    unsigned tma_parity = warp_ws_init<U, NX, CAP>(ws, lane);
    const int nitems = ctl->nitems;
    for (;;) {
        const long long it = grab(&ctl->cur[1], lane);
        if (it >= nitems) break;
        const int2 ec = wv.items[it];
        const int nd = wv.node[ec.x];
        const SolveOut o = solve_simplex<U, NX, CAP>(tb, ec.y, ns.verts + (size_t)nd * (NX + 1) * NX,
                                                    ns.vcost + (size_t)nd * (NX + 1), EHM_SX_TQP, true, ws,
                                                    wv.it_sol + (size_t)it * (NX + 1), lane, tma_parity,
                                                    wv.it_pl + (size_t)it * (NX + 1));
        account(cnt, o, lane, 2, ec.y, nd);
        if (lane == 0) {
            const size_t ix = (size_t)ec.x * tb.ncand + ec.y;
            if (o.status == 1) {
                wv.tval[ix] = o.val;
                wv.lb[ix] = f32_down(o.val);
            } else {
                // certified dead on this simplex: inherited by every descendant.  A numerical
                // failure (-1) is no certificate: the inherited bound stays and the descendants
                // evaluate the candidate again (oracle/ehm_oracle.c:ehmo_lcss_eval)
                if (o.status == 0) wv.lb[ix] = INFINITY;
                wv.it_pl[(size_t)it * (NX + 1)] = -INFINITY;  // no plane
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// k_lcss_select: per entry -- convergence, ranked improving candidates, split geometry and
// the list of follow-up problems.  One warp per entry.
// ------------------------------------------------------------------------------------------
template <int NX>
__global__ void k_lcss_select(NodeStore ns, WaveBuf wv, int W, int ncand, Ctl *ctl)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    constexpr int PPE = (kSelT + 1) * (NX + 2) + 1;  // max follow-up problems per entry
    for (int e = gw; e < W; e += nw) {
        const int nd = wv.node[e];
        const double tref = wv.tref[e];
        const int cref = ns.cand[nd];
        const double *tv = wv.tval + (size_t)e * ncand;
        // any improving candidate at all?  how many beat delta_ref?
        int nle0 = 0, nbet = 0;
        for (int c = lane; c < ncand; c += 32) {
            const double t = tv[c];
            if (t <= 0.0) {
                ++nle0;
                if (t < tref) ++nbet;
            }
        }
        nle0 = __reduce_add_sync(kFull, nle0);
        nbet = __reduce_add_sync(kFull, nbet);
        const bool conv = !(tref <= 0.0) && nle0 == 0;
        // ranked selection: skip the first `skip` candidates (already vetted on an earlier
        // visit of this node), then take up to kSelT in ascending (t, index) order.
        const int skip = wv.skip[e];
        double last_t = -INFINITY;
        int last_c = -1, nsel = 0;
        if (!conv && nbet > skip) {
            for (int rank = 0; rank < skip + kSelT && rank < nbet; ++rank) {
                double bt = INFINITY;
                int bc = 0x7fffffff;
                for (int c = lane; c < ncand; c += 32) {
                    const double t = tv[c];
                    if (t <= 0.0 && t < tref && (t > last_t || (t == last_t && c > last_c)) &&
                        (t < bt || (t == bt && c < bc))) {
                        bt = t;
                        bc = c;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ot = __shfl_xor_sync(kFull, bt, o);
                    const int oc = __shfl_xor_sync(kFull, bc, o);
                    if (ot < bt || (ot == bt && oc < bc)) {
                        bt = ot;
                        bc = oc;
                    }
                }
                last_t = bt;
                last_c = bc;
                if (rank >= skip) {
                    if (lane == 0) wv.sel[e * kSelT + nsel] = bc;
                    ++nsel;
                }
            }
        }
        if (lane == 0) {
            wv.conv[e] = conv ? 1 : 0;
            wv.nimp[e] = nbet;
            wv.nsel[e] = nsel;
            if (!conv) {
                const double *R = ns.verts + (size_t)nd * (NX + 1) * NX;
                int bi, bj;
                longest_edge<NX>(R, bi, bj);
                wv.ij[2 * e] = bi;
                wv.ij[2 * e + 1] = bj;
                for (int k = 0; k < NX; ++k)
                    wv.vmid[e * NX + k] = __ddiv_rn(__dadd_rn(R[bi * NX + k], R[bj * NX + k]), 2.0);
                // follow-up problems: vertices + midpoint for each vetted candidate, the
                // midpoint for delta_ref, and (if any candidate) min_R V*_{delta_ref}
                const int np = nsel * (NX + 2) + 1 + (nsel > 0 ? 1 : 0);
                int pos = atomicAdd(&ctl->npitems, np);
                for (int k = 0; k < nsel; ++k)
                    for (int i = 0; i <= NX + 1; ++i) wv.pitems[pos++] = PItem{e, k, i, wv.sel[e * kSelT + k]};
                wv.pitems[pos++] = PItem{e, kSelT, NX + 1, cref};
                if (nsel > 0) wv.pitems[pos++] = PItem{e, kSelT, -1, cref};
            }
        }
        (void)PPE;
    }
}

// ------------------------------------------------------------------------------------------
// k_lcss_points: vertex / midpoint P_theta_delta and min-over-simplex problems
// ------------------------------------------------------------------------------------------
template <class SH>
__global__ void __launch_bounds__(32 * SH::WPC, SH::CPS)
k_lcss_points(DevTables tb, NodeStore ns, WaveBuf wv, Ctl *ctl, DevCounters *cnt)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int U = SH::U, NX = SH::NX, CAP = SH::CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ws = smem + (size_t)warp * WarpWs<U, NX, CAP>::kDoubles;
    unsigned tma_parity = warp_ws_init<U, NX, CAP>(ws, lane);
    const int np = ctl->npitems;
    for (;;) {
        const long long q = grab(&ctl->cur[2], lane);
        if (q >= np) break;
        const PItem pi = wv.pitems[q];
        const int nd = wv.node[pi.e];
        SolveOut o;
        if (pi.i < 0) {
            o = solve_simplex<U, NX, CAP>(tb, pi.c, ns.verts + (size_t)nd * (NX + 1) * NX, nullptr,
                                         EHM_SX_MINV, false, ws, nullptr, lane, tma_parity);
            if (lane == 0) {
                wv.minv[pi.e] = o.val;
                wv.minv_st[pi.e] = o.status;
            }
        } else {
            const double *th = (pi.i <= NX) ? ns.verts + ((size_t)nd * (NX + 1) + pi.i) * NX
                                            : wv.vmid + (size_t)pi.e * NX;
            const size_t slot = ((size_t)pi.e * (kSelT + 1) + pi.k) * (NX + 2) + pi.i;
            o = solve_point<U, NX, CAP>(tb, pi.c, th, ws, wv.ptU + slot * tb.nu, lane);
            if (lane == 0) {
                wv.ptJ[slot] = o.val;
                wv.ptS[slot] = o.status;
            }
        }
        account(cnt, o, lane, 3, pi.c, nd);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// k_lcss_final: decide and act.  One warp per entry (the bound rows are copied by the warp).
// ------------------------------------------------------------------------------------------
// Bound of t*_c from the tangent plane P (vertex values) under vertex costs V (ehm_solve.cuh).
template <int NX>
__device__ __forceinline__ float plane_bound(const float (&P)[NX + 1], const double (&V)[NX + 1], double eps_a,
                                             double k2)
{
    double a = INFINITY, b = INFINITY;
#pragma unroll
    for (int i = 0; i <= NX; ++i) {
        a = fmin(a, (double)P[i] - V[i]);
        b = fmin(b, (double)P[i] - k2 * V[i]);
    }
    return __double2float_rd(fmax(a + eps_a, b));
}

template <int NX>
__global__ void k_lcss_final(DevTables tb, NodeStore ns, Frontier fr, WaveBuf wv, int W, Ctl *ctl)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int ncand = tb.ncand, nu = tb.nu;
    const double k2 = 1.0 / (1.0 + tb.eps_r);
    for (int e = gw; e < W; e += nw) {
        const int nd = wv.node[e];
        if (wv.conv[e]) {  // worker.py:368-375 close leaf
            if (lane == 0) {
                ns.closed[nd] = 1;
                atomicAdd(&ctl->nleaves, 1);
            }
            continue;
        }
        const int cref = ns.cand[nd];
        const int nsel = wv.nsel[e];
        // first vetted candidate feasible (and convergent) at all n+1 vertices
        int kstar = -1;
        for (int k = 0; k < nsel && kstar < 0; ++k) {
            bool ok = true;
            for (int i = 0; i <= NX; ++i)
                ok = ok && wv.ptS[((size_t)e * (kSelT + 1) + k) * (NX + 2) + i] == 1;
            if (ok) kstar = k;
        }
        if (kstar < 0 && nsel == kSelT && wv.nimp[e] > wv.skip[e] + kSelT) {
            // more improving candidates than vetted this visit: revisit with a larger rank
            // offset (same decisions, the evaluated t* are deterministic)
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&ctl->ftop, 1);
            pos = __shfl_sync(kFull, pos, 0);
            if (pos >= fr.cap) {
                if (lane == 0) atomicExch(&ctl->error, EHM_ECAPACITY);
                continue;
            }
            for (int c = lane; c < ncand; c += 32) fr.lb[(size_t)pos * ncand + c] = wv.lb[(size_t)e * ncand + c];
            if (lane == 0) {
                fr.node[pos] = nd;
                fr.skip[pos] = wv.skip[e] + kSelT;
            }
            continue;
        }
        int cstar = -1;
        bool var_small = false;
        double Vr[NX + 1];
        for (int i = 0; i <= NX; ++i) Vr[i] = ns.vcost[(size_t)nd * (NX + 1) + i];
        double rebase = 0.0;
        // delta* is adopted only if its follow-up problems went through: min_R V*_ref (in_variability_ball)
        // and its own midpoint problem.  A numerical failure of either (counted in nfail by the solving
        // kernel) falls back to delta_ref for this visit -- the reference's "blacklist delta* and retry"
        // (lib/oracle.py:398-414); EHM_ENUMERIC is left for a failing midpoint of delta_ref itself.
        if (kstar >= 0 && wv.minv_st[e] == 1 &&
            wv.ptS[((size_t)e * (kSelT + 1) + kstar) * (NX + 2) + NX + 1] == 1) {
            cstar = wv.sel[e * kSelT + kstar];
            const int item = wv.titem[(size_t)e * ncand + cstar];
            const double Vstar = wv.it_sol[(size_t)item * (NX + 1) + NX];
            double maxV = Vr[0];
            for (int i = 1; i <= NX; ++i) maxV = fmax(maxV, Vr[i]);
            var_small = (maxV - wv.minv[e]) < fmax(tb.eps_a, tb.eps_r * Vstar);  // oracle.py:282-283
            for (int i = 0; i <= NX; ++i)
                rebase = fmax(rebase, wv.ptJ[((size_t)e * (kSelT + 1) + kstar) * (NX + 2) + i] - Vr[i]);
            if (rebase > 0.0) rebase += 1e-9 * (1.0 + rebase);
        }
        const int kslot = (cstar >= 0) ? kstar : kSelT;
        const int cnew = (cstar >= 0) ? cstar : cref;
        const float dreb = (rebase > 0.0) ? __double2float_ru(rebase) : 0.0f;
        const size_t pslot = ((size_t)e * (kSelT + 1) + kslot) * (NX + 2);
        if (cstar >= 0 && var_small) {
            // worker.py:396-401: adopt delta*, keep the simplex, run L-CSS on it again
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&ctl->ftop, 1);
            pos = __shfl_sync(kFull, pos, 0);
            if (pos >= fr.cap) {
                if (lane == 0) atomicExch(&ctl->error, EHM_ECAPACITY);
                continue;
            }
            double Vn[NX + 1];
            for (int i = 0; i <= NX; ++i) Vn[i] = wv.ptJ[pslot + i];
            for (int c = lane; c < ncand; c += 32) {
                float v = wv.lb[(size_t)e * ncand + c];
                if (c == cref) v = -INFINITY;
                else if (v != INFINITY && v != -INFINITY && dreb > 0.0f) v = __fadd_rd(v, -dreb);
                // this visit's tangent plane bounds t*_c under the NEW vertex costs as well
                const int item = wv.titem[(size_t)e * ncand + c];
                if (item >= 0 && v != INFINITY) {
                    float P[NX + 1];
                    for (int i = 0; i <= NX; ++i) P[i] = wv.it_pl[(size_t)item * (NX + 1) + i];
                    if (P[0] > -INFINITY) v = fmaxf(v, plane_bound<NX>(P, Vn, tb.eps_a, k2));
                }
                fr.lb[(size_t)pos * ncand + c] = v;
            }
            if (lane == 0) {
                fr.node[pos] = nd;
                fr.skip[pos] = 0;
                ns.cand[nd] = cnew;
                for (int i = 0; i <= NX; ++i) {
                    ns.vcost[(size_t)nd * (NX + 1) + i] = wv.ptJ[pslot + i];
                    for (int q = 0; q < nu; ++q)
                        ns.vinput[((size_t)nd * (NX + 1) + i) * nu + q] = wv.ptU[(pslot + i) * nu + q];
                }
            }
            continue;
        }
        // worker.py:403-417: split along the longest edge; only the midpoint is re-solved
        if (wv.ptS[pslot + NX + 1] != 1) {
            if (lane == 0) atomicExch(&ctl->error, EHM_ENUMERIC);
            continue;
        }
        int id0 = 0, pos = 0;
        if (lane == 0) {
            id0 = atomicAdd(&ctl->nnodes, 2);
            pos = atomicAdd(&ctl->ftop, 2);
        }
        id0 = __shfl_sync(kFull, id0, 0);
        pos = __shfl_sync(kFull, pos, 0);
        if (id0 + 2 > ns.cap || pos + 2 > fr.cap) {
            if (lane == 0) atomicExch(&ctl->error, EHM_ECAPACITY);
            continue;
        }
        // vertex costs of the two children (side 0 replaces vertex bi, side 1 vertex bj by the midpoint)
        const int sbi = wv.ij[2 * e], sbj = wv.ij[2 * e + 1];
        double Vc0[NX + 1], Vc1[NX + 1];
        for (int i = 0; i <= NX; ++i) {
            const double base = (cstar >= 0) ? wv.ptJ[pslot + i] : Vr[i];
            Vc0[i] = (i == sbi) ? wv.ptJ[pslot + NX + 1] : base;
            Vc1[i] = (i == sbj) ? wv.ptJ[pslot + NX + 1] : base;
        }
        for (int c = lane; c < ncand; c += 32) {
            float v = wv.lb[(size_t)e * ncand + c];
            if (cstar >= 0) {
                if (c == cref) v = -INFINITY;
                else if (v != INFINITY && v != -INFINITY && dreb > 0.0f) v = __fadd_rd(v, -dreb);
            }
            float v0 = v, v1 = v;
            // the tangent plane of this visit follows the geometry (the midpoint value is the mean
            // of the two end points) and bounds t*_c under each child's own vertex costs
            const int item = wv.titem[(size_t)e * ncand + c];
            if (item >= 0 && v != INFINITY) {
                float P[NX + 1];
                for (int i = 0; i <= NX; ++i) P[i] = wv.it_pl[(size_t)item * (NX + 1) + i];
                if (P[0] > -INFINITY) {
                    const float pm = __fmul_rd(0.5f, __fadd_rd(P[sbi], P[sbj])), pbi = P[sbi], pbj = P[sbj];
                    P[sbi] = pm;
                    v0 = fmaxf(v0, plane_bound<NX>(P, Vc0, tb.eps_a, k2));
                    P[sbi] = pbi;
                    P[sbj] = pm;
                    v1 = fmaxf(v1, plane_bound<NX>(P, Vc1, tb.eps_a, k2));
                    P[sbj] = pbj;
                }
            }
            fr.lb[(size_t)pos * ncand + c] = v0;
            fr.lb[(size_t)(pos + 1) * ncand + c] = v1;
        }
        {
            // children: vertices, vertex costs and vertex inputs are copied by the whole warp
            // (round 1: lane 0 alone -- 2 x (n+1)(n + 1 + nu) doubles, 450 for the synthetic shape)
            const int bi = sbi, bj = sbj;
            const double *R = ns.verts + (size_t)nd * (NX + 1) * NX;
            const int dep = ns.depth[nd] + 1;
            if (lane == 0) {
                atomicMax(&ctl->max_depth, dep);
                ns.left[nd] = id0;
                ns.right[nd] = id0 + 1;
            }
            for (int side = 0; side < 2; ++side) {
                const int ch = id0 + side, rep = side ? bj : bi;
                double *Rc = ns.verts + (size_t)ch * (NX + 1) * NX;
                for (int q = lane; q < (NX + 1) * NX; q += 32) {
                    const int row = q / NX, k = q - row * NX;
                    Rc[q] = row == rep ? wv.vmid[e * NX + k] : R[q];
                }
                for (int i = lane; i <= NX; i += 32) {
                    const bool mid = (i == rep);
                    const size_t src = mid ? pslot + NX + 1 : pslot + i;
                    // costs: delta* values when switching, the node's own otherwise
                    ns.vcost[(size_t)ch * (NX + 1) + i] = (cstar >= 0 || mid) ? wv.ptJ[src] : ns.vcost[(size_t)nd * (NX + 1) + i];
                }
                for (int q = lane; q < (NX + 1) * nu; q += 32) {
                    const int i = q / nu, qq = q - i * nu;
                    const bool mid = (i == rep);
                    const size_t src = mid ? pslot + NX + 1 : pslot + i;
                    ns.vinput[((size_t)ch * (NX + 1) + i) * nu + qq] =
                        (cstar >= 0 || mid) ? wv.ptU[src * nu + qq]
                                            : ns.vinput[((size_t)nd * (NX + 1) + i) * nu + qq];
                }
                if (lane == 0) {
                    ns.cand[ch] = cnew;
                    ns.left[ch] = -1;
                    ns.right[ch] = -1;
                    ns.root_id[ch] = ns.root_id[nd];
                    ns.depth[ch] = dep;
                    ns.closed[ch] = 0;
                    ns.origin[2 * ch] = -1;
                    ns.origin[2 * ch + 1] = -1;
                    fr.node[pos + side] = ch;
                    fr.skip[pos + side] = 0;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// ECC (lib/worker.py:241-291): V_R over all candidates, canonical argmin of summed costs
// ------------------------------------------------------------------------------------------
template <class SH>
__global__ void __launch_bounds__(32 * SH::WPC, SH::CPS)
k_ecc_points(DevTables tb, NodeStore ns, EccBuf eb, int q0, int ne, int centroid, Ctl *ctl, DevCounters *cnt,
             const double *__restrict__ theta = nullptr)
{
    // centroid: 0 = the n+1 vertices of every entry (V_R), 1 = the barycentre of the entries without
    // a common commutation (worker.py:264-266), 2 = the explicit point theta[e] (P_theta batch)
    extern __shared__ __align__(16) double smem[];
    constexpr int U = SH::U, NX = SH::NX, CAP = SH::CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ws = smem + (size_t)warp * WarpWs<U, NX, CAP>::kDoubles;
    unsigned tma_parity = warp_ws_init<U, NX, CAP>(ws, lane);
    const int ncand = tb.ncand;
    const long long per = centroid ? ncand : (long long)ncand * (NX + 1);
    const long long total = per * ne;
    for (;;) {
        const long long q = grab(&ctl->cur[centroid == 1 ? 4 : 3], lane);
        if (q >= total) break;
        const int e = (int)(q / per);
        const long long r = q % per;
        const int c = centroid ? (int)r : (int)(r / (NX + 1));
        const int i = centroid == 2 ? 0 : (centroid ? NX + 1 : (int)(r % (NX + 1)));
        const int nd = centroid == 2 ? 0 : eb.queue[q0 + e];
        if (centroid == 1 && eb.pick[e] != -1) continue;  // only entries without a common delta
        double th[NX];
        const double *R = ns.verts + (size_t)nd * (NX + 1) * NX;
        if (centroid == 2) {
            for (int k = 0; k < NX; ++k) th[k] = theta[(size_t)e * NX + k];
        } else if (centroid) {
            for (int k = 0; k < NX; ++k) {
                double acc = 0.0;
                for (int v = 0; v <= NX; ++v) acc += R[v * NX + k];
                th[k] = acc / (NX + 1);
            }
        } else {
            for (int k = 0; k < NX; ++k) th[k] = R[i * NX + k];
        }
        // theta must live in global/shared memory for solve_point: stage it in the warp's
        // reduction scratch
        double *thp = ws + WarpWs<U, NX, CAP>::oW2 + NX;
        if (lane == 0)
            for (int k = 0; k < NX; ++k) thp[k] = th[k];
        __syncwarp();
        const size_t slot = ((size_t)e * ncand + c) * (NX + 2) + i;
        double *up = (i <= NX) ? eb.U + (((size_t)e * ncand + c) * (NX + 1) + i) * tb.nu : nullptr;
        const SolveOut o = solve_point<U, NX, CAP>(tb, c, thp, ws, up, lane);
        account(cnt, o, lane, 4);
        if (lane == 0) {
            eb.J[slot] = o.val;
            eb.S[slot] = o.status;
        }
        __syncwarp();
    }
}

// per ECC entry: canonical V_R choice (one warp per entry)
template <int NX>
__global__ void k_ecc_pick(DevTables tb, EccBuf eb, int ne, Ctl *ctl, int nslots = NX + 1)
{
    // nslots = NX+1: V_R over the vertices; 1: P_theta (only slot 0 is filled)
    const int lane = threadIdx.x & 31;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= ne) return;
    const int ncand = tb.ncand;
    double bt = INFINITY;
    int bc = 0x7fffffff;
    for (int c = lane; c < ncand; c += 32) {
        bool ok = true;
        double sum = 0.0;
        for (int i = 0; i < nslots; ++i) {
            const size_t slot = ((size_t)e * ncand + c) * (NX + 2) + i;
            ok = ok && eb.S[slot] == 1;
            sum += eb.J[slot];
        }
        if (ok && (sum < bt || (sum == bt && c < bc))) {
            bt = sum;
            bc = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ot = __shfl_xor_sync(kFull, bt, o);
        const int oc = __shfl_xor_sync(kFull, bc, o);
        if (ot < bt || (ot == bt && oc < bc)) {
            bt = ot;
            bc = oc;
        }
    }
    if (lane == 0) {
        eb.pick[e] = (bc == 0x7fffffff) ? -1 : bc;
        if (bc == 0x7fffffff) atomicAdd(&ctl->need_centroid, 1);
    }
}

template <int NX>
__global__ void k_ecc_final(DevTables tb, NodeStore ns, Frontier fr, EccBuf eb, int q0, int ne, Ctl *ctl)
{
    const int lane = threadIdx.x & 31;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= ne) return;
    const int ncand = tb.ncand, nu = tb.nu;
    const int nd = eb.queue[q0 + e];
    const int c = eb.pick[e];
    if (c >= 0) {
        // worker.py:279-291: assign the commutation, hand the node to L-CSS
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&ctl->ftop, 1);
        pos = __shfl_sync(kFull, pos, 0);
        if (pos >= fr.cap) {
            if (lane == 0) atomicExch(&ctl->error, EHM_ECAPACITY);
            return;
        }
        for (int k = lane; k < ncand; k += 32) fr.lb[(size_t)pos * ncand + k] = -INFINITY;
        if (lane == 0) {
            fr.node[pos] = nd;
            fr.skip[pos] = 0;
            ns.cand[nd] = c;
            for (int i = 0; i <= NX; ++i) {
                ns.vcost[(size_t)nd * (NX + 1) + i] = eb.J[((size_t)e * ncand + c) * (NX + 2) + i];
                for (int q = 0; q < nu; ++q)
                    ns.vinput[((size_t)nd * (NX + 1) + i) * nu + q] =
                        eb.U[(((size_t)e * ncand + c) * (NX + 1) + i) * nu + q];
            }
        }
        return;
    }
    // no common commutation: the barycentre must still be feasible (worker.py:264-266)
    int any = 0;
    for (int k = lane; k < ncand; k += 32)
        any |= eb.S[((size_t)e * ncand + k) * (NX + 2) + NX + 1] == 1;
    any = __any_sync(kFull, any);
    if (lane != 0) return;
    if (!any) {
        atomicExch(&ctl->error, EHM_EINFEASIBLE);
        return;
    }
    const int id0 = atomicAdd(&ctl->nnodes, 2);
    const int qp = atomicAdd(&ctl->necc_next, 2);
    if (id0 + 2 > ns.cap) {
        atomicExch(&ctl->error, EHM_ECAPACITY);
        return;
    }
    const double *R = ns.verts + (size_t)nd * (NX + 1) * NX;
    int bi, bj;
    longest_edge<NX>(R, bi, bj);
    const int dep = ns.depth[nd] + 1;
    atomicMax(&ctl->max_depth, dep);
    ns.left[nd] = id0;
    ns.right[nd] = id0 + 1;
    for (int side = 0; side < 2; ++side) {
        const int ch = id0 + side, rep = side ? bj : bi;
        double *Rc = ns.verts + (size_t)ch * (NX + 1) * NX;
        for (int q = 0; q < (NX + 1) * NX; ++q) Rc[q] = R[q];
        for (int k = 0; k < NX; ++k)
            Rc[rep * NX + k] = __ddiv_rn(__dadd_rn(R[bi * NX + k], R[bj * NX + k]), 2.0);
        for (int i = 0; i <= NX; ++i) {
            ns.vcost[(size_t)ch * (NX + 1) + i] = 0.0;
            for (int q = 0; q < nu; ++q) ns.vinput[((size_t)ch * (NX + 1) + i) * nu + q] = 0.0;
        }
        ns.cand[ch] = -1;
        ns.left[ch] = -1;
        ns.right[ch] = -1;
        ns.root_id[ch] = ns.root_id[nd];
        ns.depth[ch] = dep;
        ns.closed[ch] = 0;
        ns.origin[2 * ch] = -1;
        ns.origin[2 * ch + 1] = -1;
        eb.queue[qp + side] = ch;
    }
}

__global__ void k_reset_cursors(Ctl *ctl)
{
    if (threadIdx.x < 8) ctl->cur[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) ctl->need_centroid = 0;
}

}  // namespace ehm
