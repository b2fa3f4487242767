// ehm_lib.cu -- problem object, candidate condensation and the batched oracle entry points
// of libehm_b200.so (C-ABI in include/ehm.h).  sm_100a only; no CPU fallback: every entry
// point fails with EHM_ECUDA when no CUDA device is usable.
#include <chrono>
#include <cstdio>
#include <cstring>

#include "ehm_common.cuh"

namespace ehm {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
}  // namespace ehm

using namespace ehm;

// ---------------------------------------------------------------------------------------
// compiled problem shapes: U = N*n_u inputs, NX = n, CAP = row capacity of the compacted tile
// (stage rows that survive the pre-solve + n+1 simplex rows + two epigraph rows), WPC warps
// (= problems) per CTA, CPS CTAs per SM (shared memory per problem: WarpWs<U,NX,CAP>)
// ---------------------------------------------------------------------------------------
#ifndef EHM_CPS_PEND
#define EHM_CPS_PEND 4
#endif
#define EHM_FOR_EACH_CFG(X)                                                                      \
    X(0, 5, 4, 80, 4, EHM_CPS_PEND)  /* inverted pendulum, N=5: 217 rows in, 41-68 after the pre-solve   */ \
    X(1, 2, 2, 32, 4, 4)  /* two-mode toy problem of the ECC-split test (tests/golden)        */ \
    X(2, 16, 8, 92, 1, 6) /* synthetic n=8 mp-LP: 91 rows, NV = 25; LP costs only (IpmHasH) */

template <int U_, int NX_, int CAP_, int WPC_, int CPS_>
struct Shape {
    static constexpr int U = U_, NX = NX_, CAP = CAP_, WPC = WPC_, CPS = CPS_;
    static constexpr size_t smem = (size_t)WPC_ * WarpWs<U_, NX_, CAP_>::kDoubles * sizeof(double);
};

static int find_cfg(int U, int NX)
{
#define X(ID, UU, NN, CC, WW, PP) \
    if (U == UU && NX == NN) return ID;
    EHM_FOR_EACH_CFG(X)
#undef X
    return -1;
}

static void ehm_partition_free_now(void *t);   // defined with ehm_partition below
static void ehm_oracle_scratch_free(void *s); // defined with the batched oracle calls below

__device__ __forceinline__ long long next_item(DevCounters *cnt, int lane)
{
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(&cnt->work, 1ull);
    return (long long)__shfl_sync(kFull, v, 0);
}

__device__ __forceinline__ void account(DevCounters *cnt, const SolveOut &o, int lane, int stage = 0,
                                        int cand = -1, int node = -1)
{
    if (lane == 0) {
        if (o.status < 0) {
            cnt->last_fail[0] = stage; cnt->last_fail[1] = cand; cnt->last_fail[2] = node;
            cnt->last_fail[3] = o.iters;
        }
        atomicAdd(&cnt->flops[stage], o.flops);
        atomicAdd(&cnt->iters, (unsigned long long)o.iters);
        atomicAdd(&cnt->solves, (unsigned long long)o.nsolves);
        if (o.status < 0) atomicAdd(&cnt->fails, 1ull);
    }
}

template <class SH>
__global__ void __launch_bounds__(32 * SH::WPC, SH::CPS)
k_point_batch(DevTables tb, int B, const double *__restrict__ theta, const int *__restrict__ cand,
              double *J, double *u0, int *status, DevCounters *cnt)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int U = SH::U, NX = SH::NX, CAP = SH::CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ws = smem + (size_t)warp * WarpWs<U, NX, CAP>::kDoubles;
    unsigned tma_parity = warp_ws_init<U, NX, CAP>(ws, lane);
    for (;;) {
        const long long b = next_item(cnt, lane);
        if (b >= B) break;
        const int c = cand[b];
        SolveOut o{0, 0.0, 0, 0, 0.0};
        if (c >= 0 && c < tb.ncand)
            o = solve_point<U, NX, CAP>(tb, c, theta + (size_t)b * NX, ws, u0 ? u0 + (size_t)b * tb.nu : nullptr, lane);
        account(cnt, o, lane);
        if (lane == 0) {
            status[b] = o.status;
            J[b] = o.val;
        }
        __syncwarp();
    }
}

template <class SH>
__global__ void __launch_bounds__(32 * SH::WPC, SH::CPS)
k_simplex_batch(DevTables tb, int B, const double *__restrict__ R, const double *__restrict__ Vr,
                const int *__restrict__ cand, int mode, int with_phase1, double *val, double *sol,
                int *status, DevCounters *cnt)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int U = SH::U, NX = SH::NX, CAP = SH::CAP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ws = smem + (size_t)warp * WarpWs<U, NX, CAP>::kDoubles;
    unsigned tma_parity = warp_ws_init<U, NX, CAP>(ws, lane);
    for (;;) {
        const long long b = next_item(cnt, lane);
        if (b >= B) break;
        const int c = cand[b];
        SolveOut o{0, 0.0, 0, 0, 0.0};
        if (c >= 0 && c < tb.ncand)
            o = solve_simplex<U, NX, CAP>(tb, c, R + (size_t)b * (NX + 1) * NX,
                                         Vr ? Vr + (size_t)b * (NX + 1) : nullptr, mode,
                                         with_phase1 != 0, ws,
                                         sol ? sol + (size_t)b * (NX + 1) : nullptr, lane, tma_parity);
        account(cnt, o, lane);
        if (lane == 0) {
            status[b] = o.status;
            val[b] = o.val;
        }
        __syncwarp();
    }
}

// Longest-edge bisection (lib/tools.py:224-257): one warp per simplex, coalesced.  The (n+1)n
// coordinates are read once, contiguously, into shared memory; lane q computes the length of edge q
// (itertools.combinations order, as the reference enumerates them) -- the square root of a
// left-to-right chain of fused multiply-adds, the arithmetic of the reference's la.norm(R[i]-R[j])
// (numpy: sqrt(x.dot(x)), BLAS dot of a short vector), so that ties between edges of equal length
// break exactly as in the reference; the first maximum wins (smallest edge index among the equal
// maxima); both children are written back contiguously.  HBM-bound: 3 x 8 (n+1) n + 8 bytes per simplex.
constexpr int kSplitWarps = 8;
__global__ void __launch_bounds__(32 * kSplitWarps)
k_split_batch(int n, int B, const double *__restrict__ R, double *S1, double *S2, int *ij)
{
    __shared__ double sh[kSplitWarps][EHM_SPLIT_MAXN * (EHM_SPLIT_MAXN + 1)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int len = (n + 1) * n, nedges = (n + 1) * n / 2;
    double *v = sh[warp];
    for (long long b = (long long)blockIdx.x * kSplitWarps + warp; b < B; b += (long long)gridDim.x * kSplitWarps) {
        const double *Rb = R + (size_t)b * len;
        for (int e = lane; e < len; e += 32) v[e] = Rb[e];
        __syncwarp();
        double best = -1.0;
        int bq = 0x7fffffff, bi = 0, bj = 1;
        for (int q = lane; q < nedges; q += 32) {
            // edge q -> (i, j), i < j, in combinations order
            int i = 0, rem = q;
            while (rem >= n - i) { rem -= n - i; ++i; }
            const int j = i + 1 + rem;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) {
                const double d = __dsub_rn(v[i * n + k], v[j * n + k]);
                acc = __fma_rn(d, d, acc);
            }
            const double l = __dsqrt_rn(acc);
            if (l > best) { best = l; bq = q; bi = i; bj = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ol = __shfl_xor_sync(0xffffffffu, best, o);
            const int oq = __shfl_xor_sync(0xffffffffu, bq, o), oi = __shfl_xor_sync(0xffffffffu, bi, o),
                      oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ol > best || (ol == best && oq < bq)) { best = ol; bq = oq; bi = oi; bj = oj; }
        }
        double *A = S1 + (size_t)b * len, *Bq = S2 + (size_t)b * len;
        for (int e = lane; e < len; e += 32) {
            const int row = e / n, k = e - row * n;
            const double mid = __ddiv_rn(__dadd_rn(v[bi * n + k], v[bj * n + k]), 2.0);
            A[e] = row == bi ? mid : v[e];
            Bq[e] = row == bj ? mid : v[e];
        }
        if (lane == 0) {
            ij[2 * b] = bi;
            ij[2 * b + 1] = bj;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// host side: condensation of every commutation candidate (mode sequence)
// ---------------------------------------------------------------------------------------
namespace {

// x_k = T [u; theta; 1] with T an n x Y matrix; rows/cost follow from the stage data.
struct Condenser {
    const ehm_problem_desc &d;
    int n, nu, N, K, p, U, m, MP, Y;
    explicit Condenser(const ehm_problem_desc &desc) : d(desc)
    {
        n = d.n; nu = d.nu; N = d.N; K = d.K; p = d.p;
        U = N * nu; m = N * p; MP = ((m + n + 3 + 31) / 32) * 32; Y = U + n + 1;
    }
    void run(int c, double *G, double *S, double *g, double *rho, double *M) const
    {
        std::vector<int> seq(N);
        for (int k = N - 1; k >= 0; --k) { seq[k] = c % K; c /= K; }
        std::vector<double> T((size_t)n * Y, 0.0), Tn((size_t)n * Y), WT((size_t)n * Y);
        for (int i = 0; i < n; ++i) T[(size_t)i * Y + U + i] = 1.0;
        std::fill(M, M + (size_t)Y * Y, 0.0);
        for (int col = 0; col < U; ++col) std::fill(G + (size_t)col * MP, G + (size_t)(col + 1) * MP, 0.0);
        for (int col = 0; col < n; ++col) std::fill(S + (size_t)col * MP, S + (size_t)(col + 1) * MP, 0.0);
        std::fill(g, g + MP, 1.0);
        std::fill(rho, rho + MP, 0.0);
        for (int k = 0; k < N; ++k) {
            const int i = seq[k];
            const double *Ci = d.C + (size_t)i * p * n, *Di = d.D + (size_t)i * p * nu,
                         *ei = d.e + (size_t)i * p, *Ai = d.A + (size_t)i * n * n,
                         *Bi = d.B + (size_t)i * n * nu, *wi = d.w + (size_t)i * n;
            for (int r = 0; r < p; ++r) {
                const int row = k * p + r;
                double nat = 0.0;
                for (int col = 0; col < U; ++col) {
                    double v = 0.0;
                    for (int a = 0; a < n; ++a) v += Ci[r * n + a] * T[(size_t)a * Y + col];
                    if (col / nu == k) v += Di[r * nu + col % nu];
                    G[(size_t)col * MP + row] = v;
                    nat = std::max(nat, std::fabs(v));
                }
                for (int col = 0; col < n; ++col) {
                    double v = 0.0;
                    for (int a = 0; a < n; ++a) v += Ci[r * n + a] * T[(size_t)a * Y + U + col];
                    S[(size_t)col * MP + row] = -v;
                    nat = std::max(nat, std::fabs(v) * d.theta_scale[col]);
                }
                double gc = ei[r];
                for (int a = 0; a < n; ++a) gc -= Ci[r * n + a] * T[(size_t)a * Y + Y - 1];
                g[row] = gc;
                rho[row] = nat;
            }
            for (int a = 0; a < nu; ++a) {
                for (int b = 0; b < nu; ++b) M[(size_t)(k * nu + a) * Y + k * nu + b] += d.R[a * nu + b];
                M[(size_t)(k * nu + a) * Y + Y - 1] += 0.5 * d.ru[i * nu + a];
                M[(size_t)(Y - 1) * Y + k * nu + a] += 0.5 * d.ru[i * nu + a];
            }
            for (int a = 0; a < n; ++a)
                for (int col = 0; col < Y; ++col) {
                    double v = 0.0;
                    for (int b = 0; b < n; ++b) v += Ai[a * n + b] * T[(size_t)b * Y + col];
                    if (col < U && col / nu == k) v += Bi[a * nu + col % nu];
                    if (col == Y - 1) v += wi[a];
                    Tn[(size_t)a * Y + col] = v;
                }
            T.swap(Tn);
            const double *W = (k + 1 < N) ? d.Q : d.P;  // mpc_library.py:515-517
            for (int a = 0; a < n; ++a)
                for (int col = 0; col < Y; ++col) {
                    double v = 0.0;
                    for (int b = 0; b < n; ++b) v += W[a * n + b] * T[(size_t)b * Y + col];
                    WT[(size_t)a * Y + col] = v;
                }
            for (int r2 = 0; r2 < Y; ++r2)
                for (int col = 0; col < Y; ++col) {
                    double v = 0.0;
                    for (int a = 0; a < n; ++a) v += T[(size_t)a * Y + r2] * WT[(size_t)a * Y + col];
                    M[(size_t)r2 * Y + col] += v;
                }
        }
    }
};

template <typename KernelT>
int set_smem(KernelT kernel, size_t bytes)
{
    EHM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return EHM_OK;
}

}  // namespace

extern "C" const char *ehm_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t ehm_kernel_launches(void) { return g_launches.load(); }

extern "C" int ehm_problem_create(const ehm_problem_desc *desc, ehm_problem **out)
{
    if (!desc || !out) return fail(EHM_EINVAL, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(EHM_ECUDA, "no CUDA device: libehm_b200 has no CPU path");
    Condenser cd(*desc);
    const int cfg = find_cfg(cd.U, cd.n);
    if (cfg < 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "problem shape U=%d n=%d has no compiled kernel (EHM_FOR_EACH_CFG)", cd.U,
                 cd.n);
        return fail(EHM_EINVAL, buf);
    }
    {
        bool lp = true;
        for (int i = 0; i < cd.n * cd.n && lp; ++i) lp = desc->Q[i] == 0.0 && desc->P[i] == 0.0;
        for (int i = 0; i < cd.nu * cd.nu && lp; ++i) lp = desc->R[i] == 0.0;
        bool lp_only = false;
#define X(ID, UU, NN, CC, WW, PP) \
    if (cfg == ID) lp_only = !IpmWs<UU + NN + 1, CC>::kHasH;
        EHM_FOR_EACH_CFG(X)
#undef X
        if (lp_only && !lp) return fail(EHM_EINVAL, "this problem shape is compiled for LP costs only (Q = P = R = 0)");
    }
    double nc = 1;
    for (int k = 0; k < cd.N; ++k) nc *= cd.K;
    if (nc > 2e6) return fail(EHM_EINVAL, "too many commutation candidates");
    const int ncand = (int)nc;
    ehm_problem *pb = new ehm_problem();
    pb->cfg = cfg;
    DevTables &tb = pb->tb;
    tb.n = cd.n; tb.nu = cd.nu; tb.N = cd.N; tb.K = cd.K; tb.p = cd.p;
    tb.U = cd.U; tb.m = cd.m; tb.MP = cd.MP; tb.Y = cd.Y; tb.ncand = ncand;
    tb.eps_a = desc->eps_a; tb.eps_r = desc->eps_r;
    {
        bool lp = true;
        for (int i = 0; i < cd.n * cd.n && lp; ++i) lp = desc->Q[i] == 0.0 && desc->P[i] == 0.0;
        for (int i = 0; i < cd.nu * cd.nu && lp; ++i) lp = desc->R[i] == 0.0;
        tb.is_lp = lp ? 1 : 0;
    }
    const size_t szG = (size_t)ncand * cd.U * cd.MP, szS = (size_t)ncand * cd.n * cd.MP,
                 szg = (size_t)ncand * cd.MP, szM = (size_t)ncand * cd.Y * cd.Y;
    std::vector<double> hG(szG), hS(szS), hg(szg), hr(szg), hM(szM);
    for (int c = 0; c < ncand; ++c)
        cd.run(c, hG.data() + (size_t)c * cd.U * cd.MP, hS.data() + (size_t)c * cd.n * cd.MP,
               hg.data() + (size_t)c * cd.MP, hr.data() + (size_t)c * cd.MP,
               hM.data() + (size_t)c * cd.Y * cd.Y);
    cudaError_t e = cudaSuccess;
    auto up = [&](double **dst, const std::vector<double> &src) {
        if (e != cudaSuccess) return;
        e = cudaMalloc((void **)dst, src.size() * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice);
    };
    up(&pb->dG, hG); up(&pb->dS, hS); up(&pb->dg, hg); up(&pb->drho, hr); up(&pb->dM, hM);
    if (e == cudaSuccess) e = cudaMalloc((void **)&pb->dcnt, sizeof(DevCounters));
    if (e == cudaSuccess) e = cudaMemset(pb->dcnt, 0, sizeof(DevCounters));
    if (e != cudaSuccess) {
        ehm_problem_destroy(pb);
        return fail(EHM_ECUDA, std::string("problem upload: ") + cudaGetErrorString(e));
    }
    tb.G = pb->dG; tb.S = pb->dS; tb.g = pb->dg; tb.rho = pb->drho; tb.M = pb->dM;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&pb->sm_count, cudaDevAttrMultiProcessorCount, dev);
    *out = pb;
    return EHM_OK;
}

extern "C" int ehm_problem_destroy(ehm_problem *p)
{
    if (!p) return EHM_OK;
    if (p->parked) {
        ehm_partition_free_now(p->parked);
        p->parked = nullptr;
    }
    cudaFree(p->dG); cudaFree(p->dS); cudaFree(p->dg); cudaFree(p->drho); cudaFree(p->dM);
    cudaFree(p->dcnt);
    ehm_oracle_scratch_free(p->oscr);
    delete p;
    return EHM_OK;
}

extern "C" int ehm_set_eps(ehm_problem *p, double eps_a, double eps_r)
{
    if (!p) return fail(EHM_EINVAL, "null problem");
    p->tb.eps_a = eps_a;
    p->tb.eps_r = eps_r;
    return EHM_OK;
}

extern "C" int ehm_num_candidates(const ehm_problem *p) { return p ? p->tb.ncand : 0; }

template <class SH>
static int grid_for(const ehm_problem *p, long long items)
{
    const long long ctas = (items + SH::WPC - 1) / SH::WPC;
    const long long cap = (long long)p->sm_count * SH::CPS;
    return (int)std::max(1ll, std::min(ctas, cap));
}

extern "C" int ehm_qp_theta_delta_batch(ehm_problem *p, int B, const double *d_theta,
                                        const int *d_cand, double *d_J, double *d_u0,
                                        int *d_status, void *stream)
{
    if (!p || B < 0 || !d_theta || !d_cand || !d_J || !d_status) return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    EHM_CUDA(cudaMemsetAsync(&p->dcnt->work, 0, sizeof(unsigned long long), st));
    switch (p->cfg) {
#define X(ID, UU, NN, CC, WW, PP)                                                                 \
    case ID: {                                                                                \
        using SH = Shape<UU, NN, CC, WW, PP>;                                                 \
        auto kern = k_point_batch<SH>;                                                        \
        int rc = set_smem(kern, SH::smem);                                                    \
        if (rc) return rc;                                                                    \
        kern<<<grid_for<SH>(p, B), 32 * SH::WPC, SH::smem, st>>>(p->tb, B, d_theta, d_cand, d_J, \
                                                                 d_u0, d_status, p->dcnt);    \
        break;                                                                                \
    }
        EHM_FOR_EACH_CFG(X)
#undef X
    default: return fail(EHM_EINVAL, "no kernel for this problem shape");
    }
    g_launches++;
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

extern "C" int ehm_simplex_batch(ehm_problem *p, int B, const double *d_R, const double *d_Vr,
                                 const int *d_cand, int mode, int with_phase1, double *d_val,
                                 double *d_sol, int *d_status, void *stream)
{
    if (!p || B < 0 || !d_R || !d_cand || !d_val || !d_status) return fail(EHM_EINVAL, "bad argument");
    if (mode < 0 || mode > 2 || (mode == EHM_SX_TQP && !d_Vr)) return fail(EHM_EINVAL, "bad mode");
    if (B == 0) return EHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    EHM_CUDA(cudaMemsetAsync(&p->dcnt->work, 0, sizeof(unsigned long long), st));
    switch (p->cfg) {
#define X(ID, UU, NN, CC, WW, PP)                                                                 \
    case ID: {                                                                                \
        using SH = Shape<UU, NN, CC, WW, PP>;                                                 \
        auto kern = k_simplex_batch<SH>;                                                      \
        int rc = set_smem(kern, SH::smem);                                                    \
        if (rc) return rc;                                                                    \
        kern<<<grid_for<SH>(p, B), 32 * SH::WPC, SH::smem, st>>>(p->tb, B, d_R, d_Vr, d_cand, \
                                                                 mode, with_phase1, d_val,    \
                                                                 d_sol, d_status, p->dcnt);   \
        break;                                                                                \
    }
        EHM_FOR_EACH_CFG(X)
#undef X
    default: return fail(EHM_EINVAL, "no kernel for this problem shape");
    }
    g_launches++;
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

extern "C" int ehm_split_batch(int n, int B, const double *d_R, double *d_S1, double *d_S2,
                               int *d_ij, void *stream)
{
    if (n < 1 || n > EHM_SPLIT_MAXN || B < 0 || !d_R || !d_S1 || !d_S2 || !d_ij) return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    const int grid = (int)std::min<long long>(((long long)B + kSplitWarps - 1) / kSplitWarps, 148ll * 16);
    k_split_batch<<<grid, 32 * kSplitWarps, 0, (cudaStream_t)stream>>>(n, B, d_R, d_S1, d_S2, d_ij);
    g_launches++;
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

// counters (host view) -- used by the Python layer and bench.py
extern "C" int ehm_last_failure(ehm_problem *p, int32_t *info4)
{
    if (!p || !info4) return fail(EHM_EINVAL, "bad argument");
    DevCounters h;
    EHM_CUDA(cudaMemcpy(&h, p->dcnt, sizeof h, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; ++k) info4[k] = h.last_fail[k];
    return EHM_OK;
}

extern "C" int ehm_read_counters(ehm_problem *p, int64_t *iters, int64_t *solves, int64_t *fails)
{
    if (!p) return fail(EHM_EINVAL, "null problem");
    DevCounters h;
    EHM_CUDA(cudaMemcpy(&h, p->dcnt, sizeof h, cudaMemcpyDeviceToHost));
    if (iters) *iters = (int64_t)h.iters;
    if (solves) *solves = (int64_t)h.solves;
    if (fails) *fails = (int64_t)h.fails;
    return EHM_OK;
}

// =========================================================================================
// partition engine (host side)
// =========================================================================================
#include "ehm_partition.cuh"
#include "ehm_locate.cuh"
#include "ehm_oracle_batch.cuh"

struct ehm_partition {
    ehm_problem *pb = nullptr;
    NodeStore ns{};
    Frontier fr{};
    WaveBuf wv{};
    EccBuf eb{};
    Ctl *dctl = nullptr;
    Ctl hctl{};
    std::vector<void *> allocs;
    int n = 0, nu = 0;
    ehm_partition_opts opts{};
    int ecc_q0 = 0;                 // ECC queue cursor
    DevCounters c0{};               // counters at creation
    double sms[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sbytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double device_ms = 0;
    int64_t launches = 0, waves = 0;
    template <typename T>
    int alloc(T **p, size_t count)
    {
        cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) return fail(EHM_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        allocs.push_back((void *)*p);
        return EHM_OK;
    }
    ~ehm_partition()
    {
        for (void *p : allocs) cudaFree(p);
    }
};

static void ehm_partition_free_now(void *t) { delete (ehm_partition *)t; }

namespace {

void fill_stats(ehm_partition *t, ehm_partition_stats *stats)
{
    if (!stats) return;
    DevCounters c1;
    cudaMemcpy(&c1, t->pb->dcnt, sizeof c1, cudaMemcpyDeviceToHost);
    const Ctl &h = t->hctl;
    const DevCounters &c0 = t->c0;
    stats->nnodes = h.nnodes; stats->nleaves = h.nleaves;
    stats->nsolves = (int64_t)(c1.solves - c0.solves); stats->niters = (int64_t)(c1.iters - c0.iters);
    stats->nfail = (int64_t)(c1.fails - c0.fails); stats->nwaves = t->waves; stats->nlaunches = t->launches;
    stats->open_left = h.ftop + (h.necc_next - t->ecc_q0); stats->frontier = h.ftop; stats->max_depth = h.max_depth;
    stats->device_ms = t->device_ms;
    for (int k = 0; k < 8; ++k) { stats->stage_ms[k] = t->sms[k]; stats->stage_bytes[k] = t->sbytes[k]; stats->stage_flops[k] = 0; }
    stats->stage_flops[1] = c1.flops[1] - c0.flops[1]; stats->stage_flops[3] = c1.flops[2] - c0.flops[2];
    stats->stage_flops[5] = c1.flops[3] - c0.flops[3]; stats->stage_flops[7] = c1.flops[4] - c0.flops[4];
}

template <class SH>
int advance_t(ehm_partition *t, int64_t max_waves, double budget_ms)
{
    // budget_ms > 0: stop after the first wave that ends past the budget (every wave ends with a
    // stream synchronisation, so the host clock follows the device)
    const auto t_in = std::chrono::steady_clock::now();
    auto in_budget = [&]() {
        return budget_ms <= 0.0 ||
               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_in).count() < budget_ms;
    };
    ehm_problem *p = t->pb;
    const DevTables &tb = p->tb;
    const ehm_partition_opts &o = t->opts;
    const int ncand = tb.ncand;
    constexpr int NX = SH::NX;
    const size_t smem = SH::smem;
    auto kref = k_lcss_ref<SH>;
    auto keval = k_lcss_eval<SH>;
    auto kpts = k_lcss_points<SH>;
    auto kecc = k_ecc_points<SH>;
    int rc;
    if ((rc = set_smem(kref, smem)) || (rc = set_smem(keval, smem)) || (rc = set_smem(kpts, smem)) ||
        (rc = set_smem(kecc, smem)))
        return rc;
    const int gsolve = p->sm_count * SH::CPS, tsolve = 32 * SH::WPC;
    const int gaux = p->sm_count * 4, taux = 256;
    cudaStream_t st = 0;
    cudaEvent_t ev0, ev1, sev[8];
    EHM_CUDA(cudaEventCreate(&ev0));
    EHM_CUDA(cudaEventCreate(&ev1));
    for (auto &e : sev) EHM_CUDA(cudaEventCreate(&e));
    EHM_CUDA(cudaEventRecord(ev0, st));
    int64_t launches = 0, waves = 0;
    Ctl &h = t->hctl;
    // ---- ECC phase: roots (and their ECC descendants) -> L-CSS entries on the frontier ----
    while (t->ecc_q0 < h.necc_next && h.error == 0 && (max_waves <= 0 || waves < max_waves) && in_budget()) {
        const int q0 = t->ecc_q0;
        const int ne = std::min(kEccWave, h.necc_next - q0);
        EHM_CUDA(cudaEventRecord(sev[0], st));
        k_reset_cursors<<<1, 32, 0, st>>>(t->dctl);
        kecc<<<gsolve, tsolve, smem, st>>>(tb, t->ns, t->eb, q0, ne, 0, t->dctl, p->dcnt, nullptr);
        k_ecc_pick<NX><<<(ne * 32 + 127) / 128, 128, 0, st>>>(tb, t->eb, ne, t->dctl);
        kecc<<<gsolve, tsolve, smem, st>>>(tb, t->ns, t->eb, q0, ne, 1, t->dctl, p->dcnt, nullptr);
        k_ecc_final<NX><<<(ne * 32 + 127) / 128, 128, 0, st>>>(tb, t->ns, t->fr, t->eb, q0, ne, t->dctl);
        EHM_CUDA(cudaEventRecord(sev[1], st));
        launches += 5;
        EHM_CUDA(cudaMemcpyAsync(&h, t->dctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        EHM_CUDA(cudaStreamSynchronize(st));
        float ems = 0;
        cudaEventElapsedTime(&ems, sev[0], sev[1]);
        t->sms[7] += ems;
        t->ecc_q0 += ne;
        ++waves;
        if (o.verbose) fprintf(stderr, "[ehm] ecc wave: queue %d/%d frontier %d nodes %d\n", t->ecc_q0, h.necc_next, h.ftop, h.nnodes);
    }
    // ---- L-CSS phase ------------------------------------------------------------------------
    while (h.ftop > 0 && h.error == 0 && (max_waves <= 0 || waves < max_waves) && in_budget()) {
        if (o.node_budget > 0 && h.nnodes >= o.node_budget) break;
        const int W = std::min(t->wv.W, h.ftop);
        const int base = h.ftop - W;
        EHM_CUDA(cudaEventRecord(sev[0], st));
        k_pop<<<gaux, taux, 0, st>>>(t->fr, t->wv, base, W, ncand, t->dctl);
        EHM_CUDA(cudaEventRecord(sev[1], st));
        kref<<<gsolve, tsolve, smem, st>>>(tb, t->ns, t->wv, W, t->dctl, p->dcnt);
        EHM_CUDA(cudaEventRecord(sev[2], st));
        k_lcss_items<<<gaux, taux, 0, st>>>(t->ns, t->wv, W, ncand, t->dctl);
        EHM_CUDA(cudaEventRecord(sev[3], st));
        keval<<<gsolve, tsolve, smem, st>>>(tb, t->ns, t->wv, t->dctl, p->dcnt);
        EHM_CUDA(cudaEventRecord(sev[4], st));
        k_lcss_select<NX><<<gaux, taux, 0, st>>>(t->ns, t->wv, W, ncand, t->dctl);
        EHM_CUDA(cudaEventRecord(sev[5], st));
        kpts<<<gsolve, tsolve, smem, st>>>(tb, t->ns, t->wv, t->dctl, p->dcnt);
        EHM_CUDA(cudaEventRecord(sev[6], st));
        k_lcss_final<NX><<<gaux, taux, 0, st>>>(tb, t->ns, t->fr, t->wv, W, t->dctl);
        EHM_CUDA(cudaEventRecord(sev[7], st));
        launches += 7;
        const int ftop_before = base;
        EHM_CUDA(cudaMemcpyAsync(&h, t->dctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        EHM_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 7; ++k) {
            float ems = 0;
            cudaEventElapsedTime(&ems, sev[k], sev[k + 1]);
            t->sms[k] += ems;
        }
        // algorithmic HBM bytes of the streaming stages: pop moves W bound rows (read+write),
        // items reads them and writes tval(8)+titem(4), final reads one row and writes one
        // row per pushed entry
        t->sbytes[0] += 2.0 * W * ncand * 4.0;
        t->sbytes[2] += (double)W * ncand * 16.0;
        t->sbytes[6] += (double)(h.ftop - ftop_before) * ncand * 8.0;
        ++waves;
        if (o.verbose)
            fprintf(stderr, "[ehm] wave %lld: popped %d items %d pitems %d -> frontier %d nodes %d leaves %d\n",
                    (long long)(t->waves + waves), W, h.nitems, h.npitems, h.ftop, h.nnodes, h.nleaves);
    }
    EHM_CUDA(cudaEventRecord(ev1, st));
    EHM_CUDA(cudaEventSynchronize(ev1));
    EHM_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    for (auto &e : sev) cudaEventDestroy(e);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    t->device_ms += ms;
    t->launches += launches;
    t->waves += waves;
    g_launches += launches;
    if (h.error == EHM_EINFEASIBLE) return fail(EHM_EINFEASIBLE, "STOP, Theta contains infeasible regions");
    if (h.error == EHM_ECAPACITY) return fail(EHM_ECAPACITY, "node store or frontier capacity exhausted");
    if (h.error == EHM_ENUMERIC) return fail(EHM_ENUMERIC, "midpoint vertex problem infeasible");
    return EHM_OK;
}

}  // namespace

extern "C" int ehm_partition_create(ehm_problem *p, int nroots, const double *roots,
                                    const ehm_partition_opts *opts, ehm_partition **out)
{
    if (!p || nroots < 0 || (nroots > 0 && !roots) || !out) return fail(EHM_EINVAL, "bad argument");
    *out = nullptr;
    ehm_partition_opts o{};
    if (opts) o = *opts;
    const DevTables &tb = p->tb;
    const int n = tb.n, nu = tb.nu, ncand = tb.ncand;
    const int W = o.wave > 0 ? o.wave : 4096;
    const int64_t maxn = o.max_nodes > 0 ? o.max_nodes : (int64_t)1 << 22;
    const int64_t fcap = o.max_frontier > 0 ? o.max_frontier : std::max<int64_t>(65536, 64 * (int64_t)W);
    if (maxn > 0x7ffffff0ll || fcap > 0x7ffffff0ll || nroots > maxn) return fail(EHM_EINVAL, "capacity too large");
    // device buffers (several GB) are recycled: ehm_partition_free parks one object per problem
    ehm_partition *t = nullptr;
    if (p->parked) {
        ehm_partition *c = (ehm_partition *)p->parked;
        p->parked = nullptr;
        if (c->ns.cap == (int)maxn && c->fr.cap == (int)fcap && c->wv.W == W) t = c;
        else delete c;
    }
    const bool recycled = (t != nullptr);
    if (!t) t = new ehm_partition();
    t->pb = p; t->n = n; t->nu = nu; t->opts = o;
    t->ecc_q0 = 0; t->device_ms = 0; t->launches = 0; t->waves = 0;
    for (int k = 0; k < 8; ++k) { t->sms[k] = 0; t->sbytes[k] = 0; }
    int rc = EHM_OK;
    auto A = [&](int r) { if (rc == EHM_OK) rc = r; };
    NodeStore &ns = t->ns;
    Frontier &fr = t->fr;
    WaveBuf &wv = t->wv;
    EccBuf &eb = t->eb;
    if (!recycled) {
    ns.cap = (int)maxn;
    A(t->alloc(&ns.verts, (size_t)maxn * (n + 1) * n)); A(t->alloc(&ns.vcost, (size_t)maxn * (n + 1)));
    A(t->alloc(&ns.vinput, (size_t)maxn * (n + 1) * nu)); A(t->alloc(&ns.cand, maxn));
    A(t->alloc(&ns.left, maxn)); A(t->alloc(&ns.right, maxn)); A(t->alloc(&ns.root_id, maxn));
    A(t->alloc(&ns.depth, maxn)); A(t->alloc(&ns.closed, maxn)); A(t->alloc(&ns.origin, (size_t)2 * maxn));
    fr.cap = (int)fcap;
    A(t->alloc(&fr.node, fcap)); A(t->alloc(&fr.skip, fcap)); A(t->alloc(&fr.lb, (size_t)fcap * ncand));
    wv.W = W;
    const size_t WN = (size_t)W * ncand;
    A(t->alloc(&wv.node, W)); A(t->alloc(&wv.skip, W)); A(t->alloc(&wv.lb, WN)); A(t->alloc(&wv.tref, W));
    A(t->alloc(&wv.tval, WN)); A(t->alloc(&wv.titem, WN)); A(t->alloc(&wv.items, WN));
    A(t->alloc(&wv.it_sol, WN * (n + 1))); A(t->alloc(&wv.it_pl, WN * (n + 1))); A(t->alloc(&wv.sel, (size_t)W * kSelT)); A(t->alloc(&wv.nsel, W));
    A(t->alloc(&wv.nimp, W)); A(t->alloc(&wv.conv, W)); A(t->alloc(&wv.ij, (size_t)2 * W));
    A(t->alloc(&wv.vmid, (size_t)W * n));
    const size_t PP = (size_t)W * ((kSelT + 1) * (n + 2) + 1);
    A(t->alloc(&wv.pitems, PP)); A(t->alloc(&wv.ptJ, (size_t)W * (kSelT + 1) * (n + 2)));
    A(t->alloc(&wv.ptU, (size_t)W * (kSelT + 1) * (n + 2) * nu)); A(t->alloc(&wv.ptS, (size_t)W * (kSelT + 1) * (n + 2)));
    A(t->alloc(&wv.minv, W)); A(t->alloc(&wv.minv_st, W));
    A(t->alloc(&eb.queue, maxn)); A(t->alloc(&eb.J, (size_t)kEccWave * ncand * (n + 2)));
    A(t->alloc(&eb.U, (size_t)kEccWave * ncand * (n + 1) * nu)); A(t->alloc(&eb.S, (size_t)kEccWave * ncand * (n + 2)));
    A(t->alloc(&eb.pick, kEccWave));
    A(t->alloc(&t->dctl, 1));
    }
    if (rc != EHM_OK) { delete t; return rc; }
    // root nodes
    {
        std::vector<int> iz(std::max(nroots, 1), -1), id(std::max(nroots, 1)), zero(std::max(nroots, 1), 0);
        std::vector<int> org((size_t)2 * std::max(nroots, 1), -1);
        for (int i = 0; i < nroots; ++i) id[i] = i;
        std::vector<unsigned char> cz(std::max(nroots, 1), 0);
        cudaError_t e = cudaSuccess;
        if (nroots > 0) {
            e = cudaMemcpy(ns.verts, roots, (size_t)nroots * (n + 1) * n * sizeof(double), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemset(ns.vcost, 0, (size_t)nroots * (n + 1) * sizeof(double));
            if (e == cudaSuccess) e = cudaMemset(ns.vinput, 0, (size_t)nroots * (n + 1) * nu * sizeof(double));
            if (e == cudaSuccess) e = cudaMemcpy(ns.cand, iz.data(), nroots * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ns.left, iz.data(), nroots * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ns.right, iz.data(), nroots * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ns.root_id, id.data(), nroots * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ns.depth, zero.data(), nroots * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ns.closed, cz.data(), nroots, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ns.origin, org.data(), (size_t)2 * nroots * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(eb.queue, id.data(), nroots * sizeof(int), cudaMemcpyHostToDevice);
        }
        memset(&t->hctl, 0, sizeof(Ctl));
        t->hctl.nnodes = nroots;
        t->hctl.necc_next = nroots;
        if (e == cudaSuccess) e = cudaMemcpy(t->dctl, &t->hctl, sizeof(Ctl), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(&t->c0, p->dcnt, sizeof(DevCounters), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { delete t; return fail(EHM_ECUDA, std::string("root upload: ") + cudaGetErrorString(e)); }
    }
    *out = t;
    return EHM_OK;
}

extern "C" int ehm_partition_set_root_ids(ehm_partition *t, int nroots, const int32_t *ids)
{
    if (!t || nroots < 0 || nroots > t->hctl.nnodes || (nroots > 0 && !ids)) return fail(EHM_EINVAL, "bad argument");
    if (nroots == 0) return EHM_OK;
    EHM_CUDA(cudaMemcpy(t->ns.root_id, ids, (size_t)nroots * sizeof(int), cudaMemcpyHostToDevice));
    return EHM_OK;
}

extern "C" int ehm_partition_advance_for(ehm_partition *t, int64_t max_waves, double budget_ms,
                                         ehm_partition_stats *stats)
{
    if (!t) return fail(EHM_EINVAL, "null partition");
    int rc;
    switch (t->pb->cfg) {
#define X(ID, UU, NN, CC, WW, PP) \
    case ID: rc = advance_t<Shape<UU, NN, CC, WW, PP>>(t, max_waves, budget_ms); break;
        EHM_FOR_EACH_CFG(X)
#undef X
    default: rc = fail(EHM_EINVAL, "no kernel for this problem shape");
    }
    fill_stats(t, stats);
    return rc;
}

extern "C" int ehm_partition_advance(ehm_partition *t, int64_t max_waves, ehm_partition_stats *stats)
{
    return ehm_partition_advance_for(t, max_waves, 0.0, stats);
}

extern "C" int ehm_partition_run(ehm_problem *p, int nroots, const double *roots,
                                 const ehm_partition_opts *opts, ehm_partition **out,
                                 ehm_partition_stats *stats)
{
    if (!out || nroots <= 0) return fail(EHM_EINVAL, "bad argument");
    int rc = ehm_partition_create(p, nroots, roots, opts, out);
    if (rc != EHM_OK) return rc;
    return ehm_partition_advance(*out, 0, stats);  // *out stays valid on error (partial tree)
}

// ---- frontier exchange between ranks (multi-GPU rebalance) --------------------------------
// Records: meta int32[k][6] = {origin node id, cand, depth, root_id, skip, 0},
//          geom f64[k][(n+1)n + (n+1) + (n+1)nu], lb f32[k][ncand].
__global__ void k_export(NodeStore ns, Frontier fr, int ftop, int k, int n, int nu, int ncand,
                         int *meta, double *geom, float *lb)
{
    // exported entries are the k OLDEST (bottom of the stack: largest sub-trees); the holes are
    // refilled with the youngest entries so that the stack stays dense
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int ng = (n + 1) * n + (n + 1) + (n + 1) * nu;
    for (int e = gw; e < k; e += nw) {
        const int nd = fr.node[e];
        for (int c = lane; c < ncand; c += 32) lb[(size_t)e * ncand + c] = fr.lb[(size_t)e * ncand + c];
        for (int q = lane; q < ng; q += 32) {
            double v;
            if (q < (n + 1) * n) v = ns.verts[(size_t)nd * (n + 1) * n + q];
            else if (q < (n + 1) * n + (n + 1)) v = ns.vcost[(size_t)nd * (n + 1) + q - (n + 1) * n];
            else v = ns.vinput[(size_t)nd * (n + 1) * nu + q - (n + 1) * n - (n + 1)];
            geom[(size_t)e * ng + q] = v;
        }
        if (lane == 0) {
            meta[e * 6 + 0] = nd; meta[e * 6 + 1] = ns.cand[nd]; meta[e * 6 + 2] = ns.depth[nd];
            meta[e * 6 + 3] = ns.root_id[nd]; meta[e * 6 + 4] = fr.skip[e]; meta[e * 6 + 5] = 0;
            ns.closed[nd] = 2;  // exported: continued on another rank
        }
    }
}

__global__ void k_refill(Frontier fr, int ftop, int k, int ncand)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = gw; e < k; e += nw) {
        const int src = ftop - 1 - e;
        if (src < k) continue;  // that slot is itself being exported
        for (int c = lane; c < ncand; c += 32) fr.lb[(size_t)e * ncand + c] = fr.lb[(size_t)src * ncand + c];
        if (lane == 0) {
            fr.node[e] = fr.node[src];
            fr.skip[e] = fr.skip[src];
        }
    }
}

__global__ void k_import(NodeStore ns, Frontier fr, int node0, int pos0, int k, int n, int nu, int ncand,
                         const int *meta, const double *geom, const float *lb, const int *origin_rank)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int ng = (n + 1) * n + (n + 1) + (n + 1) * nu;
    for (int e = gw; e < k; e += nw) {
        const int nd = node0 + e, pos = pos0 + e;
        for (int c = lane; c < ncand; c += 32) fr.lb[(size_t)pos * ncand + c] = lb[(size_t)e * ncand + c];
        for (int q = lane; q < ng; q += 32) {
            const double v = geom[(size_t)e * ng + q];
            if (q < (n + 1) * n) ns.verts[(size_t)nd * (n + 1) * n + q] = v;
            else if (q < (n + 1) * n + (n + 1)) ns.vcost[(size_t)nd * (n + 1) + q - (n + 1) * n] = v;
            else ns.vinput[(size_t)nd * (n + 1) * nu + q - (n + 1) * n - (n + 1)] = v;
        }
        if (lane == 0) {
            ns.cand[nd] = meta[e * 6 + 1]; ns.depth[nd] = meta[e * 6 + 2]; ns.root_id[nd] = meta[e * 6 + 3];
            ns.left[nd] = -1; ns.right[nd] = -1; ns.closed[nd] = 0;
            ns.origin[2 * nd] = origin_rank[e]; ns.origin[2 * nd + 1] = meta[e * 6 + 0];
            fr.node[pos] = nd; fr.skip[pos] = meta[e * 6 + 4];
        }
    }
}

extern "C" int ehm_partition_export(ehm_partition *t, int k, int32_t *d_meta, double *d_geom, float *d_lb)
{
    if (!t || k < 0 || (k > 0 && (!d_meta || !d_geom || !d_lb))) return fail(EHM_EINVAL, "bad argument");
    Ctl &h = t->hctl;
    if (k > h.ftop) return fail(EHM_EINVAL, "export count exceeds frontier size");
    if (k == 0) return EHM_OK;
    const int ncand = t->pb->tb.ncand;
    k_export<<<t->pb->sm_count * 4, 256>>>(t->ns, t->fr, h.ftop, k, t->n, t->nu, ncand, d_meta, d_geom, d_lb);
    k_refill<<<t->pb->sm_count * 4, 256>>>(t->fr, h.ftop, k, ncand);
    g_launches += 2;
    h.ftop -= k;
    EHM_CUDA(cudaMemcpy(&t->dctl->ftop, &h.ftop, sizeof(int), cudaMemcpyHostToDevice));
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

extern "C" int ehm_partition_import(ehm_partition *t, int k, const int32_t *d_meta, const double *d_geom,
                                    const float *d_lb, const int32_t *d_origin_rank)
{
    if (!t || k < 0 || (k > 0 && (!d_meta || !d_geom || !d_lb || !d_origin_rank))) return fail(EHM_EINVAL, "bad argument");
    if (k == 0) return EHM_OK;
    Ctl &h = t->hctl;
    if (h.nnodes + k > t->ns.cap || h.ftop + k > t->fr.cap) return fail(EHM_ECAPACITY, "import exceeds capacity");
    k_import<<<t->pb->sm_count * 4, 256>>>(t->ns, t->fr, h.nnodes, h.ftop, k, t->n, t->nu, t->pb->tb.ncand, d_meta,
                                            d_geom, d_lb, d_origin_rank);
    g_launches += 1;
    h.nnodes += k;
    h.ftop += k;
    EHM_CUDA(cudaMemcpy(&t->dctl->ftop, &h.ftop, sizeof(int), cudaMemcpyHostToDevice));
    EHM_CUDA(cudaMemcpy(&t->dctl->nnodes, &h.nnodes, sizeof(int), cudaMemcpyHostToDevice));
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

extern "C" int ehm_partition_fetch_origin(ehm_partition *t, int32_t *origin)
{
    if (!t || !origin) return fail(EHM_EINVAL, "bad argument");
    EHM_CUDA(cudaMemcpy(origin, t->ns.origin, (size_t)2 * t->hctl.nnodes * sizeof(int), cudaMemcpyDeviceToHost));
    return EHM_OK;
}

extern "C" int ehm_partition_fetch(ehm_partition *t, int32_t *left, int32_t *right, int32_t *root_id,
                                   int32_t *depth, int32_t *cand, uint8_t *closed, double *vertices,
                                   double *vcost, double *vinput)
{
    if (!t) return fail(EHM_EINVAL, "null partition");
    const size_t nn = (size_t)t->hctl.nnodes, n = t->n, nu = t->nu;
    auto dl = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        return dst ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : cudaSuccess;
    };
    EHM_CUDA(dl(left, t->ns.left, nn * 4)); EHM_CUDA(dl(right, t->ns.right, nn * 4));
    EHM_CUDA(dl(root_id, t->ns.root_id, nn * 4)); EHM_CUDA(dl(depth, t->ns.depth, nn * 4));
    EHM_CUDA(dl(cand, t->ns.cand, nn * 4)); EHM_CUDA(dl(closed, t->ns.closed, nn));
    EHM_CUDA(dl(vertices, t->ns.verts, nn * (n + 1) * n * 8)); EHM_CUDA(dl(vcost, t->ns.vcost, nn * (n + 1) * 8));
    EHM_CUDA(dl(vinput, t->ns.vinput, nn * (n + 1) * nu * 8));
    return EHM_OK;
}

extern "C" int ehm_partition_free(ehm_partition *t)
{
    if (!t) return EHM_OK;
    ehm_problem *p = t->pb;
    if (p && !p->parked) p->parked = t;   // keep the device buffers for the next run
    else delete t;
    return EHM_OK;
}

// =========================================================================================
// batched reference-shaped oracle calls (kernels in ehm_oracle_batch.cuh)
// =========================================================================================
namespace {

int oracle_scratch(ehm_problem *p, int B, OracleScratch **out)
{
    OracleScratch *sc = (OracleScratch *)p->oscr;
    if (sc && sc->cap >= B) { *out = sc; return EHM_OK; }
    delete sc;
    p->oscr = nullptr;
    sc = new OracleScratch();
    const int cap = std::max(B, 8);
    const DevTables &tb = p->tb;
    const size_t n = tb.n, nu = tb.nu, nc = tb.ncand, W = cap, WN = W * nc;
    cudaError_t e = cudaSuccess;
    auto A = [&](auto **ptr, size_t count) {
        if (e != cudaSuccess) return;
        e = cudaMalloc((void **)ptr, std::max<size_t>(count, 1) * sizeof(**ptr));
        if (e == cudaSuccess) sc->allocs.push_back((void *)*ptr);
    };
    WaveBuf &wv = sc->wv;
    EccBuf &eb = sc->eb;
    wv.W = cap;
    A(&wv.node, W); A(&wv.skip, W); A(&wv.lb, WN); A(&wv.tref, W); A(&wv.tval, WN); A(&wv.titem, WN);
    A(&wv.items, WN); A(&wv.it_sol, WN * (n + 1)); A(&wv.it_pl, WN * (n + 1)); A(&wv.sel, W * kSelT);
    A(&wv.nsel, W); A(&wv.nimp, W); A(&wv.conv, W); A(&wv.ij, 2 * W); A(&wv.vmid, W * n);
    A(&wv.pitems, W * ((kSelT + 1) * (n + 2) + 1)); A(&wv.ptJ, W * (kSelT + 1) * (n + 2));
    A(&wv.ptU, W * (kSelT + 1) * (n + 2) * nu); A(&wv.ptS, W * (kSelT + 1) * (n + 2));
    A(&wv.minv, W); A(&wv.minv_st, W);
    A(&eb.queue, W); A(&eb.J, WN * (n + 2)); A(&eb.U, WN * (n + 1) * nu); A(&eb.S, WN * (n + 2)); A(&eb.pick, W);
    A(&sc->dctl, 1); A(&sc->more, 1); A(&sc->neg1, W);
    if (e != cudaSuccess) {
        delete sc;
        return fail(EHM_ENOMEM, std::string("oracle scratch: ") + cudaGetErrorString(e));
    }
    sc->cap = cap;
    p->oscr = sc;
    *out = sc;
    return EHM_OK;
}

// the caller's arrays viewed as a node store: entry e = node e
NodeStore view_of(int B, const double *R, const double *Vr, const int *cref)
{
    NodeStore ns{};
    ns.cap = B;
    ns.verts = const_cast<double *>(R);
    ns.vcost = const_cast<double *>(Vr);
    ns.cand = const_cast<int *>(cref);
    return ns;
}

template <class SH>
int pick_t(ehm_problem *p, OracleScratch *sc, int B, const double *d_R, const double *d_theta, int nslots,
           int *d_cand, double *d_J, double *d_U, int *d_status, cudaStream_t st)
{
    constexpr int NX = SH::NX;
    auto kecc = k_ecc_points<SH>;
    int rc = set_smem(kecc, SH::smem);
    if (rc) return rc;
    const NodeStore ns = view_of(B, d_R, nullptr, nullptr);
    const int gaux = p->sm_count * 4;
    k_batch_init<<<gaux, 256, 0, st>>>(sc->wv, sc->eb, B, p->tb.ncand, sc->dctl, sc->more, sc->neg1);
    kecc<<<p->sm_count * SH::CPS, 32 * SH::WPC, SH::smem, st>>>(p->tb, ns, sc->eb, 0, B, d_theta ? 2 : 0, sc->dctl,
                                                                 p->dcnt, d_theta);
    k_ecc_pick<NX><<<(B * 32 + 127) / 128, 128, 0, st>>>(p->tb, sc->eb, B, sc->dctl, nslots);
    k_batch_pick_out<NX><<<(B + 127) / 128, 128, 0, st>>>(p->tb, sc->eb, B, nslots, d_cand, d_J, d_U, d_status);
    g_launches += 4;
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

template <class SH>
int bar_t(ehm_problem *p, OracleScratch *sc, int B, const double *d_R, const double *d_Vr, const int *d_cref,
          int *d_conv, double *d_tmin, int *d_cstar, double *d_theta, double *d_Jv, double *d_Uv, int *d_vs,
          int *d_status, cudaStream_t st)
{
    constexpr int NX = SH::NX;
    auto kref = k_lcss_ref<SH>;
    auto keval = k_lcss_eval<SH>;
    auto kpts = k_lcss_points<SH>;
    int rc;
    if ((rc = set_smem(kref, SH::smem)) || (rc = set_smem(keval, SH::smem)) || (rc = set_smem(kpts, SH::smem))) return rc;
    const DevTables &tb = p->tb;
    const NodeStore ns = view_of(B, d_R, d_Vr, d_cref ? d_cref : sc->neg1);
    const int gaux = p->sm_count * 4, gsolve = p->sm_count * SH::CPS, tsolve = 32 * SH::WPC;
    k_batch_init<<<gaux, 256, 0, st>>>(sc->wv, sc->eb, B, tb.ncand, sc->dctl, sc->more, sc->neg1);
    if (d_cref) kref<<<gsolve, tsolve, SH::smem, st>>>(tb, ns, sc->wv, B, sc->dctl, p->dcnt);
    k_lcss_items<<<gaux, 256, 0, st>>>(ns, sc->wv, B, tb.ncand, sc->dctl);
    keval<<<gsolve, tsolve, SH::smem, st>>>(tb, ns, sc->wv, sc->dctl, p->dcnt);
    g_launches += 3 + (d_cref ? 1 : 0);
    if (!d_cstar) {   // bar_E
        k_lcss_select<NX><<<gaux, 256, 0, st>>>(ns, sc->wv, B, tb.ncand, sc->dctl);
        k_batch_bare_out<<<(B * 32 + 127) / 128, 128, 0, st>>>(sc->wv, B, tb.ncand, d_conv, d_tmin);
        g_launches += 2;
        EHM_CUDA(cudaGetLastError());
        return EHM_OK;
    }
    for (int round = 0; round < 1 + tb.ncand / kSelT; ++round) {   // bar_D: vet kSelT candidates per round
        k_batch_round<<<1, 32, 0, st>>>(sc->dctl, sc->more);
        k_lcss_select<NX><<<gaux, 256, 0, st>>>(ns, sc->wv, B, tb.ncand, sc->dctl);
        kpts<<<gsolve, tsolve, SH::smem, st>>>(tb, ns, sc->wv, sc->dctl, p->dcnt);
        k_batch_bard_out<NX><<<(B + 127) / 128, 128, 0, st>>>(tb, ns, sc->wv, B, d_cstar, d_theta, d_Jv, d_Uv, d_vs,
                                                             d_status, sc->more);
        g_launches += 4;
        int more = 0;
        EHM_CUDA(cudaMemcpyAsync(&more, sc->more, sizeof(int), cudaMemcpyDeviceToHost, st));
        EHM_CUDA(cudaStreamSynchronize(st));
        if (more == 0) break;
    }
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}

}  // namespace

static void ehm_oracle_scratch_free(void *s) { delete (ehm::OracleScratch *)s; }

#define EHM_DISPATCH(CALL)                                               \
    switch (p->cfg) {                                                    \
        EHM_FOR_EACH_CFG(X)                                              \
    default: return fail(EHM_EINVAL, "no kernel for this problem shape"); \
    }

extern "C" int ehm_p_theta_batch(ehm_problem *p, int B, const double *d_theta, int *d_cand, double *d_J,
                                 double *d_u0, int *d_status, void *stream)
{
    if (!p || B < 0 || !d_theta || !d_cand) return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    OracleScratch *sc;
    int rc = oracle_scratch(p, B, &sc);
    if (rc) return rc;
#define X(ID, UU, NN, CC, WW, PP) \
    case ID: return pick_t<Shape<UU, NN, CC, WW, PP>>(p, sc, B, nullptr, d_theta, 1, d_cand, d_J, d_u0, d_status, (cudaStream_t)stream);
    EHM_DISPATCH(X)
#undef X
}

extern "C" int ehm_v_r_batch(ehm_problem *p, int B, const double *d_R, int *d_cand, double *d_Jv, double *d_Uv,
                             int *d_status, void *stream)
{
    if (!p || B < 0 || !d_R || !d_cand) return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    OracleScratch *sc;
    int rc = oracle_scratch(p, B, &sc);
    if (rc) return rc;
#define X(ID, UU, NN, CC, WW, PP) \
    case ID: return pick_t<Shape<UU, NN, CC, WW, PP>>(p, sc, B, d_R, nullptr, NN + 1, d_cand, d_Jv, d_Uv, d_status, (cudaStream_t)stream);
    EHM_DISPATCH(X)
#undef X
}

extern "C" int ehm_bar_e_batch(ehm_problem *p, int B, const double *d_R, const double *d_Vr, int *d_converged,
                               double *d_tmin, void *stream)
{
    if (!p || B < 0 || !d_R || !d_Vr || !d_converged) return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    OracleScratch *sc;
    int rc = oracle_scratch(p, B, &sc);
    if (rc) return rc;
#define X(ID, UU, NN, CC, WW, PP) \
    case ID: return bar_t<Shape<UU, NN, CC, WW, PP>>(p, sc, B, d_R, d_Vr, nullptr, d_converged, d_tmin, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    EHM_DISPATCH(X)
#undef X
}

extern "C" int ehm_bar_d_batch(ehm_problem *p, int B, const double *d_R, const double *d_Vr, const int *d_cand_ref,
                               int *d_cand_star, double *d_theta_star, double *d_Jv, double *d_Uv,
                               int *d_var_small, int *d_status, void *stream)
{
    if (!p || B < 0 || !d_R || !d_Vr || !d_cand_ref || !d_cand_star || !d_theta_star || !d_Jv || !d_Uv || !d_var_small)
        return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    OracleScratch *sc;
    int rc = oracle_scratch(p, B, &sc);
    if (rc) return rc;
#define X(ID, UU, NN, CC, WW, PP) \
    case ID: return bar_t<Shape<UU, NN, CC, WW, PP>>(p, sc, B, d_R, d_Vr, d_cand_ref, nullptr, nullptr, d_cand_star, d_theta_star, d_Jv, d_Uv, d_var_small, d_status, (cudaStream_t)stream);
    EHM_DISPATCH(X)
#undef X
}

// =========================================================================================
// measurement helpers
// =========================================================================================
__global__ void k_dfma_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int ehm_fp64_peak(double *tflops)
{
    if (!tflops) return fail(EHM_EINVAL, "null");
    int dev = 0, sms = 148;
    EHM_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 4, threads = 512, iters = 1 << 16;
    double *d = nullptr;
    EHM_CUDA(cudaMalloc((void **)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0) best = std::max(best, tf);
    }
    g_launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    EHM_CUDA(cudaGetLastError());
    *tflops = best;
    return EHM_OK;
}

__global__ void k_copy_peak(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

extern "C" int ehm_hbm_peak(int64_t bytes, double *gbs)
{
    if (!gbs || bytes < (1 << 20)) return fail(EHM_EINVAL, "bad argument");
    const size_t n = (size_t)bytes / sizeof(double2);
    double2 *a = nullptr, *b = nullptr;
    EHM_CUDA(cudaMalloc((void **)&a, n * sizeof(double2)));
    EHM_CUDA(cudaMalloc((void **)&b, n * sizeof(double2)));
    EHM_CUDA(cudaMemset(a, 0, n * sizeof(double2)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_copy_peak<<<148 * 16, 512>>>(a, b, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::max(best, 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9);
    }
    g_launches += 5;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    EHM_CUDA(cudaGetLastError());
    *gbs = best;
    return EHM_OK;
}

// ---- debugging aid -----------------------------------------------------------------------
extern "C" int ehm_debug_trace(int on)
{
    EHM_CUDA(cudaMemcpyToSymbol(g_trace_on, &on, sizeof(int)));
    if (on) {
        static double zeros[8 * 128];
        EHM_CUDA(cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros)));
    }
    return EHM_OK;
}
extern "C" int ehm_debug_fetch(double *out)
{
    EHM_CUDA(cudaMemcpyFromSymbol(out, g_trace, sizeof(double) * 8 * 128));
    return EHM_OK;
}


// ---- online evaluation of the explicit control law (lib/mpc_library.py:662-792) ------------
struct ehm_locator {
    ehm::LocTree t{};
    int *left = nullptr, *right = nullptr;
    double *cm = nullptr, *vinput = nullptr;
};

extern "C" int ehm_locator_destroy(ehm_locator *l)
{
    if (!l) return EHM_OK;
    cudaFree(l->left); cudaFree(l->right); cudaFree(l->cm); cudaFree(l->vinput);
    delete l;
    return EHM_OK;
}

extern "C" int ehm_locator_create(int n, int nu, int64_t nnodes, const int32_t *left, const int32_t *right,
                                  const double *vertices, const double *vinput, int32_t top, ehm_locator **out)
{
    if (!out || n < 1 || n > ehm::kLocMaxN || nu < 1 || nnodes < 1 || !left || !right || !vertices || !vinput ||
        top < 0 || top >= nnodes)
        return fail(EHM_EINVAL, "bad argument");
    *out = nullptr;
    ehm_locator *l = new ehm_locator();
    double *dv = nullptr;
    const size_t nv = (size_t)nnodes * (n + 1) * n, nc = (size_t)nnodes * (n + n * n),
                 ni = (size_t)nnodes * (n + 1) * nu;
    cudaError_t e = cudaMalloc(&l->left, nnodes * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&l->right, nnodes * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&l->cm, nc * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&l->vinput, ni * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&dv, nv * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(l->left, left, nnodes * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(l->right, right, nnodes * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(l->vinput, vinput, ni * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dv, vertices, nv * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        ehm::k_loc_prepare<<<(unsigned)((nnodes + 127) / 128), 128>>>(n, nnodes, dv, l->cm);
        g_launches += 1;
        e = cudaDeviceSynchronize();
    }
    cudaFree(dv);
    if (e != cudaSuccess) {
        ehm_locator_destroy(l);
        return fail(EHM_ECUDA, std::string("locator: ") + cudaGetErrorString(e));
    }
    l->t.n = n; l->t.nu = nu; l->t.top = top; l->t.nnodes = nnodes;
    l->t.left = l->left; l->t.right = l->right; l->t.cm = l->cm; l->t.vinput = l->vinput;
    *out = l;
    return EHM_OK;
}

extern "C" int ehm_locator_eval(ehm_locator *l, int64_t B, const double *d_theta, double *d_u, int32_t *d_leaf,
                                void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!l || B < 0 || (B > 0 && (!d_theta || !d_u))) return fail(EHM_EINVAL, "bad argument");
    if (B == 0) return EHM_OK;
    const unsigned grid = (unsigned)((B + 127) / 128);
    switch (l->t.n) {
#define X(NN) case NN: ehm::k_locate<NN><<<grid, 128, 0, stream>>>(l->t, B, d_theta, d_u, d_leaf); break;
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
    default: return fail(EHM_EINVAL, "unsupported dimension");
    }
    g_launches += 1;
    EHM_CUDA(cudaGetLastError());
    return EHM_OK;
}
