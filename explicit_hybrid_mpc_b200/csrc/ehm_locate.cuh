// ehm_locate.cuh -- online evaluation of the explicit control law (included by ehm_lib.cu).
//
// Replaces the reference's ExplicitMPC (lib/mpc_library.py:662-792): point location in the
// binary partition tree by barycentric containment tests of the LEFT child, then barycentric
// interpolation of the leaf's vertex inputs.  One thread per query; the tree is a flat array
// of 8(n + n*n)-byte records {c = v0, Minv = inv([v1-v0 ... vn-v0])} (160 B for n = 4; the
// whole pendulum tree is 33 MB and stays in the 126 MB L2).  Latency/L2 bound: a query touches
// (#root simplices before its own) + depth records.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ehm {

constexpr int kLocMaxN = 8;

struct LocTree {
    int n, nu, top;
    int64_t nnodes;
    const int *left, *right;
    const double *cm;      // [nnodes][n + n*n]: c, then Minv row-major
    const double *vinput;  // [nnodes][(n+1)*nu]
};

// one thread per node: c = v0, Minv by Gauss-Jordan with partial pivoting (singular -> NaN)
__global__ void k_loc_prepare(int n, int64_t nnodes, const double *__restrict__ verts, double *__restrict__ cm)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    const double *v = verts + i * (n + 1) * n;
    double A[kLocMaxN * kLocMaxN], I[kLocMaxN * kLocMaxN];
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            A[r * n + c] = v[(c + 1) * n + r] - v[r];   // column c = v_{c+1} - v_0
            I[r * n + c] = (r == c) ? 1.0 : 0.0;
        }
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[k * n + k]);
        for (int r = k + 1; r < n; ++r)
            if (fabs(A[r * n + k]) > best) { best = fabs(A[r * n + k]); p = r; }
        if (!(best > 0.0)) { ok = false; break; }
        if (p != k)
            for (int c = 0; c < n; ++c) {
                double t = A[k * n + c]; A[k * n + c] = A[p * n + c]; A[p * n + c] = t;
                t = I[k * n + c]; I[k * n + c] = I[p * n + c]; I[p * n + c] = t;
            }
        const double inv = 1.0 / A[k * n + k];
        for (int c = 0; c < n; ++c) { A[k * n + c] *= inv; I[k * n + c] *= inv; }
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            const double f = A[r * n + k];
            if (f == 0.0) continue;
            for (int c = 0; c < n; ++c) {
                A[r * n + c] = fma(-f, A[k * n + c], A[r * n + c]);
                I[r * n + c] = fma(-f, I[k * n + c], I[r * n + c]);
            }
        }
    }
    double *o = cm + i * (n + n * n);
    for (int c = 0; c < n; ++c) o[c] = v[c];
    for (int e = 0; e < n * n; ++e) o[n + e] = ok ? I[e] : NAN;
}

template <int N>
__device__ __forceinline__ bool loc_alpha(const double *__restrict__ rec, const double (&x)[N], double (&al)[N + 1])
{
    double d[N];
#pragma unroll
    for (int j = 0; j < N; ++j) d[j] = x[j] - rec[j];
    double sum = 0.0;
    const double eps = 2.220446049250313e-16;
    bool in = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) a = fma(rec[N + i * N + j], d[j], a);
        al[i + 1] = a;
        sum += a;
        in = in && (a >= -eps) && (a <= 1.0 + eps);
    }
    al[0] = 1.0 - sum;
    in = in && (al[0] >= -eps) && (al[0] <= 1.0 + eps);
    return in;
}

template <int N>
__global__ void k_locate(LocTree t, int64_t B, const double *__restrict__ theta, double *__restrict__ u,
                         int *__restrict__ leaf)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    double x[N], al[N + 1];
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = theta[q * N + j];
    constexpr int REC = N + N * N;
    int cur = t.top;
    for (;;) {
        const int L = t.left[cur];
        if (L < 0) break;
        cur = loc_alpha<N>(t.cm + (int64_t)L * REC, x, al) ? L : t.right[cur];
    }
    loc_alpha<N>(t.cm + (int64_t)cur * REC, x, al);
    const double *U = t.vinput + (int64_t)cur * (N + 1) * t.nu;
    for (int k = 0; k < t.nu; ++k) {
        double v = al[0] * U[k];
#pragma unroll
        for (int i = 1; i <= N; ++i) v = fma(al[i], U[i * t.nu + k], v);
        u[q * t.nu + k] = v;
    }
    if (leaf) leaf[q] = cur;
}

}  // namespace ehm
