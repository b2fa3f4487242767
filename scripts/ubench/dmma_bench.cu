// Micro-benchmark (B200, sm_100a): throughput of mma.sync.m8n8k4.f64 (DMMA) vs DFMA, and whether
// they overlap.  Decides how ipm_solve builds its normal matrix (profiles/README.md, round 2).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC, int NFMA>
__global__ void k_mix(double *out, int iters)
{
    double c[NACC > 0 ? NACC : 1][2], f[NFMA > 0 ? NFMA : 1];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < NFMA; ++i) f[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NFMA; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chain latency
__global__ void k_lat(double *out, int iters, int mode)
{
    double c0 = 0, c1 = 0, a = 1.0 + threadIdx.x * 1e-9, b = 1.0;
    long long t0 = clock64();
    if (mode == 0) for (int it = 0; it < iters; ++it) dmma(c0, c1, a, b);
    else for (int it = 0; it < iters; ++it) c0 = fma(c0, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) out[64] = double(t1 - t0) / iters;
}

// 64-bit shuffle and LDS throughput (warp instructions per clock per SM)
__global__ void k_shfl(double *out, int iters)
{
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
    for (int it = 0; it < iters; ++it) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, 4); v3 += __shfl_xor_sync(0xffffffffu, v3, 8);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    return best;
}

template <int NACC, int NFMA>
void run(double *d, int blocks, int threads, int iters, const char *name)
{
    float ms = timeit([&] { k_mix<NACC, NFMA><<<blocks, threads>>>(d, iters); });
    const double warps = (double)blocks * threads / 32;
    const double mma_flop = warps * iters * NACC * 512.0, fma_flop = (double)blocks * threads * iters * NFMA * 2.0;
    printf("%-28s blocks %d thr %d: %.3f ms  dmma %.2f TF/s  dfma %.2f TF/s  total %.2f\n", name, blocks, threads, ms,
           mma_flop / ms / 1e9, fma_flop / ms / 1e9, (mma_flop + fma_flop) / ms / 1e9);
}

int main()
{
    double *d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
    int sms = 148;
    const int it = 1 << 14;
    for (int thr : {128, 256, 512, 1024}) {
        run<8, 0>(d, sms * 1, thr, it, "dmma x8");
        run<0, 8>(d, sms * 1, thr, it, "dfma x8");
        run<4, 4>(d, sms * 1, thr, it, "dmma x4 + dfma x4");
        run<2, 8>(d, sms * 1, thr, it, "dmma x2 + dfma x8");
        run<1, 16>(d, sms * 1, thr, it, "dmma x1 + dfma x16");
    }
    k_lat<<<1, 32>>>(d, 4096, 0); cudaDeviceSynchronize();
    double h[65]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("dmma dependent latency %.1f cycles\n", h[64]);
    k_lat<<<1, 32>>>(d, 4096, 1); cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("dfma dependent latency %.1f cycles\n", h[64]);
    for (int thr : {128, 512}) {
        float ms = timeit([&] { k_shfl<<<sms, thr>>>(d, it); });
        printf("shfl64 thr %d: %.3f ms -> %.3f 64-bit shuffles/clk/SM (at 1.9 GHz)\n", thr, ms,
               (double)thr / 32 * it * 4 / (ms * 1e-3 * 1.9e9));
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
