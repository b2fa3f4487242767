// ehm_common.cuh -- host-side problem object shared by the translation units of libehm_b200.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/ehm.h"
#include "ehm_solve.cuh"

namespace ehm {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define EHM_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return ::ehm::fail(EHM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// device work counters shared by all kernels of one problem
struct DevCounters {
    unsigned long long work;    // dynamic work-distribution cursor (reset before each launch)
    unsigned long long iters;   // IPM iterations
    unsigned long long solves;  // IPM solves
    unsigned long long fails;   // numerical failures
    double flops[8];            // algorithmic flops per pipeline stage (0 = batch API)
    int last_fail[4];           // {stage, candidate, node, iterations} of the most recent failure
};

}  // namespace ehm

struct ehm_problem {
    ehm::DevTables tb;
    int cfg;  // index of the compiled <U,NX,MR> instantiation
    double *dG = nullptr, *dS = nullptr, *dg = nullptr, *drho = nullptr, *dM = nullptr;
    ehm::DevCounters *dcnt = nullptr;
    int sm_count = 148;
    void *parked = nullptr;   // one retired ehm_partition whose device buffers are reused
    void *oscr = nullptr;     // ehm::OracleScratch of the batched oracle calls (grow-only)
    std::vector<double> u_dummy;
};

