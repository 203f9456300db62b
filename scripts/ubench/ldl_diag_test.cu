// Development aid: runs ldl_diag_kernel on one random SPD tile, checks L D L' = A, L Linv = I and y = L^-1 b against a host
// computation and prints per-phase clock64() stamps (compile with -DLDL_PROFILE).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../nllssolver.jl_b200/csrc/reduced.cuh"
using namespace nlls;
int main(int argc, char** argv) {
    const int n = ST;
    const int NG = argc > 1 ? atoi(argv[1]) : 1;   // number of identical tiles factored concurrently (profiling aid)
    std::vector<double> A(n * n), b(n);
    srand(1);
    std::vector<double> G(n * n);
    for (auto& g : G) g = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += G[i + n * k] * G[j + n * k]; A[i + n * j] = s + (i == j ? 1.0 : 0.0); }
    for (int i = 0; i < n; ++i) b[i] = rand() / (double)RAND_MAX;
    double *dS, *dL, *dx; RedTask* dt; long long* dprof;
    cudaMalloc(&dS, 8 * n * n * NG); cudaMalloc(&dL, 8 * n * n * NG); cudaMalloc(&dx, 8 * n * NG); cudaMalloc(&dt, sizeof(RedTask) * NG); cudaMalloc(&dprof, 8 * 1024);
    cudaMemset(dL, 0, 8 * n * n * NG); cudaMemset(dprof, 0, 8 * 1024);
    std::vector<RedTask> tks(NG);
    for (int g = 0; g < NG; ++g) tks[g] = RedTask{g, g, g, g};
    cudaMemcpy(dt, tks.data(), sizeof(RedTask) * NG, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(ldl_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
    {   // latency of one tile with every SM busy on its own copy: 20 launches back to back between two events
        const int NREP = 20;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int g = 0; g < NG; ++g) { cudaMemcpy(dS + (size_t)g * n * n, A.data(), 8 * n * n, cudaMemcpyHostToDevice); cudaMemcpy(dx + (size_t)g * n, b.data(), 8 * n, cudaMemcpyHostToDevice); }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            for (int r = 0; r < NREP; ++r)
#ifdef LDL_PROFILE
                ldl_diag_kernel<<<NG, DIAG_THREADS, DIAG_SMEM>>>(dS, dL, dt, dx, dprof);
#else
                ldl_diag_kernel<<<NG, DIAG_THREADS, DIAG_SMEM>>>(dS, dL, dt, dx);
#endif
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("ablate %d: %d launches of %d tiles: %.2f us per launch\n", (int)LDL_ABLATE, NREP, NG, ms * 1e3 / NREP);
        }
    }
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        for (int g = 0; g < NG; ++g) { cudaMemcpy(dS + (size_t)g * n * n, A.data(), 8 * n * n, cudaMemcpyHostToDevice); cudaMemcpy(dx + (size_t)g * n, b.data(), 8 * n, cudaMemcpyHostToDevice); }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
#ifdef LDL_PROFILE
        cudaMemset(dprof, 0, 8 * 1024);
        ldl_diag_kernel<<<NG, DIAG_THREADS, DIAG_SMEM>>>(dS, dL, dt, dx, dprof);
#else
        ldl_diag_kernel<<<NG, DIAG_THREADS, DIAG_SMEM>>>(dS, dL, dt, dx);
#endif
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    printf("err %s  kernel %.2f us\n", cudaGetErrorString(cudaGetLastError()), best * 1e3);
    std::vector<double> T(n * n), Li(n * n), y(n);
    cudaMemcpy(T.data(), dS, 8 * n * n, cudaMemcpyDeviceToHost); cudaMemcpy(Li.data(), dL, 8 * n * n, cudaMemcpyDeviceToHost); cudaMemcpy(y.data(), dx, 8 * n, cudaMemcpyDeviceToHost);
    // checks
    double e1 = 0, e2 = 0, e3 = 0;
    auto L = [&](int i, int j) { return i == j ? 1.0 : (i > j ? T[i + n * j] : 0.0); };
    for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L(i, k) * T[k + n * k] * L(j, k); e1 = fmax(e1, fabs(s - A[i + n * j])); }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += L(i, k) * Li[k + n * j]; e2 = fmax(e2, fabs(s - (i == j))); }
    for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < n; ++k) s += L(i, k) * y[k]; e3 = fmax(e3, fabs(s - b[i])); }
    printf("max |LDL' - A| = %.3e   max |L Linv - I| = %.3e   max |L y - b| = %.3e\n", e1, e2, e3);
#ifdef LDL_PROFILE
    std::vector<long long> pr(1024); cudaMemcpy(pr.data(), dprof, 8 * 1024, cudaMemcpyDeviceToHost);
    printf("prologue %lld  loop %lld  epilogue %lld (cycles)\n", pr[1] - pr[0], pr[2] - pr[1], pr[3] - pr[2]);
    for (int w = 0; w < 8; ++w) printf("warp %d: A2 %lld  sync %lld  tasks (or A1) %lld  sync %lld  [warp 0: look-ahead update %lld  8x8 factor %lld]\n", w, pr[64 + 8 * w], pr[65 + 8 * w], pr[66 + 8 * w], pr[67 + 8 * w], pr[68 + 8 * w], pr[69 + 8 * w]);
#endif
    return 0;
}
