// Development aid: mma.sync.m8n8k4.f64 (DMMA) dependent latency and issue rate vs. plain DFMA outer products on one SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma_dep(double* out, long long* cyc, int n) {
    double c0 = out[0], c1 = out[1], a = out[2], b = out[3];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); }
    long long t1 = clock64();
    out[8 + threadIdx.x] = c0 + c1; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NI>
__global__ void k_dmma_ind(double* out, long long* cyc, int n) {
    double c0[NI], c1[NI], a[4], b[4];
#pragma unroll
    for (int k = 0; k < NI; ++k) { c0[k] = out[k]; c1[k] = out[k + 1]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] = out[2 + k]; b[k] = out[5 + k]; }
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < NI; ++k) dmma(c0[k], c1[k], a[k & 3], b[(k >> 2) & 3]);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < NI; ++k) s += c0[k] + c1[k];
    out[8 + threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 8 * 4096);
    long long c; const int n = 1000;
    k_dmma_dep<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA m8n8k4 dependent latency: %.1f cycles\n", (double)c / (4.0 * n));
    for (int threads : {32, 128, 256, 512}) {
        k_dmma_ind<16><<<1, threads>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("16 independent DMMA per warp, %3d threads: %.2f cycles per DMMA per warp (256 FMA each)\n", threads, (double)c / (16.0 * n));
    }
    return 0;
}
