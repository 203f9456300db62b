// Micro-benchmarks (development aid): FP64 latency / issue rate, reciprocal chain, barrier and LDS latency on one SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_fast(double d) {
    double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0); x = fma(x, e, x); e = fma(-d, x, 1.0); x = fma(x, e, x); return x;
}
__global__ void k_dep(double* out, long long* cyc, int n) {
    double a = out[0], b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, b, b); a = fma(a, b, b); a = fma(a, b, b); a = fma(a, b, b); }
    long long t1 = clock64();
    out[2] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_indep(double* out, long long* cyc, int n) {
    double a[36]; double b = out[1];
#pragma unroll
    for (int k = 0; k < 36; ++k) a[k] = out[0] + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 36; ++k) a[k] = fma(a[k], b, b);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 36; ++k) s += a[k];
    out[2 + threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc, int n) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = rcp_fast(a); }
    long long t1 = clock64();
    out[2] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_div(double* out, long long* cyc, int n) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = 1.0 / a; }
    long long t1 = clock64();
    out[2] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { p = s[p]; }
    long long t1 = clock64();
    out[2 + threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_stsbarlds(double* out, long long* cyc, int n) {   // publish -> barrier -> read: the per-pivot handshake
    __shared__ double s[2][256];
    double v = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { s[i & 1][threadIdx.x] = v; __syncthreads(); v = s[i & 1][(threadIdx.x + 1) % blockDim.x] + 1.0; }
    long long t1 = clock64();
    out[2 + threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 64);
    double h[2] = {1.000001, 0.999999}; cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
    long long c; const int n = 1000;
    auto rep = [&](const char* name, double per) { cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-44s %8.1f cycles\n", name, (double)c / per); };
    k_dep<<<1, 32>>>(out, cyc, n); rep("DFMA dependent latency (1 warp)", 4.0 * n);
    for (int w : {32, 64, 128, 160, 256, 512}) { k_indep<<<1, w>>>(out, cyc, n); char b[64]; sprintf(b, "36 indep DFMA / thread, %d threads: per 36", w); rep(b, n); }
    k_rcp<<<1, 32>>>(out, cyc, n); rep("rcp_fast chain", n);
    k_div<<<1, 32>>>(out, cyc, n); rep("1.0/x chain", n);
    for (int w : {32, 96, 160, 256}) { k_bar<<<1, w>>>(out, cyc, n); char b[64]; sprintf(b, "__syncthreads, %d threads", w); rep(b, n); }
    k_lds<<<1, 32>>>(out, cyc, n); rep("LDS dependent latency", n);
    for (int w : {96, 160, 224}) { k_stsbarlds<<<1, w>>>(out, cyc, n); char b[64]; sprintf(b, "STS+BAR+LDS+DADD round, %d threads", w); rep(b, n); }
    return 0;
}
