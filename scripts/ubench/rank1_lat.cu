// Development aid: cost of one 6x6 rank-1 register update per thread (the pivot step of ldl_diag_kernel) for one warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_rank1(double* out, long long* cyc, int n, int mode) {
    __shared__ __align__(16) double col[2][72];
    for (int i = threadIdx.x; i < 144; i += blockDim.x) (&col[0][0])[i] = 1e-3 * (i + 1);
    __syncthreads();
    double B[6][6];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) B[a][b] = out[a * 6 + b];
    const int ty = (threadIdx.x & 31) % 12, tx = (threadIdx.x & 31) % 7;
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
        const int buf = it & 1;
        double li[6], ck[6];
        if (mode >= 1) {
            const double rd = col[buf][71];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double2 q = reinterpret_cast<const double2*>(&col[buf][6 * ty])[a];
                li[2 * a] = q.x * rd; li[2 * a + 1] = q.y * rd;
                const double2 r2 = reinterpret_cast<const double2*>(&col[buf][6 * tx])[a];
                ck[2 * a] = r2.x; ck[2 * a + 1] = r2.y;
            }
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) { li[a] = B[a][0] * 1e-9; ck[a] = B[0][a] * 1e-9; }
        }
#pragma unroll
        for (int b = 0; b < 6; ++b)
#pragma unroll
            for (int a = 0; a < 6; ++a) B[a][b] = fma(-li[a], ck[b], B[a][b]);
        if (mode >= 2) {
            double2* dst = reinterpret_cast<double2*>(&col[buf ^ 1][6 * ty]);
#pragma unroll
            for (int a = 0; a < 3; ++a) dst[a] = make_double2(B[2 * a][it % 6 == 0 ? 1 : 2], B[2 * a + 1][1]);
        }
        if (mode >= 3) __syncthreads();
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) s += B[a][b];
    out[64 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 8 * 4096);
    long long c; const int n = 720;
    for (int threads : {32, 128, 224})
        for (int mode = 0; mode < 4; ++mode) {
            k_rank1<<<1, threads>>>(out, cyc, n, mode); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("threads %3d mode %d (0 regs only, 1 +LDS/DMUL, 2 +STS publish, 3 +barrier): %7.1f cycles / step\n", threads, mode, (double)c / n);
        }
    return 0;
}
