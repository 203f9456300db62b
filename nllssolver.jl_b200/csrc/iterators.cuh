// Device pieces of the Dogleg and gradient-descent iterators (src/iterators.jl:30-115,177-208): they run on the same six
// operations as Levenberg-Marquardt plus a few vector operations on linsystem.x / b ([cameras | points] layout) and g' H g.
#pragma once
#include "common.cuh"
#include "residuals.cuh"

namespace nlls {

// partial sums of a . b (one per CTA; summed in order by reduce_partials_kernel)
__global__ void __launch_bounds__(256) dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ partials) {
    __shared__ double s_red[8];
    double v = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) v += a[i] * b[i];
    v = block_sum(v, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
}
// x = sx * x + sy * y   (sx == 0: x is not read)
__global__ void axpby_kernel(double* __restrict__ x, double sx, const double* __restrict__ y, double sy, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (sx == 0.0) ? sy * y[i] : sx * x[i] + sy * y[i];
}

// v' H v with the undamped block-sparse H (fast_bAb, src/utils.jl:71-106), v in [cameras | points] layout:
//   sum_c v_c' U_c v_c  +  sum_p ( v_p' V_p v_p + 2 v_p' sum_c W_pc v_c )
// one thread per camera, then one thread per point; per CTA partials.
template <int DC>
__global__ void __launch_bounds__(128) quadform_kernel(DevProblem p, const double* __restrict__ v, double* __restrict__ partials) {
    __shared__ double s_red[4];
    const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
    double acc = 0.0;
    if (idx < p.nA) {
        const double* U = p.H + (size_t)DC * DC * idx;
        const double* vc = v + (size_t)DC * idx;
        for (int b = 0; b < DC; ++b) {
            double s = 0.0;
            for (int a = 0; a < DC; ++a) s += U[a + DC * b] * vc[a];
            acc += vc[b] * s;
        }
    } else if (idx < (long long)p.nA + p.nB) {
        const long long pt = idx - p.nA;
        const int ob0 = p.obs_start[pt], ob1 = p.obs_start[pt + 1];
        const double* row = p.H + (size_t)p.hB + (size_t)3 * DC * ob0 + (size_t)9 * pt;
        const double* vp = v + p.gB + (size_t)3 * pt;
        const double x0 = vp[0], x1 = vp[1], x2 = vp[2];
        double u0 = 0, u1 = 0, u2 = 0;
        for (int j = ob0; j < ob1; ++j) {
            const double* W = row + (size_t)3 * DC * (j - ob0);
            const double* vc = v + (size_t)DC * p.obs_cam[j];
            for (int a = 0; a < DC; ++a) { u0 += W[3 * a] * vc[a]; u1 += W[3 * a + 1] * vc[a]; u2 += W[3 * a + 2] * vc[a]; }
        }
        const double* V = row + (size_t)3 * DC * (ob1 - ob0);
        const double w0 = V[0] * x0 + V[3] * x1 + V[6] * x2, w1 = V[1] * x0 + V[4] * x1 + V[7] * x2, w2 = V[2] * x0 + V[5] * x1 + V[8] * x2;
        acc = (x0 * w0 + x1 * w1 + x2 * w2) + 2.0 * (x0 * u0 + x1 * u1 + x2 * u2);
    }
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// update!(varnext, variables, linsystem) for an arbitrary step x (src/linearsystem.jl:206-213) + max|x| / sum x^2 partials per CTA
// (partials[b] = max, partials[gridDim.x + b] = sum of squares).  One thread per variable.
template <class R>
__global__ void __launch_bounds__(128) apply_step_kernel(DevProblem p, const double* __restrict__ x, const double* __restrict__ cams, const double* __restrict__ pts,
                                                         double* __restrict__ cams_next, double* __restrict__ pts_next, double* __restrict__ partials) {
    constexpr int DC = R::DC;
    __shared__ double s_red[8];
    const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
    double mx = 0.0, sq = 0.0;
    if (idx < p.nA) {
        double xc[DC];
        for (int a = 0; a < DC; ++a) { xc[a] = x[(size_t)DC * idx + a]; mx = nanmax(mx, fabs(xc[a])); sq += xc[a] * xc[a]; }
        R::update_cam(cams + (size_t)idx * R::CS, xc, cams_next + (size_t)idx * R::CS);
    } else if (idx < (long long)p.nA + p.nB) {
        const long long pt = idx - p.nA;
        for (int b = 0; b < 3; ++b) {
            const double xv = x[p.gB + (size_t)3 * pt + b];
            pts_next[(size_t)3 * pt + b] = pts[(size_t)3 * pt + b] + xv;
            mx = nanmax(mx, fabs(xv)); sq += xv * xv;
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    mx = warp_nanmax(mx); sq = warp_sum(sq);
    if (lane == 0) { s_red[w] = mx; s_red[4 + w] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 4; ++i) { a = nanmax(a, s_red[i]); b += s_red[4 + i]; }
        partials[blockIdx.x] = a; partials[gridDim.x + blockIdx.x] = b;
    }
}

}  // namespace nlls
