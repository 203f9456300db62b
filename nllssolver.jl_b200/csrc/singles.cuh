// optimizesingles!(problem, options, <point type>)  (src/optimize.jl:60-76,183-205): every point on its own, cameras fixed, over the
// costs that depend on it.  The sub-problems are independent 3-DoF Levenberg-Marquardt solves (UniVariateLSstatic + LevMarData): one
// thread per point runs the reference's whole outer loop — optimizeinternal! (src/optimize.jl:109-180) around iterate!(::LevMarData)
// (src/iterators.jl:139-172) — with the static solve of src/linearsolver.jl:7-18 (Cholesky, general solve if not positive definite).
#pragma once
#include "common.cuh"
#include "residuals.cuh"

namespace nlls {

struct SinglesOpts {
    double reldcost, absdcost, dstep;
    long long maxfails, maxiters;
    int timeup;   // maxtime == 0: bit 9 is set after the first iteration (test/functional.jl:51-54)
};

// x = A^-1 b for a symmetric 3 x 3 matrix (lower triangle a00 a10 a20 a11 a21 a22): Cholesky when positive definite
__device__ __forceinline__ void solve_sym3(const double a[6], const double b[3], double x[3]) {
    const double l00s = a[0];
    bool pd = l00s > 0.0;
    double l00 = sqrt(l00s), l10 = a[1] / l00, l20 = a[2] / l00;
    const double l11s = a[3] - l10 * l10;
    pd = pd && l11s > 0.0;
    double l11 = sqrt(l11s), l21 = (a[4] - l20 * l10) / l11;
    const double l22s = a[5] - l20 * l20 - l21 * l21;
    pd = pd && l22s > 0.0;
    if (pd) {
        const double l22 = sqrt(l22s);
        const double y0 = b[0] / l00, y1 = (b[1] - l10 * y0) / l11, y2 = (b[2] - l20 * y0 - l21 * y1) / l22;
        x[2] = y2 / l22;
        x[1] = (y1 - l21 * x[2]) / l11;
        x[0] = (y0 - l10 * x[1] - l20 * x[2]) / l00;
    } else {
        double inv[6];
        inv_sym3(a, inv);
        x[0] = inv[0] * b[0] + inv[1] * b[1] + inv[2] * b[2];
        x[1] = inv[1] * b[0] + inv[3] * b[1] + inv[4] * b[2];
        x[2] = inv[2] * b[0] + inv[4] * b[1] + inv[5] * b[2];
    }
}

template <class R, bool MS = false>
__global__ void __launch_bounds__(128) singles_point_kernel(DevProblem p, const double* __restrict__ cams, double* __restrict__ pts, SinglesOpts o,
                                                            unsigned long long* __restrict__ iters) {
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= p.nB) return;
    const int ob0 = p.obs_start[pt], ob1 = p.obs_start[pt + 1];
    if (ob1 == ob0) return;                                       // no cost depends on this variable: nothing to optimise
    double X[3] = {pts[(size_t)3 * pt], pts[(size_t)3 * pt + 1], pts[(size_t)3 * pt + 2]};
    double H[6], g[3];
    // costgradhess! over the point's costs, point block only (cameras fixed: varflags = point bit)   src/cost.jl:29-52
    auto linearize = [&](const double Xc[3]) {
        double c = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) H[i] = 0.0;
        g[0] = g[1] = g[2] = 0.0;
        for (int j = ob0; j < ob1; ++j) {
            double cv[R::NC];
            R::load_cam(cams, p.obs_cam[j], cv);
            const double2 z = p.obs_z[j];
            double r[2], Jc[2][R::DC], Jp[2][3];
            R::resjac(cv, Xc, z.x, z.y, r, Jc, Jp);
            const double s = r[0] * r[0] + r[1] * r[1];
            double rho, d1, d2;
            robustifydcost(rk_point<MS>(p, j), s, rho, d1, d2);
            c += 0.5 * rho;
            double gp[3];
#pragma unroll
            for (int b = 0; b < 3; ++b) gp[b] = fma(Jp[1][b], r[1], Jp[0][b] * r[0]);
            int q = 0;
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2)
#pragma unroll
                for (int b = b2; b < 3; ++b) {
                    double h = fma(Jp[1][b], Jp[1][b2], Jp[0][b] * Jp[0][b2]);
                    if (d1 != 1.0) h *= d1;
                    if (d2 != 0.0) h = fma((2 * d2) * gp[b], gp[b2], h);
                    H[q++] += h;
                }
#pragma unroll
            for (int b = 0; b < 3; ++b) g[b] += (d1 != 1.0) ? gp[b] * d1 : gp[b];
        }
        return c;
    };
    auto cost_at = [&](const double Xc[3]) {
        double c = 0.0;
        for (int j = ob0; j < ob1; ++j) {
            double cv[R::NC];
            R::load_cam(cams, p.obs_cam[j], cv);
            const double2 z = p.obs_z[j];
            double r[2];
            R::residual(cv, Xc, z.x, z.y, r);
            c += 0.5 * robustify(rk_point<MS>(p, j), r[0] * r[0] + r[1] * r[1]);
        }
        return c;
    };
    double cost = linearize(X);                                   // src/optimize.jl:118
    double bestcost = cost, lambda = 0.0;
    double Xbest[3] = {X[0], X[1], X[2]};
    long long fails = 0, iternum = 0;
    while (true) {
        iternum += 1;
        // ---- iterate!(::LevMarData)                               src/iterators.jl:139-172
        if (lambda == 0) lambda = 1e-6 * fmax(fabs(H[0]), fmax(fabs(H[3]), fabs(H[5])));   // H lower triangle: a00 a10 a20 a11 a21 a22
        double mu = 2.0, cost_, maxstep, Xn[3];
        while (true) {
            const double A[6] = {H[0] + lambda, H[1], H[2], H[3] + lambda, H[4], H[5] + lambda};
            double x[3];
            solve_sym3(A, g, x);
            x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2];
            Xn[0] = X[0] + x[0]; Xn[1] = X[1] + x[1]; Xn[2] = X[2] + x[2];
            cost_ = cost_at(Xn);
            maxstep = nanmax(nanmax(fabs(x[0]), fabs(x[1])), fabs(x[2]));
            if (!(cost_ > bestcost) || maxstep < o.dstep) {
                const double hx0 = H[0] * x[0] + H[1] * x[1] + H[2] * x[2], hx1 = H[1] * x[0] + H[3] * x[1] + H[4] * x[2],
                             hx2 = H[2] * x[0] + H[4] * x[1] + H[5] * x[2];
                const double xhx = x[0] * hx0 + x[1] * hx1 + x[2] * hx2, gx = g[0] * x[0] + g[1] * x[1] + g[2] * x[2];
                const double q = (cost_ - bestcost) / (0.5 * xhx + gx);
                const double t = 2 * q - 1;
                lambda *= q < 0.983 ? 1 - t * t * t : 0.1;
                break;
            }
            lambda *= mu;
            mu *= 2.0;
        }
        cost = cost_;
        // ---- optimizeinternal! bookkeeping                         src/optimize.jl:130-165
        double dcost = bestcost - cost;
        if (dcost >= 0) { bestcost = cost; fails = 0; }
        else {
            dcost = cost;
            fails += 1;
            if (fails == 1) { Xbest[0] = X[0]; Xbest[1] = X[1]; Xbest[2] = X[2]; }
        }
        X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2];
        int conv = 0;
        conv |= (int)isinf(cost) << 0;
        conv |= (int)isnan(cost) << 1;
        conv |= (int)(dcost < bestcost * o.reldcost) << 2;
        conv |= (int)(dcost < o.absdcost) << 3;
        conv |= (int)isinf(maxstep) << 4;
        conv |= (int)isnan(maxstep) << 5;
        conv |= (int)(maxstep < o.dstep) << 6;
        conv |= (int)(fails > o.maxfails) << 7;
        conv |= (int)(iternum >= o.maxiters) << 8;
        conv |= o.timeup << 9;
        if (conv) break;
        linearize(X);
    }
    if (!(bestcost >= cost)) { X[0] = Xbest[0]; X[1] = Xbest[1]; X[2] = Xbest[2]; }
    pts[(size_t)3 * pt] = X[0]; pts[(size_t)3 * pt + 1] = X[1]; pts[(size_t)3 * pt + 2] = X[2];
    atomicAdd(iters, (unsigned long long)iternum);
}

}  // namespace nlls
