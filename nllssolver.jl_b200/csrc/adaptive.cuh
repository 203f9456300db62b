// Adaptive robust residuals: r = mean - data under a ContaminatedGaussian kernel that is itself a variable
// (AbstractAdaptiveResidual: examples/adaptivekernel.jl:9-18, test/adaptivecost.jl:3-13; kernel src/robustadaptive.jl:3-33).
//
// Every residual touches the one kernel variable (3 DoF) and one scalar mean variable, so the Hessian is a small dense
// (3 + M) x (3 + M) matrix (the reference takes its dense path, src/linearsystem.jl:105-123) and linearisation is a pure
// reduction over the residuals.  Residuals are sorted by mean variable at prepare time and cut into chunks of one mean;
// one CTA reduces a chunk with a fixed tree, a single CTA then adds the chunk partials in order (deterministic).
//
// The kernel's gradient / Hessian with respect to (x1, x2, x3, cost) is what the reference obtains with nested ForwardDiff
// duals through update(kernel, x) (src/robust.jl:15, src/autodiff.jl:164-165): the same arithmetic is done here with an
// exact second-order forward-mode number (value, 4 partials, 10 second partials) per thread — no analytic shortcut, so
// that parity with the CPU restatement is a matter of rounding only.
#pragma once
#include <cfloat>
#include "common.cuh"

namespace nlls {

constexpr int AD_THREADS = 256;
constexpr int AD_NP = 16;        // per chunk: H_kk (6), g_k (3), H_mk (3), H_mm, g_m, cost, pad
constexpr int AD_MAXDOF = 16;    // 3 + M <= 16

struct Jet2 {                    // second-order forward-mode number in 4 variables; h = lower triangle (i >= j), index i (i + 1) / 2 + j
    double v, g[4], h[10];
};
__device__ __forceinline__ int jh(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
__device__ __forceinline__ Jet2 jconst(double x) {
    Jet2 r; r.v = x;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) r.h[i] = 0.0;
    return r;
}
__device__ __forceinline__ Jet2 jvar(double x, int k) { Jet2 r = jconst(x); r.g[k] = 1.0; return r; }
__device__ __forceinline__ Jet2 jadd(const Jet2& a, const Jet2& b) {
    Jet2 r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] + b.g[i];
#pragma unroll
    for (int i = 0; i < 10; ++i) r.h[i] = a.h[i] + b.h[i];
    return r;
}
__device__ __forceinline__ Jet2 jsub(const Jet2& a, const Jet2& b) {
    Jet2 r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] - b.g[i];
#pragma unroll
    for (int i = 0; i < 10; ++i) r.h[i] = a.h[i] - b.h[i];
    return r;
}
__device__ __forceinline__ Jet2 jmul(const Jet2& a, const Jet2& b) {
    Jet2 r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) r.h[jh(i, j)] = a.h[jh(i, j)] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[jh(i, j)];
    return r;
}
__device__ __forceinline__ Jet2 jexp(const Jet2& a) {
    Jet2 r; const double e = exp(a.v); r.v = e;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = e * a.g[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) r.h[jh(i, j)] = e * (a.h[jh(i, j)] + a.g[i] * a.g[j]);
    return r;
}
__device__ __forceinline__ Jet2 jlog(const Jet2& a) {
    Jet2 r; r.v = log(a.v); const double inv = 1.0 / a.v;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) r.h[jh(i, j)] = a.h[jh(i, j)] * inv - a.g[i] * a.g[j] * inv * inv;
    return r;
}
__device__ __forceinline__ Jet2 jinv(const Jet2& a) {
    Jet2 r; const double inv = 1.0 / a.v; r.v = inv;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.g[i] = -a.g[i] * inv * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) r.h[jh(i, j)] = -a.h[jh(i, j)] * inv * inv + 2 * a.g[i] * a.g[j] * inv * inv * inv;
    return r;
}

// The part of x -> robustify(update(kernel, x), cost + x4) that does not depend on the residual: the updated kernel
// parameters as functions of (x1, x2, x3)   src/variable.jl:18-32, src/robustadaptive.jl:12-22 (no re-sort under duals).
struct KernelJets { Jet2 wa, omwb, hd, hs2; };
__device__ __forceinline__ KernelJets kernel_jets(const double* k) {
    const Jet2 a = jmul(jconst(k[0] > 0 ? k[0] : DBL_MIN), jexp(jvar(0.0, 0)));
    const Jet2 b = jmul(jconst(k[1] > 0 ? k[1] : DBL_MIN), jexp(jvar(0.0, 1)));
    const Jet2 vv = jmul(jconst(k[2] > 0 ? k[2] : DBL_MIN), jexp(jvar(0.0, 2)));
    const Jet2 w = jmul(vv, jinv(jadd(jconst(1.0), jsub(vv, jconst(k[2])))));
    const Jet2 s1sq = jmul(a, a), s2sq = jmul(b, b);
    KernelJets r;
    r.hd = jmul(jconst(0.5), jsub(s2sq, s1sq));
    r.hs2 = jmul(jconst(0.5), s2sq);
    r.wa = jmul(w, a);
    r.omwb = jmul(jsub(jconst(1.0), w), b);
    return r;
}
// rho(cost + x4) with the kernel jets   src/robustadaptive.jl:25
__device__ __forceinline__ Jet2 cg_jet(const KernelJets& kj, double cost) {
    const Jet2 c = jadd(jconst(cost), jvar(0.0, 3));
    return jsub(jmul(c, kj.hs2), jlog(jadd(jmul(kj.wa, jexp(jmul(c, kj.hd))), kj.omwb)));
}
// robustify(kernel, cost)   src/robustadaptive.jl:25
__device__ __forceinline__ double cg_robustify(const double* k, double cost) {
    const double a = k[0], b = k[1], w = k[2];
    const double s1sq = a * a, s2sq = b * b;
    const double hd = 0.5 * (s2sq - s1sq), hs2 = 0.5 * s2sq;
    return cost * hs2 - log(w * a * exp(cost * hd) + (1 - w) * b);
}

struct AdaptDev {
    const double* data;     // residual data, sorted by mean variable
    const int4* chunks;     // (local mean index, first residual, end, 0)
    int nchunks, nmeans, dof;
    const int* moff;        // [nmeans] 0-based offset of each mean in the gradient / Hessian
    int koff;               // offset of the kernel variable's 3 DoF
    const unsigned char* fixdof;   // optimize!(problem, options, unfixed): per DoF 1 = FIXED (nullptr: none), frozen in place like the BA path
};

// K_A1  linearisation: per chunk the 15 sums of computerescostgradhess (src/residual.jl:57-111, adaptive branch :81-88,103-107)
__global__ void __launch_bounds__(AD_THREADS) adapt_lin_kernel(AdaptDev p, const double* __restrict__ kern, const double* __restrict__ means,
                                                               double* __restrict__ partials) {
    __shared__ double s_red[AD_THREADS / 32][AD_NP];
    const int4 ch = p.chunks[blockIdx.x];
    const double k3[3] = {kern[0], kern[1], kern[2]};
    const KernelJets kj = kernel_jets(k3);
    const double mean = means[ch.x];
    double acc[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = 0.0;
    for (int j = ch.y + threadIdx.x; j < ch.z; j += AD_THREADS) {
        const double r = mean - p.data[j];                      // computeresidual; J = 1   test/adaptivecost.jl:10-11
        const double s = r * r;                                  // sqnorm                   src/residual.jl:72
        const Jet2 rho = cg_jet(kj, s);                          // robustifydkernel         :81
        const double dc = rho.g[3], d2c = rho.h[jh(3, 3)];       //                          :82-83
        double gr = r, Hr = 1.0;                                 // g = J' r, H = J' J       :73-74
        const double dk0 = gr * rho.h[jh(3, 0)], dk1 = gr * rho.h[jh(3, 1)], dk2 = gr * rho.h[jh(3, 2)];   // :87
        if (dc != 1) Hr *= dc;                                   //                          :91-93
        if (d2c != 0) Hr += ((2 * d2c) * gr) * gr;               //                          :95-97
        if (dc != 1) gr *= dc;                                   //                          :99-101
        acc[0] += rho.h[jh(0, 0)]; acc[1] += rho.h[jh(1, 0)]; acc[2] += rho.h[jh(2, 0)];
        acc[3] += rho.h[jh(1, 1)]; acc[4] += rho.h[jh(2, 1)]; acc[5] += rho.h[jh(2, 2)];
        acc[6] += rho.g[0]; acc[7] += rho.g[1]; acc[8] += rho.g[2];
        acc[9] += dk0; acc[10] += dk1; acc[11] += dk2;
        acc[12] += Hr; acc[13] += gr;
        acc[14] += 0.5 * rho.v;                                  //                          :110
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 15; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) s_red[w][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 15) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < AD_THREADS / 32; ++k) t += s_red[k][threadIdx.x];
        partials[(size_t)blockIdx.x * AD_NP + threadIdx.x] = t;
    }
}

// K_A2  add the chunk partials in order and assemble the dense system (lower blocks mirrored like symmetrifyfull,
// src/BlockDenseMatrix.jl:24-34); H is dof x dof column-major, g has length dof.  Single CTA.
__global__ void __launch_bounds__(64) adapt_assemble_kernel(AdaptDev p, const double* __restrict__ partials, double* __restrict__ H, double* __restrict__ g,
                                                            double* __restrict__ cost_out) {
    __shared__ double s_k[10];   // kernel-only sums: H_kk (6), g_k (3), cost
    const int tid = threadIdx.x, d = p.dof;
    for (int i = tid; i < d * d; i += 64) H[i] = 0.0;
    __syncthreads();
    if (tid < 10) {               // totals over all chunks, in chunk order
        const int src = tid < 9 ? tid : 14;
        double t = 0.0;
        for (int c = 0; c < p.nchunks; ++c) t += partials[(size_t)c * AD_NP + src];
        s_k[tid] = t;
    }
    if (tid >= 16 && tid < 16 + 5) {   // per-mean sums, in chunk order
        const int e = tid - 16;        // 0..2 cross, 3 H_mm, 4 g_m
        for (int m = 0; m < p.nmeans; ++m) {
            double t = 0.0;
            for (int c = 0; c < p.nchunks; ++c) if (p.chunks[c].x == m) t += partials[(size_t)c * AD_NP + 9 + e];
            const int mo = p.moff[m];
            if (e < 3) { H[mo + d * (p.koff + e)] = t; H[(p.koff + e) + d * mo] = t; }
            else if (e == 3) H[mo + d * mo] = t;
            else g[mo] = t;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int ko = p.koff;
        const double hk[9] = {s_k[0], s_k[1], s_k[2], s_k[1], s_k[3], s_k[4], s_k[2], s_k[4], s_k[5]};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[(ko + a) + d * (ko + b)] = hk[a + 3 * b];
        for (int a = 0; a < 3; ++a) g[ko + a] = s_k[6 + a];
        *cost_out = s_k[9];
    }
}

// K_A3  cost: sum 0.5 rho(r^2) per chunk (fixed tree); the chunk partials are added by reduce_partials_kernel.
__global__ void __launch_bounds__(AD_THREADS) adapt_cost_kernel(AdaptDev p, const double* __restrict__ kern, const double* __restrict__ means,
                                                                double* __restrict__ cost_partials) {
    __shared__ double s_red[AD_THREADS / 32];
    const int4 ch = p.chunks[blockIdx.x];
    const double k3[3] = {kern[0], kern[1], kern[2]};
    const double mean = means[ch.x];
    double c = 0.0;
    for (int j = ch.y + threadIdx.x; j < ch.z; j += AD_THREADS) {
        const double r = mean - p.data[j];
        c += 0.5 * cg_robustify(k3, r * r);                      // src/residual.jl:49-55
    }
    const double t = block_sum(c, s_red);
    if (threadIdx.x == 0) cost_partials[blockIdx.x] = t;
}

// ZeroToInfScalar / ZeroToOneScalar updates   src/variable.jl:18-32
__device__ __forceinline__ double update_zerotoinf(double val, double x) { return (val > 0 ? val : DBL_MIN) * exp(x); }
__device__ __forceinline__ double update_zerotoone(double v, double x) {
    const double val = (v > 0 ? v : DBL_MIN) * exp(x);
    return val < __longlong_as_double(0x7ff0000000000000LL) ? val / (1 + (val - v)) : 1.0;
}

// K_A4  damped dense solve + update + step statistics (single thread; dof <= 16).
//   x = -(H + lambda I)^-1 g   : Cholesky, or LU with partial pivoting when the matrix is not positive definite (the reference
//   falls back to QR, src/linearsolver.jl:20-26 — the same solution);  varnext = update(variables, x)  src/linearsystem.jl:206-213;
//   out4 = {max|x|, sum x^2, x'Hx (undamped H, src/iterators.jl:162-163), g.x}
__global__ void adapt_solve_kernel(AdaptDev p, const double* __restrict__ H, const double* __restrict__ g, double lambda, const double* __restrict__ kern,
                                   const double* __restrict__ means, double* __restrict__ kern_next, double* __restrict__ means_next,
                                   double* __restrict__ x, double* __restrict__ out4) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int d = p.dof;
    double A[AD_MAXDOF * AD_MAXDOF], b[AD_MAXDOF], y[AD_MAXDOF];
    for (int i = 0; i < d * d; ++i) A[i] = H[i];
    for (int i = 0; i < d; ++i) { A[i + d * i] += lambda; b[i] = g[i]; }   // uniformscaling!  src/iterators.jl:149
    if (p.fixdof != nullptr)     // fixed variables: identity row / column and zero right-hand side — their step is exactly zero and the
        for (int i = 0; i < d; ++i) if (p.fixdof[i]) {   // others see the reference's reduced system (src/linearsystem.jl:93-102)
            for (int k = 0; k < d; ++k) { A[i + d * k] = 0.0; A[k + d * i] = 0.0; }
            A[i + d * i] = 1.0; b[i] = 0.0;
        }
    bool pd = true;
    {   // Cholesky A = L L' (lower, in place)
        double L[AD_MAXDOF * AD_MAXDOF];
        for (int i = 0; i < d * d; ++i) L[i] = A[i];
        for (int j = 0; j < d && pd; ++j) {
            double s = L[j + d * j];
            for (int k = 0; k < j; ++k) s -= L[j + d * k] * L[j + d * k];
            if (!(s > 0)) { pd = false; break; }
            const double ljj = sqrt(s);
            L[j + d * j] = ljj;
            for (int i = j + 1; i < d; ++i) {
                double t = L[i + d * j];
                for (int k = 0; k < j; ++k) t -= L[i + d * k] * L[j + d * k];
                L[i + d * j] = t / ljj;
            }
        }
        if (pd) {
            for (int i = 0; i < d; ++i) { double t = b[i]; for (int k = 0; k < i; ++k) t -= L[i + d * k] * y[k]; y[i] = t / L[i + d * i]; }
            for (int i = d - 1; i >= 0; --i) { double t = y[i]; for (int k = i + 1; k < d; ++k) t -= L[k + d * i] * b[k]; b[i] = t / L[i + d * i]; }
        }
    }
    if (!pd) {   // LU with partial pivoting
        for (int i = 0; i < d; ++i) b[i] = (p.fixdof != nullptr && p.fixdof[i]) ? 0.0 : g[i];
        for (int j = 0; j < d; ++j) {
            int piv = j; double best = fabs(A[j + d * j]);
            for (int i = j + 1; i < d; ++i) if (fabs(A[i + d * j]) > best) { best = fabs(A[i + d * j]); piv = i; }
            if (piv != j) { for (int k = 0; k < d; ++k) { const double t = A[j + d * k]; A[j + d * k] = A[piv + d * k]; A[piv + d * k] = t; } const double t = b[j]; b[j] = b[piv]; b[piv] = t; }
            const double inv = 1.0 / A[j + d * j];
            for (int i = j + 1; i < d; ++i) {
                const double l = A[i + d * j] * inv;
                for (int k = j + 1; k < d; ++k) A[i + d * k] -= l * A[j + d * k];
                b[i] -= l * b[j];
            }
        }
        for (int i = d - 1; i >= 0; --i) { double t = b[i]; for (int k = i + 1; k < d; ++k) t -= A[i + d * k] * b[k]; b[i] = t / A[i + d * i]; }
    }
    double mx = 0.0, sq = 0.0, xhx = 0.0, gx = 0.0;
    for (int i = 0; i < d; ++i) { const double xi = -b[i]; x[i] = xi; y[i] = xi; mx = nanmax(mx, fabs(xi)); sq += xi * xi; gx += g[i] * xi; }   // negate!  :152
    for (int j = 0; j < d; ++j) { double s = 0.0; for (int i = 0; i < d; ++i) s += H[i + d * j] * y[i]; xhx += y[j] * s; }
    out4[0] = mx; out4[1] = sq; out4[2] = xhx; out4[3] = gx;
    // update(ContaminatedGaussian, x): component-wise, then the constructor's re-sort (w is not flipped)  src/robustadaptive.jl:12-22
    double a = update_zerotoinf(kern[0], y[p.koff]), bb = update_zerotoinf(kern[1], y[p.koff + 1]);
    const double w = update_zerotoone(kern[2], y[p.koff + 2]);
    if (!(a >= bb)) { const double t = a; a = bb; bb = t; }
    // update! leaves the fixed entries of varnext alone (src/linearsystem.jl:206-213: only blockindices != 0), and the outer loop
    // swaps the two vectors — a fixed variable that a callback changed in varnext (the EM refit of test/adaptivecost.jl:19) is
    // therefore two iterations stale when it comes back.  Restated as is.
    if (p.fixdof == nullptr || !p.fixdof[p.koff]) { kern_next[0] = a; kern_next[1] = bb; kern_next[2] = w; }
    for (int m = 0; m < p.nmeans; ++m) if (p.fixdof == nullptr || !p.fixdof[p.moff[m]]) means_next[m] = means[m] + y[p.moff[m]];   // src/variable.jl:5
}

// max_i |H_ii| (initlambda, src/iterators.jl:131-137)
__global__ void adapt_maxdiag_kernel(const double* __restrict__ H, int d, double* out, const unsigned char* __restrict__ fixdof = nullptr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double m = 0.0;
    for (int i = 0; i < d; ++i) if (fixdof == nullptr || !fixdof[i]) m = nanmax(m, fabs(H[i + d * i]));
    *out = m;
}

// ---------------------------------------------------------------------------------------------------
// optimize(kernel::ContaminatedGaussian, squarederrors, maxiters)  (src/robustadaptive.jl:48-73): Expectation-Maximisation refit of
// the kernel from the squared residuals — what the EM callback of test/adaptivecost.jl:15-25 calls between Newton steps on the means.
// Per iteration: adapt_em_partial_kernel (one CTA per chunk: sum w err, sum w, sum err, count; squared errors are recomputed from
// the means, never stored), adapt_em_sums_kernel (the chunk partials in order -> 4 totals; all-reduced over the ranks by the host
// code), adapt_em_update_kernel (one thread: new parameters, the constructor's re-sort, isapprox stop -> `done` flag that turns the
// remaining iterations into no-ops, so the whole refit is enqueued without a host round trip).
// state[0..2] = oldparams (sigma1, sigma2, w), state[3] = done flag (0 / 1).
// ---------------------------------------------------------------------------------------------------
__global__ void adapt_em_init_kernel(const double* __restrict__ kern, double* __restrict__ state) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    state[0] = 1.0 / kern[0]; state[1] = 1.0 / kern[1]; state[2] = kern[2]; state[3] = 0.0;   // params(kernel)  :23,51
}
__global__ void __launch_bounds__(AD_THREADS) adapt_em_partial_kernel(AdaptDev p, const double* __restrict__ kern, const double* __restrict__ means,
                                                                      const double* __restrict__ state, double* __restrict__ partials) {
    __shared__ double s_red[AD_THREADS / 32];
    if (state[3] != 0.0) return;
    const int4 ch = p.chunks[blockIdx.x];
    const double is1 = kern[0], is2 = kern[1], w = kern[2];
    const double wratio = ((1 - w) * is2) / (is1 * w);                       // :53
    const double hd = -(0.5 * (is2 * is2 - is1 * is1));                      // :54
    const double mean = means[ch.x];
    double a = 0.0, b = 0.0, c = 0.0, n = 0.0;
    for (int j = ch.y + threadIdx.x; j < ch.z; j += AD_THREADS) {
        const double r = mean - p.data[j];
        const double err = r * r;
        const double wi = 1 / (1 + wratio * exp(hd * err));                  // :59
        a += wi * err; b += wi; c += err; n += 1.0;                          // :61-62, :50
    }
    a = block_sum(a, s_red); __syncthreads();
    b = block_sum(b, s_red); __syncthreads();
    c = block_sum(c, s_red); __syncthreads();
    n = block_sum(n, s_red);
    if (threadIdx.x == 0) { double* o = partials + (size_t)4 * blockIdx.x; o[0] = a; o[1] = b; o[2] = c; o[3] = n; }
}
__global__ void adapt_em_sums_kernel(const double* __restrict__ partials, int nchunks, const double* __restrict__ state, double* __restrict__ sums) {
    const int i = threadIdx.x;
    if (i >= 4 || state[3] != 0.0) return;
    double t = 0.0;
    for (int c = 0; c < nchunks; ++c) t += partials[(size_t)4 * c + i];
    sums[i] = t;
}
__global__ void adapt_em_update_kernel(const double* __restrict__ sums, double* __restrict__ kern, double* __restrict__ state) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || state[3] != 0.0) return;
    const double sigma1 = sums[0], tw = sums[1], total = sums[2], n = sums[3];
    const double np3[3] = {sqrt(sigma1 / tw), sqrt((total - sigma1) / (n - tw)), tw / n};     // :65
    double a = 1.0 / np3[0], b = 1.0 / np3[1];                                                // :66 (constructor :21, re-sort :13-15)
    if (!(a >= b)) { const double t = a; a = b; b = t; }
    kern[0] = a; kern[1] = b; kern[2] = np3[2];
    double dn = 0, no = 0, nn = 0;                                                            // :67 isapprox(oldparams, newparams; rtol = 1e-6)
    for (int i = 0; i < 3; ++i) { const double d = state[i] - np3[i]; dn += d * d; no += state[i] * state[i]; nn += np3[i] * np3[i]; }
    if (sqrt(dn) <= 1e-6 * fmax(sqrt(no), sqrt(nn))) state[3] = 1.0;
    for (int i = 0; i < 3; ++i) state[i] = np3[i];                                            // :70
}

}  // namespace nlls
