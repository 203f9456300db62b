// ORACLE — TEST INFRASTRUCTURE ONLY (see nlls_oracle.hpp header comment).
// CPU restatement of the NLLSsolver.jl LM hot path; sequential, single thread, reference
// summation order.  Each function cites the reference file:line it restates.
#include "nlls_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

namespace orc {

static inline uint64_t time_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

// =======================================================================================
// Robust kernels                                                        src/robust.jl
// =======================================================================================
static double robustify_inner(const RobustSpec& k, double c) {
    switch (k.kind) {
        case RK_NONE: return c;                                            // :11
        case RK_HUBER:
        case RK_HUBER2O: {                                                 // :47
            double w2 = k.width * k.width;
            return c < w2 ? c : std::sqrt(c) * (k.width * 2) - w2;
        }
        case RK_GEMANMCCLURE: {                                            // :72
            double w2 = k.width * k.width;
            return c * w2 / (c + w2);
        }
    }
    return std::numeric_limits<double>::quiet_NaN();
}
double robustify(const RobustSpec& k, double c) {
    double r = robustify_inner(k, c);
    return k.scaled ? r * k.height : r;                                    // :26
}
void robustifydcost(const RobustSpec& k, double c, double& rho, double& d1, double& d2) {
    switch (k.kind) {
        case RK_NONE: rho = c; d1 = 1.0; d2 = 0.0; break;                  // :12
        case RK_HUBER:
        case RK_HUBER2O: {                                                 // :48-55
            double w2 = k.width * k.width;
            if (c < w2) { rho = c; d1 = 1.0; d2 = 0.0; break; }
            double sq = std::sqrt(c);
            rho = sq * (k.width * 2) - w2;
            d1 = k.width / sq;
            d2 = (k.kind == RK_HUBER2O) ? (-0.5 * k.width) / (c * sq) : 0.0;
            break;
        }
        case RK_GEMANMCCLURE: {                                            // :73-77
            double w2 = k.width * k.width;
            double r = 1.0 / (c + w2);
            double w = w2 * r;
            double ww = w * w;
            rho = c * w; d1 = ww; d2 = -2 * ww * r;
            break;
        }
        default: rho = d1 = d2 = std::numeric_limits<double>::quiet_NaN();
    }
    if (k.scaled) { rho *= k.height; d1 *= k.height; d2 *= k.height; }      // :28-31
}

// =======================================================================================
// Forward-mode autodiff numbers (stand-in for ForwardDiff, src/autodiff.jl)
// =======================================================================================
// First order, up to NP partials.
constexpr int NPMAX = 12;
struct Dual {
    double v = 0;
    double d[NPMAX] = {0};
    Dual() {}
    Dual(double x) : v(x) {}
};
static inline Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v; for (int i = 0; i < NPMAX; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
static inline Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v; for (int i = 0; i < NPMAX; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
static inline Dual operator-(const Dual& a) { Dual r; r.v = -a.v; for (int i = 0; i < NPMAX; ++i) r.d[i] = -a.d[i]; return r; }
static inline Dual operator*(const Dual& a, const Dual& b) { Dual r; r.v = a.v * b.v; for (int i = 0; i < NPMAX; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
static inline Dual operator/(const Dual& a, const Dual& b) {
    Dual r; double inv = 1.0 / b.v; r.v = a.v * inv;
    for (int i = 0; i < NPMAX; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}

// Second order in 4 variables (nested duals in the reference: ForwardDiff.hessian!, src/autodiff.jl:126-130)
struct Jet2 {
    double v = 0, g[4] = {0, 0, 0, 0}, h[16] = {0};
    Jet2() {}
    Jet2(double x) : v(x) {}
};
static Jet2 jvar(double x, int i) { Jet2 r(x); r.g[i] = 1.0; return r; }
static Jet2 operator+(const Jet2& a, const Jet2& b) { Jet2 r; r.v = a.v + b.v; for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] + b.g[i]; for (int i = 0; i < 16; ++i) r.h[i] = a.h[i] + b.h[i]; return r; }
static Jet2 operator-(const Jet2& a, const Jet2& b) { Jet2 r; r.v = a.v - b.v; for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] - b.g[i]; for (int i = 0; i < 16; ++i) r.h[i] = a.h[i] - b.h[i]; return r; }
static Jet2 operator*(const Jet2& a, const Jet2& b) {
    Jet2 r; r.v = a.v * b.v;
    for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
        r.h[i + 4 * j] = a.h[i + 4 * j] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[i + 4 * j];
    return r;
}
static Jet2 jexp(const Jet2& a) {
    Jet2 r; double e = std::exp(a.v); r.v = e;
    for (int i = 0; i < 4; ++i) r.g[i] = e * a.g[i];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.h[i + 4 * j] = e * (a.h[i + 4 * j] + a.g[i] * a.g[j]);
    return r;
}
static Jet2 jlog(const Jet2& a) {
    Jet2 r; r.v = std::log(a.v); double inv = 1.0 / a.v;
    for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] * inv;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.h[i + 4 * j] = a.h[i + 4 * j] * inv - a.g[i] * a.g[j] * inv * inv;
    return r;
}
static Jet2 jinv(const Jet2& a) {
    Jet2 r; double inv = 1.0 / a.v; r.v = inv;
    for (int i = 0; i < 4; ++i) r.g[i] = -a.g[i] * inv * inv;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
        r.h[i + 4 * j] = -a.h[i + 4 * j] * inv * inv + 2 * a.g[i] * a.g[j] * inv * inv * inv;
    return r;
}

// =======================================================================================
// Variables                                              src/variable.jl, src/robustadaptive.jl
// =======================================================================================
static const double FLOATMIN = std::numeric_limits<double>::min();

static double update_zerotoinf(double val, double x) {                      // src/variable.jl:22
    return (val > 0 ? val : FLOATMIN) * std::exp(x);
}
static double update_zerotoone(double v, double x) {                        // src/variable.jl:29-32
    double val = (v > 0 ? v : FLOATMIN) * std::exp(x);
    return val < std::numeric_limits<double>::infinity() ? val / (1 + (val - v)) : 1.0;
}

Variable make_contaminated_gaussian(double s1, double s2, double w) {       // src/robustadaptive.jl:12-20
    Variable k;
    k.type = VT_CONTAMGAUSS; k.nstore = 3; k.ndof = 3;
    double a = 1.0 / s1, b = 1.0 / s2;
    if (!(a >= b)) std::swap(a, b);   // narrowest Gaussian first; w is NOT changed
    k.v[0] = a; k.v[1] = b; k.v[2] = w;
    return k;
}

static void so3_exp(const double* w, double* E) {  // Rodrigues formula, column-major 3x3
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = std::sqrt(th2);
    double A, B;
    if (th < 1e-5) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
    else { A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; }
    double K[9] = {0, w[2], -w[1], -w[2], 0, w[0], w[1], -w[0], 0};  // [w]x column-major
    double K2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += K[i + 3 * k] * K[k + 3 * j];
        K2[i + 3 * j] = s;
    }
    for (int i = 0; i < 9; ++i) E[i] = A * K[i] + B * K2[i];
    E[0] += 1; E[4] += 1; E[8] += 1;
}

Variable make_pinhole(const double* rod, const double* t, double f, double k1, double k2) {
    Variable c;
    c.type = VT_PINHOLE; c.nstore = 15; c.ndof = 9;
    so3_exp(rod, c.v);
    c.v[9] = t[0]; c.v[10] = t[1]; c.v[11] = t[2];
    c.v[12] = f; c.v[13] = k1; c.v[14] = k2;
    return c;
}

// optimize(kernel::ContaminatedGaussian{T}, squarederrors, maxiters=10)            src/robustadaptive.jl:48-73
void em_optimize(double k[3], const double* sq, int64_t n, int maxiters) {
    double total = 0;                                                        // :50  totalsquarederror = sum(squarederrors)
    for (int64_t i = 0; i < n; ++i) total += sq[i];
    double oldp[3] = {1.0 / k[0], 1.0 / k[1], k[2]};                         // :51  params(kernel)  (src/robustadaptive.jl:23)
    for (int iter = 0; iter < maxiters; ++iter) {                            // :52
        const double is1 = k[0], is2 = k[1], w = k[2];
        const double wratio = ((1 - w) * is2) / (is1 * w);                   // :53
        const double halfs1sqminuss2sq = -(0.5 * (is2 * is2 - is1 * is1));   // :54  -kernel.halfs2sqminuss1sq  (:19)
        double sigma1 = 0, totalweight = 0;
        for (int64_t i = 0; i < n; ++i) {                                    // :57-63
            const double wi = 1 / (1 + wratio * std::exp(halfs1sqminuss2sq * sq[i]));
            sigma1 += wi * sq[i];
            totalweight += wi;
        }
        const double newp[3] = {std::sqrt(sigma1 / totalweight), std::sqrt((total - sigma1) / ((double)n - totalweight)), totalweight / (double)n};   // :65
        double a = 1.0 / newp[0], b = 1.0 / newp[1];                         // :66  ContaminatedGaussian(newparams...)  (:21, re-sort :13-15)
        if (!(a >= b)) std::swap(a, b);
        k[0] = a; k[1] = b; k[2] = newp[2];
        double dn = 0, no = 0, nn = 0;                                       // :67  isapprox(oldparams, newparams; rtol = 1e-6)
        for (int i = 0; i < 3; ++i) { dn += (oldp[i] - newp[i]) * (oldp[i] - newp[i]); no += oldp[i] * oldp[i]; nn += newp[i] * newp[i]; }
        if (std::sqrt(dn) <= 1e-6 * std::max(std::sqrt(no), std::sqrt(nn))) break;
        for (int i = 0; i < 3; ++i) oldp[i] = newp[i];                       // :70
    }
}

Variable update(const Variable& var, const double* x) {
    Variable out = var;
    switch (var.type) {
        case VT_EUCLID:                                                      // src/variable.jl:5,10
            for (int i = 0; i < var.ndof; ++i) out.v[i] = var.v[i] + x[i];
            break;
        case VT_CONTAMGAUSS: {                                               // src/robustadaptive.jl:22,12-19
            double a = update_zerotoinf(var.v[0], x[0]);
            double b = update_zerotoinf(var.v[1], x[1]);
            double w = update_zerotoone(var.v[2], x[2]);
            if (!(a >= b)) std::swap(a, b);
            out.v[0] = a; out.v[1] = b; out.v[2] = w;
            break;
        }
        case VT_PINHOLE: {  // repo-defined: R <- Exp(x[0:3]) * R ; remaining 6 parameters additive
            double E[9];
            so3_exp(x, E);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                double s = 0; for (int k = 0; k < 3; ++k) s += E[i + 3 * k] * var.v[k + 3 * j];
                out.v[i + 3 * j] = s;
            }
            for (int i = 0; i < 6; ++i) out.v[9 + i] = var.v[9 + i] + x[3 + i];
            break;
        }
    }
    return out;
}

// ContaminatedGaussian value / derivatives                                src/robustadaptive.jl:25-33
double cg_robustify(const Variable& k, double cost) {
    double a = k.v[0], b = k.v[1], w = k.v[2];
    double s1sq = a * a, s2sq = b * b;
    double hd = 0.5 * (s2sq - s1sq), hs2 = 0.5 * s2sq;
    return cost * hs2 - std::log(w * a * std::exp(cost * hd) + (1 - w) * b);
}
void cg_robustifydcost(const Variable& k, double cost, double& rho, double& d1, double& d2) {
    double a = k.v[0], b = k.v[1], w = k.v[2];
    double s1sq = a * a, s2sq = b * b;
    double hd = 0.5 * (s2sq - s1sq), hs2 = 0.5 * s2sq;
    double c = cost * hs2;
    double s = w * a * std::exp(cost * hd);
    double t = (1 - w) * b;
    double den = 1 / (s + t);
    s *= hd;
    rho = c + std::log(den);
    d1 = hs2 - s * den;
    d2 = -s * hd * t * den * den;
}
// x -> robustify(update(kernel, x), cost + x[4]) differentiated twice at 0 with exact forward-mode
// arithmetic — what autorobustifydkernel does with nested ForwardDiff duals (src/autodiff.jl:164-165,
// src/robust.jl:15).  Under duals the constructor does not re-sort (src/robustadaptive.jl:13).
void cg_robustifydkernel(const Variable& k, double cost, double& val, double g[4], double H[16]) {
    Jet2 x0 = jvar(0.0, 0), x1 = jvar(0.0, 1), x2 = jvar(0.0, 2), x3 = jvar(0.0, 3);
    Jet2 a = Jet2(k.v[0] > 0 ? k.v[0] : FLOATMIN) * jexp(x0);              // src/variable.jl:22
    Jet2 b = Jet2(k.v[1] > 0 ? k.v[1] : FLOATMIN) * jexp(x1);
    Jet2 vv = Jet2(k.v[2] > 0 ? k.v[2] : FLOATMIN) * jexp(x2);             // src/variable.jl:30
    Jet2 w = vv * jinv(Jet2(1.0) + (vv - Jet2(k.v[2])));                   // src/variable.jl:31
    Jet2 s1sq = a * a, s2sq = b * b;                                       // src/robustadaptive.jl:16-18
    Jet2 hd = Jet2(0.5) * (s2sq - s1sq), hs2 = Jet2(0.5) * s2sq;
    Jet2 c = Jet2(cost) + x3;
    Jet2 r = c * hs2 - jlog(w * a * jexp(c * hd) + (Jet2(1.0) - w) * b);   // src/robustadaptive.jl:25
    val = r.v;
    for (int i = 0; i < 4; ++i) g[i] = r.g[i];
    for (int i = 0; i < 16; ++i) H[i] = r.h[i];
}

// =======================================================================================
// Residuals                                       src/residual.jl:4-14, test/optimizeba.jl:4
// =======================================================================================
template <class T>
static void res_affine(const T* pose, const T* X, const double* z, T* r) {  // test/optimizeba.jl:4, src/residual.jl:13
    r[0] = (pose[0] * X[0] + pose[1] * X[1] + pose[2] * X[2]) - T(z[0]);
    r[1] = (pose[3] * X[0] + pose[4] * X[1] + pose[5] * X[2]) - T(z[1]);
}
// repo-defined pinhole (BAL convention): P = R X + t; p = -P.xy / P.z; proj = f (1 + k1 |p|^2 + k2 |p|^4) p
template <class T>
static void res_pinhole(const T* R, const T* t, const T& f, const T& k1, const T& k2, const T* X, const double* z, T* r) {
    T P0 = R[0] * X[0] + R[3] * X[1] + R[6] * X[2] + t[0];
    T P1 = R[1] * X[0] + R[4] * X[1] + R[7] * X[2] + t[1];
    T P2 = R[2] * X[0] + R[5] * X[1] + R[8] * X[2] + t[2];
    T iz = T(-1.0) / P2;
    T px = P0 * iz, py = P1 * iz;
    T n2 = px * px + py * py;
    T dist = T(1.0) + n2 * (k1 + k2 * n2);
    T s = f * dist;
    r[0] = s * px - T(z[0]);
    r[1] = s * py - T(z[1]);
}

static Dual seed(double v, int idx) { Dual d(v); d.d[idx] = 1.0; return d; }

void computeresidual(const Cost& c, const Variable* const* vars, int& m, double* r) {
    switch (c.type) {
        case RT_AFFINE_BA: m = 2; res_affine<double>(vars[0]->v, vars[1]->v, c.data, r); break;
        case RT_PINHOLE_BA: {
            m = 2; const double* cv = vars[0]->v;
            res_pinhole<double>(cv, cv + 9, cv[12], cv[13], cv[14], vars[1]->v, c.data, r);
            break;
        }
        case RT_ADAPTIVE_OFFSET: m = 1; r[0] = vars[1]->v[0] - c.data[0]; break;      // examples/adaptivekernel.jl:16
        case RT_ROSENBROCK_A: m = 1; r[0] = c.data[0] * (1 - vars[0]->v[0]); break;   // test/functional.jl:12
        case RT_ROSENBROCK_B: m = 1; r[0] = c.data[0] * (vars[0]->v[0] * vars[0]->v[0] - vars[1]->v[0]); break;  // :24
        default: m = 0;
    }
}

// (r, J) with J = d r(update(v, delta)) / d delta at delta = 0, columns in getvars order.
// Static ForwardDiff path of the reference (src/autodiff.jl:81-93,57-61,70-75); user-supplied analytic
// computeresjac for the adaptive residuals (examples/adaptivekernel.jl:17, test/adaptivecost.jl:11).
void computeresjac(const Cost& c, const Variable* const* vars, int& m, int& P, double* r, double* J) {
    switch (c.type) {
        case RT_AFFINE_BA: {
            m = 2; P = 9;
            Dual pose[6], X[3], rd[2];
            for (int i = 0; i < 6; ++i) pose[i] = seed(vars[0]->v[i], i);      // update(EuclideanVector) = v + delta
            for (int i = 0; i < 3; ++i) X[i] = seed(vars[1]->v[i], 6 + i);
            res_affine<Dual>(pose, X, c.data, rd);
            for (int i = 0; i < 2; ++i) { r[i] = rd[i].v; for (int j = 0; j < 9; ++j) J[i + 2 * j] = rd[i].d[j]; }
            break;
        }
        case RT_PINHOLE_BA: {
            m = 2; P = 12;
            const double* cv = vars[0]->v;
            // R(delta) = Exp(w) R ; first-order in w: (I + [w]x) R  (exact first derivatives at w = 0)
            Dual w[3] = {seed(0, 0), seed(0, 1), seed(0, 2)};
            Dual R[9], t[3], X[3], rd[2];
            for (int j = 0; j < 3; ++j) {
                Dual c0(cv[0 + 3 * j]), c1(cv[1 + 3 * j]), c2(cv[2 + 3 * j]);
                R[0 + 3 * j] = c0 + (w[1] * c2 - w[2] * c1);
                R[1 + 3 * j] = c1 + (w[2] * c0 - w[0] * c2);
                R[2 + 3 * j] = c2 + (w[0] * c1 - w[1] * c0);
            }
            for (int i = 0; i < 3; ++i) t[i] = seed(cv[9 + i], 3 + i);
            Dual f = seed(cv[12], 6), k1 = seed(cv[13], 7), k2 = seed(cv[14], 8);
            for (int i = 0; i < 3; ++i) X[i] = seed(vars[1]->v[i], 9 + i);
            res_pinhole<Dual>(R, t, f, k1, k2, X, c.data, rd);
            for (int i = 0; i < 2; ++i) { r[i] = rd[i].v; for (int j = 0; j < 12; ++j) J[i + 2 * j] = rd[i].d[j]; }
            break;
        }
        case RT_ADAPTIVE_OFFSET:  // vars[0] is the kernel; Jacobian is w.r.t. the mean only
            m = 1; P = 1; r[0] = vars[1]->v[0] - c.data[0]; J[0] = 1.0; break;
        case RT_ROSENBROCK_A: m = 1; P = 1; r[0] = c.data[0] * (1 - vars[0]->v[0]); J[0] = -c.data[0]; break;
        case RT_ROSENBROCK_B: {
            m = 1; P = 2; double xx = vars[0]->v[0];
            r[0] = c.data[0] * (xx * xx - vars[1]->v[0]);
            J[0] = c.data[0] * (xx + xx);  // d(x*x) under duals = x*dx + dx*x
            J[1] = -c.data[0];
            break;
        }
        default: m = 0; P = 0;
    }
}

static inline bool is_adaptive(int type) { return type == RT_ADAPTIVE_OFFSET; }

static inline double sqnorm(const double* r, int m) {                        // src/utils.jl:29-36
    double t = 0; for (int i = 0; i < m; ++i) t += r[i] * r[i]; return t;
}

// computerescost                                                            src/residual.jl:49-55
static double computecost(const Cost& c, const RobustSpec& k, const Variable* const* vars) {
    double r[4]; int m;
    computeresidual(c, vars, m, r);
    double s = sqnorm(r, m);
    if (is_adaptive(c.type)) return 0.5 * cg_robustify(*vars[0], s);
    return 0.5 * robustify(k, s);
}

// computerescostgradhess with all variables unfixed                         src/residual.jl:57-111
// g has length Pt, H is Pt x Pt column-major, Pt = (adaptive ? 3 : 0) + P.
static double computecostgradhess(const Cost& c, const RobustSpec& k, const Variable* const* vars, int& Pt, double* g, double* H) {
    double r[4], J[4 * 16];
    int m, P;
    computeresjac(c, vars, m, P, r, J);                                      // :69
    double cost = sqnorm(r, m);                                              // :72
    double gr[16], Hr[16 * 16];
    for (int j = 0; j < P; ++j) {                                            // g = J' r  :73
        double s = 0; for (int i = 0; i < m; ++i) s += J[i + m * j] * r[i];
        gr[j] = s;
    }
    for (int a = 0; a < P; ++a) for (int b = 0; b < P; ++b) {                // H = J' J  :74
        double s = 0; for (int i = 0; i < m; ++i) s += J[i + m * a] * J[i + m * b];
        Hr[a + P * b] = s;
    }
    double dc, d2c, dck[4], d2ck[16], dkdv[3 * 16];
    const bool adaptive = is_adaptive(c.type);
    if (!adaptive) {
        robustifydcost(k, cost, cost, dc, d2c);                              // :78
    } else {
        double val;
        cg_robustifydkernel(*vars[0], cost, val, dck, d2ck);                 // :81
        cost = val;
        dc = dck[3]; d2c = d2ck[3 + 4 * 3];                                  // :82-83
        for (int j = 0; j < P; ++j) for (int q = 0; q < 3; ++q)              // dkdv = g * d2c_[1:3,4]'  :87
            dkdv[j + P * q] = gr[j] * d2ck[q + 4 * 3];
    }
    if (dc != 1) for (int i = 0; i < P * P; ++i) Hr[i] *= dc;                // :91-93
    if (d2c != 0) {                                                          // :95-97
        for (int a = 0; a < P; ++a) for (int b = 0; b < P; ++b) Hr[a + P * b] += ((2 * d2c) * gr[a]) * gr[b];
    }
    if (dc != 1) for (int j = 0; j < P; ++j) gr[j] *= dc;                    // :99-101
    if (!adaptive) {
        Pt = P;
        for (int j = 0; j < P; ++j) g[j] = gr[j];
        for (int i = 0; i < P * P; ++i) H[i] = Hr[i];
    } else {                                                                 // :103-107
        Pt = 3 + P;
        for (int q = 0; q < 3; ++q) g[q] = dck[q];
        for (int j = 0; j < P; ++j) g[3 + j] = gr[j];
        for (int a = 0; a < Pt; ++a) for (int b = 0; b < Pt; ++b) {
            double v;
            if (a < 3 && b < 3) v = d2ck[a + 4 * b];
            else if (a >= 3 && b < 3) v = dkdv[(a - 3) + P * b];
            else if (a < 3 && b >= 3) v = dkdv[(b - 3) + P * a];
            else v = Hr[(a - 3) + P * (b - 3)];
            H[a + Pt * b] = v;
        }
    }
    return 0.5 * cost;                                                       // :110
}

// =======================================================================================
// Problem container                                                         src/problem.jl
// =======================================================================================
int64_t Problem::addvariable(const Variable& v) {
    variables.push_back(v);
    lsready = false;
    return (int64_t)variables.size();
}
void Problem::addcost(const Cost& c, const RobustSpec& k) {
    // a VectorRepo slot per concrete residual TYPE (src/VectorRepo.jl:3): in Julia robustkernel(res) belongs to the type, so the same
    // residual struct with another kernel is another type with its own vector (test/functional.jl:14-24 registers two types)
    size_t t = 0;
    for (; t < costtypes.size(); ++t)
        if (costtypes[t] == c.type && kernels[t].kind == k.kind && kernels[t].width == k.width && kernels[t].scaled == k.scaled && kernels[t].height == k.height) break;
    if (t == costtypes.size()) { costtypes.push_back(c.type); costs.emplace_back(); kernels.push_back(k); }
    costs[t].push_back(c);
    lsready = false;
}
// cost(vars, costs): sequential left fold per type, then across types     src/cost.jl:11, src/VectorRepo.jl:64-69
double Problem::cost(const std::vector<Variable>& vars) const {
    double total = 0.0;
    for (size_t t = 0; t < costs.size(); ++t) {
        double sub = 0.0;
        for (const Cost& c : costs[t]) {
            const Variable* v[4];
            for (int i = 0; i < c.ndeps; ++i) v[i] = &vars[c.vi[i] - 1];
            sub += computecost(c, kernels[t], v);
        }
        total = (t == 0) ? sub : total + sub;
    }
    return total;
}

// =======================================================================================
// BlockSparseMatrix                                                src/BlockSparseMatrix.jl
// =======================================================================================
void BSM::build(const std::vector<int64_t>& colptr, const std::vector<int64_t>& rowval,
                const std::vector<int>& rowblocksizes, const std::vector<int>& colblocksizes) {
    // `colptr/rowval` (1-based) describe sparsitytransposed: size (ncolblocks x nrowblocks)   :30-47
    rbs = rowblocksizes; cbs = colblocksizes;
    t_colptr = colptr; t_rowval = rowval;
    t_nzval.assign(rowval.size(), 0);
    int64_t start = 1, ind = 0;
    for (size_t row = 0; row < rbs.size(); ++row) {
        int64_t rowwidth = rbs[row];
        for (int64_t p = colptr[row] - 1; p < colptr[row + 1] - 1; ++p) {
            int64_t col = rowval[p];
            t_nzval[ind++] = start;
            start += rowwidth * (int64_t)cbs[col - 1];
        }
    }
    data.assign((size_t)(start - 1), 0.0);
    m = 0; for (int s : rbs) m += s;
    n = 0; for (int s : cbs) n += s;
    i_colptr.clear(); i_rowval.clear(); i_nzval.clear();
}
int64_t BSM::start(int64_t i, int64_t j) const {  // indicestransposed[j, i]               :102
    auto b = t_rowval.begin() + (t_colptr[i - 1] - 1), e = t_rowval.begin() + (t_colptr[i] - 1);
    auto it = std::lower_bound(b, e, j);
    if (it == e || *it != j) return 0;
    return t_nzval[it - t_rowval.begin()];
}
void BSM::cacheindices() {  // transpose of indicestransposed                             :53-61
    if (!i_colptr.empty()) return;
    size_t ncb = cbs.size(), nrb = rbs.size(), nz = t_rowval.size();
    i_colptr.assign(ncb + 1, 0); i_rowval.assign(nz, 0); i_nzval.assign(nz, 0);
    for (size_t p = 0; p < nz; ++p) i_colptr[t_rowval[p]]++;
    int64_t acc = 1;
    for (size_t c = 0; c < ncb; ++c) { int64_t cnt = i_colptr[c + 1]; i_colptr[c] = acc; acc += cnt; }
    i_colptr[ncb] = acc;
    std::vector<int64_t> next(i_colptr.begin(), i_colptr.end() - 1);
    for (size_t r = 0; r < nrb; ++r)
        for (int64_t p = t_colptr[r] - 1; p < t_colptr[r + 1] - 1; ++p) {
            int64_t c = t_rowval[p];
            int64_t q = next[c - 1]++ - 1;
            i_rowval[q] = (int64_t)r + 1;
            i_nzval[q] = t_nzval[p];
        }
}
void BSM::uniformscaling(double k) {                                                    // :90-99
    for (size_t i = 0; i < rbs.size(); ++i) {
        int64_t ind = start((int64_t)i + 1, (int64_t)i + 1);
        int64_t bs = rbs[i];
        for (int64_t j = ind; j <= ind + bs * bs - 1; j += bs + 1) data[j - 1] += k;  // Julia range ind:(bs+1):(ind+bs^2) stops at the last diagonal
    }
}
void BSM::todense(std::vector<double>& out) const {                                     // :245-264
    std::vector<int64_t> rs(rbs.size() + 1, 0), cs(cbs.size() + 1, 0);
    for (size_t i = 0; i < rbs.size(); ++i) rs[i + 1] = rs[i] + rbs[i];
    for (size_t i = 0; i < cbs.size(); ++i) cs[i + 1] = cs[i] + cbs[i];
    out.assign((size_t)(m * n), 0.0);
    for (size_t r = 0; r < rbs.size(); ++r)
        for (int64_t p = t_colptr[r] - 1; p < t_colptr[r + 1] - 1; ++p) {
            int64_t c = t_rowval[p] - 1, idx = t_nzval[p] - 1;
            int64_t r_ = rbs[r], c_ = cbs[c];
            for (int64_t jj = 0; jj < c_; ++jj) for (int64_t ii = 0; ii < r_; ++ii)
                out[(size_t)((rs[r] + ii) + m * (cs[c] + jj))] = data[(size_t)(idx + ii + r_ * jj)];
        }
}
void BSM::symmetrifyfull(std::vector<double>& out) const {                              // :199-243
    std::vector<int64_t> rs(rbs.size() + 1, 0);
    for (size_t i = 0; i < rbs.size(); ++i) rs[i + 1] = rs[i] + rbs[i];
    out.assign((size_t)(m * m), 0.0);
    for (size_t r = 0; r < rbs.size(); ++r)
        for (int64_t p = t_colptr[r] - 1; p < t_colptr[r + 1] - 1; ++p) {
            int64_t c = t_rowval[p] - 1, idx = t_nzval[p] - 1;
            int64_t r_ = rbs[r], c_ = rbs[c];
            for (int64_t jj = 0; jj < c_; ++jj) for (int64_t ii = 0; ii < r_; ++ii) {
                double v = data[(size_t)(idx + ii + r_ * jj)];
                out[(size_t)((rs[r] + ii) + m * (rs[c] + jj))] = v;
                if ((int64_t)r != c) out[(size_t)((rs[c] + jj) + m * (rs[r] + ii))] = v;
            }
        }
}

// makesparseindices: BSM -> CSC index map, optionally symmetrified                      :141-191
CSCIndex makesparseindices(BSM& bsm, bool symmetrify) {
    bsm.cacheindices();
    size_t nrb = bsm.rbs.size(), ncb = bsm.cbs.size();
    std::vector<int64_t> startrow(nrb + 1);
    startrow[0] = 1;
    for (size_t i = 0; i < nrb; ++i) startrow[i + 1] = startrow[i] + bsm.rbs[i];
    int64_t diagspace = 0;                                                               // :123-139
    if (symmetrify) {
        for (size_t col = 0; col < ncb; ++col) {
            int64_t ind = bsm.i_colptr[col];
            if (bsm.i_colptr[col + 1] > ind) {
                int64_t row = bsm.i_rowval[ind - 1];
                if (row == (int64_t)col + 1) diagspace += (int64_t)bsm.rbs[row - 1] * bsm.rbs[row - 1];
            }
        }
    }
    int64_t nzvals = symmetrify ? (int64_t)bsm.data.size() * 2 - diagspace : (int64_t)bsm.data.size();
    CSCIndex out;
    out.rowval.resize((size_t)nzvals); out.nzval.resize((size_t)nzvals);
    int64_t ncolscalar = 0; for (int s : bsm.cbs) ncolscalar += s;
    out.colptr.assign((size_t)ncolscalar + 1, 0);
    int64_t ind = 1, col = 1;
    out.colptr[0] = 1;
    for (size_t col_ = 0; col_ < ncb; ++col_) {
        int64_t colblocksize = bsm.cbs[col_];
        int64_t lo0 = bsm.i_colptr[col_], lo1 = bsm.i_colptr[col_ + 1] - 1;  // lower_rows (1-based inclusive)
        int64_t up0 = 1, up1 = 0;
        if (symmetrify) {
            bool hasdiag = (lo1 >= lo0) && bsm.i_rowval[lo0 - 1] == (int64_t)col_ + 1;
            up0 = bsm.t_colptr[col_]; up1 = bsm.t_colptr[col_ + 1] - 1 - (hasdiag ? 1 : 0);
        }
        for (int64_t innercol = 0; innercol < colblocksize; ++innercol) {
            for (int64_t r = up0; r <= up1; ++r) {            // above-diagonal blocks (transposed)
                int64_t row = bsm.t_rowval[r - 1];
                int64_t s = startrow[row - 1], c = bsm.rbs[row - 1];
                int64_t v = bsm.t_nzval[r - 1] + innercol;
                for (int64_t i = 0; i < c; ++i) {
                    out.rowval[ind - 1] = s + i; out.nzval[ind - 1] = v; ++ind; v += colblocksize;
                }
            }
            for (int64_t r = lo0; r <= lo1; ++r) {            // diagonal and below
                int64_t row = bsm.i_rowval[r - 1];
                int64_t s = startrow[row - 1], c = bsm.rbs[row - 1];
                int64_t v = bsm.i_nzval[r - 1] + innercol * c;
                for (int64_t i = 0; i < c; ++i) {
                    out.rowval[ind - 1] = s + i; out.nzval[ind - 1] = v + i; ++ind;
                }
            }
            out.colptr[col] = ind; ++col;
        }
    }
    out.m = startrow[nrb] - 1; out.n = col - 1;
    return out;
}

std::vector<int64_t> runlengthencodesortedints(const std::vector<int64_t>& s) {          // src/utils.jl:38-52
    std::vector<int64_t> run((size_t)s.back() + 2);
    int64_t ind = 1, currval = 1;
    run[currval - 1] = ind;
    for (int64_t val : s) {
        while (val >= currval) { currval += 1; run[currval - 1] = ind; }
        ind += 1;
    }
    run[currval] = ind;
    return run;
}

// =======================================================================================
// Linear solvers                                                       src/linearsolver.jl
// =======================================================================================
// try_cholesky!: cholesky(A; check=false), if it fails ldiv!(x, qr(A), b)              :20-26
int solve_dense(int n, const double* A, const double* b, double* x) {
    std::vector<double> L(A, A + (size_t)n * n);
    bool ok = true;
    for (int j = 0; j < n && ok; ++j) {  // lower Cholesky, column by column (LAPACK dpotrf semantics)
        double d = L[j + (size_t)n * j];
        for (int k = 0; k < j; ++k) d -= L[j + (size_t)n * k] * L[j + (size_t)n * k];
        if (!(d > 0.0)) { ok = false; break; }
        d = std::sqrt(d);
        L[j + (size_t)n * j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i + (size_t)n * j];
            for (int k = 0; k < j; ++k) s -= L[i + (size_t)n * k] * L[j + (size_t)n * k];
            L[i + (size_t)n * j] = s / d;
        }
    }
    if (ok) {
        for (int i = 0; i < n; ++i) {
            double s = b[i];
            for (int k = 0; k < i; ++k) s -= L[i + (size_t)n * k] * x[k];
            x[i] = s / L[i + (size_t)n * i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = x[i];
            for (int k = i + 1; k < n; ++k) s -= L[k + (size_t)n * i] * x[k];
            x[i] = s / L[i + (size_t)n * i];
        }
        return 0;
    }
    // Householder QR (unpivoted, as LinearAlgebra.qr for dense matrices), x = R^-1 Q' b
    std::vector<double> R(A, A + (size_t)n * n), y(b, b + n), v(n);
    for (int k = 0; k < n; ++k) {
        double nrm = 0; for (int i = k; i < n; ++i) nrm += R[i + (size_t)n * k] * R[i + (size_t)n * k];
        nrm = std::sqrt(nrm);
        if (nrm == 0) continue;
        double alpha = R[k + (size_t)n * k] > 0 ? -nrm : nrm;
        for (int i = k; i < n; ++i) v[i] = R[i + (size_t)n * k];
        v[k] -= alpha;
        double vn = 0; for (int i = k; i < n; ++i) vn += v[i] * v[i];
        if (vn == 0) continue;
        for (int j = k; j < n; ++j) {
            double s = 0; for (int i = k; i < n; ++i) s += v[i] * R[i + (size_t)n * j];
            s = 2 * s / vn;
            for (int i = k; i < n; ++i) R[i + (size_t)n * j] -= s * v[i];
        }
        double s = 0; for (int i = k; i < n; ++i) s += v[i] * y[i];
        s = 2 * s / vn;
        for (int i = k; i < n; ++i) y[i] -= s * v[i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= R[i + (size_t)n * k] * x[k];
        x[i] = s / R[i + (size_t)n * i];
    }
    return 1;
}

// Sparse LDL^T. LDLFactorizations.jl (Project.toml compat 0.10; not vendored in /root/reference) is a
// Julia port of T. Davis' LDL (ACM TOMS Alg. 849): symbolic = elimination tree + column counts,
// numeric = up-looking sparse triangular solves; no pivoting; uses the upper triangle of P A P'.
// Call sites: src/linearsystem.jl:68 (ldl_analyze), src/linearsolver.jl:29 (ldl_factorize! + ldiv!).
void ldl_analyze(const CSCIndex& A, const std::vector<int64_t>& perm, LDLSymbolic& sym) {
    int64_t n = A.n;
    sym.n = n; sym.P = perm; sym.Pinv.assign(n, 0);
    for (int64_t k = 0; k < n; ++k) sym.Pinv[perm[k]] = k;
    sym.Parent.assign(n, -1); sym.Lnz0.assign(n, 0); sym.Lp.assign(n + 1, 0);
    std::vector<int64_t> Flag(n);
    for (int64_t k = 0; k < n; ++k) {
        Flag[k] = k;
        int64_t kk = perm[k];
        for (int64_t p = A.colptr[kk] - 1; p < A.colptr[kk + 1] - 1; ++p) {
            int64_t i = sym.Pinv[A.rowval[p] - 1];
            if (i < k) {
                for (; Flag[i] != k; i = sym.Parent[i]) {
                    if (sym.Parent[i] == -1) sym.Parent[i] = k;
                    sym.Lnz0[i]++;
                    Flag[i] = k;
                }
            }
        }
    }
    for (int64_t k = 0; k < n; ++k) sym.Lp[k + 1] = sym.Lp[k] + sym.Lnz0[k];
}
bool ldl_factor_solve(const CSCIndex& A, const double* Ax, const LDLSymbolic& sym, const double* b, double* x) {
    int64_t n = sym.n;
    std::vector<int64_t> Li((size_t)sym.Lp[n]), Lnz(n, 0), Flag(n), Pattern(n);
    std::vector<double> Lx((size_t)sym.Lp[n]), D(n), Y(n, 0.0);
    bool ok = true;
    for (int64_t k = 0; k < n; ++k) {
        Y[k] = 0.0;
        int64_t top = n;
        Flag[k] = k;
        Lnz[k] = 0;
        int64_t kk = sym.P[k];
        for (int64_t p = A.colptr[kk] - 1; p < A.colptr[kk + 1] - 1; ++p) {
            int64_t i = sym.Pinv[A.rowval[p] - 1];
            if (i <= k) {
                Y[i] += Ax[p];
                int64_t len = 0;
                for (; Flag[i] != k; i = sym.Parent[i]) { Pattern[len++] = i; Flag[i] = k; }
                while (len > 0) Pattern[--top] = Pattern[--len];
            }
        }
        D[k] = Y[k]; Y[k] = 0.0;
        for (; top < n; ++top) {
            int64_t i = Pattern[top];
            double yi = Y[i]; Y[i] = 0.0;
            int64_t p2 = sym.Lp[i] + Lnz[i];
            for (int64_t p = sym.Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
            double l_ki = yi / D[i];
            D[k] -= l_ki * yi;
            Li[p2] = k; Lx[p2] = l_ki; Lnz[i]++;
        }
        if (D[k] == 0.0) ok = false;
    }
    std::vector<double> w(n);
    for (int64_t k = 0; k < n; ++k) w[k] = b[sym.P[k]];
    for (int64_t j = 0; j < n; ++j) { int64_t p2 = sym.Lp[j] + Lnz[j]; for (int64_t p = sym.Lp[j]; p < p2; ++p) w[Li[p]] -= Lx[p] * w[j]; }
    for (int64_t j = 0; j < n; ++j) w[j] /= D[j];
    for (int64_t j = n - 1; j >= 0; --j) { int64_t p2 = sym.Lp[j] + Lnz[j]; for (int64_t p = sym.Lp[j]; p < p2; ++p) w[j] -= Lx[p] * w[Li[p]]; }
    for (int64_t k = 0; k < n; ++k) x[sym.P[k]] = w[k];
    return ok;
}

double fast_bAb_sparse(const CSCIndex& A, const double* nz, const double* b) {            // src/utils.jl:95-106
    double total = 0;
    for (int64_t i = 0; i < A.n; ++i) {
        double col = 0;
        for (int64_t j = A.colptr[i] - 1; j < A.colptr[i + 1] - 1; ++j) col += nz[j] * b[A.rowval[j] - 1];
        col *= b[i];
        total += col;
    }
    return total;
}
double fast_bAb_dense(int n, const double* A, const double* b) {                          // src/utils.jl:71-81
    double total = 0;
    for (int i = 0; i < n; ++i) {
        double sub = 0;
        for (int j = 0; j < n; ++j) sub += A[j + (size_t)n * i] * b[j];
        total += b[i] * sub;
    }
    return total;
}

// =======================================================================================
// Linear system                                                        src/linearsystem.jl
// =======================================================================================
void Problem::makesymmvls() {                                                             // :91-124
    // blockindices / blocksizes over the unfixed variables                               :93-102
    blockindices.assign(variables.size(), 0);
    std::vector<int> bs;
    for (size_t i = 0; i < variables.size(); ++i)
        if (unfixed.empty() || unfixed[i]) { bs.push_back(variables[i].ndof); blockindices[i] = (int64_t)bs.size(); }
    const size_t nb = bs.size();
    nblocks = (int64_t)nb;
    boffsets.assign(nb + 1, 1);
    for (size_t i = 0; i < nb; ++i) boffsets[i + 1] = boffsets[i] + bs[i];
    dof = boffsets[nb] - 1;
    sparse = false;
    std::vector<int64_t> colptr, rowval;
    if (dof >= 40) {
        // lower-triangular block pattern of V V' > 0, V = varcostmap                      :107-110
        std::vector<uint64_t> pairs;
        size_t ncost = 0; for (auto& v : costs) ncost += v.size();
        pairs.reserve(ncost * 3 + nb);
        for (auto& vec : costs) for (const Cost& c : vec)
            for (int a = 0; a < c.ndeps; ++a) for (int bb = 0; bb <= a; ++bb) {
                const int64_t ba = blockindices[(size_t)c.vi[a] - 1], b2 = blockindices[(size_t)c.vi[bb] - 1];
                if (ba == 0 || b2 == 0) continue;                                         // sparsity[unfixed, :]  :108
                uint64_t i = (uint64_t)std::max(ba, b2), j = (uint64_t)std::min(ba, b2);
                pairs.push_back((i << 32) | j);  // key sorts by block row then block column
            }
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        colptr.assign(nb + 1, 0);
        rowval.resize(pairs.size());
        for (size_t p = 0; p < pairs.size(); ++p) { colptr[(pairs[p] >> 32)]++; rowval[p] = (int64_t)(pairs[p] & 0xffffffffu); }
        int64_t acc = 1;
        for (size_t r = 0; r < nb; ++r) { int64_t cnt = colptr[r + 1]; colptr[r] = acc; acc += cnt; }
        colptr[nb] = acc;
        // block_sparse_nnz + sparse_dense_decision                          src/utils.jl:108-120
        int64_t nnz = 0;
        for (size_t r = 0; r < nb; ++r) for (int64_t p = colptr[r] - 1; p < colptr[r + 1] - 1; ++p) nnz += (int64_t)bs[r] * bs[rowval[p] - 1];
        sparse = (nnz * 64) < (25 * dof * (dof - 40));
    }
    b.assign((size_t)dof, 0.0); x.assign((size_t)dof, 0.0);
    if (sparse) {
        A.build(colptr, rowval, bs, bs);                                                  // :115
        hess = makesparseindices(A, true);                                                // :58-60
        hessval.assign(hess.nzval.size(), 0.0);
        // Fill-reducing order.  The reference calls AMD inside ldl_analyze; AMD is not restated here.
        // Blocks are ordered by ascending degree (stable, degrees above 64 tied so that high-degree
        // camera blocks keep their natural, band-preserving order), which like AMD eliminates the
        // low-degree point blocks before the camera blocks.  The permutation changes rounding only.
        std::vector<int64_t> deg(nb, 0), order(nb);
        for (size_t r = 0; r < nb; ++r) for (int64_t p = colptr[r] - 1; p < colptr[r + 1] - 1; ++p) { deg[r]++; if ((size_t)(rowval[p] - 1) != r) deg[rowval[p] - 1]++; }
        std::iota(order.begin(), order.end(), 0);
        if (elimination_order == 1) std::reverse(order.begin(), order.end());
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t c) { return std::min<int64_t>(deg[a], 64) < std::min<int64_t>(deg[c], 64); });
        std::vector<int64_t> perm; perm.reserve((size_t)dof);
        for (int64_t blk : order) for (int k = 0; k < bs[(size_t)blk]; ++k) perm.push_back(boffsets[(size_t)blk] - 1 + k);
        ldl_analyze(hess, perm, ldl);                                                     // :68
    } else {
        Adense.assign((size_t)(dof * dof), 0.0);                                          // :80-86
    }
    lsready = true;
}
void Problem::zero() {                                                                    // :192-195
    std::fill(b.begin(), b.end(), 0.0);
    if (sparse) std::fill(A.data.begin(), A.data.end(), 0.0); else std::fill(Adense.begin(), Adense.end(), 0.0);
}

// costgradhess! over all blocks, all variables unfixed            src/cost.jl:29-54, src/linearsystem.jl:132-175
double Problem::costgradhess() {
    double total = 0.0;
    for (size_t t = 0; t < costs.size(); ++t) {
        double sub = 0.0;
        for (const Cost& c : costs[t]) {
            const Variable* v[4];
            bool any = false;
            for (int i = 0; i < c.ndeps; ++i) { v[i] = &variables[c.vi[i] - 1]; any = any || blockindices[(size_t)c.vi[i] - 1] != 0; }
            if (!any) { sub += computecost(c, kernels[t], v); continue; }                 // no unfixed variable: cost only  src/cost.jl:51
            double g[20], H[20 * 20]; int Pt;
            // the reference differentiates only w.r.t. the unfixed variables (varflags, src/cost.jl:36-47); the blocks it obtains are
            // the corresponding sub-blocks of the all-unfixed (g, H) computed here (J'r, rho' J'J + 2 rho'' g g' are separable)
            sub += computecostgradhess(c, kernels[t], v, Pt, g, H);
            // updateb!                                                                    :159-170
            int loff = 0;
            for (int i = 0; i < c.ndeps; ++i) {
                int nv = v[i]->ndof;
                const int64_t bi = blockindices[(size_t)c.vi[i] - 1];
                if (bi != 0) {
                    int64_t off = boffsets[bi - 1] - 1;
                    for (int k = 0; k < nv; ++k) b[(size_t)(off + k)] += g[loff + k];
                }
                loff += nv;
            }
            // updatesymA!                                                                 :132-157
            int loffi = 0;
            for (int i = 0; i < c.ndeps; ++i) {
                int nvi = v[i]->ndof; int64_t bi = blockindices[(size_t)c.vi[i] - 1];
                if (bi == 0) { loffi += nvi; continue; }
                auto add = [&](int64_t brow, int64_t bcol, int nr, int nc, int ro, int co) {
                    // block(A, brow, bcol) .+= H[ro.., co..]
                    if (sparse) {
                        int64_t st = A.start(brow, bcol) - 1;
                        for (int cc = 0; cc < nc; ++cc) for (int rr = 0; rr < nr; ++rr) A.data[(size_t)(st + rr + nr * cc)] += H[(ro + rr) + Pt * (co + cc)];
                    } else {
                        int64_t r0 = boffsets[brow - 1] - 1, c0 = boffsets[bcol - 1] - 1;
                        for (int cc = 0; cc < nc; ++cc) for (int rr = 0; rr < nr; ++rr) Adense[(size_t)((r0 + rr) + dof * (c0 + cc))] += H[(ro + rr) + Pt * (co + cc)];
                    }
                };
                add(bi, bi, nvi, nvi, loffi, loffi);
                int loffj = 0;
                for (int j = 0; j < i; ++j) {
                    int nvj = v[j]->ndof; int64_t bj = blockindices[(size_t)c.vi[j] - 1];
                    if (bj != 0) {
                        if (bi >= bj) add(bi, bj, nvi, nvj, loffi, loffj);
                        else add(bj, bi, nvj, nvi, loffj, loffi);
                    }
                    loffj += nvj;
                }
                loffi += nvi;
            }
        }
        total = (t == 0) ? sub : total + sub;
    }
    return total;
}
void Problem::gethessian() {                                                              // :180-190
    if (sparse) {
        for (size_t i = 0; i < hess.nzval.size(); ++i) hessval[i] = A.data[(size_t)(hess.nzval[i] - 1)];
    } else {                                                    // src/BlockDenseMatrix.jl:24-34
        for (int64_t r = 1; r < dof; ++r) for (int64_t c = 0; c < r; ++c) Adense[(size_t)(c + dof * r)] = Adense[(size_t)(r + dof * c)];
    }
}

// =======================================================================================
// iterators + outer loop                        src/iterators.jl:11-208, src/optimize.jl:109-180
// (Newton :11-27, Dogleg :30-115, Levenberg-Marquardt :120-172, gradient descent :177-208)
// =======================================================================================
Result Problem::optimize(const Options& opt, std::vector<IterRecord>* trace) {
    uint64_t starttime = time_ns();
    Result res;
    uint64_t t_init = 0, t_cost = 0, t_grad = 0, t_solver = 0;
    if (!lsready) makesymmvls();                               // src/optimize.jl:16
    zero();
    if (varnext.size() != variables.size()) varnext = variables;   // :80-82
    double lambda = 0.0;                                       // LevMarData(0.0)  src/iterators.jl:124
    double trustradius = 0.0;                                  // DoglegData       :36,42
    double stepsize = 1.0;                                     // GradientDescentData :181,184
    std::vector<double> cauchy;                                // DoglegData.cauchy
    int64_t fails = 0, iternum = 0;
    uint64_t stoptime = starttime + opt.maxtime_ns;            // :115
    t_init += time_ns() - starttime;
    uint64_t t0 = time_ns();
    double cost = costgradhess();                              // :118
    t_grad += time_ns() - t0;
    res.gradientcomputations += 1;
    double bestcost = cost;
    res.startcost = cost;                                      // max(cost, -Inf)   :121
    int64_t converged = 0;
    auto diag = [&](int64_t i) -> double& { return sparse ? hessval[0] : Adense[(size_t)(i + dof * i)]; };
    (void)diag;
    // position of the diagonal entries in the CSC values (sparse path)
    std::vector<int64_t> diagpos;
    if (sparse) {
        diagpos.resize((size_t)dof);
        for (int64_t j = 0; j < dof; ++j)
            for (int64_t p = hess.colptr[j] - 1; p < hess.colptr[j + 1] - 1; ++p)
                if (hess.rowval[p] - 1 == j) { diagpos[(size_t)j] = p; break; }
    }
    auto scale_diag = [&](double k) {                          // uniformscaling!(::AbstractMatrix) src/BlockSparseMatrix.jl:83-88
        if (sparse) for (int64_t j = 0; j < dof; ++j) hessval[(size_t)diagpos[(size_t)j]] += k;
        else for (int64_t j = 0; j < dof; ++j) Adense[(size_t)(j + dof * j)] += k;
    };
    while (true) {
        iternum += 1;
        int64_t ntries = 0;
        double cost_ = 0.0;
        double maxstep = 0;
        auto solve_now = [&]() {                                // negate!(solve!(linsystem, options))  src/linearsolver.jl:20-32
            t0 = time_ns();
            if (sparse) ldl_factor_solve(hess, hessval.data(), ldl, b.data(), x.data());
            else solve_dense((int)dof, Adense.data(), b.data(), x.data());
            for (auto& xi : x) xi = -xi;
            t_solver += time_ns() - t0;
            res.linearsolvers += 1; ntries += 1;
        };
        auto update_and_cost = [&]() {                          // update! + cost(varnext)
            for (size_t i = 0; i < variables.size(); ++i)          // update!  src/linearsystem.jl:206-213 (fixed variables are left alone)
                if (blockindices[i]) varnext[i] = update(variables[i], x.data() + (boffsets[(size_t)blockindices[i] - 1] - 1));
            t0 = time_ns();
            const double c = this->cost(varnext);
            t_cost += time_ns() - t0;
            res.costcomputations += 1;
            return c;
        };
        auto maxabs_x = [&]() {                                 // maximum(abs, x), NaN-propagating
            double m = 0; bool nan = false;
            for (double xi : x) { if (std::isnan(xi)) nan = true; m = std::max(m, std::fabs(xi)); }
            return nan ? std::numeric_limits<double>::quiet_NaN() : m;
        };
        auto norm_x = [&]() { double sq = 0; for (double xi : x) sq += xi * xi; return std::sqrt(sq); };
        if (opt.iterator == 0) {
            // ---- iterate!(::NewtonData)                                       src/iterators.jl:17-27
            gethessian();
            solve_now();
            cost_ = update_and_cost();
        } else if (opt.iterator == 2) {
            // ---- iterate!(::DoglegData)                                       src/iterators.jl:48-113
            gethessian();                                      // gethessgrad :49
            double gnorm2 = 0; for (double g : b) gnorm2 += g * g;                                  // :52
            const double bAb = sparse ? fast_bAb_sparse(hess, hessval.data(), b.data()) : fast_bAb_dense((int)dof, Adense.data(), b.data());
            const double a = gnorm2 / (bAb + std::numeric_limits<double>::min());                   // :53 (floatmin)
            cauchy.resize((size_t)dof);
            for (int64_t j = 0; j < dof; ++j) cauchy[(size_t)j] = -a * b[(size_t)j];               // :54
            const double alpha2 = a * a * gnorm2, alpha = std::sqrt(alpha2);                        // :55-56
            if (trustradius == 0) trustradius = alpha;                                              // :57-60
            double beta = 0;
            if (alpha < trustradius) { solve_now(); beta = norm_x(); }                              // :61-66
            cost_ = bestcost;                                                                       // :68
            while (true) {
                double linear_approx;
                if (!(alpha < trustradius)) {                                                       // first leg :71-74
                    for (int64_t j = 0; j < dof; ++j) x[(size_t)j] = (trustradius / alpha) * cauchy[(size_t)j];
                    linear_approx = trustradius * (2 * alpha - trustradius) / (2 * a);
                } else if (beta <= trustradius) {                                                   // full Newton step :77-79
                    linear_approx = cost_;
                } else {                                                                            // second leg :80-95
                    double sq_leg = 0, c = 0;
                    for (int64_t j = 0; j < dof; ++j) { x[(size_t)j] -= cauchy[(size_t)j]; }
                    for (int64_t j = 0; j < dof; ++j) { sq_leg += x[(size_t)j] * x[(size_t)j]; c += cauchy[(size_t)j] * x[(size_t)j]; }
                    const double trsq = trustradius * trustradius - alpha2;
                    double step = std::sqrt(c * c + sq_leg * trsq);
                    step = (c <= 0) ? (-c + step) / sq_leg : trsq / (c + step);
                    for (int64_t j = 0; j < dof; ++j) x[(size_t)j] = x[(size_t)j] * step + cauchy[(size_t)j];
                    linear_approx = 0.5 * (a * (1 - step) * (1 - step) * gnorm2) + step * (2 - step) * cost_;
                }
                cost_ = update_and_cost();                                                          // :98-101
                const double mu_ = (bestcost - cost_) / linear_approx;                              // :103
                if (mu_ > 0.375) trustradius = std::max(trustradius, 3 * norm_x());                 // :104-105
                else if (mu_ < 0.125) trustradius *= 0.5;                                           // :106-107
                maxstep = maxabs_x();
                if (!(cost_ > bestcost) || maxstep < opt.dstep) break;                              // :110-113
            }
        } else if (opt.iterator == 3) {
            // ---- iterate!(::GradientDescentData)                              src/iterators.jl:187-206
            for (int64_t j = 0; j < dof; ++j) x[(size_t)j] = -b[(size_t)j] * stepsize;             // :190
            cost_ = update_and_cost();
            while (cost_ > bestcost) {                                                              // :195
                double coststep = 0; for (int64_t j = 0; j < dof; ++j) coststep += x[(size_t)j] * b[(size_t)j];   // :197
                const double costdiff = bestcost + coststep - cost_;                                // :198
                stepsize *= 0.5 * coststep / costdiff;                                              // :200
                for (int64_t j = 0; j < dof; ++j) x[(size_t)j] = -b[(size_t)j] * stepsize;         // :202
                cost_ = update_and_cost();
            }
            stepsize *= 2;                                                                          // :207
        } else {
            // ---- iterate!(::LevMarData)                                       src/iterators.jl:139-172
            gethessian();                                          // :141
            if (lambda == 0) {                                     // initlambda :131-137,142-144
                double mx = 0;
                for (int64_t j = 0; j < dof; ++j) {
                    double d = sparse ? hessval[(size_t)diagpos[(size_t)j]] : Adense[(size_t)(j + dof * j)];
                    mx = std::max(mx, std::fabs(d));
                }
                lambda = mx * 1e-6;
            }
            double lastlambda = 0.0, mu = 2.0;
            while (true) {
                scale_diag(lambda - lastlambda);                   // :149
                lastlambda = lambda;
                t0 = time_ns();
                if (sparse) ldl_factor_solve(hess, hessval.data(), ldl, b.data(), x.data());     // src/linearsolver.jl:29
                else solve_dense((int)dof, Adense.data(), b.data(), x.data());                  // :30
                for (auto& xi : x) xi = -xi;                       // negate!  :152
                t_solver += time_ns() - t0;
                res.linearsolvers += 1; ntries += 1;
                for (size_t i = 0; i < variables.size(); ++i)      // update!  src/linearsystem.jl:206-213
                    if (blockindices[i]) varnext[i] = update(variables[i], x.data() + (boffsets[(size_t)blockindices[i] - 1] - 1));
                t0 = time_ns();
                cost_ = this->cost(varnext);                       // :157
                t_cost += time_ns() - t0;
                res.costcomputations += 1;
                maxstep = 0; bool stepnan = false;
                for (double xi : x) { if (std::isnan(xi)) stepnan = true; maxstep = std::max(maxstep, std::fabs(xi)); }
                if (stepnan) maxstep = std::numeric_limits<double>::quiet_NaN();  // maximum() propagates NaN
                if (!(cost_ > bestcost) || maxstep < opt.dstep) {  // :160
                    scale_diag(-lastlambda);                       // :162
                    double bAb = sparse ? fast_bAb_sparse(hess, hessval.data(), x.data()) : fast_bAb_dense((int)dof, Adense.data(), x.data());
                    double gx = 0; for (int64_t j = 0; j < dof; ++j) gx += b[(size_t)j] * x[(size_t)j];
                    double q = (cost_ - bestcost) / (0.5 * bAb + gx);      // :163
                    double t = 2 * q - 1;
                    lambda *= q < 0.983 ? 1 - t * t * t : 0.1;             // :164
                    break;
                }
                lambda *= mu;                                      // :169-170
                mu *= 2.0;
            }
        }
        if (opt.iterator != 1) maxstep = maxabs_x();            // maximum(abs, linsystem.x)  src/optimize.jl:149
        cost = cost_;
        // ---- back in optimizeinternal!                                      src/optimize.jl:128-165
        int64_t terminate = opt.callback_terminate;            // callback(cost, ...) -> (cost, terminate)
        if (callback_kind == 1) {                              // emcallback  test/adaptivecost.jl:15-25
            std::vector<double> sq;                            // :17  squared errors at varnext, cost storage order
            int64_t kvar = 0;
            for (size_t t = 0; t < costs.size(); ++t) if (costtypes[t] == RT_ADAPTIVE_OFFSET)
                for (const Cost& c : costs[t]) { const double r = varnext[(size_t)c.vi[1] - 1].v[0] - c.data[0]; sq.push_back(r * r); kvar = c.vi[0]; }
            if (kvar > 0) {
                em_optimize(varnext[(size_t)kvar - 1].v, sq.data(), (int64_t)sq.size());   // :19
                t0 = time_ns();
                cost = this->cost(varnext);                    // :21
                t_cost += time_ns() - t0;
                res.costcomputations += 1;                     // :22
            }
        }
        double dcost = bestcost - cost;
        if (dcost >= 0) { bestcost = cost; fails = 0; }
        else {
            dcost = cost;                                      // sic  :135
            fails += 1;
            if (fails == 1) varbest = variables;               // :137-144 (a swap when sizes match; content equal for our purposes)
        }
        std::swap(variables, varnext);                         // updatefromnext!  :147
        if (trace) trace->push_back(IterRecord{cost, opt.iterator == 2 ? trustradius : (opt.iterator == 3 ? stepsize : lambda), maxstep, ntries});
        converged = 0;
        converged |= (int64_t)std::isinf(cost) << 0;
        converged |= (int64_t)std::isnan(cost) << 1;
        converged |= (int64_t)(dcost < bestcost * opt.reldcost) << 2;
        converged |= (int64_t)(dcost < opt.absdcost) << 3;
        converged |= (int64_t)std::isinf(maxstep) << 4;
        converged |= (int64_t)std::isnan(maxstep) << 5;
        converged |= (int64_t)(maxstep < opt.dstep) << 6;
        converged |= (int64_t)(fails > opt.maxfails) << 7;
        converged |= (int64_t)(iternum >= opt.maxiters) << 8;
        converged |= (int64_t)(time_ns() > stoptime) << 9;
        converged |= terminate << 16;
        if (converged != 0) break;
        t0 = time_ns();
        zero();                                                // :168
        costgradhess();                                        // :169
        t_grad += time_ns() - t0;
        res.gradientcomputations += 1;
    }
    if (!(bestcost >= cost)) std::swap(variables, varbest);   // :173-176
    res.bestcost = bestcost;
    res.termination = converged;
    res.niterations = iternum;
    res.timetotal = (time_ns() - starttime) * 1e-9;
    res.timeinit = t_init * 1e-9; res.timecost = t_cost * 1e-9; res.timegradient = t_grad * 1e-9; res.timesolver = t_solver * 1e-9;
    return res;
}

// optimizesingles!                                                  src/optimize.jl:60-76,183-205
// Every variable in `indices` (the reference sorts them by size; callers pass one size) is optimised on its own: all other variables
// fixed, only the costs that depend on it (selectcosts!), a fresh iterator (reset!) per variable.  Restated with a sub-problem per
// variable that holds just those costs and the variables they touch.
int64_t Problem::optimizesingles(const Options& opt, const std::vector<int64_t>& indices) {
    std::vector<std::vector<std::pair<int, int64_t>>> dep(variables.size());   // var -> (cost type, position)
    for (size_t t = 0; t < costs.size(); ++t)
        for (size_t q = 0; q < costs[t].size(); ++q)
            for (int i = 0; i < costs[t][q].ndeps; ++i) dep[(size_t)costs[t][q].vi[i] - 1].push_back({(int)t, (int64_t)q});
    int64_t iters = 0;
    for (int64_t ind : indices) {
        Problem sp;
        std::vector<int64_t> l2g;
        auto local = [&](int64_t gidx) {
            for (size_t k = 0; k < l2g.size(); ++k) if (l2g[k] == gidx) return (int64_t)k + 1;
            l2g.push_back(gidx);
            sp.variables.push_back(variables[(size_t)gidx - 1]);
            return (int64_t)l2g.size();
        };
        for (const auto& d : dep[(size_t)ind - 1]) {
            Cost c = costs[(size_t)d.first][(size_t)d.second];
            for (int i = 0; i < c.ndeps; ++i) c.vi[i] = local(c.vi[i]);
            sp.addcost(c, kernels[(size_t)d.first]);
        }
        if (sp.variables.empty()) continue;
        sp.unfixed.assign(sp.variables.size(), 0);
        sp.unfixed[(size_t)local(ind) - 1] = 1;
        Result r = sp.optimize(opt, nullptr);
        iters += r.niterations;
        variables[(size_t)ind - 1] = sp.variables[(size_t)local(ind) - 1];
    }
    lsready = false;
    return iters;
}

}  // namespace orc
