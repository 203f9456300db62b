// Residual types with a registered fused kernel: residual, analytic Jacobian, variable update.
// Each functor replaces, for one concrete residual type, the reference's generic
//   computeresidual (src/residual.jl:13) + ForwardDiff computeresjac (src/autodiff.jl:78-93) + update (src/variable.jl).
// Jacobian columns follow getvars order: camera DoF first, then point DoF.
#pragma once
#include "common.cuh"

namespace nlls {

// SimpleError2{2,Float64,EuclideanVector{6},EuclideanVector{3}} with the affine generatemeasurement of
// test/optimizeba.jl:4:  r = (pose[1:3].X - z1, pose[4:6].X - z2);  J = [X' 0 p1'; 0 X' p2'].
struct AffineBA {
    static constexpr int M = 2;    // nres
    static constexpr int DC = 6;   // camera DoF
    static constexpr int NC = 6;   // camera stored doubles
    static constexpr int CS = 6;   // camera stride in device memory (doubles)
    // the only residual row that depends on camera DoF a (-1: both do).  The assembly kernels skip the products with the
    // structural zeros of Jc — exactly the values the dense formulas produce (x * 0 + y = y), a third fewer FP64 instructions
    __host__ __device__ static constexpr int jc_row(int a) { return a < 3 ? 0 : 1; }

    __device__ static __forceinline__ void load_cam(const double* __restrict__ cams, int cam, double c[NC]) {
        const double2* p = reinterpret_cast<const double2*>(cams + (size_t)cam * CS);
        double2 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
        c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y; c[4] = d.x; c[5] = d.y;
    }
    __device__ static __forceinline__ void residual(const double c[NC], const double X[3], double zx, double zy, double r[2]) {
        r[0] = (c[0] * X[0] + c[1] * X[1] + c[2] * X[2]) - zx;
        r[1] = (c[3] * X[0] + c[4] * X[1] + c[5] * X[2]) - zy;
    }
    // Jc[i][a] = d r_i / d cam_a ; Jp[i][b] = d r_i / d X_b
    __device__ static __forceinline__ void resjac(const double c[NC], const double X[3], double zx, double zy, double r[2],
                                                  double Jc[2][DC], double Jp[2][3]) {
        residual(c, X, zx, zy, r);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            Jc[0][a] = X[a]; Jc[0][a + 3] = 0.0;
            Jc[1][a] = 0.0;  Jc[1][a + 3] = X[a];
            Jp[0][a] = c[a]; Jp[1][a] = c[a + 3];
        }
    }
    // update(EuclideanVector, x) = v + x                                   src/variable.jl:10
    __device__ static __forceinline__ void update_cam(const double* c, const double* x, double* out) {
#pragma unroll
        for (int i = 0; i < 6; ++i) out[i] = c[i] + x[i];
    }
};

// Products with the camera Jacobian that skip its structural zeros (a, a2 are compile-time after unrolling).
//   jtr:  (Jc' r)[a]        jtj_cc:  (Jc' Jc)[a][a2]  (*zero = true: structurally zero)        jtj_pc:  (Jp' Jc)[b][a]
template <class R>
__device__ __forceinline__ double jtr(const double (*Jc)[R::DC], const double r[2], int a) {
    const int ra = R::jc_row(a);
    return ra >= 0 ? Jc[ra][a] * r[ra] : fma(Jc[1][a], r[1], Jc[0][a] * r[0]);
}
template <class R>
__device__ __forceinline__ double jtj_cc(const double (*Jc)[R::DC], int a, int a2, bool* zero) {
    const int ra = R::jc_row(a), rb = R::jc_row(a2);
    *zero = ra >= 0 && rb >= 0 && ra != rb;
    if (*zero) return 0.0;
    if (ra >= 0) return Jc[ra][a] * Jc[ra][a2];
    if (rb >= 0) return Jc[rb][a] * Jc[rb][a2];
    return fma(Jc[1][a], Jc[1][a2], Jc[0][a] * Jc[0][a2]);
}
template <class R>
__device__ __forceinline__ double jtj_pc(const double (*Jp)[3], const double (*Jc)[R::DC], int b, int a) {
    const int ra = R::jc_row(a);
    return ra >= 0 ? Jp[ra][b] * Jc[ra][a] : fma(Jp[1][b], Jc[1][a], Jp[0][b] * Jc[0][a]);
}

// Rodrigues formula, column-major 3x3 (same series switch as the oracle so that updates agree to rounding).
__device__ __forceinline__ void so3_exp(const double w[3], double E[9]) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    double A, B;
    if (th < 1e-5) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
    else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; }
    const double K[9] = {0, w[2], -w[1], -w[2], 0, w[0], w[1], -w[0], 0};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += K[i + 3 * k] * K[k + 3 * j];
            E[i + 3 * j] = A * K[i + 3 * j] + B * s + (i == j ? 1.0 : 0.0);
        }
}

// Repo-defined pinhole reprojection residual (the reference ships no SO(3)/pinhole types — SURVEY F2).
// Camera = R (world->camera rotation, column-major), t, f, k1, k2; BAL convention:
//   P = R X + t;  p = -P.xy / P.z;  r = f (1 + k1 |p|^2 + k2 |p|^4) p - z.
// Minimal 9-DoF update: R <- Exp(x[0:3]) R, (t, f, k1, k2) additive.
struct PinholeBA {
    static constexpr int M = 2;
    static constexpr int DC = 9;
    static constexpr int NC = 15;
    static constexpr int CS = 16;
    __host__ __device__ static constexpr int jc_row(int) { return -1; }   // dense camera Jacobian

    __device__ static __forceinline__ void load_cam(const double* __restrict__ cams, int cam, double c[NC]) {
        const double2* p = reinterpret_cast<const double2*>(cams + (size_t)cam * CS);
#pragma unroll
        for (int i = 0; i < 7; ++i) { double2 v = __ldg(p + i); c[2 * i] = v.x; c[2 * i + 1] = v.y; }
        c[14] = __ldg(cams + (size_t)cam * CS + 14);
    }
    __device__ static __forceinline__ void residual(const double c[NC], const double X[3], double zx, double zy, double r[2]) {
        const double P0 = c[0] * X[0] + c[3] * X[1] + c[6] * X[2] + c[9];
        const double P1 = c[1] * X[0] + c[4] * X[1] + c[7] * X[2] + c[10];
        const double P2 = c[2] * X[0] + c[5] * X[1] + c[8] * X[2] + c[11];
        const double iz = -1.0 / P2;
        const double px = P0 * iz, py = P1 * iz;
        const double n2 = px * px + py * py;
        const double dist = 1.0 + n2 * (c[13] + c[14] * n2);
        const double s = c[12] * dist;
        r[0] = s * px - zx;
        r[1] = s * py - zy;
    }
    __device__ static __forceinline__ void resjac(const double c[NC], const double X[3], double zx, double zy, double r[2],
                                                  double Jc[2][DC], double Jp[2][3]) {
        const double Q0 = c[0] * X[0] + c[3] * X[1] + c[6] * X[2];
        const double Q1 = c[1] * X[0] + c[4] * X[1] + c[7] * X[2];
        const double Q2 = c[2] * X[0] + c[5] * X[1] + c[8] * X[2];
        const double P0 = Q0 + c[9], P1 = Q1 + c[10], P2 = Q2 + c[11];
        const double f = c[12], k1 = c[13], k2 = c[14];
        const double iz = -1.0 / P2;
        const double px = P0 * iz, py = P1 * iz;
        const double n2 = px * px + py * py;
        const double dist = 1.0 + n2 * (k1 + k2 * n2);
        const double s = f * dist;
        r[0] = s * px - zx;
        r[1] = s * py - zy;
        const double fdd = f * 2.0 * (k1 + 2.0 * k2 * n2);
        // d r / d p
        const double a00 = s + fdd * px * px, a01 = fdd * px * py, a11 = s + fdd * py * py;
        // G = (d r / d p)(d p / d P),  d p / d P = [[iz, 0, px iz], [0, iz, py iz]]
        double G[2][3];
        G[0][0] = a00 * iz; G[0][1] = a01 * iz; G[0][2] = (a00 * px + a01 * py) * iz;
        G[1][0] = a01 * iz; G[1][1] = a11 * iz; G[1][2] = (a01 * px + a11 * py) * iz;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            // d P / d w = -[Q]x = [[0, Q2, -Q1], [-Q2, 0, Q0], [Q1, -Q0, 0]]
            Jc[i][0] = -G[i][1] * Q2 + G[i][2] * Q1;
            Jc[i][1] = G[i][0] * Q2 - G[i][2] * Q0;
            Jc[i][2] = -G[i][0] * Q1 + G[i][1] * Q0;
            Jc[i][3] = G[i][0]; Jc[i][4] = G[i][1]; Jc[i][5] = G[i][2];
            const double p = (i == 0) ? px : py;
            Jc[i][6] = dist * p;
            Jc[i][7] = f * n2 * p;
            Jc[i][8] = f * n2 * n2 * p;
#pragma unroll
            for (int b = 0; b < 3; ++b) Jp[i][b] = G[i][0] * c[0 + 3 * b] + G[i][1] * c[1 + 3 * b] + G[i][2] * c[2 + 3 * b];
        }
    }
    __device__ static __forceinline__ void update_cam(const double* c, const double* x, double* out) {
        double E[9];
        const double w[3] = {x[0], x[1], x[2]};
        so3_exp(w, E);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += E[i + 3 * k] * c[k + 3 * j];
                out[i + 3 * j] = s;
            }
#pragma unroll
        for (int i = 0; i < 6; ++i) out[9 + i] = c[9 + i] + x[3 + i];
        out[15] = 0.0;
    }
};

}  // namespace nlls
