// Reduced camera system S x = rhs: tile-sparse storage + level-scheduled tile LDL' (no pivoting) + triangular solves.
//
// S (DC nA square, symmetric) is cut into ST x ST tiles aligned to camera blocks (ST = 72 = 12 affine or 8 pinhole cameras).
// Camera tiles are renumbered by a fill-reducing / parallelism-exposing order computed on the host (nested dissection by
// index when the tile pattern is banded, identity otherwise).  Only tiles that are structurally non-zero after symbolic
// tile-level fill-in are stored (lower triangle in the permuted numbering, column-major inside a tile, tile `id` at
// S + id*ST*ST).  The factorisation S = L D L' uses no pivoting — the algebra of the reference's sparse path
// (LDLFactorizations.ldl_factorize!, src/linearsolver.jl:29), so indefinite systems (Triggs-corrected robust Hessians)
// follow the same trajectory instead of failing over.
//
// Tile columns are grouped into levels of the elimination tree; all columns of a level are independent and are processed
// by the same launches (task lists are built at prepare time):
//   ldl_diag_kernel    T_JJ = L_JJ D_J L_JJ'                      (1 CTA per column of the level)
//   ldl_trsm_kernel    L_IJ = T_IJ L_JJ^-T D_J^-1  for I > J      (1 CTA per tile)
//   ldl_update_kernel  T_{I1,I2} -= L_{I1,J} D_J L_{I2,J}'        (1 CTA per pair; tasks hitting the same target tile are
//                                                                 split into rounds = separate launches, fixed order)
// then ldl_inv_kernel (all diagonal tiles) and the level-scheduled sweeps ldl_fwd_kernel / ldl_bwd_kernel.
//
// Thread layout of the tile kernels: 256 threads as 16 x 16; thread (ty, tx) owns the 5 x 5 elements
// (ty + 16 a, tx + 16 b) of a 72 x 72 tile in registers (a = 4 exists only for ty < 8, likewise b for tx).
#pragma once
#include "common.cuh"

namespace nlls {

constexpr int ST = 72;
constexpr int ST2 = ST * ST;
constexpr int RED_THREADS = 256;
constexpr int LDT = ST + 1;   // padded leading dimension of smem tiles
constexpr int RB = 5;         // register block edge: ceil(72 / 16)

struct RedSolveLists {        // device pointers for the triangular sweeps
    const int* diag_tile;     // [NT] tile id of (J, J), permuted numbering
    const int* rowptr;        // [NT + 1] tiles (J, K), K < J, of block row J
    const int* row_tile;
    const int* row_col;
    const int* colptr;        // [NT + 1] tiles (I, J), I > J, of block column J
    const int* col_tile;
    const int* col_row;
};

constexpr int NB = 8;            // panel width of the blocked tile algorithms
constexpr int LDP = NB + 1;      // padded leading dimension of panel buffers

// ---------------------------------------------------------------------------------------------------
// In-place LDL' of diagonal tiles, blocked by panels of NB columns.  On exit the strict lower triangle holds L (unit
// diagonal implied) and the diagonal holds D; the upper triangle is not referenced.
// Per panel: (1) owners publish the panel, (2) 8 lanes factor the 8 x 8 diagonal block, (3) one thread per row below it
// solves its row of L21, (4) the panel is written out, (5) all threads apply the rank-8 update to their registers.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) ldl_diag_kernel(double* __restrict__ S, const int* __restrict__ tasks) {
    __shared__ double Up[ST * LDP];   // panel of A, then U = L D
    __shared__ double Lp[ST * LDP];   // panel of L
    __shared__ double dd[NB], rdd[NB], col[NB];
    double* T = S + (size_t)tasks[blockIdx.x] * ST2;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double R[RB][RB];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int i = ty + 16 * a, k = tx + 16 * b;
            R[a][b] = (i < ST && k < ST && i >= k) ? T[i + ST * k] : 0.0;
        }
    for (int j0 = 0; j0 < ST; j0 += NB) {
        const int jb = j0 >> 4, jx0 = j0 & 15;
        // (1) owners of columns j0 .. j0+7 publish rows >= j0
        if (tx >= jx0 && tx < jx0 + NB) {
            const int t = tx - jx0;
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if (b == jb) {
#pragma unroll
                    for (int a = 0; a < RB; ++a) { const int i = ty + 16 * a; if (i >= j0 && i < ST) Up[i * LDP + t] = R[a][b]; }
                }
        }
        __syncthreads();
        // (2) 8 x 8 diagonal block: lane r owns row r
        if (tid < NB) {
            const int r = tid;
            double A[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) A[c] = (c <= r) ? Up[(j0 + r) * LDP + c] : 0.0;
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                if (r >= t) col[r] = A[t];
                __syncwarp(0xffu);
                const double d = col[t];
                if (r > t) {
                    const double l = A[t] / d;
#pragma unroll
                    for (int c = t + 1; c < NB; ++c) if (c <= r) A[c] = fma(-l, col[c], A[c]);
                    A[t] = l;
                }
                __syncwarp(0xffu);
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) if (c < r) Lp[(j0 + r) * LDP + c] = A[c];
            dd[r] = A[r];
            rdd[r] = 1.0 / A[r];
        }
        __syncthreads();
        // (3) rows below the block: U21 L11' = A21, L21 = U21 D11^-1
        if (tid < ST && tid >= j0 + NB) {
            const int i = tid;
            double uu[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                double v = Up[i * LDP + t];
#pragma unroll
                for (int q = 0; q < t; ++q) v = fma(-uu[q], Lp[(j0 + t) * LDP + q], v);
                uu[t] = v;
            }
#pragma unroll
            for (int t = 0; t < NB; ++t) { Up[i * LDP + t] = uu[t]; Lp[i * LDP + t] = uu[t] * rdd[t]; }
        }
        __syncthreads();
        // (4) write the finished panel columns
        for (int e = tid; e < ST * NB; e += RED_THREADS) {
            const int i = e % ST, t = e / ST, k = j0 + t;
            if (i > k) T[i + ST * k] = Lp[i * LDP + t];
            else if (i == k) T[i + ST * k] = dd[t];
        }
        // (5) trailing update A22 -= L21 U21'
        if (j0 + NB < ST) {
#pragma unroll
            for (int a = 0; a < RB; ++a) {
                const int i = ty + 16 * a;
                if (16 * a + 15 < j0 + NB || i >= ST) continue;
                double lv[NB];
#pragma unroll
                for (int t = 0; t < NB; ++t) lv[t] = (i >= j0 + NB) ? Lp[i * LDP + t] : 0.0;
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int k = tx + 16 * b;
                    if (b > a || 16 * b + 15 < j0 + NB) continue;
                    if (k >= j0 + NB && k <= i) {
                        double acc = R[a][b];
#pragma unroll
                        for (int t = 0; t < NB; ++t) acc = fma(-lv[t], Up[k * LDP + t], acc);
                        R[a][b] = acc;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// L_IJ = T_IJ L_JJ^-T D_J^-1, blocked: per panel the final columns X[:, P] = (X[:, P] raw) L11^-T (one thread per row), then
// all threads apply X[:, >P] -= X[:, P] L[>P, P]'.  Finally column k is scaled by 1 / D_k.
__global__ void __launch_bounds__(RED_THREADS) ldl_trsm_kernel(double* __restrict__ S, const int2* __restrict__ tasks) {
    extern __shared__ double sm[];
    double* L = sm;                  // L_JJ: L[k * LDT + j] (strict lower) with D on the diagonal
    double* Xp = sm + ST * LDT;      // [ST][LDP] panel of X
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int2 tk = tasks[blockIdx.x];
    double* T = S + (size_t)tk.x * ST2;
    const double* D = S + (size_t)tk.y * ST2;
    for (int e = tid; e < ST2; e += RED_THREADS) { const int r = e % ST, c = e / ST; L[r * LDT + c] = D[e]; }
    double R[RB][RB];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            R[a][b] = (r < ST && c < ST) ? T[r + ST * c] : 0.0;
        }
    __syncthreads();
    for (int j0 = 0; j0 < ST; j0 += NB) {
        const int jb = j0 >> 4, jx0 = j0 & 15;
        const bool owner = tx >= jx0 && tx < jx0 + NB;
        if (owner) {
            const int t = tx - jx0;
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if (b == jb) {
#pragma unroll
                    for (int a = 0; a < RB; ++a) { const int r = ty + 16 * a; if (r < ST) Xp[r * LDP + t] = R[a][b]; }
                }
        }
        __syncthreads();
        if (tid < ST) {              // row tid: x_t = raw_t - sum_{q<t} x_q L[j0+t][j0+q]
            double x[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                double v = Xp[tid * LDP + t];
#pragma unroll
                for (int q = 0; q < t; ++q) v = fma(-x[q], L[(j0 + t) * LDT + j0 + q], v);
                x[t] = v;
            }
#pragma unroll
            for (int t = 0; t < NB; ++t) Xp[tid * LDP + t] = x[t];
        }
        __syncthreads();
        if (owner) {                 // owners take the final panel values back
            const int t = tx - jx0;
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if (b == jb) {
#pragma unroll
                    for (int a = 0; a < RB; ++a) { const int r = ty + 16 * a; if (r < ST) R[a][b] = Xp[r * LDP + t]; }
                }
        }
        if (j0 + NB < ST) {          // X[:, k] -= sum_t X[:, j0+t] L[k][j0+t] for k >= j0 + NB
#pragma unroll
            for (int a = 0; a < RB; ++a) {
                const int r = ty + 16 * a;
                if (r >= ST) continue;
                double xv[NB];
#pragma unroll
                for (int t = 0; t < NB; ++t) xv[t] = Xp[r * LDP + t];
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int k = tx + 16 * b;
                    if (16 * b + 15 < j0 + NB) continue;
                    if (k >= j0 + NB && k < ST) {
                        double acc = R[a][b];
#pragma unroll
                        for (int t = 0; t < NB; ++t) acc = fma(-xv[t], L[k * LDT + j0 + t], acc);
                        R[a][b] = acc;
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            if (r < ST && c < ST) T[r + ST * c] = R[a][b] / L[c * LDT + c];
        }
}

// T_{I1,I2} -= (L_{I1,J} D_J) L_{I2,J}'
__global__ void __launch_bounds__(RED_THREADS) ldl_update_kernel(double* __restrict__ S, const int4* __restrict__ tasks) {
    extern __shared__ double sm[];
    double* A = sm;                 // (L_{I1,J} D)  A[k * LDT + r]
    double* B = sm + ST * LDT;      // L_{I2,J}      B[k * LDT + c]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int4 tk = tasks[blockIdx.x];
    const double* Ta = S + (size_t)tk.x * ST2;
    const double* Tb = S + (size_t)tk.y * ST2;
    double* Tc = S + (size_t)tk.z * ST2;
    const double* D = S + (size_t)tk.w * ST2;
    for (int e = tid; e < ST2; e += RED_THREADS) {
        const int r = e % ST, k = e / ST;
        A[k * LDT + r] = Ta[e] * D[k * ST + k];
        B[k * LDT + r] = Tb[e];
    }
    __syncthreads();
    double acc[RB][RB];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) acc[a][b] = 0.0;
    for (int k = 0; k < ST; ++k) {
        double av[RB], bv[RB];
#pragma unroll
        for (int a = 0; a < RB; ++a) { const int r = ty + 16 * a; av[a] = (r < ST) ? A[k * LDT + r] : 0.0; }
#pragma unroll
        for (int b = 0; b < RB; ++b) { const int c = tx + 16 * b; bv[b] = (c < ST) ? B[k * LDT + c] : 0.0; }
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b = 0; b < RB; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            if (r < ST && c < ST) Tc[r + ST * c] -= acc[a][b];
        }
}

// Linv_J = L_JJ^-1 (unit lower triangular, explicit ones on the diagonal, zeros above) for every diagonal tile, in parallel
// after the factorisation; turns the triangular sweeps of the solve into mat-vecs.  Blocked forward substitution of
// L X = I by row panels: X[P, :] = L11^-1 X[P, :] (one thread per column), then X[>P, :] -= L[>P, P] X[P, :].
__global__ void __launch_bounds__(RED_THREADS) ldl_inv_kernel(const double* __restrict__ S, const int* __restrict__ diag_tile, double* __restrict__ Linv) {
    extern __shared__ double sm[];
    double* L = sm;                       // L[i * LDT + j]
    double* Xp = sm + ST * LDT;           // [NB][LDT] row panel of X
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, J = blockIdx.x;
    const double* T = S + (size_t)diag_tile[J] * ST2;
    for (int e = tid; e < ST2; e += RED_THREADS) { const int r = e % ST, c = e / ST; L[r * LDT + c] = T[e]; }
    double R[RB][RB];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) R[a][b] = (ty + 16 * a == tx + 16 * b) ? 1.0 : 0.0;
    __syncthreads();
    for (int j0 = 0; j0 < ST; j0 += NB) {
        const int ja = j0 >> 4, jy0 = j0 & 15;
        const bool owner = ty >= jy0 && ty < jy0 + NB;
        if (owner) {
            const int t = ty - jy0;
#pragma unroll
            for (int a = 0; a < RB; ++a)
                if (a == ja) {
#pragma unroll
                    for (int b = 0; b < RB; ++b) { const int k = tx + 16 * b; if (k < ST) Xp[t * LDT + k] = R[a][b]; }
                }
        }
        __syncthreads();
        if (tid < ST) {              // column tid: x_t = raw_t - sum_{q<t} L[j0+t][j0+q] x_q
            double x[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                double v = Xp[t * LDT + tid];
#pragma unroll
                for (int q = 0; q < t; ++q) v = fma(-L[(j0 + t) * LDT + j0 + q], x[q], v);
                x[t] = v;
            }
#pragma unroll
            for (int t = 0; t < NB; ++t) Xp[t * LDT + tid] = x[t];
        }
        __syncthreads();
        if (owner) {
            const int t = ty - jy0;
#pragma unroll
            for (int a = 0; a < RB; ++a)
                if (a == ja) {
#pragma unroll
                    for (int b = 0; b < RB; ++b) { const int k = tx + 16 * b; if (k < ST) R[a][b] = Xp[t * LDT + k]; }
                }
        }
        if (j0 + NB < ST) {          // X[i, :] -= sum_t L[i][j0+t] X[j0+t, :] for i >= j0 + NB
#pragma unroll
            for (int a = 0; a < RB; ++a) {
                const int i = ty + 16 * a;
                if (16 * a + 15 < j0 + NB || i >= ST || i < j0 + NB) continue;
                double lv[NB];
#pragma unroll
                for (int t = 0; t < NB; ++t) lv[t] = L[i * LDT + j0 + t];
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int k = tx + 16 * b;
                    if (k < ST) {
                        double acc = R[a][b];
#pragma unroll
                        for (int t = 0; t < NB; ++t) acc = fma(-lv[t], Xp[t * LDT + k], acc);
                        R[a][b] = acc;
                    }
                }
            }
        }
        __syncthreads();
    }
    double* out = Linv + (size_t)J * ST2;
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int i = ty + 16 * a, k = tx + 16 * b;
            if (i < ST && k < ST) out[i + ST * k] = R[a][b];
        }
}

// ---------------------------------------------------------------------------------------------------
// Triangular sweeps, one CTA per tile column of the current level ("pull" form: a CTA only writes its own x_J).
//   forward : y_J = Linv_J (x_J - sum_{K<J} L_JK y_K)
//   backward: x_J = Linv_J' (y_J / D_J - sum_{I>J} L_IJ' x_I)
// 216 threads = (row, k-third); x lives in global memory (length NT*ST, permuted numbering).
// ---------------------------------------------------------------------------------------------------
constexpr int SOLVE_THREADS = 3 * ST;

__device__ __forceinline__ double gemv_part(const double* __restrict__ M, const double* xin, int r, int seg, bool transposed) {
    double s = 0.0;
    if (!transposed) { for (int c = seg * 24; c < seg * 24 + 24; ++c) s = fma(M[r + ST * c], xin[c], s); }
    else { for (int k = seg * 24; k < seg * 24 + 24; ++k) s = fma(M[k + ST * r], xin[k], s); }
    return s;
}

__global__ void __launch_bounds__(SOLVE_THREADS) ldl_fwd_kernel(const double* __restrict__ S, const double* __restrict__ Linv, RedSolveLists t,
                                                                const int* __restrict__ cols, double* __restrict__ x) {
    __shared__ double xj[ST], xk[ST], part[SOLVE_THREADS];
    const int tid = threadIdx.x, r = tid % ST, seg = tid / ST;
    const int J = cols[blockIdx.x];
    if (tid < ST) xj[tid] = x[(size_t)J * ST + tid];
    for (int q = t.rowptr[J]; q < t.rowptr[J + 1]; ++q) {
        __syncthreads();
        if (tid < ST) xk[tid] = x[(size_t)t.row_col[q] * ST + tid];
        __syncthreads();
        part[tid] = gemv_part(S + (size_t)t.row_tile[q] * ST2, xk, r, seg, false);
        __syncthreads();
        if (tid < ST) xj[tid] -= (part[tid] + part[tid + ST]) + part[tid + 2 * ST];
    }
    __syncthreads();
    part[tid] = gemv_part(Linv + (size_t)J * ST2, xj, r, seg, false);
    __syncthreads();
    if (tid < ST) x[(size_t)J * ST + tid] = (part[tid] + part[tid + ST]) + part[tid + 2 * ST];
}

__global__ void __launch_bounds__(SOLVE_THREADS) ldl_bwd_kernel(const double* __restrict__ S, const double* __restrict__ Linv, RedSolveLists t,
                                                                const int* __restrict__ cols, double* __restrict__ x) {
    __shared__ double xj[ST], xi[ST], part[SOLVE_THREADS];
    const int tid = threadIdx.x, r = tid % ST, seg = tid / ST;
    const int J = cols[blockIdx.x];
    if (tid < ST) xj[tid] = x[(size_t)J * ST + tid] / S[(size_t)t.diag_tile[J] * ST2 + tid + ST * tid];
    for (int q = t.colptr[J]; q < t.colptr[J + 1]; ++q) {
        __syncthreads();
        if (tid < ST) xi[tid] = x[(size_t)t.col_row[q] * ST + tid];
        __syncthreads();
        part[tid] = gemv_part(S + (size_t)t.col_tile[q] * ST2, xi, r, seg, true);
        __syncthreads();
        if (tid < ST) xj[tid] -= (part[tid] + part[tid + ST]) + part[tid + 2 * ST];
    }
    __syncthreads();
    part[tid] = gemv_part(Linv + (size_t)J * ST2, xj, r, seg, true);
    __syncthreads();
    if (tid < ST) x[(size_t)J * ST + tid] = (part[tid] + part[tid + ST]) + part[tid + 2 * ST];
}

// natural <-> permuted tile numbering of the right-hand side / solution
__global__ void red_permute_kernel(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ pos, int NT, int to_permuted) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= NT * ST) return;
    const int o = k / ST, r = k - o * ST;
    if (to_permuted) dst[(size_t)pos[o] * ST + r] = src[k];
    else dst[k] = src[(size_t)pos[o] * ST + r];
}

// S tiles <- 0 (memset) except diagonal tiles: lower triangle of U_c + lambda I (rank 0 only in a multi-rank run);
// padding rows get a unit diagonal.  rhs (natural numbering) <- g_c.
template <int DC>
__global__ void red_init_kernel(double* __restrict__ S, const int* __restrict__ diag_tile_nat, const double* __restrict__ H, const double* __restrict__ g,
                                double* __restrict__ rhs, int nA, double lambda, int add_u) {
    constexpr int TC = ST / DC;
    const int o = blockIdx.x;   // natural tile index
    double* T = S + (size_t)diag_tile_nat[o] * ST2;
    for (int e = threadIdx.x; e < ST2; e += blockDim.x) {
        const int r = e % ST, c = e / ST;
        const int cr = r / DC, cc = c / DC;
        const long long cam = (long long)o * TC + cr;
        double v = 0.0;
        if (cam >= nA) { v = (r == c && add_u) ? 1.0 : 0.0; }
        else if (cr == cc && r >= c && add_u) {
            v = H[(size_t)DC * DC * cam + (r - cr * DC) + DC * (c - cc * DC)];
            if (r == c) v += lambda;
        }
        T[e] = v;
    }
    for (int r = threadIdx.x; r < ST; r += blockDim.x) {
        const long long k = (long long)o * ST + r;
        rhs[k] = (add_u && k < (long long)DC * nA) ? g[k] : 0.0;
    }
}

}  // namespace nlls
