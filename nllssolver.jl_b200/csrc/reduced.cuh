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

// ---------------------------------------------------------------------------------------------------
// In-place LDL' of diagonal tiles.  On exit the strict lower triangle holds L (unit diagonal implied) and the diagonal
// holds D; the upper triangle is not referenced.  Register-resident right-looking elimination, one barrier per pivot.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) ldl_diag_kernel(double* __restrict__ S, const int* __restrict__ tasks) {
    __shared__ double u[2][ST];      // column j below the diagonal before scaling (= L D), double-buffered
    __shared__ double dj[2];
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
    for (int j = 0; j < ST; ++j) {
        const int jb = j >> 4, jx = j & 15, buf = j & 1;
        if (tx == jx) {              // owners of column j publish it
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if (b == jb) {
#pragma unroll
                    for (int a = 0; a < RB; ++a) {
                        const int i = ty + 16 * a;
                        if (i < ST && i > j) u[buf][i] = R[a][b];
                        if (i == j) dj[buf] = R[a][b];
                    }
                }
        }
        __syncthreads();
        const double rd = 1.0 / dj[buf];
        double li[RB], uk[RB];
#pragma unroll
        for (int a = 0; a < RB; ++a) { const int i = ty + 16 * a; li[a] = (i < ST && i > j) ? u[buf][i] * rd : 0.0; }
#pragma unroll
        for (int b = 0; b < RB; ++b) { const int k = tx + 16 * b; uk[b] = (k < ST && k > j) ? u[buf][k] : 0.0; }
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const int i = ty + 16 * a, k = tx + 16 * b;
                if (i >= k) R[a][b] = fma(-li[a], uk[b], R[a][b]);   // uk = 0 for k <= j, li = 0 for i <= j
                if (k == j && i > j && i < ST) R[a][b] = li[a];       // store L in column j
            }
    }
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int i = ty + 16 * a, k = tx + 16 * b;
            if (i < ST && k < ST && i >= k) T[i + ST * k] = R[a][b];
        }
}

// L_IJ = T_IJ L_JJ^-T D_J^-1 : column sweep X[:,k] -= X[:,j] L[k][j] (X in registers), then scale column k by 1/D_k.
__global__ void __launch_bounds__(RED_THREADS) ldl_trsm_kernel(double* __restrict__ S, const int2* __restrict__ tasks) {
    extern __shared__ double sm[];
    double* L = sm;                  // L_JJ: L[k * LDT + j] (strict lower) with D on the diagonal
    double* xc = sm + ST * LDT;      // [2][ST] published column of X
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
    for (int j = 0; j < ST - 1; ++j) {
        const int jb = j >> 4, jx = j & 15, buf = j & 1;
        if (tx == jx) {
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if (b == jb) {
#pragma unroll
                    for (int a = 0; a < RB; ++a) { const int r = ty + 16 * a; if (r < ST) xc[buf * ST + r] = R[a][b]; }
                }
        }
        __syncthreads();
        double xr[RB], lk[RB];
#pragma unroll
        for (int a = 0; a < RB; ++a) { const int r = ty + 16 * a; xr[a] = (r < ST) ? xc[buf * ST + r] : 0.0; }
#pragma unroll
        for (int b = 0; b < RB; ++b) { const int k = tx + 16 * b; lk[b] = (k < ST && k > j) ? L[k * LDT + j] : 0.0; }
#pragma unroll
        for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b = 0; b < RB; ++b) R[a][b] = fma(-xr[a], lk[b], R[a][b]);
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
// after the factorisation; turns the triangular sweeps of the solve into mat-vecs.
__global__ void __launch_bounds__(RED_THREADS) ldl_inv_kernel(const double* __restrict__ S, const int* __restrict__ diag_tile, double* __restrict__ Linv) {
    extern __shared__ double sm[];
    double* L = sm;
    double* X = sm + ST * LDT;            // X[i * LDT + c]: column c of the inverse
    const int tid = threadIdx.x, J = blockIdx.x;
    const double* T = S + (size_t)diag_tile[J] * ST2;
    for (int e = tid; e < ST2; e += RED_THREADS) { const int r = e % ST, c = e / ST; L[r * LDT + c] = T[e]; }
    __syncthreads();
    if (tid < ST) {                       // column tid of the inverse: forward substitution of L x = e_tid
        const int c = tid;
        for (int i = 0; i < ST; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s = fma(-L[i * LDT + k], X[k * LDT + c], s);
            X[i * LDT + c] = (i < c) ? 0.0 : s;
        }
    }
    __syncthreads();
    double* out = Linv + (size_t)J * ST2;
    for (int e = tid; e < ST2; e += RED_THREADS) { const int r = e % ST, c = e / ST; out[e] = X[r * LDT + c]; }
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
