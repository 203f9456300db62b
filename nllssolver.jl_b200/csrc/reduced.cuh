// Reduced camera system S x = rhs: tile-sparse storage + level-scheduled, left-looking tile LDL' (no pivoting) + sweeps.
//
// S (DC nA square, symmetric) is cut into ST x ST tiles aligned to camera blocks (ST = 72 = 12 affine or 8 pinhole cameras).
// Camera tiles are renumbered by a fill-reducing / parallelism-exposing order computed on the host (nested dissection by
// index when the tile pattern is banded, identity otherwise).  Only tiles that are structurally non-zero after symbolic
// tile-level fill-in are stored (lower triangle in the permuted numbering, column-major inside a tile, tile `id` at
// S + id*ST*ST).  The factorisation S = L D L' uses no pivoting — the algebra of the reference's sparse path
// (LDLFactorizations.ldl_factorize!, src/linearsolver.jl:29), so indefinite systems (Triggs-corrected robust Hessians)
// follow the same trajectory instead of failing over.
//
// Tile columns are grouped into levels of the elimination tree; per level two launches (task lists built at prepare time):
//   ldl_tile_kernel<true>   one CTA per column J of the level:  T_JJ -= sum_K L_JK D_K L_JK'  (left-looking gather of every
//                           update, accumulated in registers), then in-register LDL' of the 72 x 72 tile together with
//                           Linv_J = L_JJ^-1 (the row operations applied to an identity), then the forward substitution of the
//                           right-hand side  y_J = Linv_J (b_J - sum_{K<J} L_JK y_K)
//   ldl_tile_kernel<false>  one CTA per tile (I, J), I > J:  T_IJ -= sum_K L_IK D_K L_JK';  L_IJ = T_IJ Linv_J' D_J^-1  (a GEMM)
// then the backward sweep ldl_bwd_kernel, one launch per level in reverse order.
//
// Thread layout of the tile kernels: 144 threads as 12 x 12; thread (ty, tx) owns the 6 x 6 block (6 ty + a, 6 tx + b).
#pragma once
#include "common.cuh"

namespace nlls {

constexpr int ST = 72;
constexpr int ST2 = ST * ST;
constexpr int TB = 6;                 // register block edge
constexpr int TG = ST / TB;           // 12 x 12 thread grid
constexpr int RED_THREADS = TG * TG;  // 144

struct RedSolveLists {        // device pointers for the triangular sweeps
    const int* diag_tile;     // [NT] tile id of (J, J), permuted numbering
    const int* rowptr;        // [NT + 1] tiles (J, K), K < J, of block row J
    const int* row_tile;
    const int* row_col;
    const int* colptr;        // [NT + 1] tiles (I, J), I > J, of block column J
    const int* col_tile;
    const int* col_row;
};

struct RedTask {              // one CTA of ldl_tile_kernel
    int tile;                 // target tile id
    int dtile;                // diagonal tile of the target's column J
    int col;                  // J (permuted numbering)
    int upd0, upd1;           // range of its updates in the RedUpd array
    int pad0, pad1, pad2;
};
struct RedUpd { int a, b, dk, pad; };   // tiles L_IK, L_JK and the diagonal tile of K

// reciprocal to ~1 ulp: MUFU seed + two Newton steps (the IEEE division is ~3x longer and sits on the 72-pivot critical path)
__device__ __forceinline__ double rcp_fast(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    return x;
}

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// whole tile (41 472 B, contiguous) global -> shared, 16 bytes per request
__device__ __forceinline__ void tile_to_smem(double* sdst, const double* gsrc) {
    for (int q = threadIdx.x; q < ST2 / 2; q += RED_THREADS) cp_async16(sdst + 2 * q, gsrc + 2 * q);
}

struct RedSmem {
    static constexpr size_t bytes = (size_t)(3 * ST2 + 8 * ST) * sizeof(double);
};

template <bool DIAG>
__global__ void __launch_bounds__(RED_THREADS) ldl_tile_kernel(double* __restrict__ S, double* __restrict__ Linv, const RedTask* __restrict__ tasks,
                                                               const RedUpd* __restrict__ upds, RedSolveLists lists, double* __restrict__ xp) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                  // [k][i]  (a tile as stored: column-major)
    double* Bs = sm + ST2;
    double* Cs = sm + 2 * ST2;        // Linv_J (off-diagonal tasks) / partial sums (diagonal tasks)
    double* Ds = sm + 3 * ST2;        // [ST] diagonal of D_K / D_J
    double* colA = Ds + ST;           // [2][ST] published column of A
    double* rowM = colA + 2 * ST;     // [2][ST] published row of M
    double* vs = rowM + 2 * ST;       // [ST] right-hand side segment
    const int tid = threadIdx.x, tx = tid % TG, ty = tid / TG;
    const RedTask tk = tasks[blockIdx.x];
    double* T = S + (size_t)tk.tile * ST2;
    const bool work = DIAG ? (ty >= tx) : true;

    if (!DIAG) tile_to_smem(Cs, Linv + (size_t)tk.col * ST2);   // ready since the previous launch; lands while the updates run

    // ---- forward substitution, first half: v = b_J - sum_{K<J} L_JK y_K (rows coalesced, k split in two halves)
    if (DIAG) {
        const int r = tid % ST, h = tid / ST;
        double acc = 0.0;
        for (int q = lists.rowptr[tk.col]; q < lists.rowptr[tk.col + 1]; ++q) {
            const double* M = S + (size_t)lists.row_tile[q] * ST2 + r + (size_t)ST * (ST / 2) * h;
            const double* y = xp + (size_t)lists.row_col[q] * ST + (ST / 2) * h;
            double m[ST / 2];
#pragma unroll
            for (int k = 0; k < ST / 2; ++k) m[k] = M[(size_t)ST * k];
#pragma unroll
            for (int k = 0; k < ST / 2; ++k) acc = fma(m[k], y[k], acc);
        }
        Cs[h * ST + r] = acc;
    }

    // ---- target block into registers
    double R[TB][TB];
#pragma unroll
    for (int b = 0; b < TB; ++b) {
        const double2* src = reinterpret_cast<const double2*>(T + (size_t)ST * (TB * tx + b) + TB * ty);
#pragma unroll
        for (int a = 0; a < TB / 2; ++a) { const double2 v = work ? src[a] : make_double2(0.0, 0.0); R[2 * a][b] = v.x; R[2 * a + 1][b] = v.y; }
    }

    // ---- left-looking updates:  R -= (L_IK D_K) L_JK'
    for (int u = tk.upd0; u < tk.upd1; ++u) {
        const RedUpd up = upds[u];
        __syncthreads();                                   // previous As / Bs / Ds readers are done
        tile_to_smem(As, S + (size_t)up.a * ST2);
        if (up.b != up.a) tile_to_smem(Bs, S + (size_t)up.b * ST2);
        if (tid < ST) Ds[tid] = S[(size_t)up.dk * ST2 + (size_t)(ST + 1) * tid];
        cp_async_wait_all();
        __syncthreads();
        const double* Bt = (up.b != up.a) ? Bs : As;
        if (work) {
#pragma unroll 4
            for (int k = 0; k < ST; ++k) {
                const double dk = Ds[k];
                const double2* ap = reinterpret_cast<const double2*>(As + k * ST + TB * ty);
                const double2* bp = reinterpret_cast<const double2*>(Bt + k * ST + TB * tx);
                double av[TB], bv[TB];
#pragma unroll
                for (int a = 0; a < TB / 2; ++a) { const double2 v = ap[a]; av[2 * a] = v.x * dk; av[2 * a + 1] = v.y * dk; }
#pragma unroll
                for (int b = 0; b < TB / 2; ++b) { const double2 v = bp[b]; bv[2 * b] = v.x; bv[2 * b + 1] = v.y; }
#pragma unroll
                for (int a = 0; a < TB; ++a)
#pragma unroll
                    for (int b = 0; b < TB; ++b) R[a][b] = fma(-av[a], bv[b], R[a][b]);
            }
        }
    }

    if constexpr (!DIAG) {
        // ---- L_IJ = R Linv_J' D_J^-1 :  X[i][c] = sum_{k <= c} R[i][k] Linv[c][k] / d_c
        __syncthreads();
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            double2* dst = reinterpret_cast<double2*>(As + (TB * tx + b) * ST + TB * ty);
#pragma unroll
            for (int a = 0; a < TB / 2; ++a) dst[a] = make_double2(R[2 * a][b], R[2 * a + 1][b]);
        }
        if (tid < ST) Ds[tid] = S[(size_t)tk.dtile * ST2 + (size_t)(ST + 1) * tid];
        cp_async_wait_all();
        __syncthreads();
        double X[TB][TB];
#pragma unroll
        for (int a = 0; a < TB; ++a)
#pragma unroll
            for (int b = 0; b < TB; ++b) X[a][b] = 0.0;
        const int kend = TB * tx + TB;                     // Linv is lower triangular: Linv[c][k] = 0 for k > c
#pragma unroll 4
        for (int k = 0; k < kend; ++k) {
            const double2* ap = reinterpret_cast<const double2*>(As + k * ST + TB * ty);
            const double2* bp = reinterpret_cast<const double2*>(Cs + k * ST + TB * tx);
            double av[TB], bv[TB];
#pragma unroll
            for (int a = 0; a < TB / 2; ++a) { const double2 v = ap[a]; av[2 * a] = v.x; av[2 * a + 1] = v.y; }
#pragma unroll
            for (int b = 0; b < TB / 2; ++b) { const double2 v = bp[b]; bv[2 * b] = v.x; bv[2 * b + 1] = v.y; }
#pragma unroll
            for (int a = 0; a < TB; ++a)
#pragma unroll
                for (int b = 0; b < TB; ++b) X[a][b] = fma(av[a], bv[b], X[a][b]);
        }
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            const double rd = rcp_fast(Ds[TB * tx + b]);
            double2* dst = reinterpret_cast<double2*>(T + (size_t)ST * (TB * tx + b) + TB * ty);
#pragma unroll
            for (int a = 0; a < TB / 2; ++a) dst[a] = make_double2(X[2 * a][b] * rd, X[2 * a + 1][b] * rd);
        }
    } else {
    // ---- diagonal tile: in-register LDL' with the row operations mirrored on M (-> M = L^-1).  One barrier per pivot:
    // the owners of column j / row j publish them (double buffered), everybody below applies the rank-1 update.
    double M[TB][TB];
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int b = 0; b < TB; ++b) M[a][b] = (ty == tx && a == b) ? 1.0 : 0.0;
    __syncthreads();
    if (tid < ST) vs[tid] = xp[(size_t)tk.col * ST + tid] - (Cs[tid] + Cs[ST + tid]);
    for (int jb = 0; jb < TG; ++jb) {
#pragma unroll
        for (int jj = 0; jj < TB; ++jj) {
            const int j = TB * jb + jj, buf = j & 1;
            if (tx == jb && ty >= jb) {
#pragma unroll
                for (int a = 0; a < TB; ++a) colA[buf * ST + TB * ty + a] = R[a][jj];
            }
            if (ty == jb && tx <= jb) {
#pragma unroll
                for (int b = 0; b < TB; ++b) rowM[buf * ST + TB * tx + b] = M[jj][b];
            }
            __syncthreads();
            if (ty >= jb && tx <= ty) {
                const double rd = rcp_fast(colA[buf * ST + j]);
                double li[TB];
#pragma unroll
                for (int a = 0; a < TB; ++a) li[a] = (TB * ty + a > j) ? colA[buf * ST + TB * ty + a] * rd : 0.0;
                if (tx >= jb) {
#pragma unroll
                    for (int b = 0; b < TB; ++b) {
                        const double ck = (TB * tx + b > j) ? colA[buf * ST + TB * tx + b] : 0.0;
#pragma unroll
                        for (int a = 0; a < TB; ++a) R[a][b] = fma(-li[a], ck, R[a][b]);
                    }
                    if (tx == jb) {
#pragma unroll
                        for (int a = 0; a < TB; ++a) if (TB * ty + a > j) R[a][jj] = li[a];
                    }
                }
                if (tx <= jb) {
#pragma unroll
                    for (int b = 0; b < TB; ++b) {
                        const double mr = rowM[buf * ST + TB * tx + b];
#pragma unroll
                        for (int a = 0; a < TB; ++a) M[a][b] = fma(-li[a], mr, M[a][b]);
                    }
                }
            }
        }
    }
    // ---- write L (strict lower) + D (diagonal) and Linv; forward substitution, second half: y_J = Linv_J v
    double* Li = Linv + (size_t)tk.col * ST2;
    if (ty >= tx) {
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            double2* dst = reinterpret_cast<double2*>(T + (size_t)ST * (TB * tx + b) + TB * ty);
            double2* dsi = reinterpret_cast<double2*>(Li + (size_t)ST * (TB * tx + b) + TB * ty);
#pragma unroll
            for (int a = 0; a < TB / 2; ++a) { dst[a] = make_double2(R[2 * a][b], R[2 * a + 1][b]); dsi[a] = make_double2(M[2 * a][b], M[2 * a + 1][b]); }
        }
#pragma unroll
        for (int a = 0; a < TB; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < TB; ++b) s = fma(M[a][b], vs[TB * tx + b], s);
            As[tx * ST + TB * ty + a] = s;
        }
    }
    __syncthreads();
    if (tid < ST) {
        double s = 0.0;
        for (int c = 0; c <= tid / TB; ++c) s += As[c * ST + tid];
        xp[(size_t)tk.col * ST + tid] = s;
    }
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward sweep, one CTA per tile column of the current level ("pull" form: a CTA only writes its own x_J):
//   x_J = Linv_J' (y_J / D_J - sum_{I>J} L_IJ' x_I)
// Tiles are staged in shared memory (coalesced 16-byte requests, double buffered); thread (c, half) owns column c.
// ---------------------------------------------------------------------------------------------------
constexpr size_t BWD_SMEM = (size_t)(2 * ST2 + 5 * ST) * sizeof(double);
__global__ void __launch_bounds__(RED_THREADS) ldl_bwd_kernel(const double* __restrict__ S, const double* __restrict__ Linv, RedSolveLists t,
                                                              const int* __restrict__ cols, double* __restrict__ x) {
    extern __shared__ __align__(16) double sm[];
    double* Ts = sm;                 // [2][ST2]
    double* xi = sm + 2 * ST2;       // [2][ST]
    double* part = xi + 2 * ST;      // [2][ST]
    double* w = part + 2 * ST;       // [ST]
    const int tid = threadIdx.x, c = tid % ST, h = tid / ST;
    const int J = cols[blockIdx.x];
    const int q0 = t.colptr[J], q1 = t.colptr[J + 1];
    double acc = 0.0;
    if (q0 < q1) {
        tile_to_smem(Ts, S + (size_t)t.col_tile[q0] * ST2);
        if (tid < ST) xi[tid] = x[(size_t)t.col_row[q0] * ST + tid];
    }
    for (int q = q0; q < q1; ++q) {
        const int st = (q - q0) & 1;
        cp_async_wait_all();
        __syncthreads();
        if (q + 1 < q1) {
            tile_to_smem(Ts + (st ^ 1) * ST2, S + (size_t)t.col_tile[q + 1] * ST2);
            if (tid < ST) xi[(st ^ 1) * ST + tid] = x[(size_t)t.col_row[q + 1] * ST + tid];
        }
        const double* M = Ts + st * ST2 + (size_t)ST * c + (ST / 2) * h;
        const double* xv = xi + st * ST + (ST / 2) * h;
#pragma unroll 12
        for (int i = 0; i < ST / 2; ++i) acc = fma(M[i], xv[i], acc);
    }
    __syncthreads();
    tile_to_smem(Ts, Linv + (size_t)J * ST2);
    part[h * ST + c] = acc;
    cp_async_wait_all();
    __syncthreads();
    if (tid < ST) w[tid] = x[(size_t)J * ST + tid] / S[(size_t)t.diag_tile[J] * ST2 + (size_t)(ST + 1) * tid] - (part[tid] + part[ST + tid]);
    __syncthreads();
    {   // x_J[c] = sum_{i >= c} Linv[i][c] w[i]
        const double* M = Ts + (size_t)ST * c + (ST / 2) * h;
        const double* wv = w + (ST / 2) * h;
        double s = 0.0;
#pragma unroll 12
        for (int i = 0; i < ST / 2; ++i) s = fma(((ST / 2) * h + i >= c) ? M[i] : 0.0, wv[i], s);
        part[h * ST + c] = s;
    }
    __syncthreads();
    if (tid < ST) x[(size_t)J * ST + tid] = part[tid] + part[ST + tid];
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
