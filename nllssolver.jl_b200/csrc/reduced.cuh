// Reduced camera system S x = rhs: tile-sparse storage + level-scheduled tile LDL' (no pivoting) + sweeps.
//
// S (DC nA square, symmetric) is cut into ST x ST tiles aligned to camera blocks (ST = 72 = 12 affine or 8 pinhole cameras).
// Camera tiles are renumbered by a fill-reducing / parallelism-exposing order computed on the host (nested dissection by
// index when the tile pattern is banded, identity otherwise).  Only tiles that are structurally non-zero after symbolic
// tile-level fill-in are stored (lower triangle in the permuted numbering, column-major inside a tile, tile `id` at
// S + id*ST*ST).  The factorisation S = L D L' uses no pivoting — the algebra of the reference's sparse path
// (LDLFactorizations.ldl_factorize!, src/linearsolver.jl:29), so indefinite systems (Triggs-corrected robust Hessians)
// follow the same trajectory instead of failing over.
//
// Tile columns are grouped into levels of the elimination tree; per level three launches (task lists built at prepare time):
//   ldl_diag_kernel  one CTA per column J of the level: in-register LDL' of the 72 x 72 tile, Linv_J = L_JJ^-1 obtained by
//                    mirroring the row operations on an identity, and the forward substitution y_J = L_JJ^-1 b_J carried as
//                    an extra column.  One barrier per pivot; the pivot reciprocal is computed one step ahead by its owner.
//   ldl_off_kernel   two CTAs per tile (I, J), I > J:  L_IJ = T_IJ Linv_J' D_J^-1 (a triangular GEMM), then the right-hand
//                    side push  b_I -= L_IJ y_J
//   ldl_upd_kernel   two CTAs per pair (a >= b) of rows of J:  T_{ab} -= (L_aJ D_J) L_bJ'  (FP64 RED into the target tile)
// then the backward sweep ldl_bwd_kernel, one launch per level in reverse order.
//
// Measured on B200 (scripts/ubench/fp64_lat.cu): DFMA 8.4 cycles dependent / 2.07 cycles issue per warp and SM sub-partition,
// STS+BAR+LDS handshake 60 cycles, reciprocal chain 48 cycles — the kernels below are laid out around those numbers.
#pragma once
#include "common.cuh"

namespace nlls {

constexpr int ST = 72;
constexpr int ST2 = ST * ST;
constexpr int TB = 6;                 // register block edge
constexpr int TG = ST / TB;           // 12 x 12 blocks per tile
constexpr int NLB = TG * (TG + 1) / 2;  // 78 lower-triangular blocks
constexpr int DIAG_THREADS = 224;     // 3 warps of L blocks, 3 warps of Linv blocks, 1 right-hand-side warp
constexpr int RED_THREADS = 144;      // backward sweep

struct RedSolveLists {        // device pointers for the backward sweep
    const int* diag_tile;     // [NT] tile id of (J, J), permuted numbering
    const int* colptr;        // [NT + 1] tiles (I, J), I > J, of block column J
    const int* col_tile;
    const int* col_row;
};

struct RedTask { int tile, dtile, col, row; };        // diag: (tile, -, J, -); off-diagonal: (tile (I,J), diag tile of J, J, I)
struct RedUpd { int a, b, dk, target; };              // tiles L_aJ, L_bJ, diagonal tile of J, target tile (a, b)

// reciprocal to ~1 ulp: MUFU seed + two Newton steps
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
template <int NT_>
__device__ __forceinline__ void tile_to_smem(double* sdst, const double* gsrc) {
    for (int q = threadIdx.x; q < ST2 / 2; q += NT_) cp_async16(sdst + 2 * q, gsrc + 2 * q);
}

// ---------------------------------------------------------------------------------------------------
// Diagonal tile.  Warp roles (warp w runs on SM sub-partition w % 4; early-finishing and late-starting warps share one):
//   w0 = L blocks 52..77   w1 = Linv blocks 52..77   w2 = L blocks 26..51   w3 = Linv blocks 26..51
//   w4 = L blocks 0..25    w5 = Linv blocks 0..25    w6 = right-hand side
// L blocks are numbered by (tx, ty) ascending — block (ty, tx) is touched by pivots j < 6 tx + 6, so low numbers retire first;
// Linv blocks by (ty, tx) ascending — block (ty, tx) is touched by pivots 6 tx <= j < 6 ty + 6.
// A single warp issues dependent instructions ~5 cycles apart, so the per-pivot instruction count is what matters:
//   * finished entries are "retired" to shared memory the moment they are final (column j of L, row j+1 of Linv, d_j), after
//     which their registers may hold garbage — the rank-1 updates then need no row / column masks at all;
//   * every role has its own loop (named barrier, explicit thread count), so nothing role-dependent is re-evaluated per pivot;
//   * the next pivot and its reciprocal are computed by every thread in the shadow of the 36 updates, only the owner stores.
// ---------------------------------------------------------------------------------------------------
constexpr int LDM = ST + 1;           // padded row stride of the Linv staging tile (conflict-free transposed read-out)
constexpr size_t DIAG_SMEM = (size_t)(ST2 + ST * LDM + 5 * ST + 4) * sizeof(double);

__device__ __forceinline__ void diag_bar() { asm volatile("bar.sync 1, %0;" ::"n"(DIAG_THREADS) : "memory"); }

__global__ void __launch_bounds__(DIAG_THREADS) ldl_diag_kernel(double* __restrict__ S, double* __restrict__ Linv, const RedTask* __restrict__ tasks,
                                                                double* __restrict__ xp) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                  // the tile as stored (column-major); retired columns of L overwrite it
    double* Ms = sm + ST2;            // [ST][LDM] retired rows of M = L^-1
    double* colA = Ms + ST * LDM;     // [2][ST] published column of A
    double* rowM = colA + 2 * ST;     // [2][ST] published row of M
    double* dbuf = rowM + 2 * ST;     // [ST] pivots
    double* rdb = dbuf + ST;          // [2] published pivot reciprocal
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const RedTask tk = tasks[blockIdx.x];
    double* T = S + (size_t)tk.tile * ST2;
    tile_to_smem<DIAG_THREADS>(As, T);

    const int kind = (w == 6) ? 2 : (w & 1);              // 0: L blocks, 1: Linv blocks, 2: right-hand side
    const int grp = 2 - (w >> 1);                         // block group 0..2
    const bool has_blk = kind < 2 && lane < NLB / 3;
    int ty = 0, tx = 0;
    if (kind < 2) {
        int n = (NLB / 3) * grp + min(lane, NLB / 3 - 1);
        if (kind == 0) { while (n >= TG - tx) { n -= TG - tx; ++tx; } ty = tx + n; }      // by tx, then ty
        else { while (n >= ty + 1) { n -= ty + 1; ++ty; } tx = n; }                          // by ty, then tx
    }
    double v[3] = {0.0, 0.0, 0.0};
    if (kind == 2) {
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) if (lane + 32 * s3 < ST) v[s3] = xp[(size_t)tk.col * ST + lane + 32 * s3];
    }
    cp_async_wait_all();
    __syncthreads();
    double B[TB][TB];                                     // L block (kind 0) or Linv block (kind 1)
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int b = 0; b < TB; ++b) B[a][b] = (kind == 0) ? As[(TB * tx + b) * ST + TB * ty + a] : ((ty == tx && a == b) ? 1.0 : 0.0);
    if (kind == 0 && has_blk && tx == 0) {
#pragma unroll
        for (int a = 0; a < TB; ++a) colA[TB * ty + a] = B[a][0];
        if (ty == 0) { rdb[0] = rcp_fast(B[0][0]); dbuf[0] = B[0][0]; }
    }
    if (kind == 1 && has_blk && ty == 0) {
#pragma unroll
        for (int b = 0; b < TB; ++b) { rowM[b] = B[0][b]; Ms[b] = B[0][b]; }
    }
    __syncthreads();                                      // every thread holds its block: As may now receive retired columns

    if (kind == 0) {
        // ---------------- L blocks: rank-1 update, retire column j, publish column j + 1 and its pivot ----------------
        const int last_j = grp == 0 ? 17 : (grp == 1 ? 35 : ST - 1);
        const double* cty = colA + TB * ty;
        const double* ctx = colA + TB * tx;
        for (int jb = 0; jb < TG; ++jb) {
            const bool live = has_blk && tx >= jb;
            const bool own = live && tx == jb;
#pragma unroll
            for (int jj = 0; jj < TB; ++jj) {
                const int j = TB * jb + jj, buf = j & 1, nb = buf ^ 1;
                const int jn = (jj + 1) % TB;
                diag_bar();
                if (j > last_j || !live) continue;
                const double rd = rdb[buf];
                double li[TB], ck[TB];
#pragma unroll
                for (int a = 0; a < TB / 2; ++a) {
                    const double2 q = reinterpret_cast<const double2*>(cty + buf * ST)[a];
                    li[2 * a] = q.x * rd; li[2 * a + 1] = q.y * rd;
                    const double2 r2 = reinterpret_cast<const double2*>(ctx + buf * ST)[a];
                    ck[2 * a] = r2.x; ck[2 * a + 1] = r2.y;
                }
                // next pivot (meaningful on its diagonal block only) and its reciprocal, in the shadow of the updates
                const double dn = fma(-li[jn], ck[jn], B[jn][jn]);
                const double rn = rcp_fast(dn);
                const bool own_next = (jj + 1 < TB) ? own : (has_blk && tx == jb + 1);
                if (own_next && ty == tx) { rdb[nb] = rn; dbuf[j + 1 < ST ? j + 1 : j] = dn; }
#pragma unroll
                for (int b = 0; b < TB; ++b)
#pragma unroll
                    for (int a = 0; a < TB; ++a) B[a][b] = fma(-li[a], ck[b], B[a][b]);
                if (own) {                                  // column j of L is final: retire it
                    double2* dst = reinterpret_cast<double2*>(As + j * ST + TB * ty);
#pragma unroll
                    for (int a = 0; a < TB / 2; ++a) dst[a] = make_double2(li[2 * a], li[2 * a + 1]);
                }
                if (own_next) {
                    double2* dst = reinterpret_cast<double2*>(colA + nb * ST + TB * ty);
#pragma unroll
                    for (int a = 0; a < TB / 2; ++a) dst[a] = make_double2(B[2 * a][jn], B[2 * a + 1][jn]);
                }
            }
        }
    } else if (kind == 1) {
        // ---------------- Linv blocks: the same row operations on the identity; retire / publish row j + 1 ----------------
        const int last_j = grp == 0 ? 41 : (grp == 1 ? 59 : ST - 1);
        const double* cty = colA + TB * ty;
        const double* rtx = rowM + TB * tx;
        for (int jb = 0; jb < TG; ++jb) {
            const bool live = has_blk && tx <= jb && ty >= jb;
#pragma unroll
            for (int jj = 0; jj < TB; ++jj) {
                const int j = TB * jb + jj, buf = j & 1, nb = buf ^ 1;
                const int jn = (jj + 1) % TB;
                diag_bar();
                if (j > last_j) continue;
                if (live) {
                    const double rd = rdb[buf];
                    double li[TB], mr[TB];
#pragma unroll
                    for (int a = 0; a < TB / 2; ++a) {
                        const double2 q = reinterpret_cast<const double2*>(cty + buf * ST)[a];
                        li[2 * a] = q.x * rd; li[2 * a + 1] = q.y * rd;
                        const double2 r2 = reinterpret_cast<const double2*>(rtx + buf * ST)[a];
                        mr[2 * a] = r2.x; mr[2 * a + 1] = r2.y;
                    }
#pragma unroll
                    for (int b = 0; b < TB; ++b)
#pragma unroll
                        for (int a = 0; a < TB; ++a) B[a][b] = fma(-li[a], mr[b], B[a][b]);
                }
                // row j + 1 of M is final now: retire and publish it
                const bool own_next = has_blk && ((jj + 1 < TB) ? (live && ty == jb) : (ty == jb + 1 && tx <= jb + 1));
                if (own_next) {
#pragma unroll
                    for (int b = 0; b < TB; ++b) { rowM[nb * ST + TB * tx + b] = B[jn][b]; Ms[(j + 1) * LDM + TB * tx + b] = B[jn][b]; }
                }
            }
        }
    } else {
        // ---------------- right-hand side: forward substitution carried as an extra column, v_i -= l_ij v_j ----------------
        for (int j = 0; j < ST; ++j) {
            const int buf = j & 1;
            diag_bar();
            const double rd = rdb[buf];
            const int sj = j >> 5;
            const double vsel = (sj == 0) ? v[0] : (sj == 1 ? v[1] : v[2]);
            const double vj = __shfl_sync(0xffffffffu, vsel, j & 31);
#pragma unroll
            for (int s3 = 0; s3 < 3; ++s3) {
                const int i = lane + 32 * s3;
                if (i > j && i < ST) v[s3] = fma(-(colA[buf * ST + i] * rd), vj, v[s3]);
            }
        }
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) if (lane + 32 * s3 < ST) xp[(size_t)tk.col * ST + lane + 32 * s3] = v[s3];
    }
    __syncthreads();
    // ---- write L (strict lower) + D (diagonal) and Linv (lower incl. the unit diagonal; the upper triangle stays zero)
    double* Li = Linv + (size_t)tk.col * ST2;
    for (int e = tid; e < ST2; e += DIAG_THREADS) {
        const int i = e % ST, k = e / ST;
        if (i > k) { T[e] = As[e]; Li[e] = Ms[i * LDM + k]; }
        else if (i == k) { T[e] = dbuf[k]; Li[e] = 1.0; }
    }
}

// ---------------------------------------------------------------------------------------------------
// FP64 tensor-core GEMMs on 72 x 72 tiles (mma.sync.m8n8k4.f64, SASS DMMA).  Measured on B200: a register outer product
// with three distinct operands issues one DFMA per ~4.7 cycles per SM sub-partition, DMMA sustains 256 FMA per 16.7 cycles
// (61 FMA/clk/SM, the FP64 peak) — so every tile GEMM of the factorisation goes through DMMA.
// Both operands are read with the same pattern from column-major tiles staged in shared memory with a row stride of 76
// doubles (76 / 2 = 6 (mod 16) puts the 4 x 4 fragment rows of a half-warp on 16 distinct 8-byte bank pairs):
//   A fragment (8 x 4, row):  lane l holds A[l >> 2][l & 3]  =  tileA[(k0 + (l & 3)) * 76 + i0 + (l >> 2)]
//   B fragment (4 x 8, col):  lane l holds B[l & 3][l >> 2]  =  tileB[(k0 + (l & 3)) * 76 + c0 + (l >> 2)]     (B = tileB')
//   C fragment (8 x 8):       lane l holds C[l >> 2][2 (l & 3) + {0, 1}]
// A CTA has 3 warps; warp w of CTA r owns the 8 rows of row block 3 r + w and all 9 column blocks (18 accumulators).
// ---------------------------------------------------------------------------------------------------
constexpr int LDT = 76;
constexpr int GEMM_THREADS = 96;
constexpr int GEMM_CTAS = 3;          // CTAs per tile task
constexpr int NBLK = ST / 8;          // 9 blocks of 8 per tile edge

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// tile (contiguous, column-major, leading dimension 72) -> shared memory with leading dimension LDT
template <int NT_>
__device__ __forceinline__ void tile_to_smem_ld(double* sdst, const double* gsrc) {
    for (int q = threadIdx.x; q < ST2 / 2; q += NT_) {
        const int k = q / (ST / 2), c = q - k * (ST / 2);
        cp_async16(sdst + k * LDT + 2 * c, gsrc + k * ST + 2 * c);
    }
}

// ---------------------------------------------------------------------------------------------------
// Off-diagonal tile (I, J):  L_IJ = T_IJ Linv_J' D_J^-1, i.e. X[i][c] = sum_{k <= c} T[i][k] Linv[c][k] / d_c ; then the
// right-hand side push b_I -= L_IJ y_J.
// ---------------------------------------------------------------------------------------------------
constexpr size_t OFF_SMEM = (size_t)(2 * ST * LDT + 2 * ST) * sizeof(double);

__global__ void __launch_bounds__(GEMM_THREADS) ldl_off_kernel(double* __restrict__ S, const double* __restrict__ Linv, const RedTask* __restrict__ tasks,
                                                               double* __restrict__ xp) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                  // T_IJ   [k][i]
    double* Bs = sm + ST * LDT;       // Linv_J [k][c]
    double* Ds = Bs + ST * LDT;       // [ST] 1 / D_J
    double* ys = Ds + ST;             // [ST] y_J
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const RedTask tk = tasks[blockIdx.x / GEMM_CTAS];
    const int I = GEMM_CTAS * (blockIdx.x % GEMM_CTAS) + w;      // row block of this warp
    double* T = S + (size_t)tk.tile * ST2;
    tile_to_smem_ld<GEMM_THREADS>(As, T);
    tile_to_smem_ld<GEMM_THREADS>(Bs, Linv + (size_t)tk.col * ST2);
    if (tid < ST) {
        Ds[tid] = rcp_fast(S[(size_t)tk.dtile * ST2 + (size_t)(ST + 1) * tid]);
        ys[tid] = xp[(size_t)tk.col * ST + tid];
    }
    cp_async_wait_all();
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;
    double c0[NBLK], c1[NBLK];
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) { c0[cb] = 0.0; c1[cb] = 0.0; }
    const double* ap = As + fk * LDT + 8 * I + fr;
    const double* bp = Bs + fk * LDT + fr;
#pragma unroll
    for (int ks = 0; ks < ST / 4; ++ks) {
        const double a = ap[ks * 4 * LDT];
#pragma unroll
        for (int cb = 0; cb < NBLK; ++cb) {
            if (cb < ks / 2) continue;                        // Linv is lower triangular: Linv[c][k] = 0 for k > c
            dmma884(c0[cb], c1[cb], a, bp[ks * 4 * LDT + 8 * cb]);
        }
    }
    // scale by 1 / d_c, store L_IJ, and push the right-hand side:  b_I -= L_IJ y_J
    double pr = 0.0;
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) {
        const int c = 8 * cb + 2 * fk;
        const double x0 = c0[cb] * Ds[c], x1 = c1[cb] * Ds[c + 1];
        T[(size_t)ST * c + 8 * I + fr] = x0;
        T[(size_t)ST * (c + 1) + 8 * I + fr] = x1;
        pr = fma(x0, ys[c], pr);
        pr = fma(x1, ys[c + 1], pr);
    }
    pr += __shfl_xor_sync(0xffffffffu, pr, 1);
    pr += __shfl_xor_sync(0xffffffffu, pr, 2);
    if (fk == 0) atomicAdd(xp + (size_t)tk.row * ST + 8 * I + fr, -pr);
}

// ---------------------------------------------------------------------------------------------------
// Update  T_{ab} -= (L_aJ D_J) L_bJ'  for one pair of rows of column J; results are subtracted from the target tile with FP64
// reductions (several columns of a level can hit the same tile).
// ---------------------------------------------------------------------------------------------------
constexpr size_t UPD_SMEM = (size_t)(2 * ST * LDT + ST) * sizeof(double);

__global__ void __launch_bounds__(GEMM_THREADS) ldl_upd_kernel(double* __restrict__ S, const RedUpd* __restrict__ upds) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                  // L_aJ [k][i]
    double* Bs = sm + ST * LDT;       // L_bJ [k][c]
    double* Ds = Bs + ST * LDT;       // [ST] D_J
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const RedUpd up = upds[blockIdx.x / GEMM_CTAS];
    const int I = GEMM_CTAS * (blockIdx.x % GEMM_CTAS) + w;
    const bool diag = up.b == up.a;
    tile_to_smem_ld<GEMM_THREADS>(As, S + (size_t)up.a * ST2);
    if (!diag) tile_to_smem_ld<GEMM_THREADS>(Bs, S + (size_t)up.b * ST2);
    if (tid < ST) Ds[tid] = S[(size_t)up.dk * ST2 + (size_t)(ST + 1) * tid];
    cp_async_wait_all();
    __syncthreads();
    const double* Bt = diag ? As : Bs;
    const int fr = lane >> 2, fk = lane & 3;
    const int ncb = diag ? I + 1 : NBLK;                      // diagonal targets: lower blocks only
    double c0[NBLK], c1[NBLK];
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) { c0[cb] = 0.0; c1[cb] = 0.0; }
    const double* ap = As + fk * LDT + 8 * I + fr;
    const double* bp = Bt + fk * LDT + fr;
#pragma unroll
    for (int ks = 0; ks < ST / 4; ++ks) {
        const double a = ap[ks * 4 * LDT] * Ds[4 * ks + fk];
#pragma unroll
        for (int cb = 0; cb < NBLK; ++cb) {
            if (cb >= ncb) continue;
            dmma884(c0[cb], c1[cb], a, bp[ks * 4 * LDT + 8 * cb]);
        }
    }
    double* T = S + (size_t)up.target * ST2;
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) {
        if (cb >= ncb) continue;
        const int c = 8 * cb + 2 * fk;
        atomicAdd(T + (size_t)ST * c + 8 * I + fr, -c0[cb]);
        atomicAdd(T + (size_t)ST * (c + 1) + 8 * I + fr, -c1[cb]);
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
        tile_to_smem<RED_THREADS>(Ts, S + (size_t)t.col_tile[q0] * ST2);
        if (tid < ST) xi[tid] = x[(size_t)t.col_row[q0] * ST + tid];
    }
    for (int q = q0; q < q1; ++q) {
        const int st = (q - q0) & 1;
        cp_async_wait_all();
        __syncthreads();
        if (q + 1 < q1) {
            tile_to_smem<RED_THREADS>(Ts + (st ^ 1) * ST2, S + (size_t)t.col_tile[q + 1] * ST2);
            if (tid < ST) xi[(st ^ 1) * ST + tid] = x[(size_t)t.col_row[q + 1] * ST + tid];
        }
        const double* M = Ts + st * ST2 + (size_t)ST * c + (ST / 2) * h;
        const double* xv = xi + st * ST + (ST / 2) * h;
#pragma unroll 12
        for (int i = 0; i < ST / 2; ++i) acc = fma(M[i], xv[i], acc);
    }
    __syncthreads();
    tile_to_smem<RED_THREADS>(Ts, Linv + (size_t)J * ST2);
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
