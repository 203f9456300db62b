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
//   ldl_off_kernel   three CTAs per tile (I, J), I > J:  L_IJ = T_IJ Linv_J' D_J^-1 (a triangular GEMM)
//   ldl_upd_kernel   three CTAs per TARGET tile (a, b) of the level:  T_{ab} -= sum_J (L_aJ D_J) L_bJ' in a fixed order, written once
//                    by its owner (no reductions: the solve is bitwise reproducible); diagonal targets also push b_I -= L_IJ y_J
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
constexpr int RED_THREADS = 144;      // backward sweep

struct RedSolveLists {        // device pointers for the backward sweep
    const int* diag_tile;     // [NT] tile id of (J, J), permuted numbering
    const int* colptr;        // [NT + 1] tiles (I, J), I > J, of block column J
    const int* col_tile;
    const int* col_row;
};

struct RedTask { int tile, dtile, col, row; };        // diag: (tile, -, J, -); off-diagonal: (tile (I,J), diag tile of J, J, I)
struct RedUpd { int a, b, dk, col; };                 // tiles L_aJ, L_bJ, diagonal tile of J, column J (its y_J feeds the right-hand side push)
struct RedTarget { int target, u0, u1, row; };        // target tile, its updates [u0, u1) in a fixed order, block row I of a DIAGONAL target (-1 otherwise)

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

// Programmatic dependent launch (the reduced solve is ~40 short dependent kernels): a kernel of the chain waits for its predecessor
// with pdl_wait() and immediately allows its successor to be scheduled, so the successor's launch latency and whatever it does before
// its own pdl_wait() overlap this kernel.  Because the trigger comes after the wait, a kernel that is running knows that everything up
// to its predecessor's predecessor is complete and visible — that is what may be touched before pdl_wait().  Without the launch
// attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// rows [row0, row0 + 24) of a tile only (what the three warps of one CTA of a tile task read of their A operand: a third of the bytes)
template <int NT_>
__device__ __forceinline__ void tile_rows24_to_smem_ld(double* sdst, const double* gsrc, int row0) {
    for (int q = threadIdx.x; q < ST * 12; q += NT_) {
        const int k = q / 12, c = q - k * 12;
        cp_async16(sdst + k * LDT + row0 + 2 * c, gsrc + k * ST + row0 + 2 * c);
    }
}

// ---------------------------------------------------------------------------------------------------
// Diagonal tile: blocked right-looking LDL' with panels of 8 columns (9 panels), everything in shared memory, with look-ahead:
//   A1 (warp 0)        LDL' of the next 8 x 8 diagonal block in registers (lane = row; pivot row / reciprocal broadcast by
//                      shuffles, the next reciprocal issued before the bulk of the current update) together with
//                      N11 = L11^-1 (the same row operations on an 8 x 8 identity) — runs while warps 1-7 do B / C of the
//                      current panel;
//   A2 (all warps)     L21 = A21 N11' D^-1  as DMMA products, one 8-row block per warp;
//   B  (warps 1-7)     trailing update  A22 -= L21 D L21'  as DMMA rank-8 updates (warp 0 updates the next diagonal block itself);
//   C  (warps 1-7)     R = [ I | b_J ] is carried along:  Mfin[panel rows] = N11 R[panel rows],  R[rows below] -= L21 Mfin[panel rows].
// Mfin ends up as [ Linv_J | y_J ]: L_JJ^-1 and the forward-substituted right-hand side.  Two barriers per panel; a per-pivot
// register formulation measured 25 us per tile (one warp issues dependent FP64 instructions ~5 cycles apart).
// ---------------------------------------------------------------------------------------------------
constexpr int DIAG_THREADS = 512;                          // many warps: one warp issues dependent instructions only every ~5 cycles
constexpr int DIAG_WARPS = DIAG_THREADS / 32;
constexpr int MCOLS = ST + 8;                              // Linv columns + one block holding the right-hand side

// Static task table of the B / C phase (element offsets precomputed: a warp spends its time in dependent integer
// instructions otherwise).  Every task is one 8 x 8 block update  C -= A B  over the panel's 8 columns (two DMMAs); offsets are
// relative to the start of shared memory (As, Rs and Mf are consecutive).  For panel p, entry t = {kind, oC, oA, oB}:
//   kind 0  trailing block (I, J), p < J <= I:  C = A(I, J),  A = L21[I] D (scaled on load),  B = L21[J]'     (both operands in As)
//   kind 1  R[I][J] -= L21[I] Mfin[p][J], I > p, J in {0..p, rhs}:  C in Rs,  A = L21[I] in As,  B = Mfin[p][J] in Mf
// Mfin[p][J] = N11 R[p][J] itself is computed in the A2 phase by the warps that have no L21 block there, so the B / C tasks are
// uniform and need no layout change through scratch memory.
// (lane-dependent parts are added by the kernel: fragment row / k offsets.)
constexpr int DIAG_RS_OFF = ST * LDT;
constexpr int DIAG_MF_OFF = DIAG_RS_OFF + (ST + 8) * LDT;
struct DiagTaskTable {
    alignas(16) int v[NBLK][64][4];
    int n[NBLK];
    constexpr DiagTaskTable() : v(), n() {
        for (int p = 0; p < NBLK; ++p) {
            const int nb = NBLK - 1 - p, c0 = 8 * p;
            int c = 0;
            for (int I = 0; I < nb; ++I)
                for (int J = (I == 0) ? 1 : 0; J <= I; ++J) {       // (0, 0) = the next diagonal block, left to warp 0
                    const int gi = 8 * (p + 1 + I), gj = 8 * (p + 1 + J);
                    v[p][c][0] = 0; v[p][c][1] = gj * LDT + gi; v[p][c][2] = c0 * LDT + gi; v[p][c][3] = c0 * LDT + gj; ++c;
                }
            for (int jj = 0; jj < p + 2; ++jj)
                for (int ii = 1; ii <= nb; ++ii) {
                    const int J = (jj == p + 1) ? NBLK : jj, I = p + ii;
                    v[p][c][0] = 1; v[p][c][1] = DIAG_RS_OFF + 8 * J * LDT + 8 * I; v[p][c][2] = c0 * LDT + 8 * I; v[p][c][3] = DIAG_MF_OFF + 8 * J * LDT + c0; ++c;
                }
            n[p] = c;
        }
    }
};
// The table lives in global memory and is copied into shared memory in the kernel's prologue (before the dependency wait): indexed
// constant-bank loads (LDC with a per-warp register index) measured ~40 % of a worker warp's time in the B / C phase (mio / short
// scoreboard stalls on the four LDCs of every task).
__device__ const DiagTaskTable g_diag_tasks = DiagTaskTable();
constexpr int DIAG_TAB_INT4 = NBLK * 64;
constexpr size_t DIAG_SMEM = (size_t)(ST * LDT + 2 * MCOLS * LDT + 2 * 128 + ST + 16) * sizeof(double) + (size_t)(DIAG_TAB_INT4 + 4) * sizeof(int4);

#ifndef LDL_ABLATE
#define LDL_ABLATE 0      // development aid (scripts/ubench): bit mask of phases of ldl_diag_kernel to skip, timing only
#endif
#ifdef LDL_PROFILE
#define LDL_STAMP(i) do { if (threadIdx.x == 0) prof[i] = clock64(); } while (0)
#define LDL_T0() t_ph = clock64()
#define LDL_ACC(i) do { const long long t_now = clock64(); t_acc[(i) - 64] += t_now - t_ph; t_ph = t_now; } while (0)
#else
#define LDL_STAMP(i)
#define LDL_T0()
#define LDL_ACC(i)
#endif

// LDL' of the 8 x 8 block at (c0, c0) of As, N11 = L11^-1 and W = N11' D^-1 (one warp).  Lane (r, g) = 4 r + g owns columns
// 2 g, 2 g + 1 of row r of both the block and N11.  This routine is the critical path of the whole reduced solve (72 pivots per
// tile, one tile per level): a pivot-at-a-time loop measured 190 cycles per pivot — shuffle, reciprocal (MUFU + two Newton steps),
// multiplier, update, all dependent.  Pivots are therefore taken in PAIRS (j, j + 1):
//     d_j = a,   d_{j+1} = c - b^2 / a = det / a   with det = a c - b^2,   so   1 / d_j = rcp(a)  and  1 / d_{j+1} = a rcp(det)
// — two reciprocals that start together from values known at the start of the pair — and both columns of multipliers and the rank-2
// update of the remaining block follow from one round of shuffles:
//     l_rj = a_rj / d_j,   t_r = a_r,j+1 - l_rj b  (= the once-updated column j + 1),   l_r,j+1 = t_r / d_{j+1},
//     a_rk -= l_rj a_kj + l_r,j+1 t_k.
// Same factorisation (no pivoting, same L and D up to rounding), about 2.3 x shorter dependent chain.
__device__ __forceinline__ void diag_block_factor(double* As, int c0, double* N11, double* W11, double* dpan, double* dall) {
    const int lane = threadIdx.x & 31, r = lane >> 2, g = lane & 3;
    const unsigned FULL = 0xffffffffu;
    double P[2], N[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) { P[e] = As[(c0 + 2 * g + e) * LDT + c0 + r]; N[e] = (r == 2 * g + e) ? 1.0 : 0.0; }
    double rdrow = 0.0;                                                     // 1 / d_r
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = 2 * q;                                                // columns j, j + 1 live in lanes (., q)
        const double a = __shfl_sync(FULL, P[0], 4 * j + q);                // A[j][j]
        const double b = __shfl_sync(FULL, P[0], 4 * (j + 1) + q);          // A[j+1][j]
        const double c = __shfl_sync(FULL, P[1], 4 * (j + 1) + q);          // A[j+1][j+1]
        const double ar0 = __shfl_sync(FULL, P[0], (lane & ~3) | q);        // A[r][j], A[r][j+1]
        const double ar1 = __shfl_sync(FULL, P[1], (lane & ~3) | q);
        double ak0[2], ak1[2], nj0[2], nj1[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            ak0[e] = __shfl_sync(FULL, P[0], 4 * (2 * g + e) + q);          // A[2g+e][j], A[2g+e][j+1]
            ak1[e] = __shfl_sync(FULL, P[1], 4 * (2 * g + e) + q);
            nj0[e] = __shfl_sync(FULL, N[e], 4 * j + g);                    // N[j][2g+e], N[j+1][2g+e]
            nj1[e] = __shfl_sync(FULL, N[e], 4 * (j + 1) + g);
        }
        const double rd0 = rcp_fast(a);
        const double det = fma(a, c, -(b * b));
        const double rd1 = a * rcp_fast(det);
        const double l10 = b * rd0;                                         // l_{j+1,j}
        const double d1 = fma(-l10, b, c);                                  // d_{j+1} as the pivot-at-a-time recurrence rounds it
        if (lane == 0) { dpan[j] = a; dpan[j + 1] = d1; dall[c0 + j] = a; dall[c0 + j + 1] = d1; }
        if (r == j) rdrow = rd0;
        if (r == j + 1) rdrow = rd1;
        const double l0 = ar0 * rd0;
        const double l1 = fma(-l0, b, ar1) * rd1;
        if (r > j) {
            // N_r -= l_rj N_j ;  rows below j + 1 also  N_r -= l_r,j+1 (N_{j+1} - l_{j+1,j} N_j)     (entries right of the diagonal are exact zeros)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double n = fma(-l0, nj0[e], N[e]);
                if (r > j + 1) n = fma(-l1, fma(-l10, nj0[e], nj1[e]), n);
                N[e] = n;
            }
            if (g == q) {                                                   // the two finished columns of L
                P[0] = l0;
                if (r > j + 1) P[1] = l1;
            } else if (g > q && r > j + 1) {                                // rank-2 update of the remaining block
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double tk = fma(-(ak0[e] * rd0), b, ak1[e]);
                    P[e] = fma(-l1, tk, fma(-l0, ak0[e], P[e]));
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int k = 2 * g + e;
        if (k < r) As[(c0 + k) * LDT + c0 + r] = P[e];                      // L11, strict lower
        N11[r * 8 + k] = N[e];
        W11[k * 8 + r] = N[e] * rdrow;                                      // W = N11' D^-1  (L21 = A21 W)
    }
}

__global__ void __launch_bounds__(DIAG_THREADS) ldl_diag_kernel(double* __restrict__ S, double* __restrict__ Linv, const RedTask* __restrict__ tasks,
                                                                double* __restrict__ xp
#ifdef LDL_PROFILE
                                                                , long long* prof
#endif
                                                                ) {
    extern __shared__ __align__(16) double sm[];
    LDL_STAMP(0);
#ifdef LDL_PROFILE
    long long t_ph = 0, t_acc[6] = {0, 0, 0, 0, 0, 0};
#endif
    double* As = sm;                     // [k][i], leading dimension LDT
    double* Rs = As + ST * LDT;          // R (working right-hand sides [ I | b ]), column-major: Rs[col * LDT + row], MCOLS columns
    double* Mf = Rs + MCOLS * LDT;       // finished rows of [ Linv | y ], same layout
    double* NW = Mf + MCOLS * LDT;       // 2 x { N11 [8][8] row-major, W11 [8][8] (W[k * 8 + c] = N11[c][k] / d_c) }
    double* dall = NW + 2 * 128;         // [ST] pivots
    double* dpan2 = dall + ST;           // 2 x [8] pivots of a panel
    int4* tab = reinterpret_cast<int4*>(dpan2 + 16);   // [NBLK][64] task table, then the per-panel task counts
    int* tabn = reinterpret_cast<int*>(tab + DIAG_TAB_INT4);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const RedTask tk = tasks[blockIdx.x];
    double* T = S + (size_t)tk.tile * ST2;
    for (int q = tid; q < MCOLS * LDT / 2; q += DIAG_THREADS) reinterpret_cast<double2*>(Rs)[q] = make_double2(0.0, 0.0);
    for (int q = tid; q < DIAG_TAB_INT4; q += DIAG_THREADS) tab[q] = reinterpret_cast<const int4*>(&g_diag_tasks.v[0][0][0])[q];
    if (tid < NBLK) tabn[tid] = g_diag_tasks.n[tid];
    pdl_wait(); pdl_trigger();           // the tile and b_J were last written by the previous level's update kernel
    tile_to_smem_ld<DIAG_THREADS>(As, T);
    const double bval = (tid < ST) ? xp[(size_t)tk.col * ST + tid] : 0.0;
    __syncthreads();
    if (tid < ST) { Rs[tid * LDT + tid] = 1.0; Rs[ST * LDT + tid] = bval; }
    cp_async_wait_all();
    __syncthreads();
    LDL_STAMP(1);
    if (w == 0) diag_block_factor(As, 0, NW, NW + 64, dpan2, dall);
    __syncthreads();

    for (int p = 0; p < NBLK; ++p) {
        const int c0 = 8 * p, pb = p & 1;
        const double* N11 = NW + pb * 128;
        const double* W11 = N11 + 64;
        const double* dpan = dpan2 + 8 * pb;
        LDL_T0();
        // ---------------- A2: L21 = A21 W, one 8-row block per warp (in place)
        if (!(LDL_ABLATE & 4))
        for (int I = p + 1 + w; I < NBLK; I += DIAG_WARPS) {
            double c0r = 0.0, c1r = 0.0, aq[2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) aq[ks] = As[(c0 + 4 * ks + fk) * LDT + 8 * I + fr];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) dmma884(c0r, c1r, aq[ks], W11[(4 * ks + fk) * 8 + fr]);
            __syncwarp();
            double* cp = As + (c0 + 2 * fk) * LDT + 8 * I + fr;
            cp[0] = c0r; cp[LDT] = c1r;
        }
        {   // Mfin[p][J] = N11 R[p][J], J in {0..p, rhs}: one task per warp that has no L21 block in this panel
            const int jj = w - (NBLK - 1 - p);
            if (jj >= 0 && jj < p + 2) {
                const int J = (jj == p + 1) ? NBLK : jj;
                const double* rp = Rs + 8 * J * LDT + c0 + fr * LDT + fk;
                double m0 = 0.0, m1 = 0.0;
                dmma884(m0, m1, N11[fr * 8 + fk], rp[0]);
                dmma884(m0, m1, N11[fr * 8 + 4 + fk], rp[4]);
                double* cp = Mf + 8 * J * LDT + c0 + 2 * fk * LDT + fr;
                cp[0] = m0; cp[LDT] = m1;
            }
        }
        LDL_ACC(64);
        __syncthreads();
        LDL_ACC(65);
        const int nb = NBLK - 1 - p;                                            // block rows below the panel
        if (w == 0) {
            // ---------------- look-ahead: update the next diagonal block, then factor it (A1 of panel p + 1)
            if (nb > 0 && !(LDL_ABLATE & 8)) {
                const int g = c0 + 8;
                double* cp = As + (g + 2 * fk) * LDT + g + fr;
                double c0r = cp[0], c1r = cp[LDT];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const double a = -(As[(c0 + 4 * ks + fk) * LDT + g + fr] * dpan[4 * ks + fk]);
                    dmma884(c0r, c1r, a, As[(c0 + 4 * ks + fk) * LDT + g + fr]);
                }
                cp[0] = c0r; cp[LDT] = c1r;
                __syncwarp();
#ifdef LDL_PROFILE
                const long long t_f0 = clock64();
                t_acc[4] += t_f0 - t_ph;
#endif
                if (!(LDL_ABLATE & 1)) diag_block_factor(As, g, NW + (pb ^ 1) * 128, NW + (pb ^ 1) * 128 + 64, dpan2 + 8 * (pb ^ 1), dall);
#ifdef LDL_PROFILE
                t_acc[5] += clock64() - t_f0;
#endif
            }
        } else if ((w & 3) != 0 && !(LDL_ABLATE & 2)) {
            // ---------------- B / C: the panel's block updates (table above), two per warp in flight.  Warps 4, 8 and 12 share warp 0's
            // SM sub-partition and stay out: the 8 x 8 factorisation of the look-ahead is a chain of dependent FP64 instructions that
            // measured 265 cycles per pivot while DMMAs of three other warps queued on the same FP64 pipe, against ~100 alone.
            constexpr int NWORK = 12;
            const int wi = w - 1 - (w >> 2);
            double dpf[2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) dpf[ks] = dpan[4 * ks + fk];
            const int ntask = tabn[p];
            const int laneA = fk * LDT + fr, laneC = 2 * fk * LDT + fr, laneR = fr * LDT + fk;
            constexpr int U = 3;         // block updates in flight per warp (51 tasks at most: two rounds)
            for (int t0 = wi; t0 < ntask; t0 += U * NWORK) {
                int4 td[U];
                double cr0[U], cr1[U], av[U][2], bv[U][2];
                double* cp[U];
#pragma unroll
                for (int u = 0; u < U; ++u) td[u] = tab[p * 64 + min(t0 + u * NWORK, ntask - 1)];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    cp[u] = sm + td[u].y + laneC;
                    const double* ap = sm + td[u].z + laneA;
                    const double* bp = sm + td[u].w + (td[u].x ? laneR : laneA);
                    const int bstep = td[u].x ? 4 : 4 * LDT;
                    cr0[u] = cp[u][0]; cr1[u] = cp[u][LDT];
                    av[u][0] = ap[0]; av[u][1] = ap[4 * LDT];
                    bv[u][0] = bp[0]; bv[u][1] = bp[bstep];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double s0 = td[u].x ? 1.0 : dpf[0], s1 = td[u].x ? 1.0 : dpf[1];
                    dmma884(cr0[u], cr1[u], -(av[u][0] * s0), bv[u][0]);
                    dmma884(cr0[u], cr1[u], -(av[u][1] * s1), bv[u][1]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (u == 0 || t0 + u * NWORK < ntask) { cp[u][0] = cr0[u]; cp[u][LDT] = cr1[u]; }
            }
        }
        LDL_ACC(66);
        __syncthreads();
        LDL_ACC(67);
    }
    LDL_STAMP(2);
    // ---- write L (strict lower) + D (diagonal), Linv (lower incl. the unit diagonal; the upper triangle stays zero) and y_J
    double* Li = Linv + (size_t)tk.col * ST2;
    if (!(LDL_ABLATE & 16))
    for (int k = w; k < ST; k += DIAG_WARPS)
        for (int i = k + lane; i < ST; i += 32) {
            if (i > k) { T[k * ST + i] = As[k * LDT + i]; Li[k * ST + i] = Mf[k * LDT + i]; }
            else { T[k * ST + i] = dall[k]; Li[k * ST + i] = 1.0; }
        }
    if (tid < ST) xp[(size_t)tk.col * ST + tid] = Mf[ST * LDT + tid];
    LDL_STAMP(3);
#ifdef LDL_PROFILE
    if (lane == 0) for (int i = 0; i < 6; ++i) prof[64 + 8 * w + i] = t_acc[i];
#endif
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
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const RedTask tk = tasks[blockIdx.x / GEMM_CTAS];
    const int I = GEMM_CTAS * (blockIdx.x % GEMM_CTAS) + w;      // row block of this warp
    double* T = S + (size_t)tk.tile * ST2;
    tile_rows24_to_smem_ld<GEMM_THREADS>(As, T, 24 * (blockIdx.x % GEMM_CTAS));   // (this CTA's 24 rows) T_IJ was last written by the previous level's update kernel, two launches back: safe before the wait
    pdl_wait(); pdl_trigger();
    tile_to_smem_ld<GEMM_THREADS>(Bs, Linv + (size_t)tk.col * ST2);
    if (tid < ST) Ds[tid] = rcp_fast(S[(size_t)tk.dtile * ST2 + (size_t)(ST + 1) * tid]);
    (void)xp;
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
    // scale by 1 / d_c and store L_IJ  (the right-hand side push b_I -= L_IJ y_J is done by the update kernel's diagonal targets,
    // which own block row I for the level: no reductions, so the solve is deterministic)
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) {
        const int c = 8 * cb + 2 * fk;
        T[(size_t)ST * c + 8 * I + fr] = c0[cb] * Ds[c];
        T[(size_t)ST * (c + 1) + 8 * I + fr] = c1[cb] * Ds[c + 1];
    }
}

// ---------------------------------------------------------------------------------------------------
// Update, one CTA triple per TARGET tile of the level:  T_{ab} -= sum_J (L_aJ D_J) L_bJ'  over the target's updates in a fixed
// order, accumulated in registers and subtracted once (owner-writes: several columns of a level hit the same tile; round 1 added
// them with FP64 reductions, which made the solve depend on scheduling).  Diagonal targets (a == b, block row I) also push the
// right-hand side  b_I -= sum_J L_IJ y_J  — every off-diagonal tile (I, J) of the level has exactly one such update.
// ---------------------------------------------------------------------------------------------------
constexpr size_t UPD_SMEM = (size_t)(2 * ST * LDT + 2 * ST) * sizeof(double);

__global__ void __launch_bounds__(GEMM_THREADS) ldl_upd_kernel(double* __restrict__ S, const RedUpd* __restrict__ upds, const RedTarget* __restrict__ targets,
                                                               double* __restrict__ xp) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                  // L_aJ [k][i]
    double* Bs = sm + ST * LDT;       // L_bJ [k][c]
    double* Ds = Bs + ST * LDT;       // [ST] D_J
    double* ys = Ds + ST;             // [ST] y_J
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const RedTarget tg = targets[blockIdx.x / GEMM_CTAS];
    const int I = GEMM_CTAS * (blockIdx.x % GEMM_CTAS) + w;
    const int fr = lane >> 2, fk = lane & 3;
    const bool diag = tg.row >= 0;
    const int ncb = diag ? I + 1 : NBLK;                      // diagonal targets: lower blocks only
    double c0[NBLK], c1[NBLK];
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) { c0[cb] = 0.0; c1[cb] = 0.0; }
    double pr = 0.0;
    pdl_wait(); pdl_trigger();
    for (int u = tg.u0; u < tg.u1; ++u) {
        const RedUpd up = upds[u];
        __syncthreads();             // the previous update is done with the staged tiles
        if (diag) tile_to_smem_ld<GEMM_THREADS>(As, S + (size_t)up.a * ST2);                       // A doubles as B
        else {
            tile_rows24_to_smem_ld<GEMM_THREADS>(As, S + (size_t)up.a * ST2, 24 * (blockIdx.x % GEMM_CTAS));   // this CTA's 24 rows of L_aJ
            tile_to_smem_ld<GEMM_THREADS>(Bs, S + (size_t)up.b * ST2);
        }
        if (tid < ST) {
            Ds[tid] = S[(size_t)up.dk * ST2 + (size_t)(ST + 1) * tid];
            if (diag) ys[tid] = xp[(size_t)up.col * ST + tid];
        }
        cp_async_wait_all();
        __syncthreads();
        const double* Bt = diag ? As : Bs;
        const double* ap = As + fk * LDT + 8 * I + fr;
        const double* bp = Bt + fk * LDT + fr;
#pragma unroll
        for (int ks = 0; ks < ST / 4; ++ks) {
            const double araw = ap[ks * 4 * LDT];
            const double a = araw * Ds[4 * ks + fk];
            if (diag) pr = fma(araw, ys[4 * ks + fk], pr);
#pragma unroll
            for (int cb = 0; cb < NBLK; ++cb) {
                if (cb >= ncb) continue;
                dmma884(c0[cb], c1[cb], a, bp[ks * 4 * LDT + 8 * cb]);
            }
        }
    }
    double* T = S + (size_t)tg.target * ST2;
#pragma unroll
    for (int cb = 0; cb < NBLK; ++cb) {
        if (cb >= ncb) continue;
        const int c = 8 * cb + 2 * fk;
        T[(size_t)ST * c + 8 * I + fr] -= c0[cb];
        T[(size_t)ST * (c + 1) + 8 * I + fr] -= c1[cb];
    }
    if (diag) {
        pr += __shfl_xor_sync(0xffffffffu, pr, 1);
        pr += __shfl_xor_sync(0xffffffffu, pr, 2);
        if (fk == 0) xp[(size_t)tg.row * ST + 8 * I + fr] -= pr;
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
    pdl_wait(); pdl_trigger();
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

// ---------------------------------------------------------------------------------------------------
// Backward sweep as ONE dataflow launch instead of one launch per level (12 dependent launches of ~11 us on the Venice shape):
// CTA b walks the columns order[b], order[b + G], ... (order = levels from last to first, so everything a column needs comes
// earlier in the list) and, instead of a kernel boundary, waits for the flag of every x_I it pulls.  All G <= #SM CTAs are
// resident, and a column only waits for columns that precede it in the list, so the walk cannot deadlock.
// flags[] is cleared (memset node) before the launch; flags[J] = 1 once x_J is in global memory.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wait_flag(const int* f) {
    while (*reinterpret_cast<const volatile int*>(f) == 0) { __nanosleep(20); }
    __threadfence();
}
__global__ void __launch_bounds__(RED_THREADS) ldl_bwd_flow_kernel(const double* __restrict__ S, const double* __restrict__ Linv, RedSolveLists t,
                                                                   const int* __restrict__ order, int ncols, int* __restrict__ flags, double* __restrict__ x) {
    extern __shared__ __align__(16) double sm[];
    double* Ts = sm;                 // [2][ST2]
    double* xi = sm + 2 * ST2;       // [2][ST]
    double* part = xi + 2 * ST;      // [2][ST]
    double* w = part + 2 * ST;       // [ST]
    const int tid = threadIdx.x, c = tid % ST, h = tid / ST;
    pdl_wait(); pdl_trigger();
    for (int pos = blockIdx.x; pos < ncols; pos += gridDim.x) {
        const int J = order[pos];
        const int q0 = t.colptr[J], q1 = t.colptr[J + 1];
        double acc = 0.0;
        __syncthreads();             // the previous column of this CTA is done with the buffers
        if (q0 < q1) {
            tile_to_smem<RED_THREADS>(Ts, S + (size_t)t.col_tile[q0] * ST2);      // the tiles do not depend on other columns: in flight while we wait
            if (tid == 0) wait_flag(flags + t.col_row[q0]);
            __syncthreads();
            if (tid < ST) xi[tid] = __ldcg(x + (size_t)t.col_row[q0] * ST + tid);
        }
        for (int q = q0; q < q1; ++q) {
            const int st = (q - q0) & 1;
            cp_async_wait_all();
            __syncthreads();
            if (q + 1 < q1) {
                tile_to_smem<RED_THREADS>(Ts + (st ^ 1) * ST2, S + (size_t)t.col_tile[q + 1] * ST2);
                if (tid == 0) wait_flag(flags + t.col_row[q + 1]);
            }
            const double* M = Ts + st * ST2 + (size_t)ST * c + (ST / 2) * h;
            const double* xv = xi + st * ST + (ST / 2) * h;
#pragma unroll 12
            for (int i = 0; i < ST / 2; ++i) acc = fma(M[i], xv[i], acc);
            if (q + 1 < q1) {
                __syncthreads();     // thread 0 has seen the flag
                if (tid < ST) xi[(st ^ 1) * ST + tid] = __ldcg(x + (size_t)t.col_row[q + 1] * ST + tid);
            }
        }
        __syncthreads();
        tile_to_smem<RED_THREADS>(Ts, Linv + (size_t)J * ST2);
        part[h * ST + c] = acc;
        cp_async_wait_all();
        __syncthreads();
        if (tid < ST) w[tid] = __ldcg(x + (size_t)J * ST + tid) / S[(size_t)t.diag_tile[J] * ST2 + (size_t)(ST + 1) * tid] - (part[tid] + part[ST + tid]);
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
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicExch(flags + J, 1);
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward sweep, dataflow form, second version (the default).  The first version above measured no faster than one launch per
// level: what a level costs is not the launch but the dependent chain INSIDE a column — wait for x_I, stage it, a 36-long FMA
// chain on bank-conflicted columns (stride 72 doubles: a half-warp hits 2 bank pairs), then a 41 KB load of Linv_J that only
// starts after the products, another 36-long chain, fence, flag.  Here everything that does not depend on other columns is in
// shared memory before the first wait (up to BW_NBUF row tiles, Linv_J, y_J / D_J), tiles are staged with a row stride of 74
// doubles (a half-warp reads 8 distinct bank pairs), x_I goes straight from L2 into registers, and 288 threads split every
// dot product into four 18-element pieces with two accumulators each (9 dependent FMAs).  The column also writes its part of
// the solution in natural numbering (out), which removes the trailing permutation launch.
// ---------------------------------------------------------------------------------------------------
constexpr int BW_LD = 74;
constexpr int BW_THREADS = 4 * ST;        // 288
constexpr int BW_NBUF = 3;
constexpr int BW_Q = ST / 4;              // 18 elements per thread
constexpr size_t BWD2_SMEM = (size_t)((BW_NBUF + 1) * ST * BW_LD + 4 * ST + 2 * ST) * sizeof(double);

template <int NT_>
__device__ __forceinline__ void tile_to_smem_bw(double* sdst, const double* gsrc) {
    for (int q = threadIdx.x; q < ST2 / 2; q += NT_) {
        const int k = q / (ST / 2), c = q - k * (ST / 2);
        cp_async16(sdst + k * BW_LD + 2 * c, gsrc + k * ST + 2 * c);
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(BW_THREADS) ldl_bwd_flow2_kernel(const double* __restrict__ S, const double* __restrict__ Linv, RedSolveLists t,
                                                                    const int* __restrict__ order, const int* __restrict__ nat_of_pos, int ncols,
                                                                    int* __restrict__ flags, double* __restrict__ x, double* __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    double* Ts = sm;                                   // [BW_NBUF][ST * BW_LD]
    double* Li = sm + BW_NBUF * ST * BW_LD;            // Linv_J
    double* part = Li + ST * BW_LD;                    // [4][ST]
    double* yd = part + 4 * ST;                        // y_J / D_J
    double* w = yd + ST;
    const int tid = threadIdx.x, c = tid % ST, h = tid / ST;
    pdl_wait(); pdl_trigger();
    for (int pos = blockIdx.x; pos < ncols; pos += gridDim.x) {
        const int J = order[pos];
        const int q0 = t.colptr[J], q1 = t.colptr[J + 1], nq = q1 - q0;
        __syncthreads();             // the previous column of this CTA is done with the buffers
        // ---- everything that does not depend on other columns: Linv_J (group 0), the first row tiles (one group each), y_J / D_J
        tile_to_smem_bw<BW_THREADS>(Li, Linv + (size_t)J * ST2);
        cp_async_commit();
#pragma unroll
        for (int b = 0; b < BW_NBUF; ++b) {
            if (b < nq) tile_to_smem_bw<BW_THREADS>(Ts + b * ST * BW_LD, S + (size_t)t.col_tile[q0 + b] * ST2);
            cp_async_commit();
        }
        if (tid < ST) yd[tid] = x[(size_t)J * ST + tid] / S[(size_t)t.diag_tile[J] * ST2 + (size_t)(ST + 1) * tid];   // y_J: forward phase, complete
        double a0 = 0.0, a1 = 0.0;
        for (int q = 0; q < nq; ++q) {
            const int row = t.col_row[q0 + q];
            if (tid == 0) wait_flag(flags + row);
            // groups are committed in order: Linv, tiles 0 .. BW_NBUF - 1, then one (possibly empty) refill group per iteration, so tile q
            // is group 1 + q of 1 + BW_NBUF + q committed: complete once at most BW_NBUF - 1 groups are pending
            cp_async_wait_group<BW_NBUF - 1>();
            __syncthreads();         // flag seen by thread 0, tile q visible to everyone
            double xv[BW_Q];
            const double2* xg = reinterpret_cast<const double2*>(x + (size_t)row * ST + BW_Q * h);
#pragma unroll
            for (int i = 0; i < BW_Q / 2; ++i) { const double2 v = __ldcg(xg + i); xv[2 * i] = v.x; xv[2 * i + 1] = v.y; }
            const double* M = Ts + (q % BW_NBUF) * ST * BW_LD + (size_t)BW_LD * c + BW_Q * h;
#pragma unroll
            for (int i = 0; i < BW_Q; i += 2) { a0 = fma(M[i], xv[i], a0); a1 = fma(M[i + 1], xv[i + 1], a1); }
            __syncthreads();         // buffer q % BW_NBUF is free
            if (q + BW_NBUF < nq) tile_to_smem_bw<BW_THREADS>(Ts + (q % BW_NBUF) * ST * BW_LD, S + (size_t)t.col_tile[q0 + q + BW_NBUF] * ST2);
            cp_async_commit();       // (possibly empty) keeps the group count in step with q
        }
        part[h * ST + c] = a0 + a1;
        cp_async_wait_group<0>();
        __syncthreads();
        if (tid < ST) w[tid] = yd[tid] - ((part[tid] + part[ST + tid]) + (part[2 * ST + tid] + part[3 * ST + tid]));
        __syncthreads();
        {   // x_J[c] = sum_{i >= c} Linv[i][c] w[i]
            const double* M = Li + (size_t)BW_LD * c + BW_Q * h;
            const double* wv = w + BW_Q * h;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < BW_Q; i += 2) {
                s0 = fma((BW_Q * h + i >= c) ? M[i] : 0.0, wv[i], s0);
                s1 = fma((BW_Q * h + i + 1 >= c) ? M[i + 1] : 0.0, wv[i + 1], s1);
            }
            part[h * ST + c] = s0 + s1;
        }
        __syncthreads();
        if (tid < ST) {
            const double v = (part[tid] + part[ST + tid]) + (part[2 * ST + tid] + part[3 * ST + tid]);
            x[(size_t)J * ST + tid] = v;
            out[(size_t)nat_of_pos[J] * ST + tid] = v;
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicExch(flags + J, 1); }
    }
}

// natural <-> permuted tile numbering of the right-hand side / solution
__global__ void red_permute_kernel(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ pos, int NT, int to_permuted,
                                   int* __restrict__ flags = nullptr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait(); pdl_trigger();
    if (flags && k < NT) flags[k] = 0;          // the dataflow backward sweep's column flags (first node of the solve)
    if (k >= NT * ST) return;
    const int o = k / ST, r = k - o * ST;
    if (to_permuted) dst[(size_t)pos[o] * ST + r] = src[k];
    else dst[k] = src[(size_t)pos[o] * ST + r];
}

// S tiles <- 0 (memset) except diagonal tiles: lower triangle of U_c + lambda I (multi-rank: added by ONE rank per tile — add_u_tile,
// the tile's owner in the exchange by ownership); padding rows get a unit diagonal.  rhs (natural numbering) <- g_c (rank 0).
template <int DC>
__global__ void red_init_kernel(double* __restrict__ S, const int* __restrict__ diag_tile_nat, const double* __restrict__ H, const double* __restrict__ g,
                                double* __restrict__ rhs, int nA, double lambda, const int* __restrict__ add_u_tile, int add_g) {
    constexpr int TC = ST / DC;
    const int o = blockIdx.x;   // natural tile index
    const int add_u = add_u_tile[o];
    double* T = S + (size_t)diag_tile_nat[o] * ST2;
    if (add_u)
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
        rhs[k] = (add_g && k < (long long)DC * nA) ? g[k] : 0.0;
    }
}

}  // namespace nlls
