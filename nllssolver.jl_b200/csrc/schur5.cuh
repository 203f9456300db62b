// Schur v5: point-wise tensor-core Schur elimination with window-aligned register accumulation (plan: schur5_plan.hpp).
//
// What bounded v4 (ncu, round 1): 121 M shared-memory wavefronts — every block contribution (i, j) of a point re-read W_i and
// Y_j from shared memory (k (k + 1) fragment loads for k (k + 1) / 2 DMMAs) after a separate phase had written Y there.  Here a
// warp takes a WHOLE point: its W_p (3 x DC k, one contiguous piece of the staged H span) is read once as B fragments (8 columns
// each) and once more to form the A fragments  Y_p' = W_p' A_p^-1  on the fly (three broadcast loads + three FMAs per lane), and
// the lower-triangular 8 x 8 output tiles of  S_p = Y_p' W_p  are accumulated in registers across all the points of a
// super-tile (window coordinates make the tiles of different points coincide).  No Y buffer, no per-contribution entries:
// ~33 M wavefronts and 12.8 M DMMAs on the Venice shape instead of 121 M and 16.5 M.
//
// Two formulations (Schur5Cfg<DC>::ZT, schur5_plan.hpp): the classic one above (dc = 6: three bands of the window, 11 consumer warps)
// and the factored one (dc = 9: two six-row bands, 7 consumer warps with 253 registers):  A_p^-1 = L_p D_p L_p' (3 x 3 LDL'),
// Z_p = L_p' W_p written over the staged W_p by the producer, S_p = Z_p' D_p Z_p — one load per 8-row tile of a point then serves
// the B fragment, the A fragment (d[kk] times the same values) and the rhs.  The variants tried around these two and their
// same-box timings are listed in DESIGN.md section 4.
//
// CTA = Schur5Cfg<DC>::CONS consumer warps + 1 producer warp, one CTA per SM, a contiguous range of point tiles each:
//   producer   TMA bulk loads (H span, g_p, the tile's entry blob) NS - 1 tiles ahead into a ring of NS stages
//              (mbarrier complete_tx), then per point A_p^-1 = (V_p + lambda I)^-1 into the stage's point table
//              ([Ainv row kk | g_kk] per inner index kk) and to global memory for the back-substitution;
//   consumers  wait for the stage (mbarrier), walk their entry list — one entry = (point, band of BR row tiles) — and release
//              the stage (mbarrier).  No CTA-wide barrier anywhere: warps drift up to NS - 1 tiles apart, which absorbs the
//              per-tile imbalance.  A FLUSH entry ends a super-tile: the warp's tiles go to S with FP64 reductions.
#pragma once
#include "common.cuh"
#include "reduced.cuh"
#include "schur5_plan.hpp"
#include <utility>

namespace nlls {

#ifndef S5_FLUSH_TABLE
#define S5_FLUSH_TABLE 1   // narrow-band configurations (12 warps) flush with the table-walk routines, the wide ones with a per-FLUSH context
#endif
#ifndef S5_NSTAGES
#define S5_NSTAGES 3
#endif
constexpr int S5_NS = S5_NSTAGES;      // stages
constexpr int S5_PAD = 3072;  // bytes in front of / behind the stages: fragment loads of window rows outside a point's own rows are
                              // not clamped (their values are discarded), they only have to stay inside the CTA's shared memory

template <int DC> struct Schur5Smem {
    using C = Schur5Cfg<DC>;
    static constexpr int ROW = C::WB * C::OBS + 9 * C::PTS;        // doubles of H span per stage
    static constexpr int ROWS = ROW + 4;                            // + slack of an 8-byte-misaligned span (stays even)
    static constexpr int PTAB = 16 * C::PTS;                        // per point 4 x [Ainv[kk][0..2], g[kk]]; kk = 3 stays zero
    static constexpr int GST = (3 * C::PTS + 4) & ~1;               // staged g_p
    static constexpr int BLOB = (S5_HDR + C::PTS / 2 + 2 + C::OBS / 4 + 2 + 2 * C::ENT_CAP + 3) & ~3;   // u32 words
    static constexpr size_t stage_bytes = (size_t)(ROWS + PTAB + GST) * sizeof(double) + (size_t)BLOB * sizeof(unsigned);
    static constexpr size_t bytes = 2 * S5_PAD + S5_NS * stage_bytes + 3 * S5_NS * sizeof(uint64_t) + 16;
    static_assert(stage_bytes % 16 == 0, "stages must stay 16-byte aligned");
    static_assert(24 * 8 * C::NTW + 256 <= S5_PAD, "unclamped fragment loads reach at most one window (24 bytes per scalar row) outside a span");
    static_assert(bytes <= 232448, "stages must fit the 227 KB of shared memory a CTA can have");
};

struct Schur5Dev {
    const int* cta_item;
    const Schur5Item* items;
    const unsigned* blob;
    const long long* ftab;   // FLUSH table (schur5_flush_table)
    int ncons;        // consumer warps in use (<= S5_CONSUMERS; the others leave at once)
    long long* dbg;   // != nullptr (NLLS_B200_S5DBG): per CTA and warp [16][4] cycle counters — consumers: {waiting for a tile, total};
                      // producer: {point phases, total, polls without work}
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
template <int OFF>
__device__ __forceinline__ double lds_f64_at(uint32_t base) {   // [base + OFF]: the offset goes into the instruction's immediate field
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(base), "n"(OFF));
    return v;
}

// ---- FLUSH helpers: one accumulator pair (window row R, window columns C0, C0 + 1) / one rhs partial into the reduced system ----
// Everything that depends only on the super-tile (window base relative to its first camera tile, number of valid cameras, the S tiles
// the window can touch) is worked out once per FLUSH by the caller; the per-pair routine is ~30 instructions.  (The first version
// re-derived all of it per pair from the table in global memory — 150 instructions per call; harmless with three bands, a quarter of
// the kernel once a warp flushes the whole window: 45 pairs per lane.)
template <int DC> struct S5FlushCtx {
    int boff;        // window base camera - first camera of camera tile I0
    int camlim;      // cameras of the window that exist: min(WC, nA - base)
    long long pe[Schur5Flush<DC>::NPAIR];   // 2 * element offset of S tile (I0 + a, I0 + b) + transposed flag, or -1
};
template <int DC>
__device__ __noinline__ void s5_flush_tile(double v0, double v1, int R, int C0, const S5FlushCtx<DC>& fc, double* __restrict__ S) {
    using F = Schur5Flush<DC>;
    const int ca = R / DC, ar = R - ca * DC;
    if (ca >= fc.camlim) return;
    const int crel = fc.boff + ca, ta = crel / F::TC, r0 = (crel - ta * F::TC) * DC + ar;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int Cc = C0 + h, cb = Cc / DC, cr = Cc - cb * DC;
        const double v = h ? v1 : v0;
        if (v != 0.0 && (cb < ca || (cb == ca && cr <= ar))) {
            const int crelb = fc.boff + cb, tb = crelb / F::TC, c0 = (crelb - tb * F::TC) * DC + cr;
            const long long pe = fc.pe[F::pair(ta, tb)];
            atomicAdd(S + (pe >> 1) + ((pe & 1) ? c0 + ST * r0 : r0 + ST * c0), -v);
        }
    }
}
template <int DC>
__device__ __noinline__ void s5_flush_rhs(double rv, int R, int base, int camlim, double* __restrict__ rhs, int kk) {
    rv += __shfl_xor_sync(0xffffffffu, rv, 1);
    rv += __shfl_xor_sync(0xffffffffu, rv, 2);
    const int ca = R / DC, ar = R - ca * DC;
    if (kk == 0 && ca < camlim && rv != 0.0) atomicAdd(rhs + (size_t)(base + ca) * DC + ar, -rv);
}

// table-walk variants (S5_FLUSH_TABLE): everything re-derived per pair from the FLUSH table in global memory
template <int DC>
__device__ __noinline__ void s5_flush_tile_t(double v0, double v1, int R, int C0, const long long* __restrict__ ft, int nA, double* __restrict__ S) {
    using F = Schur5Flush<DC>;
    const int base = (int)ft[0], I0 = (int)ft[1];
    const int ca = R / DC, ar = R - ca * DC;
    const int crow = base + ca;
    if (ca >= Schur5Cfg<DC>::WC || crow >= nA) return;
    const int ta = crow / F::TC - I0, r0 = (crow - (I0 + ta) * F::TC) * DC + ar;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int Cc = C0 + h, cb = Cc / DC, cr = Cc - cb * DC;
        const double v = h ? v1 : v0;
        if (v != 0.0 && (cb < ca || (cb == ca && cr <= ar))) {
            const int ccol = base + cb, tb = ccol / F::TC - I0, c0 = (ccol - (I0 + tb) * F::TC) * DC + cr;
            const long long pe = ft[2 + F::pair(ta, tb)];
            atomicAdd(S + (pe >> 1) + ((pe & 1) ? c0 + (long long)ST * r0 : r0 + (long long)ST * c0), -v);
        }
    }
}
template <int DC>
__device__ __noinline__ void s5_flush_rhs_t(double rv, int R, const long long* __restrict__ ft, int nA, double* __restrict__ rhs, int kk) {
    rv += __shfl_xor_sync(0xffffffffu, rv, 1);
    rv += __shfl_xor_sync(0xffffffffu, rv, 2);
    const int ca = R / DC, ar = R - ca * DC;
    const int crow = (int)ft[0] + ca;
    if (kk == 0 && ca < Schur5Cfg<DC>::WC && crow < nA && rv != 0.0) atomicAdd(rhs + (size_t)crow * DC + ar, -rv);
}

// ---- the entries of one consumer warp for one tile ------------------------------------------------------------------------------
// BAND (and with it every accumulator index) is a compile-time constant, and so is the SHAPE of an entry: the first window tile of
// the point (TLO) and how far into the band it reaches (E = min(t_hi, M0 + BR - 1) - M0).  One indirect branch per entry picks the
// shape; the code behind it is straight-line — exactly the loads, FMAs and DMMAs the shape needs, no tests, no predicates.  (The
// first version tested every (row, column) site at run time: 300 instructions per entry, most of them dependent
// LOP3 -> branch pairs, and the 3 warps a scheduler has could not hide them: ncu `wait` 2.8 + `branch_resolving` 0.6 per issue.)
template <int DC, int BAND>
struct Schur5Band {
    using C = Schur5Cfg<DC>;
    static constexpr int NTW = C::NTW, BR = C::BR, NB = C::NBANDS;
    __host__ __device__ static constexpr int MT(int r) { return C::row_tile(BAND, r); }
    static constexpr int NCOL = MT(BR - 1) + 1;
    // shape key = t_lo * (BR + 1) + number of active rows; for a static TLO the active rows are R0 .. R0 + NACT - 1
    __host__ __device__ static constexpr int R0(int tlo) { return C::r0(BAND, tlo); }

    // The staged W_p has been replaced in place by Z_p = L_p' W_p (producer), A_p^-1 = L_p D_p L_p':  S_p = Z_p' D_p Z_p.  The A fragment
    // of row tile r is then d[kk] times the B fragment of column tile r — ONE load per tile of the point serves both (round 2b formed
    // Y = A_p^-1 W on the fly: three loads, three FP64 operations and a mask per row tile and lane, the instructions ncu showed throttled
    // by the other warps' DMMAs on the shared FP64 pipe).  Columns run from the point's last tile down, so a row's A fragment exists
    // exactly when the first column that needs it comes up and only one B fragment is live at a time.
    __host__ __device__ static constexpr int row_of_tile(int mt) {   // index of window tile mt among the band's rows, or -1
        for (int r = 0; r < BR; ++r) if (MT(r) == mt) return r;
        return -1;
    }
    // ---- classic formulation (ZT == false): Y = A_p^-1 W formed on the fly from three loads per row tile
    template <int TLO, int NACT, int R>
    static __device__ __forceinline__ void row(double (&a)[BR], double (&racc)[BR], uint32_t wb, int u, int lim, const double2& pa, const double2& pb) {
        if constexpr (R >= R0(TLO) && R < R0(TLO) + NACT) {   // A fragment: Y[kk][window row 8 MT + fr] = sum_m Ainv[kk][m] W[m][row]
            constexpr int mt = MT(R);
            const double w0 = lds_f64_at<192 * mt>(wb), w1 = lds_f64_at<192 * mt + 8>(wb), w2 = lds_f64_at<192 * mt + 16>(wb);
            const double y = fma(pb.x, w2, fma(pa.y, w1, pa.x * w0));
            a[R] = ((unsigned)(u + 8 * mt) < (unsigned)lim) ? y : 0.0;   // rows of the tile outside the point's own rows (first / last tile)
            racc[R] = fma(a[R], pb.y, racc[R]);
        } else {
            a[R] = 0.0;
        }
    }
    template <int TLO, int NACT, int N>
    static __device__ __forceinline__ void col(double (&acc)[BR][NTW][2], const double (&a)[BR], uint32_t wbb, int u, int lim) {
        constexpr int last = MT(R0(TLO) + NACT - 1);           // the last active row tile: columns run up to it
        if constexpr (N >= TLO && N <= last) {                 // B fragment: W[kk][window column 8 N + fr]
            double b = lds_f64_at<192 * N>(wbb);
            if constexpr (N == TLO || N == last) b = ((unsigned)(u + 8 * N) < (unsigned)lim) ? b : 0.0;   // only the point's first / last tile can be partial
#pragma unroll
            for (int r = 0; r < BR; ++r)
                if (r >= R0(TLO) && r < R0(TLO) + NACT && MT(r) >= N) dmma884(acc[r][N][0], acc[r][N][1], a[r], b);
        }
    }
    // ---- Z formulation
    template <int TLO, int NACT, int N>
    static __device__ __forceinline__ void colz(double (&acc)[BR][NTW][2], double (&a)[BR], double (&racc)[BR], uint32_t wbb, int u, int lim, const double2& pa) {
        constexpr int last = MT(R0(TLO) + NACT - 1);           // the last active row tile: columns run up to it
        if constexpr (N >= TLO && N <= last) {                 // B fragment: Z[kk][window column 8 N + fr]
            double b = lds_f64_at<192 * N>(wbb);
            if constexpr (N == TLO || N == last) b = ((unsigned)(u + 8 * N) < (unsigned)lim) ? b : 0.0;   // only the point's first / last tile can be partial
            constexpr int rn = row_of_tile(N);
            if constexpr (rn >= 0) {                           // N is one of the band's (active) rows: its A fragment d[kk] Z[kk][row]
                a[rn] = pa.x * b;
                racc[rn] = fma(a[rn], pa.y, racc[rn]);         // rhs: (D Z)' (L' g)
            }
#pragma unroll
            for (int r = 0; r < BR; ++r)
                if (r >= R0(TLO) && r < R0(TLO) + NACT && MT(r) >= N) dmma884(acc[r][N][0], acc[r][N][1], a[r], b);
        }
    }
    template <int ID, int... Rs, int... Ns>
    static __device__ __forceinline__ void shape(double (&acc)[BR][NTW][2], double (&racc)[BR], uint32_t wb, int kk, int u, int lim, const double2& pa,
                                                 const double2& pb, std::integer_sequence<int, Rs...>, std::integer_sequence<int, Ns...>) {
        if constexpr (ID < C::nshapes(BAND)) {
            constexpr int TLO = C::shape_tlo(BAND, ID), NACT = C::shape_nact(BAND, ID);
            double a[BR];
            const uint32_t wbb = wb + 8u * (unsigned)kk;
            if constexpr (C::ZT) {
                (colz<TLO, NACT, NCOL - 1 - Ns>(acc, a, racc, wbb, u, lim, pa), ...);
            } else {
                (row<TLO, NACT, Rs>(a, racc, wb, u, lim, pa, pb), ...);
                (col<TLO, NACT, Ns>(acc, a, wbb, u, lim), ...);
            }
        }
    }
    // entry words (host: schur5_plan.hpp): x = wofs | DC k << 16 | DC delta << 24,  y = local point | shape << 8 | band << 16
    template <int ID, int... Rs, int... Ns>
    static __device__ __forceinline__ void entry(double (&acc)[BR][NTW][2], double (&racc)[BR], const uint2 ent, uint32_t rowb, uint32_t ptb, int fr, int kk,
                                                 std::integer_sequence<int, Rs...> rs, std::integer_sequence<int, Ns...> ns) {
        const int lim = (int)((ent.x >> 16) & 255u);
        const int u = fr - (int)(ent.x >> 24);
        const uint32_t wb = rowb + 8u * (ent.x & 0xffffu) + 24u * (unsigned)fr;   // W[0][window row fr]   (rowb carries - 8 BIAS)
        const uint32_t pq = ptb + 128u * (ent.y & 255u);
        const double2 pa = lds_f64x2(pq);                                      // ZT: d[kk], (L' g)[kk]; classic: Ainv[kk][0..1]   (kk = 3: zeros)
        double2 pb = make_double2(0.0, 0.0);
        if constexpr (!C::ZT) pb = lds_f64x2(pq + 16u);                        // classic: Ainv[kk][2], g[kk]
        shape<ID>(acc, racc, wb, kk, u, lim, pa, pb, rs, ns);
    }
    // end of a super-tile: add this warp's tiles to S (lower triangle of the window, cameras < nA) and clear them.  The per-tile work is
    // a shared, non-inlined routine (values passed in registers): inlined three times per band it was a third of the kernel's code,
    // and the instruction cache is what the consumers' straight-line shape code needs.
    static __device__ __forceinline__ void flush(double (&acc)[BR][NTW][2], double (&racc)[BR], const long long* __restrict__ ft, const DevProblem& p,
                                                 double* __restrict__ S, double* __restrict__ rhs, int fr, int kk) {
#if S5_FLUSH_TABLE
        if constexpr (!C::WIDE) {
#pragma unroll
            for (int r = 0; r < BR; ++r) {
                s5_flush_rhs_t<DC>(racc[r], 8 * MT(r) + fr, ft, p.nA, rhs, kk);
                racc[r] = 0.0;
#pragma unroll
                for (int n = 0; n < NCOL; ++n) {
                    if (n <= MT(r)) {
                        s5_flush_tile_t<DC>(acc[r][n][0], acc[r][n][1], 8 * MT(r) + fr, 8 * n + 2 * kk, ft, p.nA, S);
                        acc[r][n][0] = 0.0; acc[r][n][1] = 0.0;
                    }
                }
            }
            return;
        }
#endif
        using F = Schur5Flush<DC>;
        S5FlushCtx<DC> fc;
        const int base = (int)ft[0];
        fc.boff = base - (int)ft[1] * F::TC;
        fc.camlim = min((int)C::WC, p.nA - base);
#pragma unroll
        for (int q = 0; q < F::NPAIR; ++q) fc.pe[q] = ft[2 + q];
#pragma unroll
        for (int r = 0; r < BR; ++r) {
            s5_flush_rhs<DC>(racc[r], 8 * MT(r) + fr, base, fc.camlim, rhs, kk);
            racc[r] = 0.0;
#pragma unroll
            for (int n = 0; n < NCOL; ++n) {
                if (n <= MT(r)) {
                    s5_flush_tile<DC>(acc[r][n][0], acc[r][n][1], 8 * MT(r) + fr, 8 * n + 2 * kk, fc, S);
                    acc[r][n][0] = 0.0; acc[r][n][1] = 0.0;
                }
            }
        }
    }
    // The host sorts a warp's entries of a tile by shape: the switch runs once per run of equal shapes, the run itself is a tight loop
    // over straight-line code (the compiler lowers the switch to a compare tree — paying that per entry cost 14 branches each).
    static __device__ __forceinline__ void run(double (&acc)[BR][NTW][2], double (&racc)[BR], uint32_t ep, unsigned cnt, uint32_t rowb, uint32_t ptb,
                                               const DevProblem& p, const long long* __restrict__ ftab, double* __restrict__ S, double* __restrict__ rhs, int fr, int kk) {
        constexpr auto RS = std::make_integer_sequence<int, BR>{};
        constexpr auto NS_ = std::make_integer_sequence<int, NCOL>{};
        static_assert(C::nshapes(BAND) <= 48, "shape switch");
        const uint32_t eend = ep + 8u * cnt;
        uint2 ent = lds_u2(ep);
        while (ep < eend) {
            if (ent.y & S5_FLUSH) { flush(acc, racc, ftab + (size_t)ent.x * Schur5Flush<DC>::STRIDE, p, S, rhs, fr, kk); break; }   // always the last entry of a list
            const unsigned id = (ent.y >> 8) & 255u;
#define S5_CASE(K)                                                                   \
    case K:                                                                          \
        do {                                                                         \
            const uint2 cur = ent;                                                   \
            ep += 8u;                                                                \
            ent = lds_u2(ep);  /* one entry past the list is still inside the blob */ \
            entry<K>(acc, racc, cur, rowb, ptb, fr, kk, RS, NS_);                    \
        } while (ep < eend && ((ent.y >> 8) & 0x10ffu) == K);                        \
        break;
            if constexpr (C::nshapes(BAND) <= 24) {   // (the unused cases of the larger switch are not free: deeper compare tree, larger code)
                switch (id) {
                    S5_CASE(0) S5_CASE(1) S5_CASE(2) S5_CASE(3) S5_CASE(4) S5_CASE(5) S5_CASE(6) S5_CASE(7) S5_CASE(8) S5_CASE(9)
                    S5_CASE(10) S5_CASE(11) S5_CASE(12) S5_CASE(13) S5_CASE(14) S5_CASE(15) S5_CASE(16) S5_CASE(17) S5_CASE(18) S5_CASE(19)
                    S5_CASE(20) S5_CASE(21) S5_CASE(22) S5_CASE(23)
                    default: ep = eend; break;
                }
            } else {
                switch (id) {
                    S5_CASE(0) S5_CASE(1) S5_CASE(2) S5_CASE(3) S5_CASE(4) S5_CASE(5) S5_CASE(6) S5_CASE(7) S5_CASE(8) S5_CASE(9)
                    S5_CASE(10) S5_CASE(11) S5_CASE(12) S5_CASE(13) S5_CASE(14) S5_CASE(15) S5_CASE(16) S5_CASE(17) S5_CASE(18) S5_CASE(19)
                    S5_CASE(20) S5_CASE(21) S5_CASE(22) S5_CASE(23) S5_CASE(24) S5_CASE(25) S5_CASE(26) S5_CASE(27) S5_CASE(28) S5_CASE(29)
                    S5_CASE(30) S5_CASE(31) S5_CASE(32) S5_CASE(33) S5_CASE(34) S5_CASE(35) S5_CASE(36) S5_CASE(37) S5_CASE(38) S5_CASE(39)
                    S5_CASE(40) S5_CASE(41) S5_CASE(42) S5_CASE(43) S5_CASE(44) S5_CASE(45) S5_CASE(46) S5_CASE(47)
                    default: ep = eend; break;
                }
            }
#undef S5_CASE
        }
    }
};

template <int DC>
__global__ void __launch_bounds__(Schur5Cfg<DC>::THREADS, 1) schur5_kernel(DevProblem p, Schur5Dev sp, double* __restrict__ S, double* __restrict__ rhs,
                                                                double* __restrict__ Ainv_out, double lambda) {
    using C = Schur5Cfg<DC>;
    using SM = Schur5Smem<DC>;
    constexpr int WB = C::WB, NTW = C::NTW, BR = C::BR, NS = S5_NS, S5_THREADS = C::THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ka = sp.cta_item[blockIdx.x], nitem = sp.cta_item[blockIdx.x + 1] - ka;
    if (nitem <= 0) return;
    const Schur5Item* items = sp.items + ka;
    unsigned char* stages = smem_raw + S5_PAD;
    auto s_row = [&](int st) { return reinterpret_cast<double*>(stages + (size_t)st * SM::stage_bytes); };
    auto s_pt = [&](int st) { return s_row(st) + SM::ROWS; };
    auto s_g = [&](int st) { return s_pt(st) + SM::PTAB; };
    auto s_blob = [&](int st) { return reinterpret_cast<unsigned*>(s_g(st) + SM::GST); };
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(stages + (size_t)NS * SM::stage_bytes + S5_PAD);
    uint64_t* bar_ready = bar_full + NS;
    uint64_t* bar_empty = bar_ready + NS;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_ready[s], (uint32_t)C::PROD); mbar_init(&bar_empty[s], (uint32_t)sp.ncons); }
    }
    // the point tables' kk = 3 slots stay zero for the whole kernel: they make the fourth inner slot of every A fragment zero
    for (int s = 0; s < NS; ++s) {
        double* pt = s_pt(s);
        for (int i = tid; i < SM::PTAB; i += S5_THREADS) pt[i] = 0.0;
    }
    __syncthreads();

    if (warp == C::CONS) {
        // =============================================== producer ===============================================
        auto issue = [&](int j) {   // lane 0: bulk loads of item j into stage j % NS
            const int st = j % NS;
            const Schur5Item it = items[j];
            const int mis = (it.flags >> 2) & 1;
            const double* gsrc = p.H + (size_t)p.hB + (size_t)WB * it.ob0 + (size_t)9 * it.pt0 - mis;
            const uint32_t span = (uint32_t)((WB * it.nob + 9 * it.npt + mis + 1) & ~1) * 8u;
            const size_t g0 = (size_t)p.gB + (size_t)3 * it.pt0;
            const int gm = (int)(g0 & 1);
            const uint32_t gbytes = (uint32_t)((3 * it.npt + gm + 1) & ~1) * 8u;
            const uint32_t bl = it.nblob * 4u;
            mbar_expect_tx(&bar_full[st], span + gbytes + bl);
            bulk_load(s_row(st), gsrc, span, &bar_full[st]);
            bulk_load(s_g(st), p.g + g0 - gm, gbytes, &bar_full[st]);
            bulk_load(s_blob(st), sp.blob + it.blob0, bl, &bar_full[st]);
        };
        // Two independent duties, polled: refill a stage as soon as the consumers have released it, and run the point phase of a
        // tile as soon as its bulk loads have landed — the second must never wait for the first (a consumer that is ahead of the
        // others needs the NEXT tile's table while the slowest one still holds the stage the next refill wants).
        int ji = 0, jp = 0;   // next item to load / next item whose point table is due
        const long long t_begin = clock64();
        long long t_pp = 0, n_idle = 0;
        while (jp < nitem) {
            int can_issue = 0, can_pp = 0;
            if (lane == 0) {
                if (ji < nitem) can_issue = (ji < NS) ? 1 : (int)mbar_test(&bar_empty[ji % NS], (uint32_t)((ji / NS - 1) & 1));
                if (can_issue) issue(ji);
                if (jp < ji + can_issue) can_pp = (int)mbar_test(&bar_full[jp % NS], (uint32_t)((jp / NS) & 1));
            }
            can_issue = __shfl_sync(0xffffffffu, can_issue, 0);
            can_pp = __shfl_sync(0xffffffffu, can_pp, 0);
            ji += can_issue;
            if (!can_pp) { if (!can_issue) { __nanosleep(64); ++n_idle; } continue; }
            const long long t_p0 = clock64();
            const int st = jp % NS;
            const Schur5Item it = items[jp];
            const int mis = (it.flags >> 2) & 1;
            const double* row = s_row(st) + mis;
            const double* sg = s_g(st) + (int)(((size_t)p.gB + (size_t)3 * it.pt0) & 1);
            const unsigned short* ptab = reinterpret_cast<const unsigned short*>(s_blob(st) + S5_HDR);
            double* pt = s_pt(st);
            for (int q = lane; q < it.npt; q += 32) {
                const int oe = ptab[q];
                const double* V = row + WB * oe + 9 * q;
                const double a[6] = {V[0] + lambda, V[1], V[2], V[4] + lambda, V[5], V[8] + lambda};
                double inv[6];
                inv_sym3(a, inv);
                double* ao = Ainv_out + (size_t)6 * (it.pt0 + q);
#pragma unroll
                for (int e = 0; e < 6; ++e) ao[e] = inv[e];
                if constexpr (!C::ZT) {      // classic point table: [Ainv row kk | g_kk]
                    double* o = pt + 16 * q;
                    o[0] = inv[0]; o[1] = inv[1]; o[2] = inv[2]; o[3] = sg[3 * q];
                    o[4] = inv[1]; o[5] = inv[3]; o[6] = inv[4]; o[7] = sg[3 * q + 1];
                    o[8] = inv[2]; o[9] = inv[4]; o[10] = inv[5]; o[11] = sg[3 * q + 2];
                    continue;
                }
                // A_p^-1 = L D L' (unit lower L, no pivoting: indefinite point blocks keep their signs in D)
                const double d0 = inv[0], l10 = inv[1] / d0, l20 = inv[2] / d0;
                const double d1 = fma(-l10, inv[1], inv[3]);
                const double l21 = fma(-l20, inv[1], inv[4]) / d1;
                const double d2 = fma(-l21 * l21, d1, fma(-l20, inv[2], inv[5]));
                const double g0 = sg[3 * q], g1 = sg[3 * q + 1], g2 = sg[3 * q + 2];
                double* o = pt + 16 * q;     // per inner index kk: [d_kk, (L' g)_kk, and the factors the W -> Z pass needs]
                o[0] = d0; o[1] = fma(l20, g2, fma(l10, g1, g0)); o[2] = l10; o[3] = l20;
                o[4] = d1; o[5] = fma(l21, g2, g1); o[6] = l21;
                o[8] = d2; o[9] = g2;
            }
            __syncwarp();
            __syncwarp();
            if constexpr (C::ZT) {   // W -> Z = L' W in place, one observation (3 x DC block, column-major) per lane:  z0 = w0 + l10 w1 + l20 w2,  z1 = w1 + l21 w2,  z2 = w2
                // (tried and measured slower: the owning consumer warp transforming its own points one entry ahead — whole-window mode —
                //  and the tile's observations dealt to the consumer warps one tile ahead behind an mbarrier)
                const unsigned char* opt = reinterpret_cast<const unsigned char*>(s_blob(st) + s_blob(st)[13]);
                double* wrow = s_row(st) + mis;
                if constexpr (C::PROD == 2) asm volatile("bar.sync 1, 64;" ::: "memory");   // the helper warp needs this tile's factors
                for (int j = lane; j < it.nob; j += 32 * C::PROD) {
                    const int q = opt[j];
                    const double* o = pt + 16 * q;
                    const double l10 = o[2], l20 = o[3], l21 = o[6];
                    double* w = wrow + WB * j + 9 * q;
#pragma unroll
                    for (int c = 0; c < DC; ++c) {
                        const double w0 = w[3 * c], w1 = w[3 * c + 1], w2 = w[3 * c + 2];
                        w[3 * c] = fma(l20, w2, fma(l10, w1, w0));
                        w[3 * c + 1] = fma(l21, w2, w1);
                    }
                }
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&bar_ready[st]);
            ++jp;
            t_pp += clock64() - t_p0;
        }
        if (sp.dbg && lane == 0) {
            long long* d = sp.dbg + ((size_t)blockIdx.x * 16 + warp) * 4;
            d[0] = t_pp; d[1] = clock64() - t_begin; d[2] = n_idle;
        }
        return;
    }

    if constexpr (C::PROD == 2) {
        if (warp == C::CONS + 1) {
            // ========================================= second producer warp: half of the W -> Z pass =========================================
            for (int jp = 0; jp < nitem; ++jp) {
                const int st = jp % NS;
                mbar_wait(&bar_full[st], (uint32_t)((jp / NS) & 1));              // the tile's bulk loads, seen by this warp itself
                const Schur5Item it = items[jp];
                const int mis = (it.flags >> 2) & 1;
                const double* pt = s_pt(st);
                asm volatile("bar.sync 1, 64;" ::: "memory");                      // the first producer warp has written the tile's factors
                const unsigned char* opt = reinterpret_cast<const unsigned char*>(s_blob(st) + s_blob(st)[13]);
                double* wrow = s_row(st) + mis;
                for (int j = lane + 32; j < it.nob; j += 64) {
                    const int q = opt[j];
                    const double* o = pt + 16 * q;
                    const double l10 = o[2], l20 = o[3], l21 = o[6];
                    double* w = wrow + WB * j + 9 * q;
#pragma unroll
                    for (int c = 0; c < DC; ++c) {
                        const double w0 = w[3 * c], w1 = w[3 * c + 1], w2 = w[3 * c + 2];
                        w[3 * c] = fma(l20, w2, fma(l10, w1, w0));
                        w[3 * c + 1] = fma(l21, w2, w1);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_ready[st]);
            }
            return;
        }
    }
    // ================================================= consumers =================================================
    if (warp >= sp.ncons) return;
    const int fr = lane >> 2, kk = lane & 3;
    double acc[BR][NTW][2];
    double racc[BR];
#pragma unroll
    for (int r = 0; r < BR; ++r) {
        racc[r] = 0.0;
#pragma unroll
        for (int n = 0; n < NTW; ++n) { acc[r][n][0] = 0.0; acc[r][n][1] = 0.0; }
    }
    const long long t_begin = clock64();
    long long t_wait = 0;
    for (int k = 0; k < nitem; ++k) {
        const int st = k % NS;
        const long long t_w0 = clock64();
        mbar_wait(&bar_ready[st], (uint32_t)((k / NS) & 1));
        t_wait += clock64() - t_w0;
        const uint32_t blobb = smem_u32(s_blob(st));
        const unsigned hdr = lds_u32(blobb + 4u * (unsigned)warp);
        const unsigned cnt = hdr & 0xffffu;
        if (cnt) {
            const unsigned eoff = lds_u32(blobb + 4u * 14u), mis = lds_u32(blobb + 4u * 15u);
            const uint32_t rowb = smem_u32(s_row(st)) + 8u * mis - 8u * (unsigned)C::BIAS;
            const uint32_t ptb = smem_u32(s_pt(st)) + 32u * (unsigned)kk;
            const uint32_t ep = blobb + 4u * eoff + 8u * (hdr >> 16);
            const unsigned band = (lds_u32(ep + 4u) >> 16) & 15u;   // a warp keeps its band for the whole super-tile, hence for the tile
            if constexpr (C::NBANDS == 1) {
                (void)band;
                Schur5Band<DC, 0>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk);
            } else if constexpr (C::NBANDS == 2) {
                if (band == 0) Schur5Band<DC, 0>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk);
                else Schur5Band<DC, 1>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk);
            } else if constexpr (C::NBANDS == 3) {
                switch (band) {
                    case 0: Schur5Band<DC, 0>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    case 1: Schur5Band<DC, 1>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    default: Schur5Band<DC, 2>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                }
            } else {
                switch (band) {
                    case 0: Schur5Band<DC, 0>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    case 1: Schur5Band<DC, 1>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    case 2: Schur5Band<DC, 2>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    case 3: Schur5Band<DC, 3>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    case 4: Schur5Band<DC, 4>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                    default: Schur5Band<DC, 5>::run(acc, racc, ep, cnt, rowb, ptb, p, sp.ftab, S, rhs, fr, kk); break;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);
    }
    if (sp.dbg && lane == 0) {
        long long* d = sp.dbg + ((size_t)blockIdx.x * 16 + warp) * 4;
        d[0] = t_wait; d[1] = clock64() - t_begin;
    }
}

// ---------------------------------------------------------------------------------------------------
// The points the plans leave out (gaps in the camera list, tracks wider than the window; from index first_irr on: irregular points
// — tracks longer than a tile, several costs on one camera — that belong to no tile at all, so their A_p^-1 is written here): one
// warp per point straight from global memory, one lane per block pair (i >= j), FP64 reductions into S.  A few thousand points at
// most.  Two costs of a point on the SAME camera (i != j, ci == cj) add W_i' Y_j + W_j' Y_i to the diagonal block.
// ---------------------------------------------------------------------------------------------------
template <int DC>
__global__ void __launch_bounds__(128) schur_outlier_kernel(DevProblem p, const int* __restrict__ pts, int npts, double* __restrict__ S,
                                                            double* __restrict__ rhs, double lambda, int first_irr = 0x7fffffff, double* __restrict__ Ainv = nullptr) {
    constexpr int WB = 3 * DC;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= npts) return;
    const int pt = pts[wid];
    const int ob0 = p.obs_start[pt], k = p.obs_start[pt + 1] - ob0;
    const double* hrow = p.H + (size_t)p.hB + (size_t)WB * ob0 + (size_t)9 * pt;
    const double* V = hrow + (size_t)WB * k;
    const double a6[6] = {V[0] + lambda, V[1], V[2], V[4] + lambda, V[5], V[8] + lambda};
    double inv[6];
    inv_sym3(a6, inv);
    if (wid >= first_irr && lane == 0) {
#pragma unroll
        for (int e = 0; e < 6; ++e) Ainv[(size_t)6 * pt + e] = inv[e];
    }
    const double Ai[3][3] = {{inv[0], inv[1], inv[2]}, {inv[1], inv[3], inv[4]}, {inv[2], inv[4], inv[5]}};
    const double* gp = p.g + p.gB + (size_t)3 * pt;
    const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
    const double t[3] = {Ai[0][0] * g0 + Ai[0][1] * g1 + Ai[0][2] * g2, Ai[1][0] * g0 + Ai[1][1] * g1 + Ai[1][2] * g2, Ai[2][0] * g0 + Ai[2][1] * g1 + Ai[2][2] * g2};
    for (int i = lane; i < k; i += 32) {   // rhs_c -= W_c' A^-1 g_p
        const double* Wi = hrow + (size_t)WB * i;
        const int ci = p.obs_cam[ob0 + i];
#pragma unroll
        for (int a = 0; a < DC; ++a) atomicAdd(rhs + (size_t)ci * DC + a, -(Wi[3 * a] * t[0] + Wi[3 * a + 1] * t[1] + Wi[3 * a + 2] * t[2]));
    }
    const int npair = k * (k + 1) / 2;
    for (int e = lane; e < npair; e += 32) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        const double* Wi = hrow + (size_t)WB * i;
        const double* Wj = hrow + (size_t)WB * j;
        const int ci = p.obs_cam[ob0 + i], cj = p.obs_cam[ob0 + j];   // ascending inside a point: ci >= cj
#pragma unroll
        for (int b = 0; b < DC; ++b) {
            const double w0 = Wj[3 * b], w1 = Wj[3 * b + 1], w2 = Wj[3 * b + 2];
            const double y0 = Ai[0][0] * w0 + Ai[0][1] * w1 + Ai[0][2] * w2, y1 = Ai[1][0] * w0 + Ai[1][1] * w1 + Ai[1][2] * w2,
                         y2 = Ai[2][0] * w0 + Ai[2][1] * w1 + Ai[2][2] * w2;
#pragma unroll
            for (int a = 0; a < DC; ++a) {
                if (i == j && b > a) continue;
                const double v = Wi[3 * a] * y0 + Wi[3 * a + 1] * y1 + Wi[3 * a + 2] * y2;
                if (i != j && ci == cj) {   // M + M' on the lower triangle of the diagonal block
                    atomicAdd(S + schur5_soff(ci, a > b ? a : b, cj, a > b ? b : a, p.tile_id, p.tile_pos, p.NT, DC, ST), a == b ? -2.0 * v : -v);
                } else
                atomicAdd(S + schur5_soff(ci, a, cj, b, p.tile_id, p.tile_pos, p.NT, DC, ST), -v);
            }
        }
    }
}

}  // namespace nlls
