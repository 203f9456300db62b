// Host-side plan of the Schur v5 kernel (pure C++: shared by nlls_b200.cu and by the CPU plan check in tests/native/).
//
// The kernel evaluates, for every point p with cameras c_1 < ... < c_k,   S_p = Y_p' W_p   with   W_p = [W_{p,c_1} ... W_{p,c_k}]
// (3 x DC k, contiguous in H) and Y_p = (V_p + lambda I)^-1 W_p, in 8 x 8 output tiles (one mma.sync.m8n8k4.f64 each), and
// accumulates these tiles over MANY points in registers before anything is added to the reduced camera matrix:
//   * a SUPER-TILE is a run of consecutive point tiles whose points all see cameras inside one WINDOW [base, base + WC) of
//     consecutive cameras; inside it every point lives in window coordinates (scalar row = DC (camera - base) + dof), so the
//     output tiles of all its points line up and can share accumulators;
//   * the window's NTW row tiles are dealt to NBANDS interleaved row classes ("bands": band b owns the row tiles b, b + NBANDS,
//     b + 2 NBANDS, ...); a consumer warp owns one band for the whole super-tile (accumulators: BR x NTW tiles, lower triangle in
//     use) and a share of the points.  Interleaving makes every point feed all bands about equally wherever it sits in the
//     window, so the bands stay balanced tile by tile (contiguous bands were not: the warps share the tile's stage, and whoever
//     is late on a tile holds it).  The host deals warps to bands in proportion to their work and points to the band's warps
//     greedily (least loaded first, within the tile first);
//   * at the end of the super-tile every warp adds its tiles to S (FLUSH entry).
// Points that do not fit (camera list not contiguous, or wider than the window) are left to the per-chunk fallback kernel.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

namespace nlls {

constexpr int S5_CONSUMERS = 11;       // most consumer warps a CTA can have (+ 1 producer warp = 12 warps: registers are allocated per 4 warps); Schur5Cfg<DC>::CONS are used
#ifndef S5_FULL6
#define S5_FULL6 0      // DC <= 6: 1 = one band that owns the whole window (a warp takes ALL row tiles of a point: one entry per point, 184 M instead of 231 M
                        // instructions), 7 consumer warps with 254 registers; 0 = three bands, 11 consumer warps.  Venice shape, same box: 0.61 vs 0.57 ms —
                        // 1.75 warps per scheduler hide less of the DMMA / FP64 latencies than 2.75
#endif
#ifndef S5_ZT6
#define S5_ZT6 0        // DC <= 6: 1 = consumers work on Z = L' W (A_p^-1 = L D L'; the producer warp transforms the staged W in place), 0 = Y = A_p^-1 W
                        // formed on the fly.  Venice shape, same box: Z made the 11 consumer warps 13 % faster (778 K instead of 900 K busy cycles) but
                        // the producer's pass (3 700 cycles per tile) became the critical path: 0.65 vs 0.57 ms.  For dc = 9 (7 consumer warps, two per
                        // point) Z wins: 1.75 -> 1.27 ms
#endif
#ifndef S5_PROD2
#define S5_PROD2 1      // (12-warp configurations with the Z formulation) two producer warps share the W -> Z pass
#endif
#ifndef S5_BR9
#define S5_BR9 6        // DC = 9: row tiles per band (2: six interleaved bands, 11 consumer warps; 6: two interleaved bands, 7 consumer warps with 255 registers)
#endif
constexpr int S5_HDR = 16;             // header words of a tile blob: [w] (first entry << 16) | count of consumer warp w, [13] word offset of the observation -> point bytes, [14] word offset of the entries, [15] span misalignment
constexpr unsigned S5_FLUSH = 1u << 20;
#ifndef S5_OBS6
#define S5_OBS6 232
#endif
#ifndef S5_CAP3
#define S5_CAP3 1.0     // relative capacity of consumer warps 3 and 7 (their sub-partition holds two consumers + the producer)
#endif
// cost model of one entry (cycles of the FP64 pipe, which is what the consumer warps of a sub-partition share): DMMAs, A fragments
// (3 FMAs + the rhs FMA + masks), B fragments, and a shape-independent part
struct Schur5Cost { double dmma = 17.0, afrag = 10.0, bfrag = 4.0, fixed = 40.0; };
inline Schur5Cost& schur5_cost() { static Schur5Cost c; return c; }
static_assert(S5_CONSUMERS <= 13, "header layout");

#ifdef __CUDACC__
#define S5_CE __host__ __device__
#else
#define S5_CE
#endif
template <int DC> struct Schur5Cfg {
    static constexpr int NTW = (DC <= 6) ? 9 : 12;          // row tiles of the window
    // Whole-window mode (DC <= 6, round 2c).  With three bands a point became three entries (one per band, each with its own decode,
    // A_p^-1 loads and B fragments): ncu counted 20 issued instructions per DMMA and the 2.75 warps a scheduler had could issue one
    // dependent instruction every ~5.5 cycles each — the DMMA pipe was 36 % busy.  One band = one entry per point: the B fragments are
    // loaded once, the row tiles' DMMAs interleave with the next rows' A fragments, ~7 instructions per DMMA.  The price is the whole
    // lower triangle of the window in registers (45 tiles = 180 registers), hence 7 consumer warps + 1 producer at 255 registers.
    static constexpr bool FULL = (DC <= 6) && (S5_FULL6 != 0);
    static constexpr int BR = FULL ? NTW : ((DC <= 6) ? 3 : S5_BR9);   // row tiles per band
    static constexpr int NBANDS = NTW / BR;
    static constexpr bool WIDE = FULL || BR * NTW > 27;     // more accumulators than 12 warps' 168 registers hold: 8 warps with 255 registers
    static constexpr bool ZT = (DC > 6) || (S5_ZT6 != 0);   // Z = L' W formulation (schur5.cuh)
    // producer warps: with 12 warps and the Z formulation the W -> Z pass of ONE producer warp is the critical path (3 700 cycles per tile);
    // a second producer warp takes half of the tile's observations and one consumer warp is given up for it
    static constexpr int PROD = (!WIDE && ZT && S5_PROD2 != 0) ? 2 : 1;
    static constexpr int CONS = WIDE ? 7 : S5_CONSUMERS + 1 - PROD;    // consumer warps of a CTA
    static constexpr int THREADS = 32 * (CONS + PROD);
    // r-th row tile of a band.  DC = 9: interleaved (band b owns b, b + NBANDS, ...).  DC = 6: {0,1,5} {2,3,6} {4,7,8} — the window's
    // middle rows carry most of the work (points start a few cameras above the base and span ~5 tiles); with 11 consumer warps the
    // bands get 4 / 4 / 3 warps, and this split brings the three shares close to 4 : 4 : 3 on the BAL-shaped problems (the plainly
    // interleaved split left the three-warp band 19 % over the mean on the Venice shape, this one 7 %).
    S5_CE static constexpr int row_tile(int band, int r) {
        if (FULL) return r;
        if (DC <= 6) { return band == 0 ? (r == 0 ? 0 : (r == 1 ? 1 : 5)) : (band == 1 ? (r == 0 ? 2 : (r == 1 ? 3 : 6)) : (r == 0 ? 4 : (r == 1 ? 7 : 8))); }
        if (DC > 6 && BR == 6) {   // two bands for 3 + 4 warps: {0,2,4,6,7,9} (34 of the 78 tiles) and {1,3,5,8,10,11} (44) — the plainly interleaved split
                                   // (36 : 42) left the three-warp band 14 % over the four-warp one (its warps never waited, the others 25 % of the time)
            return band == 0 ? (r < 4 ? 2 * r : (r == 4 ? 7 : 9)) : (r < 3 ? 2 * r + 1 : (r == 3 ? 8 : r + 6));
        }
        return band + NBANDS * r;
    }
    // Shapes of an entry of band `band`: (first window tile of the point TLO, number of active rows NACT); the band's rows >= TLO
    // are r0(TLO) .. BR - 1 and the first NACT of them are active.  Shapes are numbered densely (TLO ascending, NACT ascending):
    // the host stores the number in the entry, the kernel switches on it (a dense switch compiles to one indirect branch).
    S5_CE static constexpr int r0(int band, int tlo) {   // number of the band's rows above tile tlo
        int n = 0;
        for (int r = 0; r < BR; ++r) if (row_tile(band, r) < tlo) ++n;
        return n;
    }
    S5_CE static constexpr int nshapes(int band) {
        int n = 0;
        for (int tlo = 0; tlo <= row_tile(band, BR - 1); ++tlo) n += BR - r0(band, tlo);
        return n;
    }
    S5_CE static constexpr int shape_id(int band, int tlo, int nact) {
        int n = 0;
        for (int t = 0; t < tlo; ++t) n += BR - r0(band, t);
        return n + nact - 1;
    }
    S5_CE static constexpr int shape_tlo(int band, int id) {
        int n = 0;
        for (int tlo = 0; tlo <= row_tile(band, BR - 1); ++tlo) { n += BR - r0(band, tlo); if (id < n) return tlo; }
        return -1;
    }
    S5_CE static constexpr int shape_nact(int band, int id) { return id - shape_id(band, shape_tlo(band, id), 1) + 1; }
    static constexpr int BIAS = 3 * 8 * NTW;                  // entry field wofs = (W_p offset in the span) - 3 (DC delta) + BIAS  >= 0
    static constexpr int WROWS = 8 * NTW;                   // scalar rows of the window
    static constexpr int WC = WROWS / DC;                   // cameras of the window
    static constexpr int OBS = (DC <= 6) ? S5_OBS6 : 128;   // tile capacity (observations / points)
    static constexpr int PTS = OBS / 2;
    static constexpr int WB = 3 * DC;
    static constexpr int ENT_CAP = PTS * NBANDS + S5_CONSUMERS + 5;   // entries per tile
    static_assert(NTW % BR == 0, "bands of equal height");
};

struct Schur5Item {       // one point tile of a CTA's range (32 bytes)
    int pt0, npt, ob0, nob;
    unsigned blob0;       // first u32 of the tile's blob
    unsigned nblob;       // u32 words of the blob (multiple of 4)
    int flags;            // bit2: the H span starts 8 bytes off a 16-byte boundary
    int pad;
};

struct Schur5Plan {
    std::vector<int> cta_item;            // [ncta + 1]
    std::vector<Schur5Item> items;
    std::vector<unsigned> blob;           // per tile: [S5_HDR header][point table: obs_end (u16) per point, padded to an even word count][local point (u8) per observation, padded to an even word count][entries: 2 words each]
    std::vector<int> outliers;            // points left to the fallback kernel
    std::vector<int> super_base;          // window base camera of every super-tile (a FLUSH entry carries the super-tile's index)
    // statistics
    long long n_entries = 0, n_dmma = 0, n_flush = 0, n_super = 0, n_frag_a = 0, n_frag_b = 0;
    double out_frac = 0.0;                // share of the block contributions that belong to outlier points
    double imbalance = 1.0;               // sum over super-tiles of (max warp load x warps) / sum of loads
};

// Element (ar, ac) of S block (camera cr, camera cc), cr >= cc, in the tile-sparse storage of reduced.cuh.
// ST: tile edge (72), TC: cameras per tile; tile_id / pos as uploaded for the reduced solve.
#ifdef __CUDACC__
#define S5_HD __host__ __device__ __forceinline__
#else
#define S5_HD inline
#endif
S5_HD long long schur5_soff(int cr, int ar, int cc, int ac, const int* tile_id, const int* pos, int NT, int DC, int ST) {
    const int TC = ST / DC;
    const int I = cr / TC, J = cc / TC;
    const int r0 = (cr - I * TC) * DC + ar, c0 = (cc - J * TC) * DC + ac;
    const int pI = pos[I], pJ = pos[J];
    if (pI >= pJ) return (long long)tile_id[(size_t)pI * NT + pJ] * ST * ST + r0 + (long long)ST * c0;
    return (long long)tile_id[(size_t)pJ * NT + pI] * ST * ST + c0 + (long long)ST * r0;
}

// FLUSH table: per super-tile [base camera, I0 = base / TC, then for every pair (a >= b) of the NTS consecutive S tiles the window can
// touch: 2 * (element offset of the S tile (I0 + a, I0 + b)) + (1 if it is stored transposed), or -1 if the tile does not exist].
template <int DC> struct Schur5Flush {
    static constexpr int ST = 72, TC = ST / DC;
    static constexpr int NTS = (Schur5Cfg<DC>::WC - 1 + TC - 1) / TC + 1;   // S tiles a window of WC cameras can span
    static constexpr int NPAIR = NTS * (NTS + 1) / 2;
    static constexpr int STRIDE = 2 + NPAIR;
    S5_CE static constexpr int pair(int a, int b) { return a * (a + 1) / 2 + b; }
};
template <int DC>
std::vector<long long> schur5_flush_table(const std::vector<int>& super_base, const std::vector<int>& tile_id, const std::vector<int>& pos, int NT) {
    using F = Schur5Flush<DC>;
    std::vector<long long> t(super_base.size() * (size_t)F::STRIDE, -1);
    for (size_t s = 0; s < super_base.size(); ++s) {
        long long* e = &t[s * (size_t)F::STRIDE];
        const int I0 = super_base[s] / F::TC;
        e[0] = super_base[s]; e[1] = I0;
        for (int a = 0; a < F::NTS; ++a) for (int b = 0; b <= a; ++b) {
            const int I = I0 + a, J = I0 + b;
            if (I >= NT) continue;
            const int pI = pos[(size_t)I], pJ = pos[(size_t)J];
            const int id = pI >= pJ ? tile_id[(size_t)pI * NT + pJ] : tile_id[(size_t)pJ * NT + pI];
            if (id < 0) continue;
            e[2 + F::pair(a, b)] = 2 * ((long long)id * F::ST * F::ST) + (pI >= pJ ? 0 : 1);
        }
    }
    return t;
}

// tile range [t_lo, t_hi] of a point with `k` cameras starting `delta` cameras above the window base
inline void schur5_tile_range(int DC, int delta, int k, int& t_lo, int& t_hi) {
    t_lo = (DC * delta) / 8;
    t_hi = (DC * (delta + k) - 1) / 8;
}

// Builds the plan.  obs_start[nB + 1] / obs_cam[nobs]: point-major observations (cameras ascending inside a point);
// hB: offset of the point rows in H (DC*DC*nA);  ncta: CTAs (one per SM);  maxrun: tiles per super-tile at most.
// irr (optional): points that must stay outside every tile (they are appended to the outliers by the caller).
template <int DC>
Schur5Plan schur5_build_plan(const std::vector<int>& obs_start, const std::vector<int>& obs_cam, long long nA, int ncta, int maxrun = 48, int ncons = Schur5Cfg<DC>::CONS,
                             const std::vector<unsigned char>* irr = nullptr) {
    using C = Schur5Cfg<DC>;
    Schur5Plan P;
    const long long nB = (long long)obs_start.size() - 1;
    const long long hB = (long long)DC * DC * nA;
    // ---- point tiles: consecutive points, <= OBS observations, <= PTS points, boundaries preferably 16-byte aligned in H
    std::vector<int> tile_pt;     // (first point, one past the last point) per tile
    {
        auto aligned = [&](long long pt) { return ((hB + (long long)C::WB * obs_start[(size_t)pt] + 9 * pt) & 1) == 0; };
        auto skip = [&](long long pt) { return irr != nullptr && (*irr)[(size_t)pt] != 0; };
        long long p0 = 0;
        while (p0 < nB) {
            if (skip(p0)) { ++p0; continue; }
            long long p1 = p0;
            while (p1 < nB && !skip(p1) && (p1 - p0) < C::PTS && (obs_start[(size_t)p1 + 1] - obs_start[(size_t)p0]) <= C::OBS) ++p1;
            if (p1 < nB && !aligned(p1) && p1 - 1 > p0 && aligned(p1 - 1)) --p1;
            if (p1 == p0) { P.cta_item.clear(); return P; }   // a point with more observations than a tile holds: no v5 plan
            tile_pt.push_back((int)p0); tile_pt.push_back((int)p1);
            p0 = p1;
        }
    }
    const int nt = (int)tile_pt.size() / 2;
    // ---- per point: start camera, track length, eligibility (contiguous camera list, not wider than the window minus slack)
    std::vector<int> pstart((size_t)nB), pk((size_t)nB);
    std::vector<unsigned char> elig((size_t)nB);
    const int kmax_fit = C::WC;
    ncons = std::max(C::NBANDS, std::min(ncons, C::CONS));
    double contrib_all = 0.0, contrib_out = 0.0;
    for (long long p = 0; p < nB; ++p) {
        const int b = obs_start[(size_t)p], e = obs_start[(size_t)p + 1], k = e - b;
        pk[(size_t)p] = k;
        pstart[(size_t)p] = k > 0 ? obs_cam[(size_t)b] : 0;
        elig[(size_t)p] = (k > 0 && k <= kmax_fit && obs_cam[(size_t)e - 1] - obs_cam[(size_t)b] + 1 == k && !(irr != nullptr && (*irr)[(size_t)p])) ? 1 : 0;
        contrib_all += 0.5 * k * (k + 1);
    }
    // ---- contiguous tile ranges of similar weight, one per CTA
    ncta = std::max(1, std::min(ncta, nt));
    std::vector<double> wsum((size_t)nt + 1, 0.0);
    {   // the consumers' own cost model (window offset unknown yet: an average misalignment of half a tile is assumed)
        const Schur5Cost& cm = schur5_cost();
        std::vector<double> kcost(258, 0.0);
        for (int k = 1; k < 258; ++k) {
            const double tiles = (DC * k + 7.0) / 8.0 + 0.5;                  // row tiles the point spans
            const double dm = 0.5 * tiles * (tiles + 1.0), nb = std::min<double>(C::NBANDS, tiles);
            kcost[(size_t)k] = cm.dmma * dm + cm.afrag * tiles + cm.bfrag * (dm / std::max(1.0, tiles)) * nb + cm.fixed * nb;
        }
        for (int t = 0; t < nt; ++t) {
            double w = 300.0;
            for (int p = tile_pt[(size_t)2 * t]; p < tile_pt[(size_t)2 * t + 1]; ++p) w += kcost[(size_t)std::min(pk[(size_t)p], 257)];
            wsum[(size_t)t + 1] = wsum[(size_t)t] + w;
        }
    }
    std::vector<int> cta_tile((size_t)ncta + 1, nt);
    cta_tile[0] = 0;
    for (int c = 1; c < ncta; ++c) cta_tile[(size_t)c] = (int)(std::lower_bound(wsum.begin(), wsum.end(), wsum[(size_t)nt] * c / ncta) - wsum.begin());
    for (int c = 1; c <= ncta; ++c) cta_tile[(size_t)c] = std::max(cta_tile[(size_t)c], cta_tile[(size_t)c - 1]);
    cta_tile[(size_t)ncta] = nt;

    P.cta_item.assign((size_t)ncta + 1, 0);
    P.items.reserve((size_t)nt);
    std::vector<std::vector<unsigned>> went(S5_CONSUMERS);   // per consumer warp: entries (2 words each) of the tile being built
    double imb_num = 0.0, imb_den = 0.0;
    struct TileOut { std::vector<unsigned> ent[S5_CONSUMERS]; };
    for (int c = 0; c < ncta; ++c) {
        const int ta = cta_tile[(size_t)c], tb = cta_tile[(size_t)c + 1];
        int t = ta;
        while (t < tb) {
            // ---- super-tile [t, u): greedy extension while every eligible point fits one window
            int base = INT32_MAX, hi = INT32_MIN;
            int u = t;
            while (u < tb && u - t < maxrun) {
                int b2 = base, h2 = hi;
                for (int p = tile_pt[(size_t)2 * u]; p < tile_pt[(size_t)2 * u + 1]; ++p) if (elig[(size_t)p]) {
                    b2 = std::min(b2, pstart[(size_t)p]); h2 = std::max(h2, pstart[(size_t)p] + pk[(size_t)p]);
                }
                if (u > t && h2 > INT32_MIN && (long long)(h2 - b2) * DC > C::WROWS) break;
                base = b2; hi = h2; ++u;
            }
            if (base == INT32_MAX) base = 0;
            // (a single tile whose own eligible points span more than the window: the late ones become outliers)
            ++P.n_super;
            P.super_base.push_back(base);
            // ---- work per band, warps per band
            double work[C::NBANDS] = {0};
            auto unit_cost = [&](int delta, int k, int band, long long* dm, int* na, int* nb) {
                int lo, hi2; schur5_tile_range(DC, delta, k, lo, hi2);
                long long d = 0;
                int a = 0, last = -1;
                for (int r = 0; r < C::BR; ++r) { const int mt = C::row_tile(band, r); if (mt >= lo && mt <= hi2) { d += mt - lo + 1; ++a; last = mt; } }
                if (a == 0) return 0.0;
                const int b = last - lo + 1;
                if (dm) *dm = d; if (na) *na = a; if (nb) *nb = b;
                const Schur5Cost& cm = schur5_cost();
                return cm.dmma * d + cm.afrag * a + cm.bfrag * b + cm.fixed;
            };
            auto fits = [&](int p) { return elig[(size_t)p] && pstart[(size_t)p] >= base && (long long)(pstart[(size_t)p] - base + pk[(size_t)p]) * DC <= C::WROWS; };
            for (int p = tile_pt[(size_t)2 * t]; p < tile_pt[(size_t)2 * (u - 1) + 1]; ++p) if (fits(p))
                for (int b = 0; b < C::NBANDS; ++b) work[b] += unit_cost(pstart[(size_t)p] - base, pk[(size_t)p], b, nullptr, nullptr, nullptr);
            // Warps of a band: counts in proportion to the work, heaviest band first.
            auto capw = [](int w) { return (w & 3) == 3 ? S5_CAP3 : 1.0; };
            std::vector<int> bw[C::NBANDS];
            {
                int nw[C::NBANDS] = {0}, used = 0;
                for (int b = 0; b < C::NBANDS; ++b) if (work[b] > 0) { nw[b] = 1; ++used; }
                while (used < ncons) {
                    int best = -1;
                    for (int b = 0; b < C::NBANDS; ++b) if (nw[b] > 0 && (best < 0 || work[b] / nw[b] > work[best] / nw[best])) best = b;
                    if (best < 0) break;
                    ++nw[best]; ++used;
                }
                int border[C::NBANDS];
                for (int b = 0; b < C::NBANDS; ++b) border[b] = b;
                std::stable_sort(border, border + C::NBANDS, [&](int x, int y) { return work[x] > work[y]; });
                // warp w runs on SM sub-partition w % 4: the warps of one band share a sub-partition where they can (they run the same
                // straight-line shape code: the instruction cache was the second largest stall with the bands mixed)
                static const int worder[11] = {0, 4, 8, 3, 1, 5, 9, 7, 2, 6, 10};
                static_assert(S5_CONSUMERS == 11, "warp order table");
                int next = 0;
                for (int i = 0; i < C::NBANDS; ++i) for (int q = 0; q < nw[border[i]]; ++q) { while (worder[next] >= ncons) ++next; bw[border[i]].push_back(worder[next++]); }
            }
            double load[S5_CONSUMERS] = {0};
            bool touched[S5_CONSUMERS] = {false};
            // ---- tiles of the super-tile
            for (int tt = t; tt < u; ++tt) {
                const int pt0 = tile_pt[(size_t)2 * tt], pt1 = tile_pt[(size_t)2 * tt + 1];
                const int ob0 = obs_start[(size_t)pt0];
                for (auto& v : went) v.clear();
                double tload[S5_CONSUMERS] = {0};
                for (int p = pt0; p < pt1; ++p) {
                    if (!fits(p)) { if (pk[(size_t)p] > 0) { P.outliers.push_back(p); contrib_out += 0.5 * pk[(size_t)p] * (pk[(size_t)p] + 1); } continue; }
                    const int delta = pstart[(size_t)p] - base, k = pk[(size_t)p];
                    const unsigned wrel = (unsigned)(C::WB * (obs_start[(size_t)p] - ob0) + 9 * (p - pt0));
                    for (int b = 0; b < C::NBANDS; ++b) {
                        long long dm = 0; int na = 0, nb = 0;
                        const double cst = unit_cost(delta, k, b, &dm, &na, &nb);
                        if (cst == 0.0) continue;
                        // least loaded warp of the band, first within this tile (the warps share the tile's stage: whoever is late
                        // holds it), then over the super-tile
                        int w = bw[b][0];
                        for (int q : bw[b]) {
                            const double lq = tload[q] / capw(q), lw = tload[w] / capw(w);
                            if (lq < lw || (lq == lw && load[q] / capw(q) < load[w] / capw(w))) w = q;
                        }
                        tload[w] += cst;
                        load[w] += cst; touched[w] = true;
                        int lo, hi2; schur5_tile_range(DC, delta, k, lo, hi2);
                        const unsigned sid = (unsigned)C::shape_id(b, lo, na);
                        went[(size_t)w].push_back((wrel + (unsigned)C::BIAS - 3u * (unsigned)(DC * delta)) | ((unsigned)(DC * k) << 16) | ((unsigned)(DC * delta) << 24));
                        went[(size_t)w].push_back((unsigned)(p - pt0) | (sid << 8) | ((unsigned)b << 16));
                        P.n_entries++; P.n_dmma += dm; P.n_frag_a += na; P.n_frag_b += nb;
                    }
                }
                // entries of one shape next to each other: the kernel picks the shape's code once per run of equal shapes
                for (auto& v : went) {
                    std::vector<std::pair<unsigned, unsigned>> tmp(v.size() / 2);
                    for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = {v[2 * i], v[2 * i + 1]};
                    std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<unsigned, unsigned>& a, const std::pair<unsigned, unsigned>& b2) { return ((a.second >> 8) & 255u) < ((b2.second >> 8) & 255u); });
                    for (size_t i = 0; i < tmp.size(); ++i) { v[2 * i] = tmp[i].first; v[2 * i + 1] = tmp[i].second; }
                }
                if (tt == u - 1)
                    for (int b = 0; b < C::NBANDS; ++b) for (int w : bw[b]) if (touched[w]) {
                        went[(size_t)w].push_back((unsigned)(P.super_base.size() - 1));
                        went[(size_t)w].push_back(((unsigned)b << 16) | S5_FLUSH);
                        P.n_flush++;
                    }
                // ---- the tile's blob
                Schur5Item it;
                it.pt0 = pt0; it.npt = pt1 - pt0; it.ob0 = ob0; it.nob = obs_start[(size_t)pt1] - ob0;
                it.flags = ((hB + (long long)C::WB * ob0 + 9ll * pt0) & 1) ? 4 : 0;
                it.pad = 0;
                it.blob0 = (unsigned)P.blob.size();
                const size_t h = P.blob.size();
                P.blob.resize(h + S5_HDR, 0u);
                for (int p = pt0; p < pt1; p += 2) {
                    const unsigned lo = (unsigned)(obs_start[(size_t)p + 1] - ob0);
                    const unsigned hi2 = (p + 1 < pt1) ? (unsigned)(obs_start[(size_t)p + 2] - ob0) : 0u;
                    P.blob.push_back(lo | (hi2 << 16));
                }
                if ((P.blob.size() - h) & 1) P.blob.push_back(0u);
                P.blob[h + 13] = (unsigned)(P.blob.size() - h);   // word offset of the observation -> local point bytes (the producer's W -> Z pass goes by observation)
                {
                    unsigned word = 0; int nb = 0;
                    for (int p = pt0; p < pt1; ++p)
                        for (int j = obs_start[(size_t)p]; j < obs_start[(size_t)p + 1]; ++j) {
                            word |= (unsigned)(p - pt0) << (8 * nb);
                            if (++nb == 4) { P.blob.push_back(word); word = 0; nb = 0; }
                        }
                    if (nb) P.blob.push_back(word);
                }
                if ((P.blob.size() - h) & 1) P.blob.push_back(0u);
                P.blob[h + 14] = (unsigned)(P.blob.size() - h);   // word offset of the entries
                P.blob[h + 15] = (unsigned)((it.flags >> 2) & 1);  // span misalignment (doubles)
                unsigned first = 0;
                for (int w = 0; w < S5_CONSUMERS; ++w) {
                    const unsigned cnt = (unsigned)(went[(size_t)w].size() / 2);
                    P.blob[h + (size_t)w] = (first << 16) | cnt;
                    P.blob.insert(P.blob.end(), went[(size_t)w].begin(), went[(size_t)w].end());
                    first += cnt;
                }
                while ((P.blob.size() - h) & 3) P.blob.push_back(0u);
                it.nblob = (unsigned)(P.blob.size() - h);
                P.items.push_back(it);
            }
            double mx = 0, sum = 0; int nact = 0;
            double capsum = 0;
            for (int w = 0; w < ncons; ++w) { mx = std::max(mx, load[w] / capw(w)); sum += load[w]; capsum += capw(w); nact += 1; }
            imb_num += mx * capsum; imb_den += sum;
            t = u;
        }
        P.cta_item[(size_t)c + 1] = (int)P.items.size();
    }
    P.out_frac = contrib_all > 0 ? contrib_out / contrib_all : 0.0;
    P.imbalance = imb_den > 0 ? imb_num / imb_den : 1.0;
    return P;
}

}  // namespace nlls
