// CPU check of the reduced-system planning (nllssolver.jl_b200/csrc/reduced_plan.hpp): elimination orders, tile-level symbolic
// factorisation against a dense boolean elimination, levels, and the storage order by rank ownership.  Test infrastructure only.
#include <cmath>
#include <cstring>
#include <random>
#include <set>

#include "../../nllssolver.jl_b200/csrc/reduced_plan.hpp"

using namespace nlls;

// Banded tile pattern of NT columns with half-bandwidth 1 plus `nlong` random edges (I, I - 2) — the shape a BAL-like problem with a
// handful of long tracks produces.  mode: 0 = graph order, 1 = band order.
// out[0] = levels, out[1] = stored tiles, out[2] = max rows per column, out[3] = 1 if the symbolic pattern equals the dense boolean
// elimination of the permuted pattern, out[4] = 1 if every column waits for all the columns it depends on (levels strictly increase).
extern "C" int reduced_plan_check(int NT, int nlong, unsigned seed, int mode, double* out) {
    std::mt19937 rng(seed);
    std::vector<unsigned char> natpat((size_t)NT * NT, 0);
    for (int I = 0; I < NT; ++I) { natpat[(size_t)I * NT + I] = 1; if (I > 0) natpat[(size_t)I * NT + I - 1] = 1; }
    for (int e = 0; e < nlong && NT > 2; ++e) { const int I = 2 + (int)(rng() % (unsigned)(NT - 2)); natpat[(size_t)I * NT + I - 2] = 1; }
    const int w = red_half_bandwidth(natpat, NT);
    const std::vector<int> order = mode == 0 ? red_order_graph(natpat, NT) : red_order_band(NT, std::max(w, 1));
    std::vector<int> seen((size_t)NT, 0);
    if ((int)order.size() != NT) return -1;
    for (int v : order) { if (v < 0 || v >= NT || seen[(size_t)v]) return -2; seen[(size_t)v] = 1; }
    const RedSymbolic s = red_symbolic(natpat, NT, order);
    // dense boolean elimination of the permuted symmetric pattern
    std::vector<unsigned char> D((size_t)NT * NT, 0);
    for (int I = 0; I < NT; ++I) for (int J = 0; J <= I; ++J) if (natpat[(size_t)I * NT + J]) { const int a = s.pos[(size_t)I], b = s.pos[(size_t)J]; D[(size_t)a * NT + b] = D[(size_t)b * NT + a] = 1; }
    for (int k = 0; k < NT; ++k) for (int i = k + 1; i < NT; ++i) if (D[(size_t)i * NT + k]) for (int j = k + 1; j < NT; ++j) if (D[(size_t)j * NT + k]) D[(size_t)i * NT + j] = 1;
    int same = 1, mono = 1, maxrows = 0;
    for (int I = 0; I < NT; ++I) for (int J = 0; J <= I; ++J) if ((s.pat[(size_t)I * NT + J] != 0) != (D[(size_t)I * NT + J] != 0)) same = 0;
    for (int J = 0; J < NT; ++J) { maxrows = std::max(maxrows, (int)s.rows[(size_t)J].size()); for (int I : s.rows[(size_t)J]) if (s.level[(size_t)I] <= s.level[(size_t)J]) mono = 0; }
    out[0] = s.nlev; out[1] = s.ntiles; out[2] = maxrows; out[3] = same; out[4] = mono;
    return 0;
}

// Ownership order: NT tile columns, `nranks` ranks owning contiguous column bands (rank r's points touch the tiles (I, J) with both
// columns in its band, |I - J| <= 1; neighbouring bands overlap by one column).  Checks: ids are a permutation of 0 .. slots-1 minus the
// padding, exclusive tiles of rank r lie in block r, shared tiles behind the blocks, fill-only tiles last, every diagonal tile has
// exactly one rank adding U_c.  out[0] = block, out[1] = shared, out[2] = fill-only, out[3] = slots.
extern "C" int reduced_owner_check(int NT, int nranks, double* out) {
    std::vector<unsigned char> natpat((size_t)NT * NT, 0);
    std::vector<int> toucher((size_t)NT * NT, -1);
    for (int r = 0; r < nranks; ++r) {
        const int lo = (int)((long long)NT * r / nranks), hi = std::min(NT, (int)((long long)NT * (r + 1) / nranks) + 1);   // one column of overlap
        for (int I = lo; I < hi; ++I) for (int J = std::max(lo, I - 1); J <= I; ++J) {
            natpat[(size_t)I * NT + J] = 1;
            int& t = toucher[(size_t)I * NT + J];
            t = (t == -1) ? r : (t == r ? r : -2);
        }
    }
    for (int I = 0; I < NT; ++I) natpat[(size_t)I * NT + I] = 1;
    const std::vector<int> order = red_order_graph(natpat, NT);
    const RedSymbolic s = red_symbolic(natpat, NT, order);
    std::vector<int> tile_id((size_t)NT * NT, -1);
    int nt = 0;
    for (int J = 0; J < NT; ++J) for (int I = J; I < NT; ++I) if (s.pat[(size_t)I * NT + J]) tile_id[(size_t)I * NT + J] = nt++;
    std::vector<int> adders((size_t)NT, 0);
    RedOwnership own;
    std::vector<int> ids;
    for (int rank = 0; rank < nranks; ++rank) {
        std::vector<int> t2 = tile_id;
        own = red_order_by_owner(t2, NT, order, toucher, nranks, rank);
        if (rank == 0) ids = t2; else if (ids != t2) return -1;                     // every rank derives the same storage order
        for (int I = 0; I < NT; ++I) adders[(size_t)I] += own.add_u[(size_t)I];
    }
    for (int I = 0; I < NT; ++I) if (adders[(size_t)I] != 1) return -2;              // exactly one rank adds U_c to every diagonal tile
    std::set<int> used;
    for (int pJ = 0; pJ < NT; ++pJ) for (int pI = pJ; pI < NT; ++pI) {
        const int id = ids[(size_t)pI * NT + pJ];
        if ((id >= 0) != (tile_id[(size_t)pI * NT + pJ] >= 0)) return -3;
        if (id < 0) continue;
        if (id >= own.nslots || !used.insert(id).second) return -4;                   // in range, no two tiles in one slot
        const int I = order[(size_t)pI], J = order[(size_t)pJ];
        int t = toucher[(size_t)std::max(I, J) * NT + std::min(I, J)];
        if (pI == pJ && t == -1) t = 0;
        if (t >= 0) { if (id < t * own.block || id >= (t + 1) * own.block) return -5; }
        else if (t == -2) { if (id < own.shared0 || id >= own.shared0 + own.nshared) return -6; }
        else if (id < own.shared0 + own.nshared) return -7;
    }
    out[0] = (double)own.block; out[1] = (double)own.nshared; out[2] = (double)own.nfill; out[3] = (double)own.nslots;
    return 0;
}
