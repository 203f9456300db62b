// Host-side planning of the reduced camera system (pure C++: shared by nlls_b200.cu and by the CPU check in tests/native/):
// elimination order of the 72 x 72 tile columns, tile-level symbolic factorisation, elimination levels, and the storage order of
// the tiles by rank ownership for the multi-rank exchange.
#pragma once
#include <algorithm>
#include <vector>

namespace nlls {

// Half-bandwidth (in tiles) of a lower-triangular tile pattern natpat[I * NT + J], I >= J.
inline int red_half_bandwidth(const std::vector<unsigned char>& natpat, int NT) {
    int w = 0;
    for (int I = 0; I < NT; ++I) for (int J = 0; J < I; ++J) if (natpat[(size_t)I * NT + J]) w = std::max(w, I - J);
    return w;
}

// Round-1/2a order: separators as wide as the half-bandwidth w, leaves of up to 2 w + w columns eliminated one after the other.
inline std::vector<int> red_order_band(int NT, int w) {
    std::vector<int> out;
    const int leaf = std::max(2 * w, 4);
    struct Rec { static void nd(int lo, int hi, int w, int leaf, std::vector<int>& out) {
        if (hi - lo <= leaf + w) { for (int i = lo; i < hi; ++i) out.push_back(i); return; }
        const int s0 = lo + (hi - lo - w) / 2;
        nd(lo, s0, w, leaf, out); nd(s0 + w, hi, w, leaf, out);
        for (int i = s0; i < s0 + w; ++i) out.push_back(i);
    } };
    Rec::nd(0, NT, w, leaf, out);
    return out;
}

// Nested dissection on the ACTUAL tile graph, down to single columns: the separator of a node list is its middle column plus, for
// every edge that still joins the two sides, the endpoint nearer to the middle.  The reduced solve is a chain of dependent levels
// (~27 us each: diagonal tile, off-diagonal tiles, updates), so what counts is the height of the elimination tree, not the fill:
// Venice shape 12 -> 8 levels with 437 instead of 443 tiles (a handful of long tracks no longer widen every separator, and the leaves
// are no longer eliminated sequentially).  Returns nat_of_pos: the natural tile column at every elimination position.
inline std::vector<int> red_order_graph(const std::vector<unsigned char>& natpat, int NT) {
    std::vector<int> out;
    struct Rec { static void nd(const std::vector<int>& nodes, const std::vector<unsigned char>& pat, int NT, std::vector<int>& out) {
        const int n = (int)nodes.size();
        if (n <= 2) { for (int v : nodes) out.push_back(v); return; }
        const int m = n / 2;
        std::vector<unsigned char> insep((size_t)n, 0);
        insep[(size_t)m] = 1;
        for (int a = m - 1; a >= 0; --a)
            for (int b = m + 1; b < n && !insep[(size_t)a]; ++b)
                if (!insep[(size_t)b] && pat[(size_t)nodes[(size_t)b] * NT + nodes[(size_t)a]]) {   // nodes ascend: (b, a) is in the lower triangle
                    if (m - a <= b - m) insep[(size_t)a] = 1; else insep[(size_t)b] = 1;
                }
        std::vector<int> left, right, sep;
        for (int i = 0; i < n; ++i) (insep[(size_t)i] ? sep : (i < m ? left : right)).push_back(nodes[(size_t)i]);
        nd(left, pat, NT, out); nd(right, pat, NT, out);
        for (int v : sep) out.push_back(v);
    } };
    std::vector<int> all((size_t)NT);
    for (int i = 0; i < NT; ++i) all[(size_t)i] = i;
    Rec::nd(all, natpat, NT, out);
    return out;
}

struct RedSymbolic {
    std::vector<int> pos;                     // elimination position of every natural tile column
    std::vector<unsigned char> pat;           // permuted numbering, lower triangle, after fill
    std::vector<std::vector<int>> rows;       // rows[J]: the rows I > J of column J (permuted numbering, ascending)
    std::vector<int> level;                   // elimination level of every column (permuted numbering)
    int nlev = 0, ntiles = 0;
};
// Tile-level symbolic factorisation for the order nat_of_pos: eliminating column J couples every pair of its rows; column I waits
// for every column J < I that has I among its rows.
inline RedSymbolic red_symbolic(const std::vector<unsigned char>& natpat, int NT, const std::vector<int>& nat_of_pos) {
    RedSymbolic s;
    s.pos.assign((size_t)NT, 0);
    for (int q = 0; q < NT; ++q) s.pos[(size_t)nat_of_pos[(size_t)q]] = q;
    s.pat.assign((size_t)NT * NT, 0);
    for (int I = 0; I < NT; ++I) for (int J = 0; J <= I; ++J) if (natpat[(size_t)I * NT + J]) {
        const int a = std::max(s.pos[(size_t)I], s.pos[(size_t)J]), b = std::min(s.pos[(size_t)I], s.pos[(size_t)J]);
        s.pat[(size_t)a * NT + b] = 1;
    }
    s.rows.assign((size_t)NT, {});
    for (int J = 0; J < NT; ++J) {
        std::vector<int>& r = s.rows[(size_t)J];
        for (int I = J + 1; I < NT; ++I) if (s.pat[(size_t)I * NT + J]) r.push_back(I);
        for (size_t a = 0; a < r.size(); ++a) for (size_t b = 0; b <= a; ++b) s.pat[(size_t)r[a] * NT + r[b]] = 1;
    }
    s.level.assign((size_t)NT, 0);
    for (int J = 0; J < NT; ++J) {
        for (int I : s.rows[(size_t)J]) s.level[(size_t)I] = std::max(s.level[(size_t)I], s.level[(size_t)J] + 1);
        s.nlev = std::max(s.nlev, s.level[(size_t)J] + 1);
    }
    for (int J = 0; J < NT; ++J) for (int I = J; I < NT; ++I) if (s.pat[(size_t)I * NT + J]) ++s.ntiles;
    return s;
}

// Storage order of the tiles by ownership (multi-rank exchange of the reduced system).  toucher[max(I,J) * NT + min(I,J)] (natural
// numbering): -1 nobody's points touch the tile, r >= 0 only rank r's do, -2 several ranks'.  tile_id (permuted numbering, -1 = not
// stored) is renumbered in place to [rank 0's exclusive tiles | rank 1's | ... (blocks of `block` tiles) | shared | fill-only];
// add_u[I] (natural camera tile) = this rank adds U_c + lambda I to the diagonal tile.  Returns the number of slots.
struct RedOwnership { long long block = 0, shared0 = 0, nshared = 0, nslots = 0; std::vector<int> add_u; std::vector<long long> nexcl; long long nfill = 0; };
inline RedOwnership red_order_by_owner(std::vector<int>& tile_id, int NT, const std::vector<int>& nat_of_pos, const std::vector<int>& toucher, int nranks, int rank) {
    RedOwnership o;
    o.add_u.assign((size_t)NT, 0);
    std::vector<std::vector<int>> excl((size_t)nranks);
    std::vector<int> shared, fill;
    int nt = 0;
    for (int pJ = 0; pJ < NT; ++pJ) for (int pI = pJ; pI < NT; ++pI) {
        const int id = tile_id[(size_t)pI * NT + pJ];
        if (id < 0) continue;
        nt = std::max(nt, id + 1);
        const int I = nat_of_pos[(size_t)pI], J = nat_of_pos[(size_t)pJ];
        int t = toucher[(size_t)std::max(I, J) * NT + std::min(I, J)];
        if (pI == pJ) {                       // diagonal tile of natural camera tile I: somebody has to add U_c
            if (t == -1) t = 0;
            o.add_u[(size_t)I] = (t >= 0 ? t : 0) == rank ? 1 : 0;
        }
        if (t >= 0) excl[(size_t)t].push_back(id);
        else if (t == -2) shared.push_back(id);
        else fill.push_back(id);
    }
    size_t maxc = 0;
    for (const auto& v : excl) maxc = std::max(maxc, v.size());
    std::vector<int> newid((size_t)nt, -1);
    for (int r = 0; r < nranks; ++r) for (size_t i = 0; i < excl[(size_t)r].size(); ++i) newid[(size_t)excl[(size_t)r][i]] = (int)((size_t)r * maxc + i);
    for (size_t i = 0; i < shared.size(); ++i) newid[(size_t)shared[i]] = (int)((size_t)nranks * maxc + i);
    for (size_t i = 0; i < fill.size(); ++i) newid[(size_t)fill[i]] = (int)((size_t)nranks * maxc + shared.size() + i);
    for (int& v : tile_id) if (v >= 0) v = newid[(size_t)v];
    o.block = (long long)maxc; o.shared0 = (long long)nranks * (long long)maxc; o.nshared = (long long)shared.size(); o.nfill = (long long)fill.size();
    o.nslots = o.shared0 + o.nshared + o.nfill;
    for (const auto& v : excl) o.nexcl.push_back((long long)v.size());
    return o;
}

}  // namespace nlls
