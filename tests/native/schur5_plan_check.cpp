// CPU check of the Schur v5 host plan (nllssolver.jl_b200/csrc/schur5_plan.hpp): runs the plan's semantics — entry streams per
// consumer warp, window coordinates, band / tile ranges, validity masks, FLUSH address mapping — in plain C++ on random problems
// and compares the accumulated reduced system with a brute-force evaluation of  S -= sum_p W_p' (V_p + lambda I)^-1 W_p.
// Test infrastructure only (built by tests/test_schur5_plan.py with g++); the lane-level fragment layout of the CUDA kernel is
// the one thing it does not model.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>

#include "../../nllssolver.jl_b200/csrc/schur5_plan.hpp"

using namespace nlls;

namespace {
void inv3(const double* V, double lambda, double* Ai) {   // V: 3x3 column-major full symmetric
    double a00 = V[0] + lambda, a10 = V[1], a20 = V[2], a11 = V[4] + lambda, a21 = V[5], a22 = V[8] + lambda;
    double c00 = a11 * a22 - a21 * a21, c10 = a20 * a21 - a10 * a22, c20 = a10 * a21 - a20 * a11;
    double det = a00 * c00 + a10 * c10 + a20 * c20, id = 1.0 / det;
    double i00 = c00 * id, i10 = c10 * id, i20 = c20 * id, i11 = (a00 * a22 - a20 * a20) * id, i21 = (a10 * a20 - a00 * a21) * id, i22 = (a00 * a11 - a10 * a10) * id;
    double M[9] = {i00, i10, i20, i10, i11, i21, i20, i21, i22};
    std::memcpy(Ai, M, sizeof(M));
}

template <int DC>
double run_case(unsigned seed, int nA, int nB, double kmean, int scatter_every, int ncta, double* stats) {
    using C = Schur5Cfg<DC>;
    constexpr int ST = 72, WB = 3 * DC;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::poisson_distribution<int> Pk(std::max(0.1, kmean - 2.0));
    // ---- problem: banded contiguous tracks (create_bal_shaped's rule), every `scatter_every`-th point gets a gap in its camera list
    std::vector<int> obs_start(1, 0), obs_cam;
    for (int p = 0; p < nB; ++p) {
        int k = std::min(nA, 2 + Pk(rng));
        double centre = 2.0 + (double)(nA - 3) * p / std::max(1, nB - 1);
        int start = (int)std::ceil(centre - k / 2.0);
        start = std::max(1, std::min(start, nA - k + 1)) - 1;
        const bool gap = scatter_every > 0 && (p % scatter_every) == scatter_every - 1 && start + k < nA && k >= 2;
        for (int j = 0; j < k; ++j) obs_cam.push_back(start + j + ((gap && j == k - 1) ? 1 : 0));
        obs_start.push_back((int)obs_cam.size());
    }
    const long long nobs = (long long)obs_cam.size();
    const long long hB = (long long)DC * DC * nA;
    std::vector<double> H((size_t)(hB + WB * nobs + 9ll * nB + 2), 0.0), g((size_t)(DC * nA + 3 * nB), 0.0);
    for (int p = 0; p < nB; ++p) {
        for (int j = obs_start[p]; j < obs_start[p + 1]; ++j)
            for (int e = 0; e < WB; ++e) H[(size_t)(hB + (long long)WB * j + 9ll * p + e)] = U(rng);
        double* V = &H[(size_t)(hB + (long long)WB * obs_start[p + 1] + 9ll * p)];
        double L[9]; for (double& v : L) v = U(rng);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = (i == j) ? 0.5 : 0.0; for (int q = 0; q < 3; ++q) s += L[i + 3 * q] * L[j + 3 * q]; V[i + 3 * j] = s; }
        for (int i = 0; i < 3; ++i) g[(size_t)(DC * nA + 3 * p + i)] = U(rng);
    }
    const double lambda = 0.37;
    // ---- tile-sparse S with every tile present, random tile permutation
    const int TC = ST / DC, NT = (nA + TC - 1) / TC;
    std::vector<int> pos((size_t)NT), tile_id((size_t)NT * NT, -1);
    for (int i = 0; i < NT; ++i) pos[(size_t)i] = i;
    std::shuffle(pos.begin(), pos.end(), rng);
    int ntile = 0;
    for (int J = 0; J < NT; ++J) for (int I = J; I < NT; ++I) tile_id[(size_t)I * NT + J] = ntile++;
    std::vector<double> S((size_t)ntile * ST * ST, 0.0), rhs((size_t)NT * ST, 0.0);
    // ---- the plan, executed
    Schur5Plan P = schur5_build_plan<DC>(obs_start, obs_cam, nA, ncta, 6);
    if (P.cta_item.empty()) return -1.0;
    std::vector<unsigned char> handled((size_t)nB, 0);
    struct Acc { double t[C::BR][C::NTW][8][8]; double r[C::BR][8]; };
    for (size_t c = 0; c + 1 < P.cta_item.size(); ++c) {
        std::vector<Acc> acc(S5_CONSUMERS);
        for (auto& a : acc) std::memset(&a, 0, sizeof(Acc));
        for (int ii = P.cta_item[c]; ii < P.cta_item[c + 1]; ++ii) {
            const Schur5Item& it = P.items[(size_t)ii];
            const unsigned* blob = &P.blob[it.blob0];
            const unsigned* ptab = blob + S5_HDR;
            const unsigned* ents = blob + blob[14];
            {   // layout: header, point table (even word count), observation -> local point bytes (even word count), entries
                const int pw = (it.npt + 1) / 2 + (((it.npt + 1) / 2) & 1), ow = (it.nob + 3) / 4 + (((it.nob + 3) / 4) & 1);
                if ((int)blob[13] != S5_HDR + pw || ents != ptab + pw + ow || (int)blob[15] != ((it.flags >> 2) & 1)) return -7.0;
                const unsigned char* opt = reinterpret_cast<const unsigned char*>(blob + blob[13]);
                for (int q = 0, j = 0; q < it.npt; ++q) {
                    const int oe = (int)((ptab[q >> 1] >> (16 * (q & 1))) & 0xffffu);
                    for (; j < oe; ++j) if (opt[j] != q) return -13.0;
                }
            }
            const double* span = &H[(size_t)(hB + (long long)WB * it.ob0 + 9ll * it.pt0)];
            if ((((hB + (long long)WB * it.ob0 + 9ll * it.pt0) & 1) ? 4 : 0) != it.flags) return -2.0;
            for (int w = 0; w < S5_CONSUMERS; ++w) {
                const unsigned first = blob[w] >> 16, cnt = blob[w] & 0xffffu;
                for (unsigned e = first; e < first + cnt; ++e) {
                    const unsigned x = ents[2 * e], y = ents[2 * e + 1];
                    const int band = (int)((y >> 16) & 15);
                    Acc& A = acc[(size_t)w];
                    if (y & S5_FLUSH) {
                        const int base = P.super_base[x];
                        const std::vector<long long> ft = schur5_flush_table<DC>(P.super_base, tile_id, pos, NT);
                        const long long* fe = &ft[(size_t)x * Schur5Flush<DC>::STRIDE];
                        if (fe[0] != base) return -10.0;
                        for (int r = 0; r < C::BR; ++r) {
                            const int mt = C::row_tile(band, r);
                            for (int fr = 0; fr < 8; ++fr) {
                                const int R = 8 * mt + fr, ca = R / DC, ar = R % DC;
                                if (ca >= C::WC || base + ca >= nA) continue;
                                rhs[(size_t)(base + ca) * DC + ar] -= A.r[r][fr];
                                for (int nt2 = 0; nt2 <= mt; ++nt2) for (int cc = 0; cc < 8; ++cc) {
                                    const int Cc = 8 * nt2 + cc, cb = Cc / DC, cr = Cc % DC;
                                    if (cb > ca || (cb == ca && cr > ar)) continue;
                                    const long long so = schur5_soff(base + ca, ar, base + cb, cr, tile_id.data(), pos.data(), NT, DC, ST);
                                    {   // the table the kernel uses must give the same address
                                        using F = Schur5Flush<DC>;
                                        const int I0 = (int)fe[1], ta = (base + ca) / F::TC - I0, tb = (base + cb) / F::TC - I0;
                                        const long long pe = fe[2 + F::pair(ta, tb)];
                                        if (pe < 0) return -11.0;
                                        const int r0 = (base + ca - (I0 + ta) * F::TC) * DC + ar, c0 = (base + cb - (I0 + tb) * F::TC) * DC + cr;
                                        const long long so2 = (pe >> 1) + ((pe & 1) ? c0 + (long long)ST * r0 : r0 + (long long)ST * c0);
                                        if (so2 != so) return -12.0;
                                    }
                                    S[(size_t)so] -= A.t[r][nt2][fr][cc];
                                }
                            }
                        }
                        std::memset(&A, 0, sizeof(Acc));
                        continue;
                    }
                    const int lim0 = (int)((x >> 16) & 255), off0 = (int)(x >> 24), q = (int)(y & 255u), sid = (int)((y >> 8) & 255u);
                    if (lim0 % DC || off0 % DC) return -8.0;
                    const int k = lim0 / DC, delta = off0 / DC, wrel = (int)(x & 0xffffu) - C::BIAS + 3 * off0;
                    const int oe = (int)((ptab[q >> 1] >> (16 * (q & 1))) & 0xffffu);
                    const double* W = span + wrel;
                    const double* V = span + WB * oe + 9 * q;
                    if (V != W + WB * k) return -3.0;                      // the point table and the entry must agree
                    const int pg = it.pt0 + q;
                    handled[(size_t)pg] |= (unsigned char)(1u << band);
                    double Ai[9]; inv3(V, lambda, Ai);
                    const double* gp = &g[(size_t)(DC * nA + 3 * pg)];
                    int lo, hi; schur5_tile_range(DC, delta, k, lo, hi);
                    const int lim = DC * k;
                    {   // the shape number the kernel switches on
                        int nact = 0;
                        for (int r = 0; r < C::BR; ++r) { const int mt = C::row_tile(band, r); if (mt >= lo && mt <= hi) ++nact; }
                        if (nact < 1 || sid != C::shape_id(band, lo, nact) || sid >= C::nshapes(band) || C::shape_tlo(band, sid) != lo || C::shape_nact(band, sid) != nact) return -9.0;
                    }
                    for (int r = 0; r < C::BR; ++r) {
                        const int mt = C::row_tile(band, r);
                        if (mt < lo || mt > hi) continue;
                        double a[8][3];
                        for (int fr = 0; fr < 8; ++fr) {
                            const int rl = 8 * mt + fr - DC * delta;
                            for (int kk = 0; kk < 3; ++kk) {
                                double yv = 0.0;
                                if (rl >= 0 && rl < lim) for (int m = 0; m < 3; ++m) yv += Ai[kk + 3 * m] * W[3 * rl + m];
                                a[fr][kk] = yv;
                                A.r[r][fr] += yv * gp[kk];
                            }
                        }
                        for (int nt2 = lo; nt2 <= std::min(hi, mt); ++nt2)
                            for (int fr = 0; fr < 8; ++fr) for (int cc = 0; cc < 8; ++cc) {
                                const int cl = 8 * nt2 + cc - DC * delta;
                                if (cl < 0 || cl >= lim) continue;
                                double s = 0.0;
                                for (int kk = 0; kk < 3; ++kk) s += a[fr][kk] * W[3 * cl + kk];
                                A.t[r][nt2][fr][cc] += s;
                            }
                    }
                }
            }
        }
        for (auto& a : acc) { const double* z = reinterpret_cast<const double*>(&a); for (size_t i = 0; i < sizeof(Acc) / 8; ++i) if (z[i] != 0.0) return -4.0; }   // every CTA ends flushed
    }
    // ---- outliers + brute force
    std::vector<unsigned char> is_out((size_t)nB, 0);
    for (int p : P.outliers) is_out[(size_t)p] = 1;
    const int n = NT * ST;
    std::vector<double> D((size_t)n * n, 0.0), rref((size_t)n, 0.0);
    double contrib = 0;
    for (int p = 0; p < nB; ++p) {
        const int b = obs_start[p], e = obs_start[p + 1];
        if (e == b) continue;
        if (!is_out[(size_t)p] && handled[(size_t)p] == 0) return -5.0;
        if (is_out[(size_t)p] && handled[(size_t)p] != 0) return -6.0;
        const double* V = &H[(size_t)(hB + (long long)WB * e + 9ll * p)];
        double Ai[9]; inv3(V, lambda, Ai);
        for (int i = b; i < e; ++i) {
            const double* Wi = &H[(size_t)(hB + (long long)WB * i + 9ll * p)];
            for (int a = 0; a < DC; ++a) {
                double t = 0;
                for (int m = 0; m < 3; ++m) for (int q = 0; q < 3; ++q) t += Wi[3 * a + m] * Ai[m + 3 * q] * g[(size_t)(DC * nA + 3 * p + q)];
                rref[(size_t)obs_cam[(size_t)i] * DC + a] -= t;
                if (is_out[(size_t)p]) rhs[(size_t)obs_cam[(size_t)i] * DC + a] -= t;
            }
            for (int j = b; j <= i; ++j) {
                const double* Wj = &H[(size_t)(hB + (long long)WB * j + 9ll * p)];
                contrib += 1;
                for (int a = 0; a < DC; ++a) for (int bb = 0; bb < DC; ++bb) {
                    if (i == j && bb > a) continue;
                    double t = 0;
                    for (int m = 0; m < 3; ++m) for (int q = 0; q < 3; ++q) t += Wi[3 * a + m] * Ai[m + 3 * q] * Wj[3 * bb + q];
                    D[(size_t)(obs_cam[(size_t)i] * DC + a) + (size_t)n * (obs_cam[(size_t)j] * DC + bb)] -= t;
                    if (is_out[(size_t)p]) S[(size_t)schur5_soff(obs_cam[(size_t)i], a, obs_cam[(size_t)j], bb, tile_id.data(), pos.data(), NT, DC, ST)] -= t;
                }
            }
        }
    }
    // ---- un-tile S independently of schur5_soff and compare
    double err = 0, mx = 0;
    std::vector<int> nat((size_t)NT);
    for (int i = 0; i < NT; ++i) nat[(size_t)pos[(size_t)i]] = i;
    std::vector<double> D2((size_t)n * n, 0.0);
    for (int pJ = 0; pJ < NT; ++pJ) for (int pI = pJ; pI < NT; ++pI) {
        const double* T = &S[(size_t)tile_id[(size_t)pI * NT + pJ] * ST * ST];
        const int I = nat[(size_t)pI], J = nat[(size_t)pJ];
        for (int cc = 0; cc < ST; ++cc) for (int r = 0; r < ST; ++r) {
            const double v = T[r + ST * cc];
            if (v == 0.0) continue;
            int gr = I * ST + r, gc = J * ST + cc;      // TC * DC == ST for DC in {6, 9}
            if (gr < gc) std::swap(gr, gc);             // tiles above the natural diagonal hold the transposed block
            D2[(size_t)gr + (size_t)n * gc] += v;
        }
    }
    for (size_t i = 0; i < D.size(); ++i) { err = std::max(err, std::fabs(D[i] - D2[i])); mx = std::max(mx, std::fabs(D[i])); }
    for (int i = 0; i < DC * nA; ++i) { err = std::max(err, std::fabs(rref[(size_t)i] - rhs[(size_t)i])); }
    if (stats) {
        stats[0] = (double)P.n_dmma / std::max(1.0, contrib); stats[1] = P.out_frac; stats[2] = P.imbalance; stats[3] = (double)P.n_super;
        stats[4] = (double)P.n_flush; stats[5] = (double)P.outliers.size(); stats[6] = (double)P.n_entries; stats[7] = (double)(P.n_frag_a * 3 + P.n_frag_b * 2 + P.n_entries * 3) / std::max(1, nB);
    }
    return err / std::max(mx, 1e-300);
}
}  // namespace

extern "C" double schur5_plan_check(int dc, unsigned seed, int nA, int nB, double kmean, int scatter_every, int ncta, double* stats) {
    if (dc == 6) return run_case<6>(seed, nA, nB, kmean, scatter_every, ncta, stats);
    if (dc == 9) return run_case<9>(seed, nA, nB, kmean, scatter_every, ncta, stats);
    return -100.0;
}

// statistics of the plan for a given problem structure (tests + offline tuning of the planner)
extern "C" int schur5_plan_stats(int dc, long long nB, const int* obs_start, const int* obs_cam, long long nA, int ncta, int maxrun, double* stats) {
    std::vector<int> os(obs_start, obs_start + nB + 1), oc(obs_cam, obs_cam + obs_start[nB]);
    Schur5Plan P = dc == 6 ? schur5_build_plan<6>(os, oc, nA, ncta, maxrun) : schur5_build_plan<9>(os, oc, nA, ncta, maxrun);
    if (P.cta_item.empty()) return -1;
    stats[0] = (double)P.n_dmma; stats[1] = P.out_frac; stats[2] = P.imbalance; stats[3] = (double)P.n_super; stats[4] = (double)P.n_flush;
    stats[5] = (double)P.outliers.size(); stats[6] = (double)P.n_entries; stats[7] = (double)P.n_frag_a; stats[8] = (double)P.n_frag_b; stats[9] = (double)P.blob.size() * 4;
    return 0;
}
