// Kernels of the LM hot path (sm_100a, FP64).  Launch wrappers live in nlls_b200.cu.
//
// Data layout in HBM:
//   observations, point-major ("reordercostsforschur!" order, src/problem.jl:177-199): obs_cam[nobs] (int32 local
//   camera), obs_pt[nobs] (int32 local point), obs_z[nobs] (double2), obs_start[npt+1] (CSR by point);
//   a second camera-major copy (cm_pt, cm_z) for the camera pass.
//   H: the reference's BlockSparseMatrix.data for cameras-before-points variable order:
//     [ U_c : nA blocks of DC x DC ][ for each point p: W_{p,c1} .. W_{p,ck} (3 x DC each, ascending camera), V_p (3x3) ]
//   so the row of point p starts at hB + 3*DC*obs_start[p] + 9*p and one tile of consecutive points is ONE contiguous span.
//   g: [ g_c : DC*nA ][ g_p : 3*nB ].
#pragma once
#include "common.cuh"
#include "residuals.cuh"
#include "reduced.cuh"

namespace nlls {

struct DevProblem {
    // point-major observations
    const int* obs_cam;
    const int* obs_pt;
    const double2* obs_z;
    const int* obs_start;
    const int* tile_pt;  // [ntiles + 1]
    int ntiles, nA, nB, nobs;
    // camera-major copy
    const int* cm_pt;
    const double2* cm_z;
    const int* item_cam;
    const int* item_beg;
    const int* item_end;
    const int* cam_item_start;  // [nA + 1]
    int nitems;
    // linear system
    double* H;
    double* g;
    long long hB;  // DC*DC*nA
    long long gB;  // DC*nA
    RobustParams rk;
    // several cost sets (addcost! with residual types that differ in their robust kernel, src/cost.jl:54): per-observation set id in
    // point-major (obs_set) and camera-major (cm_set) order and the sets' kernels; nullptr for the usual single set (rk)
    const unsigned char* obs_set;
    const unsigned char* cm_set;
    const RobustParams* rk_tab;
    int use_tma;
    int schur_stride;
    // reduced camera system: tile-sparse (tile_id[I * NT + J], -1 = structurally zero) or dense n x n
    const int* tile_id;
    const int* tile_pos;   // natural camera tile -> position in the elimination order
    int NT;
    int s_tiled;
    // optimize!(problem, options, unfixed): per camera / point 1 = FIXED (nullptr: none).  A fixed variable keeps its place in the
    // system but is frozen: its diagonal block is the identity, its gradient and every cross block with it are zero, so its step
    // is exactly zero and the other variables see the reference's reduced system (src/cost.jl:36-47, src/linearsystem.jl:93-102).
    const unsigned char* fixA;
    const unsigned char* fixB;
};

// robust kernel of point-major observation j / camera-major observation k.  MS (several cost sets) is a template parameter of the
// residual kernels: the single-set instantiations read p.rk from the parameter bank exactly as before (a run-time test on
// p.obs_set measured +13 % on the camera pass)
template <bool MS> __device__ __forceinline__ RobustParams rk_point(const DevProblem& p, int j) { if constexpr (MS) return p.rk_tab[p.obs_set[j]]; else return p.rk; }
template <bool MS> __device__ __forceinline__ RobustParams rk_cam(const DevProblem& p, int k) { if constexpr (MS) return p.rk_tab[p.cm_set[k]]; else return p.rk; }

// ---------------------------------------------------------------------------------------------------
// K1  lin_point: fused residual + analytic Jacobian + robust weights + J'WJ for tiles of points (persistent, pipelined).
// Replaces costgradhess! (src/cost.jl:29-52) -> computerescostgradhess (src/residual.jl:57-111) ->
// updatesymlinearsystem! (src/linearsystem.jl:132-175) for the point rows of H (W and V blocks) and g_p.
//
// A CTA loops over tiles t = blockIdx.x + k gridDim.x (one thread per observation of the tile):
//   * one int4 descriptor per tile (pt0, npt, ob0, nob) replaces the tile -> obs_start chain;
//   * register pipeline, three tiles deep: the observation (cam, pt, z) and CSR slice of tile k+2 and the camera / point
//     values of tile k+1 are in flight while tile k is computed, so no global load is waited for where it is issued
//     (the one-tile-per-CTA v0 kernel was latency bound on exactly that chain: ncu long_scoreboard + barrier, 35 % DRAM);
//   * the staged H span is double buffered: the bulk store of tile k drains while tile k+1 is computed;
//   * the span sits in shared memory at the parity of its global offset, so EVERY tile goes out as one 16-byte aligned
//     cp.async.bulk (SASS UBLKCP) plus at most one scalar head/tail element; W blocks are staged with 128-bit stores;
//   * per-point sums (V_p, g_p) run over (point, element) items in parallel, each in observation order (deterministic, and
//     the reference's order when costs are stored camera-major).
// ---------------------------------------------------------------------------------------------------
template <class R, int TO, int TP>
struct LinSmem {
    static constexpr int WB = 3 * R::DC;
    static constexpr int OUT = WB * TO + 9 * TP + 2;     // staged H span + parity slot (even)
    static constexpr int PC = 9 * TO;                    // per-observation point contributions (6 V + 3 g), SoA
    static constexpr size_t bytes = (size_t)(2 * OUT + PC + 16) * sizeof(double) + (size_t)(TP + 4) * sizeof(int);
};

template <class R, int TO, int TP, bool MS = false>
__global__ void __launch_bounds__(TO) lin_point_kernel(DevProblem p, const int4* __restrict__ tiles, const double* __restrict__ cams,
                                                       const double* __restrict__ pts, double* __restrict__ cost_partials) {
    constexpr int DC = R::DC, WB = 3 * DC, NP = (WB - 1) / 2, NW = TO / 32;
    using SM = LinSmem<R, TO, TP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_out0 = reinterpret_cast<double*>(smem_raw);
    double* s_pc = s_out0 + 2 * SM::OUT;
    double* s_red = s_pc + SM::PC;
    int* s_ost = reinterpret_cast<int*>(s_red + 16);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x;
    int t = blockIdx.x;
    if (t >= p.ntiles) return;
    // pipeline registers: tile k (level 2 complete), tile k+1 (level 1), tile k+2 (being loaded), descriptor of tile k+3.
    // Every prefetch is an unconditional load from a clamped (always valid) address straight into its register: a predicated
    // load or arithmetic on the loaded value would make the warp wait for it right where it is issued.
    const int last = p.ntiles - 1;
    int4 d = tiles[t];
    int4 dn = tiles[min(t + G, last)];
    int4 dn2 = tiles[min(t + 2 * G, last)];
    int cam, ptg, lo, ncam, nptg, nlo;
    double2 z, nz;
    { const int j = d.z + min(tid, max(d.w - 1, 0)); cam = p.obs_cam[j]; ptg = p.obs_pt[j]; z = p.obs_z[j]; lo = p.obs_start[d.x + min(tid, d.y)]; }
    { const int j = dn.z + min(tid, max(dn.w - 1, 0)); ncam = p.obs_cam[j]; nptg = p.obs_pt[j]; nz = p.obs_z[j]; nlo = p.obs_start[dn.x + min(tid, dn.y)]; }
    double cv[R::NC], X[3];
    R::load_cam(cams, cam, cv);
    X[0] = __ldg(pts + (size_t)3 * ptg); X[1] = __ldg(pts + (size_t)3 * ptg + 1); X[2] = __ldg(pts + (size_t)3 * ptg + 2);

    for (int it = 0;; ++it) {
        const bool has_next = t + G < p.ntiles;
        const int4 dn3 = tiles[min(t + 3 * G, last)];
        const int pt0 = d.x, npt = d.y, ob0 = d.z, nob = d.w;
        const size_t span0 = (size_t)p.hB + (size_t)WB * ob0 + (size_t)9 * pt0;
        const int par = (int)(span0 & 1);
        double* s_base = s_out0 + (it & 1) * SM::OUT + par;   // element e of the span lives at s_base[e]: same parity as H + span0 + e
        if (tid <= npt) s_ost[tid] = lo - ob0;

        double c = 0.0;
        if (tid < nob) {
            const int pl = ptg - pt0;
            double r[2], Jc[2][DC], Jp[2][3];
            R::resjac(cv, X, z.x, z.y, r, Jc, Jp);
            const double s = r[0] * r[0] + r[1] * r[1];               // sqnorm            src/residual.jl:72
            double rho, d1, d2;
            robustifydcost(rk_point<MS>(p, ob0 + tid), s, rho, d1, d2);   //                   src/residual.jl:78
            c = 0.5 * rho;                                            //                   src/residual.jl:110
            const bool fpt = p.fixB != nullptr && p.fixB[ptg];
            const bool fcross = fpt || (p.fixA != nullptr && p.fixA[cam]);
            double gc[DC], gp[3];
#pragma unroll
            for (int a = 0; a < DC; ++a) gc[a] = jtr<R>(Jc, r, a);                        // g = J' r   :73   (explicit FMAs: -fmad=false)
#pragma unroll
            for (int b = 0; b < 3; ++b) gp[b] = fma(Jp[1][b], r[1], Jp[0][b] * r[0]);
            const double td2 = 2 * d2;
            double w[WB];                                             // W block, column-major 3 x DC   src/linearsystem.jl:149
#pragma unroll
            for (int a = 0; a < DC; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    double h = jtj_pc<R>(Jp, Jc, b, a);                                // H = J' J   :74
                    if (d1 != 1.0) h *= d1;                                            // IRLS       :91-93
                    if (d2 != 0.0) h = fma(td2 * gp[b], gc[a], h);                     // Triggs     :95-97
                    w[b + 3 * a] = fcross ? 0.0 : h;
                }
            // 128-bit shared stores at the block's own parity
            const int off = WB * tid + 9 * pl;
            double* wd = s_base + off;
            const bool odd = ((par + off) & 1) != 0;
            double2* wd2 = reinterpret_cast<double2*>(wd + (odd ? 1 : 0));
#pragma unroll
            for (int k = 0; k < NP; ++k) wd2[k] = make_double2(odd ? w[2 * k + 1] : w[2 * k], odd ? w[2 * k + 2] : w[2 * k + 1]);
            if (odd) wd[0] = w[0];
#pragma unroll
            for (int e = 2 * NP; e < WB; ++e) if (e >= 2 * NP + (odd ? 1 : 0)) wd[e] = w[e];
            // this observation's contribution to V_p (lower triangle) and g_p
            int q = 0;
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2)
#pragma unroll
                for (int b = b2; b < 3; ++b) {
                    double h = fma(Jp[1][b], Jp[1][b2], Jp[0][b] * Jp[0][b2]);
                    if (d1 != 1.0) h *= d1;
                    if (d2 != 0.0) h = fma(td2 * gp[b], gp[b2], h);
                    s_pc[9 * tid + (q++)] = fpt ? 0.0 : h;
                }
#pragma unroll
            for (int b = 0; b < 3; ++b) s_pc[9 * tid + 6 + b] = fpt ? 0.0 : ((d1 != 1.0) ? gp[b] * d1 : gp[b]);  // g *= dc  :99-101
        }
        // level-2 loads of tile k+1 (their addresses arrived one iteration ago)
        double ncv[R::NC], nX[3];
        R::load_cam(cams, ncam, ncv);
        nX[0] = __ldg(pts + (size_t)3 * nptg); nX[1] = __ldg(pts + (size_t)3 * nptg + 1); nX[2] = __ldg(pts + (size_t)3 * nptg + 2);
        // level-1 loads of tile k+2 (its descriptor arrived one iteration ago)
        int n2cam, n2ptg, n2lo;
        double2 n2z;
        { const int j = dn2.z + min(tid, max(dn2.w - 1, 0)); n2cam = p.obs_cam[j]; n2ptg = p.obs_pt[j]; n2z = p.obs_z[j]; n2lo = p.obs_start[dn2.x + min(tid, dn2.y)]; }
        __syncthreads();

        // per-point sums over (point, element) items:  block(A, p, p) += ..., b[p] += ...   src/linearsystem.jl:140,166
        // (observation order; loads are issued four at a time, adding an exact 0.0 past the end of the point)
        for (int item = tid; item < 9 * npt; item += TO) {
            const int q = item / 9, i = item - 9 * q;
            const int j0 = s_ost[q], j1 = s_ost[q + 1];
            double v = 0.0;
            for (int j = j0; j < j1; j += 4) {
                double a[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) a[e] = (j + e < j1) ? s_pc[9 * (j + e) + i] : 0.0;
#pragma unroll
                for (int e = 0; e < 4; ++e) v += a[e];
            }
            if (p.fixB != nullptr && p.fixB[pt0 + q]) v = (i == 0 || i == 3 || i == 5) ? 1.0 : 0.0;   // frozen point: V_p = I, g_p = 0
            if (i < 6) {   // lower-triangle element i of V_p -> its position(s) in the full column-major 3 x 3 block
                double* V = s_base + WB * j1 + 9 * q;
                const int pa = (0x854210 >> (4 * i)) & 15, pb = (0x874630 >> (4 * i)) & 15;
                V[pa] = v;
                if (pb != pa) V[pb] = v;
            } else {
                p.g[p.gB + (size_t)3 * (pt0 + q) + (i - 6)] = v;
            }
        }
        // cost of the tile: xor-shuffle tree inside a warp, then the warps in order (the cost kernel uses the same tree)
        c = warp_sum(c);
        if (lane == 0) s_red[wid] = c;
        if (p.use_tma) {
            if (tid == 0) bulk_store_wait();     // the previous tile's store has drained: the other stage is free again
            fence_proxy_async();
        }
        __syncthreads();

        const int span = WB * nob + 9 * npt;
        double* gdst = p.H + span0;
        if (tid == 0) {
            double tsum = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) tsum += s_red[i];
            cost_partials[t] = tsum;
            if (p.use_tma) {
                const int body = (span - par) & ~1;
                if (par) gdst[0] = s_base[0];
                if (body > 0) bulk_store(gdst + par, s_base + par, (uint32_t)body * 8u);
                if (par + body < span) gdst[span - 1] = s_base[span - 1];
            }
        }
        if (!p.use_tma) { for (int i = tid; i < span; i += TO) gdst[i] = s_base[i]; }
        if (!has_next) break;
        t += G; d = dn; dn = dn2; dn2 = dn3;
        cam = ncam; ptg = nptg; z = nz; lo = nlo;
        ncam = n2cam; nptg = n2ptg; nz = n2z; nlo = n2lo;
#pragma unroll
        for (int i = 0; i < R::NC; ++i) cv[i] = ncv[i];
        X[0] = nX[0]; X[1] = nX[1]; X[2] = nX[2];
    }
    if (p.use_tma && tid == 0) bulk_store_wait();
}

// ---------------------------------------------------------------------------------------------------
// K3  cost: sum 0.5 rho(|r|^2) with the same tiling and reduction tree as K1 (bit-identical cost for identical variables).
// Replaces cost(vars, costs) (src/cost.jl:11) -> computerescost (src/residual.jl:49-55).  Persistent, register-pipelined.
// ---------------------------------------------------------------------------------------------------
template <class R, int TO>
__global__ void __launch_bounds__(TO) cost_kernel(DevProblem p, const int4* __restrict__ tiles, const double* __restrict__ cams,
                                                  const double* __restrict__ pts, double* __restrict__ cost_partials) {
    constexpr int NW = TO / 32;
    __shared__ double s_red[2 * NW];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x;
    int t = blockIdx.x;
    if (t >= p.ntiles) return;
    const int last = p.ntiles - 1;
    int4 d = tiles[t];
    int4 dn = tiles[min(t + G, last)];
    int4 dn2 = tiles[min(t + 2 * G, last)];
    int cam, ptg, ncam, nptg;
    double2 z, nz;
    { const int j = d.z + min(tid, max(d.w - 1, 0)); cam = p.obs_cam[j]; ptg = p.obs_pt[j]; z = p.obs_z[j]; }
    { const int j = dn.z + min(tid, max(dn.w - 1, 0)); ncam = p.obs_cam[j]; nptg = p.obs_pt[j]; nz = p.obs_z[j]; }
    double cv[R::NC], X[3];
    R::load_cam(cams, cam, cv);
    X[0] = __ldg(pts + (size_t)3 * ptg); X[1] = __ldg(pts + (size_t)3 * ptg + 1); X[2] = __ldg(pts + (size_t)3 * ptg + 2);
    for (int it = 0;; ++it) {
        const bool has_next = t + G < p.ntiles;
        const int4 dn3 = tiles[min(t + 3 * G, last)];
        double ncv[R::NC], nX[3];
        R::load_cam(cams, ncam, ncv);
        nX[0] = __ldg(pts + (size_t)3 * nptg); nX[1] = __ldg(pts + (size_t)3 * nptg + 1); nX[2] = __ldg(pts + (size_t)3 * nptg + 2);
        int n2cam, n2ptg;
        double2 n2z;
        { const int j = dn2.z + min(tid, max(dn2.w - 1, 0)); n2cam = p.obs_cam[j]; n2ptg = p.obs_pt[j]; n2z = p.obs_z[j]; }
        double c = 0.0;
        if (tid < d.w) {
            double r[2];
            R::residual(cv, X, z.x, z.y, r);
            c = 0.5 * robustify(p.rk, r[0] * r[0] + r[1] * r[1]);   // (single cost set only: several sets always take the camera-major pass)
        }
        c = warp_sum(c);
        if (lane == 0) s_red[(it & 1) * NW + wid] = c;
        __syncthreads();
        if (tid == 0) {
            double tsum = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) tsum += s_red[(it & 1) * NW + i];
            cost_partials[t] = tsum;
        }
        if (!has_next) break;
        t += G; d = dn; dn = dn2; dn2 = dn3;
        cam = ncam; ptg = nptg; z = nz;
        ncam = n2cam; nptg = n2ptg; nz = n2z;
#pragma unroll
        for (int i = 0; i < R::NC; ++i) cv[i] = ncv[i];
        X[0] = nX[0]; X[1] = nX[1]; X[2] = nX[2];
    }
}

// ---------------------------------------------------------------------------------------------------
// K2  lin_cam: camera diagonal blocks U_c = sum J_c' W J_c and g_c over the camera-major observation copy, plus the cost
// sum 0.5 rho(|r|^2) of the same observations (the residuals are computed anyway).
// One CTA per work item (a camera and at most CAM_CHUNK of its observations); fixed-shape reduction; per item NU + 1 partials.
// Inside the LM loop this kernel IS the cost evaluation of a try (cost(varnext, costs), src/iterators.jl:157): when the try is
// accepted, the camera blocks it produced are exactly those of the re-linearisation at the accepted point (src/optimize.jl:169),
// so the re-linearisation only runs the point pass and the camera pass costs nothing extra (launch_cost / do_linearize).
// ---------------------------------------------------------------------------------------------------
#ifndef LINCAM_OCC6
#define LINCAM_OCC6 2
#endif
template <class R, bool MS = false>
__global__ void __launch_bounds__(256, (R::DC <= 6) ? LINCAM_OCC6 : 1) lin_cam_kernel(DevProblem p, const double* __restrict__ cams, const double* __restrict__ pts,
                                                      double* __restrict__ partials) {
    constexpr int DC = R::DC, NU = DC * (DC + 1) / 2 + DC;
    __shared__ double s_red[8][NU + 1];
    const int tid = threadIdx.x, item = blockIdx.x;
    const int cam = p.item_cam[item];
    const int beg = p.item_beg[item], end = p.item_end[item];
    double cv[R::NC];
    R::load_cam(cams, cam, cv);
    double acc[NU];
    double cacc = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) acc[i] = 0.0;
    // batches of four observations per thread: the index / measurement loads and then the point gathers of a batch are all
    // issued before the first one is used (the one-at-a-time loop waited two dependent global round trips per observation)
    constexpr int UB = 4;
    for (int j0 = beg + tid; j0 < end; j0 += 256 * UB) {
        int ptb[UB];
        double2 zb[UB];
        double Xb[UB][3];
#pragma unroll
        for (int u = 0; u < UB; ++u) { const int j = min(j0 + 256 * u, end - 1); ptb[u] = p.cm_pt[j]; zb[u] = p.cm_z[j]; }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            Xb[u][0] = __ldg(pts + (size_t)3 * ptb[u]); Xb[u][1] = __ldg(pts + (size_t)3 * ptb[u] + 1); Xb[u][2] = __ldg(pts + (size_t)3 * ptb[u] + 2);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (j0 + 256 * u >= end) break;
            const double2 z = zb[u];
            const double X[3] = {Xb[u][0], Xb[u][1], Xb[u][2]};
            double r[2], Jc[2][DC], Jp[2][3];
            R::resjac(cv, X, z.x, z.y, r, Jc, Jp);
            const double s = r[0] * r[0] + r[1] * r[1];
            double rho, d1, d2;
            robustifydcost(rk_cam<MS>(p, j0 + 256 * u), s, rho, d1, d2);
            cacc += 0.5 * rho;
            // explicit FMAs (the file is compiled with -fmad=false): this pass is FP64-issue bound, not HBM bound — ncu.  The
            // weights fold into the accumulation:  acc += d1 (J'J) + (2 d2 g) g'  (exact no-ops when d1 == 1 / d2 == 0)
            double gc[DC], tg[DC];
#pragma unroll
            for (int a = 0; a < DC; ++a) { gc[a] = jtr<R>(Jc, r, a); tg[a] = (2 * d2) * gc[a]; }
            int q = 0;
#pragma unroll
            for (int a2 = 0; a2 < DC; ++a2)
#pragma unroll
                for (int a = a2; a < DC; ++a) {
                    bool zero;
                    const double h = jtj_cc<R>(Jc, a, a2, &zero);
                    acc[q] = zero ? fma(tg[a], gc[a2], acc[q]) : fma(tg[a], gc[a2], fma(d1, h, acc[q]));
                    ++q;
                }
#pragma unroll
            for (int a = 0; a < DC; ++a) { acc[q] = fma(d1, gc[a], acc[q]); ++q; }
        }
    }
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) s_red[w][i] = v;
    }
    {
        const double v = warp_sum(cacc);
        if (lane == 0) s_red[w][NU] = v;
    }
    __syncthreads();
    if (tid <= NU) {
        double tsum = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) tsum += s_red[k][tid];
        partials[(size_t)item * (NU + 1) + tid] = tsum;
    }
}

// sum of the work items' cost partials (element NU of every item) in a fixed order -> *out  (single CTA)
__global__ void __launch_bounds__(1024) cam_cost_reduce_kernel(const double* __restrict__ partials, int nitems, int stride, double* out) {
    __shared__ double s_red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < nitems; i += 1024) v += partials[(size_t)i * stride + stride - 1];
    v = block_sum(v, s_red);
    if (threadIdx.x == 0) *out = v;
}

// K2b: per camera, add its work-item partials in order and write the full symmetric block + g_c.
template <int DC>
__global__ void cam_finalize_kernel(DevProblem p, const double* __restrict__ partials) {
    constexpr int NU = DC * (DC + 1) / 2 + DC;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.nA * NU) return;
    const int cam = idx / NU, e = idx - cam * NU;
    double s = 0.0;
    for (int it = p.cam_item_start[cam]; it < p.cam_item_start[cam + 1]; ++it) s += partials[(size_t)it * (NU + 1) + e];
    constexpr int NL = DC * (DC + 1) / 2;
    if (e < NL) {
        // e enumerates the lower triangle column by column: (a >= a2)
        int a2 = 0, rem = e;
        while (rem >= DC - a2) { rem -= DC - a2; ++a2; }
        const int a = a2 + rem;
        if (p.fixA != nullptr && p.fixA[cam]) s = (a == a2) ? 1.0 : 0.0;   // frozen camera: U_c = I
        double* U = p.H + (size_t)DC * DC * cam;
        U[a + DC * a2] = s;
        U[a2 + DC * a] = s;
    } else {
        p.g[(size_t)DC * cam + (e - NL)] = (p.fixA != nullptr && p.fixA[cam]) ? 0.0 : s;
    }
}

// Sum `n` partials in a fixed order into out[slot] (single CTA).  op: 0 sum, 1 nan-propagating max.
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ partials, int n, double* out, int op) {
    __shared__ double s_red[32];
    double v = 0.0;
    if (op == 0) { for (int i = threadIdx.x; i < n; i += 1024) v += partials[i]; v = block_sum(v, s_red); }
    else { for (int i = threadIdx.x; i < n; i += 1024) v = nanmax(v, partials[i]); v = block_nanmax(v, s_red); }
    if (threadIdx.x == 0) *out = v;
}

// max_i |H_ii| over every diagonal entry (initlambda, src/iterators.jl:131-137). out must be zeroed first.  One thread per variable
// (a camera's DC / a point's 3 diagonal entries), grid-stride, one atomic per CTA: the first version (one thread per entry, one atomic
// per warp) spent 120 us on the Venice shape in 93 000 same-address atomics — 5 % of a one-iteration optimize! call.
template <int DC>
__global__ void __launch_bounds__(256) maxdiag_kernel(DevProblem p, unsigned long long* out) {
    __shared__ double s_red[8];
    const long long nv = (long long)p.nA + p.nB;
    double v = 0.0;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < nv; idx += (long long)gridDim.x * 256) {
        if (idx < p.nA) {
            if (p.fixA != nullptr && p.fixA[idx]) continue;
            const double* U = p.H + (size_t)DC * DC * idx;
#pragma unroll
            for (int a = 0; a < DC; ++a) v = nanmax(v, fabs(U[a + DC * a]));
        } else {
            const long long pt = idx - p.nA;
            if (p.fixB != nullptr && p.fixB[pt]) continue;
            const double* V = p.H + (size_t)p.hB + (size_t)3 * DC * p.obs_start[pt + 1] + (size_t)9 * pt;
            v = nanmax(v, nanmax(fabs(V[0]), nanmax(fabs(V[4]), fabs(V[8]))));
        }
    }
    v = block_nanmax(v, s_red);
    if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(v));  // non-negative doubles (and NaN above them) order like integers
}

// ---------------------------------------------------------------------------------------------------
// Schur elimination of the point blocks (new functionality, mathematically equal to the reference's
// full-system solve (H + lambda I) x = g, src/linearsolver.jl:29 — SURVEY F3).
//   A_p = V_p + lambda I ;  S = U + lambda I - sum_p W_p' A_p^-1 W_p ;  rhs = g_c - sum_p W_p' A_p^-1 g_p
// S is the tile-sparse reduced camera matrix of reduced.cuh (lower triangle in the permuted tile numbering).
// ---------------------------------------------------------------------------------------------------
// load the tile's contiguous H span into shared memory (TMA bulk load when 16-byte aligned)
__device__ __forceinline__ void load_span(double* s_dst, const double* gsrc, int span, uint64_t* bar, int use_tma) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(gsrc) & 15) == 0) && ((span & 1) == 0);
    if (use_tma && aligned && span > 0) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (uint32_t)span * 8u);
            bulk_load(s_dst, gsrc, (uint32_t)span * 8u, bar);
        }
        mbar_wait(bar, 0);
    } else {
        for (int i = threadIdx.x; i < span; i += blockDim.x) s_dst[i] = gsrc[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// Schur v2: CTA-level pre-reduction.  Larger point tiles (<= SCH_OBS observations); the host sorts, per tile, every
// contribution (point, obs i, obs j <= i) by its target block (cam_i, cam_j) and cuts the sorted list into chunks
// (<= SCH_CHUNK contributions of ONE block).  One thread owns a chunk: it accumulates the DC x DC block (and, for diagonal
// blocks, the rhs segment) in registers from shared memory in a fixed order and issues one FP64 reduction per element —
// the number of global reductions drops by the in-tile multiplicity of a block (~8x on Venice-shaped data).
// ---------------------------------------------------------------------------------------------------
constexpr int SCH_OBS = 256;
constexpr int SCH_PTS = 128;
constexpr int SCH_THREADS = 256;
constexpr int SCH_CHUNK = 16;
constexpr int SCH_CBW = 3;      // block columns per thread: a chunk is processed by DC / SCH_CBW threads
constexpr int SCH_YS = 3 * SCH_CBW + 1;   // padded stride of one column group of Y = A^-1 W (16-byte aligned rows)

struct SchurChunk {
    long long soff;   // element offset of block element (0,0) in S
    int ent0;         // first entry (global index into the entry array)
    int cam;          // camera of the block row (rhs segment) — used by diagonal blocks
    short n;          // number of entries
    short flags;      // bit0: block is stored transposed; bit1: diagonal block (cam_i == cam_j)
};
struct SchurPlan {
    const int* stile_pt;          // [2 * nstiles]: first point, one past the last point of every tile
    const int* chunk_off;         // [nstiles + 1]
    const int* ent_off;           // [nstiles + 1] entry range of a tile
    const SchurChunk* chunks;     // per tile sorted by length (long first)
    const unsigned int* ents;     // (i_local << 16) | j_local
    int nstiles;
    long long ld;                 // leading dimension of an S tile (ST)
};

template <int DC>
struct Schur2Smem {
    static constexpr int WB = 3 * DC;
    static constexpr int ROW = WB * SCH_OBS + 9 * SCH_PTS;
    static constexpr int MAXENT = 2048;    // staged contribution entries per tile (tiles with more read them from global memory)
    static constexpr int MAXCH = 256;      // staged chunk descriptors per tile
    static constexpr int YSZ = SCH_OBS * (DC / SCH_CBW) * SCH_YS;   // Y = A_p^-1 W per observation, column groups padded
    static constexpr size_t bytes = (size_t)(ROW + YSZ + 6 * SCH_PTS + 3 * SCH_PTS + 2) * sizeof(double) + (size_t)SCH_OBS * sizeof(unsigned short) + 32 +
                                    (size_t)MAXENT * sizeof(unsigned int) + (size_t)MAXCH * sizeof(SchurChunk);
};

template <int DC>
__global__ void __launch_bounds__(SCH_THREADS, 2) schur2_kernel(DevProblem p, SchurPlan sp, double* __restrict__ S, double* __restrict__ rhs,
                                                                 double* __restrict__ Ainv_out, double lambda) {
    constexpr int WB = 3 * DC;
    constexpr int NG = DC / SCH_CBW;   // column groups per block
    static_assert(DC % SCH_CBW == 0, "block columns must split evenly");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_row = reinterpret_cast<double*>(smem_raw);
    double* s_Y = s_row + Schur2Smem<DC>::ROW;
    double* s_Ai = s_Y + Schur2Smem<DC>::YSZ;
    double* s_t = s_Ai + 6 * SCH_PTS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_t + 3 * SCH_PTS);
    unsigned short* s_pl = reinterpret_cast<unsigned short*>(bar + 2);
    SchurChunk* s_ch = reinterpret_cast<SchurChunk*>(s_pl + SCH_OBS + 8);
    unsigned int* s_ent = reinterpret_cast<unsigned int*>(s_ch + Schur2Smem<DC>::MAXCH);

    const int tid = threadIdx.x;
    const int G = p.schur_stride;
    const int per = (sp.nstiles + G - 1) / G;
    int t = (blockIdx.x % G) * per + blockIdx.x / G;     // strided tile order: co-resident CTAs touch distant cameras
    if (G <= 1) t = blockIdx.x;
    if (t >= sp.nstiles) return;
    const int pt0 = sp.stile_pt[2 * t], pt1 = sp.stile_pt[2 * t + 1];
    const int ob0 = p.obs_start[pt0], ob1 = p.obs_start[pt1];
    const int npt = pt1 - pt0, nob = ob1 - ob0;
    const int c0 = sp.chunk_off[t], c1 = sp.chunk_off[t + 1];
    const int e0 = sp.ent_off[t], e1 = sp.ent_off[t + 1];
    const bool staged = (c1 - c0) <= Schur2Smem<DC>::MAXCH && (e1 - e0) <= Schur2Smem<DC>::MAXENT;
    const size_t span0 = (size_t)p.hB + (size_t)WB * ob0 + (size_t)9 * pt0;
    load_span(s_row, p.H + span0, WB * nob + 9 * npt, bar, p.use_tma);
    for (int i = tid; i < nob; i += SCH_THREADS) s_pl[i] = (unsigned short)(p.obs_pt[ob0 + i] - pt0);
    if (staged) {   // the tile's plan: coalesced copies, so the accumulation loop never waits on global memory
        for (int i = tid; i < c1 - c0; i += SCH_THREADS) s_ch[i] = sp.chunks[c0 + i];
        for (int i = tid; i < e1 - e0; i += SCH_THREADS) s_ent[i] = sp.ents[e0 + i];
    }
    __syncthreads();

    for (int q = tid; q < npt; q += SCH_THREADS) {   // A_p^-1 and t_p = A_p^-1 g_p
        const int oe = p.obs_start[pt0 + q + 1] - ob0;
        const double* V = s_row + WB * oe + 9 * q;
        const double a[6] = {V[0] + lambda, V[1], V[2], V[4] + lambda, V[5], V[8] + lambda};
        double inv[6];
        inv_sym3(a, inv);
#pragma unroll
        for (int i = 0; i < 6; ++i) { s_Ai[6 * q + i] = inv[i]; Ainv_out[(size_t)6 * (pt0 + q) + i] = inv[i]; }
        const double* gp = p.g + p.gB + (size_t)3 * (pt0 + q);
        const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
        s_t[3 * q] = inv[0] * g0 + inv[1] * g1 + inv[2] * g2;
        s_t[3 * q + 1] = inv[1] * g0 + inv[3] * g1 + inv[4] * g2;
        s_t[3 * q + 2] = inv[2] * g0 + inv[4] * g1 + inv[5] * g2;
    }
    __syncthreads();
    // Y_i = A_p^-1 W_i once per observation (every pair (i, j) of the point reuses it), stored by column group
    for (int i = tid; i < nob; i += SCH_THREADS) {
        const int pl = s_pl[i];
        const double* w = s_row + WB * i + 9 * pl;
        const double* ai = s_Ai + 6 * pl;
        const double i00 = ai[0], i10 = ai[1], i20 = ai[2], i11 = ai[3], i21 = ai[4], i22 = ai[5];
        double* y = s_Y + (size_t)i * NG * SCH_YS;
#pragma unroll
        for (int a = 0; a < DC; ++a) {
            const double w0 = w[3 * a], w1 = w[3 * a + 1], w2 = w[3 * a + 2];
            double* ya = y + (a / SCH_CBW) * SCH_YS + 3 * (a % SCH_CBW);
            ya[0] = fma(i20, w2, fma(i10, w1, i00 * w0));
            ya[1] = fma(i21, w2, fma(i11, w1, i10 * w0));
            ya[2] = fma(i22, w2, fma(i21, w1, i20 * w0));
        }
    }
    __syncthreads();

    const int nunits = (c1 - c0) * NG;
    for (int u = tid; u < nunits; u += SCH_THREADS) {
        const int c = c0 + u / NG, b0 = (u % NG) * SCH_CBW;
        const SchurChunk ck = staged ? s_ch[c - c0] : sp.chunks[c];
        const unsigned int* ents = staged ? (s_ent + (ck.ent0 - e0)) : (sp.ents + ck.ent0);
        const bool diag = (ck.flags & 2) != 0;
        const bool do_rhs = diag && b0 == 0;
        double acc[DC][SCH_CBW], racc[DC];
#pragma unroll
        for (int a = 0; a < DC; ++a) {
            racc[a] = 0.0;
#pragma unroll
            for (int b = 0; b < SCH_CBW; ++b) acc[a][b] = 0.0;
        }
        for (int e = 0; e < ck.n; ++e) {
            const unsigned int en = ents[e];
            const int i = (int)(en >> 16), j = (int)(en & 0xffffu);
            const int pl = s_pl[i];
            const double* wi = s_row + WB * i + 9 * pl;
            double T[SCH_CBW][3];   // T = A^-1 W_j (columns b0 .. b0+2), precomputed
            {
                const double2* yj = reinterpret_cast<const double2*>(s_Y + ((size_t)j * NG + b0 / SCH_CBW) * SCH_YS);
                double yv[SCH_YS + 1];
#pragma unroll
                for (int q = 0; q < (SCH_YS + 1) / 2; ++q) { const double2 v = yj[q]; yv[2 * q] = v.x; yv[2 * q + 1] = v.y; }
#pragma unroll
                for (int b = 0; b < SCH_CBW; ++b) { T[b][0] = yv[3 * b]; T[b][1] = yv[3 * b + 1]; T[b][2] = yv[3 * b + 2]; }
            }
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            if (do_rhs) { t0 = s_t[3 * pl]; t1 = s_t[3 * pl + 1]; t2 = s_t[3 * pl + 2]; }
#pragma unroll
            for (int a = 0; a < DC; ++a) {
                const double w0 = wi[3 * a], w1 = wi[3 * a + 1], w2 = wi[3 * a + 2];
#pragma unroll
                for (int b = 0; b < SCH_CBW; ++b) acc[a][b] = fma(w2, T[b][2], fma(w1, T[b][1], fma(w0, T[b][0], acc[a][b])));
                if (do_rhs) racc[a] = fma(w2, t2, fma(w1, t1, fma(w0, t0, racc[a])));
            }
        }
        double* Sb = S + ck.soff;
        const long long sa = (ck.flags & 1) ? sp.ld : 1, sb = (ck.flags & 1) ? 1 : sp.ld;
#pragma unroll
        for (int b = 0; b < SCH_CBW; ++b)
#pragma unroll
            for (int a = 0; a < DC; ++a) {
                if (diag && a < b0 + b) continue;
                atomicAdd(Sb + sa * a + sb * (b0 + b), -acc[a][b]);
            }
        if (do_rhs) {
#pragma unroll
            for (int a = 0; a < DC; ++a) atomicAdd(rhs + (size_t)ck.cam * DC + a, -racc[a]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Schur v4: persistent, software-pipelined FP64 tensor-core accumulation over runs of consecutive tiles ("super-tiles").
// ncu on v2: shared-memory wavefronts at 80 % of peak (half of them bank conflicts — every lane walks its own block's
// contributions, so the operand rows it reads are scattered), FP64 pipe 33 %, and ~11 us of exposed latency per tile
// (dependent index loads -> TMA -> inverse -> products, separated by CTA barriers).  Here:
//  * a WARP owns an S block for a whole super-tile and evaluates  S_ij -= sum_p W_pi' Y_pj  with one mma.sync.m8n8k4.f64 per
//    contribution:  A[a][kk] = W_i[kk][a]  (rows a < DC of 8),  B[kk][b] = Y_j[kk][b]  (b < DC; column DC carries
//    t_p = A_p^-1 g_p, so the rhs segment of a diagonal block falls out of the same product).  The inner index has four
//    slots for three coordinates: s_Y keeps every column as [y0 y1 y2 0], so whatever finite value the A fragment picks up
//    at kk = 3 (the first element of the next operand row of the span) is multiplied by zero.  With that, the 16 lanes of a
//    half warp read 13 consecutive doubles of ONE W row and 16 consecutive doubles of ONE Y row: no bank conflicts (the
//    earlier packing of 4 contributions into 3 DMMAs mixed two rows per load: 2 wavefronts per half warp, and shared-memory
//    wavefronts are what bounds this kernel).  The block accumulator is two registers per lane and the reductions into S
//    are issued once per (block, super-tile).  Blocks are dealt to the warps by the host (longest-processing-time first).
//  * one CTA per SM walks a contiguous range of tiles with a two-stage pipeline: while the warps run the products of tile k
//    they also compute Y for tile k+1 (both stages resident), and the TMA bulk loads of tile k+2 (its H span and its
//    contribution list / per-observation table) are in flight — ONE CTA barrier per tile.
// A super-tile with more than WARPS * NB distinct blocks (a single unusually wide tile) is walked once per round.
// ---------------------------------------------------------------------------------------------------
constexpr int SCH4_WARPS = 16;
constexpr int SCH4_THREADS = 32 * SCH4_WARPS;
template <int DC> struct Schur4Cfg {
    static constexpr int OBS = (DC <= 7) ? 232 : 128;   // tile capacity (observations / points) — two stages must fit 227 KB
    static constexpr int PTS = OBS / 2;
    static constexpr int MT = (DC + 7) / 8;         // 8-row fragments of a block
    static constexpr int NTC = (DC + 1 + 7) / 8;    // 8-column fragments (DC block columns + the rhs column)
    static constexpr int NB = (MT * NTC == 1) ? 16 : 4;   // block slots per warp (accumulators: NB * MT * NTC * 2 doubles per lane)
    static constexpr int WB = 3 * DC;
    static constexpr int YS = 4 * (DC + 1) + 1;     // doubles per observation in s_Y: (DC + 1) columns [Y_j[:, b] | 0], the last one [t_p | 0];
                                                    // + 1: an odd row stride keeps the Y phase's stores (one observation per lane pair) off each other's banks
    static constexpr int ZPADA = 3 * 8 * MT + 8;    // zeros behind the staged span / behind s_Y: operands of the padding entries
    static constexpr int ZPADB = 4 * 8 * NTC + 8;
    static constexpr int ROW = WB * OBS + 9 * PTS;  // doubles of H span per stage
    static constexpr int ROWS = ROW + 2 + ZPADA + 2; // + slack of an 8-byte-misaligned span, + zeros
    static constexpr int YSZ = YS * OBS + ZPADB;
    static constexpr int MAXENT = 4096;             // contribution entries per tile (the host cuts the tiles accordingly)
    static constexpr int BLOB = OBS + MAXENT + 8;   // u32 per stage: [per-observation table | contribution entries | padding entries]
    static constexpr size_t bytes = 2 * ((size_t)(ROWS + YSZ) * sizeof(double) + (size_t)BLOB * sizeof(unsigned int)) + 64;
    static_assert(ROWS % 2 == 0 && (2 * YSZ) % 2 == 0 && BLOB % 4 == 0, "stages must stay 16-byte aligned");
    static_assert(bytes <= 232448, "two stages must fit the 227 KB of shared memory a CTA can have");
};
struct SchurUnit {
    long long soff;   // element offset of block element (0,0) in S
    int cam;          // camera of the block row (rhs segment) — used by diagonal blocks
    int flags;        // bit0: stored transposed; bit1: diagonal block; bit3: slot in use
};
struct SchurItem {    // one (tile, round) visit of a CTA, 48 bytes
    int pt0, npt, ob0, nob;
    int blob0, ne4, wrow, urow;   // first u32 of the tile's blob, padded entry count, row of wtab, row of units to flush after this item (-1: none)
    int flags, pad0, pad1, pad2;  // bit0: first item of a (super-tile, round): clear the accumulators; bit1: round 0 (write A_p^-1)
};
struct SchurPlan4 {
    const int* cta_item;          // [ncta + 1]
    const SchurItem* items;
    const SchurUnit* units;       // [urow][WARPS * NB]
    const unsigned int* blob;     // per tile: [nob padded to 4: (first obs of its point << 31) | (obs index behind the point's W blocks << 16) | local point]
                                  //           [ne4 entries: (smem byte offset of W_i << 16) | smem byte offset of Y_j; per warp, by slot; padded with null entries]
    const unsigned int* wtab;     // [wrow][WARPS * NB + WARPS]: contributions per (warp, slot), then the first entry of every warp's run
    long long ld;
};

// Operands of one contribution (one DMMA): the lane's element of the A fragment, W_i[kk = fk][a = fr], and of the B fragment,
// Y_j[kk = fk][b = fr] — the per-lane parts of the addresses are folded into abase / bbase.
// (measured alternatives, both slower: 4 contributions packed into 3 DMMAs — two operand rows per load, twice the wavefronts;
//  DMMA step <-> coordinate with one contribution per lane — four rows per load)
template <int MT, int NTC> struct SchurAcc { double c[MT][NTC][2]; };
template <int MT, int NTC> struct SchurOps { double a[MT], b[NTC]; };
template <int MT, int NTC>
__device__ __forceinline__ void schur4_fetch(SchurOps<MT, NTC>& op, const uint32_t en, uint32_t abase, uint32_t bbase) {
#pragma unroll
    for (int m = 0; m < MT; ++m) op.a[m] = lds_f64(abase + (en >> 16) + 192u * m);        // 8 rows x 3 doubles further
#pragma unroll
    for (int n = 0; n < NTC; ++n) op.b[n] = lds_f64(bbase + (en & 0xffffu) + 256u * n);   // 8 columns x 4 doubles further
}

template <int DC>
__global__ void __launch_bounds__(SCH4_THREADS, 1) schur4_kernel(DevProblem p, SchurPlan4 sp, double* __restrict__ S, double* __restrict__ rhs,
                                                                  double* __restrict__ Ainv_out, double lambda) {
    using C = Schur4Cfg<DC>;
    constexpr int WB = C::WB, YS = C::YS, NB = C::NB, MT = C::MT, NTC = C::NTC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_row0 = reinterpret_cast<double*>(smem_raw);                 // [2][ROWS]
    double* s_Y0 = s_row0 + 2 * C::ROWS;                                  // [2][YSZ]
    unsigned int* s_blob0 = reinterpret_cast<unsigned int*>(s_Y0 + 2 * C::YSZ);   // [2][BLOB]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_blob0 + 2 * C::BLOB);   // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ka = sp.cta_item[blockIdx.x], nitem = sp.cta_item[blockIdx.x + 1] - ka;
    if (nitem <= 0) return;
    const SchurItem* items = sp.items + ka;
    const int fr = lane >> 2, fk = lane & 3;

    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    // s_Y: the fourth slot of every column and the pad behind the stage stay zero for the whole kernel; the pad behind the span
    // is the A operand of the padding entries
    for (int i = tid; i < 2 * C::YSZ; i += SCH4_THREADS) s_Y0[i] = 0.0;
    for (int i = tid; i < 2 * (C::ZPADA + 2); i += SCH4_THREADS) {
        const int st = i / (C::ZPADA + 2), k = i - st * (C::ZPADA + 2);
        s_row0[st * C::ROWS + C::ROW + 2 + k] = 0.0;
    }
    __syncthreads();

    auto issue = [&](const SchurItem& it, int k) {   // thread 0: TMA bulk loads of item k into stage k & 1
        const int st = k & 1;
        const int mis = (it.flags >> 2) & 1;   // the span starts 8 bytes off a 16-byte boundary: load from the element before it
        const double* gsrc = p.H + (size_t)p.hB + (size_t)WB * it.ob0 + (size_t)9 * it.pt0 - mis;
        const uint32_t span = (uint32_t)((WB * it.nob + 9 * it.npt + mis + 1) & ~1) * 8u;
        const uint32_t nob4 = (uint32_t)((it.nob + 3) & ~3);
        const uint32_t bl = (nob4 + (uint32_t)it.ne4 + 8u) * 4u;   // + the two padding groups behind the list
        mbar_expect_tx(&bar[st], span + bl);
        bulk_load(s_row0 + st * C::ROWS, gsrc, span, &bar[st]);
        if (bl) bulk_load(s_blob0 + st * C::BLOB, sp.blob + it.blob0, bl, &bar[st]);
    };
    auto yphase = [&](const SchurItem& it, int k) {   // Y for item k (stage k & 1); two threads per observation
        const int st = k & 1;
        const int pt0 = it.pt0, nob = it.nob, fl = it.flags;
        const double* row = s_row0 + st * C::ROWS + ((fl >> 2) & 1);
        const unsigned int* info = s_blob0 + st * C::BLOB;
        double* Y = s_Y0 + st * C::YSZ;
        for (int u = tid; u < 2 * nob; u += SCH4_THREADS) {
            const int i = u >> 1, h = u & 1;
            const unsigned int in = info[i];
            const int q = (int)(in & 0xffffu), oe = (int)((in >> 16) & 0x7fffu);
            const int pg = pt0 + q;
            // g_p first: the only global round trip of this phase overlaps the inverse instead of following it (fetching it a
            // whole product phase ahead through obs_pt measured slower: 0.69 -> 0.74 ms)
            const double* gp = p.g + p.gB + (size_t)3 * pg;
            const double g0 = ldg_f64_here(gp), g1 = ldg_f64_here(gp + 1), g2 = ldg_f64_here(gp + 2);
            const double* V = row + WB * oe + 9 * q;
            const double a[6] = {V[0] + lambda, V[1], V[2], V[4] + lambda, V[5], V[8] + lambda};
            double inv[6];
            inv_sym3(a, inv);   // recomputed by every observation of the point — latency-bound either way
            double* y = Y + (size_t)i * YS;
            if (h == 0) {
                if ((in >> 31) && (fl & 2)) {
#pragma unroll
                    for (int e = 0; e < 6; ++e) Ainv_out[(size_t)6 * pg + e] = inv[e];
                }
                y[4 * DC] = inv[0] * g0 + inv[1] * g1 + inv[2] * g2;
                y[4 * DC + 1] = inv[1] * g0 + inv[3] * g1 + inv[4] * g2;
                y[4 * DC + 2] = inv[2] * g0 + inv[4] * g1 + inv[5] * g2;
            }
            const double* w = row + WB * i + 9 * q;
#pragma unroll
            for (int c2 = 0; c2 < (DC + 1) / 2; ++c2) {
                const int c = 2 * c2 + h;
                if (c < DC) {
                    const double w0 = w[3 * c], w1 = w[3 * c + 1], w2 = w[3 * c + 2];
                    y[4 * c] = fma(inv[2], w2, fma(inv[1], w1, inv[0] * w0));
                    y[4 * c + 1] = fma(inv[4], w2, fma(inv[3], w1, inv[1] * w0));
                    y[4 * c + 2] = fma(inv[5], w2, fma(inv[4], w1, inv[2] * w0));
                }
            }
        }
    };

    // item descriptors travel in registers, fetched two items ahead (a global round trip per tile would be exposed otherwise)
    SchurItem itA = items[0], itB = items[nitem > 1 ? 1 : 0], itC = itB;
    if (tid == 0) { issue(itA, 0); if (nitem > 1) issue(itB, 1); }
    mbar_wait(&bar[0], 0);
    yphase(itA, 0);
    unsigned int wt_next = 0;
    if (lane <= NB) wt_next = sp.wtab[(size_t)itA.wrow * (SCH4_WARPS * NB + SCH4_WARPS) + (lane < NB ? warp * NB + lane : SCH4_WARPS * NB + warp)];
    __syncthreads();

    // per-lane operand bases (shared-memory byte addresses): A fragment element (row fr, inner fk) sits 3 fr + fk doubles into
    // the W block, B fragment element (inner fk, column fr) 4 fr + fk doubles into the observation's s_Y row.
    // Lanes of fragment rows >= DC (and of the columns behind the rhs column) read whatever follows the operand row — finite
    // or not, it only reaches accumulator rows / columns that are never written back.
    const uint32_t rowb0 = smem_u32(s_row0) + 8u * (3 * fr + fk), yb0 = smem_u32(s_Y0) + 8u * (4 * fr + fk);
    SchurAcc<MT, NTC> acc[NB];
    for (int k = 0; k < nitem; ++k) {
        const int st = k & 1;
        const SchurItem it = itA;
        if (k + 2 < nitem) itC = items[k + 2];
        const unsigned int wt = wt_next;
        if (k + 1 < nitem && lane <= NB) wt_next = sp.wtab[(size_t)itB.wrow * (SCH4_WARPS * NB + SCH4_WARPS) + (lane < NB ? warp * NB + lane : SCH4_WARPS * NB + warp)];
        if (it.flags & 1) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int n = 0; n < NTC; ++n) { acc[b].c[m][n][0] = 0.0; acc[b].c[m][n][1] = 0.0; }
        }
        {
            const uint32_t rowb = rowb0 + 8u * (uint32_t)(st * C::ROWS + ((it.flags >> 2) & 1));
            const uint32_t yb = yb0 + 8u * (uint32_t)(st * C::YSZ);
            const uint32_t entb = smem_u32(s_blob0 + st * C::BLOB) + 4u * (uint32_t)((it.nob + 3) & ~3);
            // The warp's contributions of this tile are ONE contiguous run of the blob, ordered by slot: the software pipeline
            // (operands of the next contribution and the entry of the one after in flight during the DMMA of the current one) runs
            // across slot boundaries, so a slot with one or two contributions costs no start-up latency.  What is fetched behind
            // the warp's last contribution (the next warp's entries, or the padding entries the host appends to every list) is a
            // valid offset and is never used — stale words of an earlier tile would not be.
            uint32_t ep = entb + 4u * __shfl_sync(0xffffffffu, wt, NB);
            SchurOps<MT, NTC> cur, nxt;
            schur4_fetch<MT, NTC>(cur, lds_u32(ep), rowb, yb);
            uint32_t en = lds_u32(ep + 4u);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int nc = (int)__shfl_sync(0xffffffffu, wt, b);
                for (int c = 0; c < nc; ++c) {
                    schur4_fetch<MT, NTC>(nxt, en, rowb, yb);
                    en = lds_u32(ep + 8u);
                    ep += 4u;
#pragma unroll
                    for (int m = 0; m < MT; ++m)
#pragma unroll
                        for (int n = 0; n < NTC; ++n) dmma884(acc[b].c[m][n][0], acc[b].c[m][n][1], cur.a[m], cur.b[n]);
                    cur = nxt;
                }
            }
        }
        if (it.urow >= 0) {   // one FP64 reduction per block element for the whole (super-tile, round)
            SchurUnit un;
            un.soff = 0; un.cam = 0; un.flags = 0;
            if (lane < NB) un = sp.units[(size_t)it.urow * (SCH4_WARPS * NB) + warp * NB + lane];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const long long soff = __shfl_sync(0xffffffffu, un.soff, b);
                const int cam = __shfl_sync(0xffffffffu, un.cam, b);
                const int flags = __shfl_sync(0xffffffffu, un.flags, b);
                if (!(flags & 8)) continue;
                const bool diag = (flags & 2) != 0;
                double* Sb = S + soff;
                const long long sa = (flags & 1) ? sp.ld : 1, sb = (flags & 1) ? 1 : sp.ld;
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int n = 0; n < NTC; ++n)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int a = 8 * m + fr, c = 8 * n + 2 * fk + h;
                            if (a >= DC) continue;
                            if (c < DC) { if (!(diag && a < c)) atomicAdd(Sb + sa * a + sb * c, -acc[b].c[m][n][h]); }
                            else if (c == DC && diag) atomicAdd(rhs + (size_t)cam * DC + a, -acc[b].c[m][n][h]);
                        }
            }
        }
        if (k + 1 < nitem) {   // the bulk loads of item k+1 had the products above to land
            mbar_wait(&bar[st ^ 1], (uint32_t)(((k + 1) >> 1) & 1));
            yphase(itB, k + 1);
        }
        __syncthreads();   // stage st is free, Y of item k+1 is complete
        if (tid == 0 && k + 2 < nitem) issue(itC, k + 2);
        itA = itB; itB = itC;
    }
}

// ---------------------------------------------------------------------------------------------------
// Back-substitution + variable update + step statistics over tiles of points (persistent, pipelined).
//   dx_p = A_p^-1 (g_p - sum_c W_pc dx_c);  x = -dx (negate!, src/iterators.jl:152);
//   varnext[p] = update(variables[p], x)  (src/linearsystem.jl:206-213, src/variable.jl:10)
// Per CTA partials (n = gridDim.x): [b] max|x_p|, [n + b] sum x_p^2, [2n + b] x'Hx terms owned by the point rows,
// [3n + b] g_p . x_p   (x'Hx uses the UNdamped H like src/iterators.jl:162-163).
// The tile's H span arrives by a double-buffered cp.async.bulk load (mbarrier complete_tx) placed at the parity of its
// global offset; observation indices, camera steps and per-point vectors ride the same register pipeline as K1.
// ---------------------------------------------------------------------------------------------------
// WB consecutive doubles from shared memory with 128-bit loads at the block's own 16-byte parity (the 144-byte block stride
// makes 64-bit accesses 2-way bank conflicted)
template <int WB>
__device__ __forceinline__ void load_wblock(const double* wd, double w[WB]) {
    constexpr int NP = (WB - 1) / 2;
    const bool odd = (reinterpret_cast<uintptr_t>(wd) & 8) != 0;
    const double2* wd2 = reinterpret_cast<const double2*>(wd + (odd ? 1 : 0));
    double2 v[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) v[k] = wd2[k];
    const double xa = wd[odd ? 0 : 2 * NP];
    w[0] = odd ? xa : v[0].x;
#pragma unroll
    for (int e = 1; e < 2 * NP; ++e) w[e] = odd ? ((e & 1) ? v[(e - 1) / 2].x : v[(e - 2) / 2].y) : ((e & 1) ? v[(e - 1) / 2].y : v[e / 2].x);
    w[2 * NP] = odd ? v[NP - 1].y : xa;
    if (WB > 2 * NP + 1) w[WB - 1] = wd[WB - 1];
}

template <int DC, int TO, int TP>
struct BacksubSmem {
    static constexpr int WB = 3 * DC;
    static constexpr int ROW = WB * TO + 9 * TP + 2;
    static constexpr size_t bytes = (size_t)(2 * ROW + 3 * TO + 4 * (TO / 32) + 2) * sizeof(double) + (size_t)(TP + 4) * sizeof(int);
};

template <int DC, int TO, int TP>
__global__ void __launch_bounds__(TO) backsub_kernel(DevProblem p, const int4* __restrict__ tiles, const double* __restrict__ dxc,
                                                     const double* __restrict__ Ainv, const double* __restrict__ pts, double* __restrict__ pts_next,
                                                     double* __restrict__ x, double* __restrict__ partials, int pstride) {
    constexpr int WB = 3 * DC, NW = TO / 32;
    using SM = BacksubSmem<DC, TO, TP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_row0 = reinterpret_cast<double*>(smem_raw);
    double* s_u = s_row0 + 2 * SM::ROW;             // 3 per observation (SoA): W_pc dx_c
    double* s_red = s_u + 3 * TO;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 4 * NW);   // 2 mbarriers
    int* s_ost = reinterpret_cast<int*>(bar + 2);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x;
    int t = blockIdx.x;
    double mx = 0.0, sq = 0.0, xhx = 0.0, gx = 0.0;   // running step statistics of this thread's points
    if (t < p.ntiles) {
        const int last = p.ntiles - 1;
        if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
        __syncthreads();
        // issue the span load of tile descriptor dd into stage st
        auto issue = [&](const int4& dd, int st) {
            const size_t span0 = (size_t)p.hB + (size_t)WB * dd.z + (size_t)9 * dd.x;
            const int par = (int)(span0 & 1), span = WB * dd.w + 9 * dd.y;
            double* sb = s_row0 + st * SM::ROW + par;
            const double* gs = p.H + span0;
            if (p.use_tma) {
                const int body = (span - par) & ~1;
                if (tid == 0) {
                    mbar_expect_tx(bar + st, (uint32_t)body * 8u);
                    bulk_load(sb + par, gs + par, (uint32_t)body * 8u, bar + st);
                }
                if (tid == 32 % TO) {
                    if (par) sb[0] = gs[0];
                    if (par + body < span) sb[span - 1] = gs[span - 1];
                }
            } else {
                for (int i = tid; i < span; i += TO) sb[i] = gs[i];
            }
        };
        int4 d = tiles[t];
        int4 dn = tiles[min(t + G, last)];
        int4 dn2 = tiles[min(t + 2 * G, last)];
        issue(d, 0);
        __syncthreads();                              // the scalar head/tail elements of stage 0 are visible
        // register pipeline as in K1: unconditional loads from clamped addresses, no arithmetic on a value in flight
        int cam, ptg, lo, ncam, nptg, nlo;
        { const int j = d.z + min(tid, max(d.w - 1, 0)); cam = p.obs_cam[j]; ptg = p.obs_pt[j]; lo = p.obs_start[d.x + min(tid, d.y)]; }
        { const int j = dn.z + min(tid, max(dn.w - 1, 0)); ncam = p.obs_cam[j]; nptg = p.obs_pt[j]; nlo = p.obs_start[dn.x + min(tid, dn.y)]; }
        double dx[DC], pg[3], pa[6], pX[3];
#pragma unroll
        for (int a = 0; a < DC; ++a) dx[a] = dxc[(size_t)cam * DC + a];
        {
            const size_t pt = (size_t)d.x + min(tid, max(d.y - 1, 0));
#pragma unroll
            for (int i = 0; i < 3; ++i) { pg[i] = p.g[p.gB + 3 * pt + i]; pX[i] = pts[3 * pt + i]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) pa[i] = Ainv[6 * pt + i];
        }
        for (int it = 0;; ++it) {
            const bool has_next = t + G < p.ntiles;
            const int4 dn3 = tiles[min(t + 3 * G, last)];
            const int pt0 = d.x, npt = d.y, ob0 = d.z, nob = d.w;
            const size_t span0 = (size_t)p.hB + (size_t)WB * ob0 + (size_t)9 * pt0;
            const int st = it & 1;
            const double* s_base = s_row0 + st * SM::ROW + (int)(span0 & 1);
            if (has_next) issue(dn, st ^ 1);          // the other stage was released by the barrier that ended iteration it-1
            if (tid <= npt) s_ost[tid] = lo - ob0;
            if (p.use_tma) mbar_wait(bar + st, (uint32_t)((it >> 1) & 1));
            else __syncthreads();
            if (tid < nob) {
                const int pl = ptg - pt0;
                double w[WB];
                load_wblock<WB>(s_base + WB * tid + 9 * pl, w);
                double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
                for (int a = 0; a < DC; ++a) { u0 += w[3 * a] * dx[a]; u1 += w[3 * a + 1] * dx[a]; u2 += w[3 * a + 2] * dx[a]; }
                s_u[tid] = u0; s_u[TO + tid] = u1; s_u[2 * TO + tid] = u2;
            }
            // level-2 loads of tile k+1, level-1 loads of tile k+2
            double ndx[DC], npg[3], npa[6], npX[3];
#pragma unroll
            for (int a = 0; a < DC; ++a) ndx[a] = dxc[(size_t)ncam * DC + a];
            {
                const size_t pt = (size_t)dn.x + min(tid, max(dn.y - 1, 0));
#pragma unroll
                for (int i = 0; i < 3; ++i) { npg[i] = p.g[p.gB + 3 * pt + i]; npX[i] = pts[3 * pt + i]; }
#pragma unroll
                for (int i = 0; i < 6; ++i) npa[i] = Ainv[6 * pt + i];
            }
            int n2cam, n2ptg, n2lo;
            { const int j = dn2.z + min(tid, max(dn2.w - 1, 0)); n2cam = p.obs_cam[j]; n2ptg = p.obs_pt[j]; n2lo = p.obs_start[dn2.x + min(tid, dn2.y)]; }
            __syncthreads();
            if (tid < npt) {
                const int j0 = s_ost[tid], j1 = s_ost[tid + 1];
                double u0 = 0, u1 = 0, u2 = 0;
                for (int j = j0; j < j1; ++j) { u0 += s_u[j]; u1 += s_u[TO + j]; u2 += s_u[2 * TO + j]; }
                const size_t pt = (size_t)pt0 + tid;
                const double r0 = pg[0] - u0, r1 = pg[1] - u1, r2 = pg[2] - u2;
                const double x0 = -(pa[0] * r0 + pa[1] * r1 + pa[2] * r2);
                const double x1 = -(pa[1] * r0 + pa[3] * r1 + pa[4] * r2);
                const double x2 = -(pa[2] * r0 + pa[4] * r1 + pa[5] * r2);
                x[p.gB + 3 * pt] = x0; x[p.gB + 3 * pt + 1] = x1; x[p.gB + 3 * pt + 2] = x2;
                pts_next[3 * pt] = pX[0] + x0;
                pts_next[3 * pt + 1] = pX[1] + x1;
                pts_next[3 * pt + 2] = pX[2] + x2;
                mx = nanmax(mx, nanmax(nanmax(fabs(x0), fabs(x1)), fabs(x2)));
                sq += x0 * x0 + x1 * x1 + x2 * x2;
                gx += pg[0] * x0 + pg[1] * x1 + pg[2] * x2;
                const double* V = s_base + WB * j1 + 9 * tid;
                const double v0 = V[0] * x0 + V[3] * x1 + V[6] * x2;
                const double v1 = V[1] * x0 + V[4] * x1 + V[7] * x2;
                const double v2 = V[2] * x0 + V[5] * x1 + V[8] * x2;
                // x_p' V x_p + 2 x_p' (sum_c W_pc x_c),  with x_c = -dx_c  =>  sum_c W_pc x_c = -u
                xhx += (x0 * v0 + x1 * v1 + x2 * v2) - 2.0 * (x0 * u0 + x1 * u1 + x2 * u2);
            }
            __syncthreads();                           // releases this stage, s_u and s_ost
            if (!has_next) break;
            t += G; d = dn; dn = dn2; dn2 = dn3;
            cam = ncam; ptg = nptg; lo = nlo; ncam = n2cam; nptg = n2ptg; nlo = n2lo;
#pragma unroll
            for (int a = 0; a < DC; ++a) dx[a] = ndx[a];
#pragma unroll
            for (int i = 0; i < 3; ++i) { pg[i] = npg[i]; pX[i] = npX[i]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) pa[i] = npa[i];
        }
    }
    // block reduction of the running statistics (fixed tree), one partial per CTA
    mx = warp_nanmax(mx); sq = warp_sum(sq); xhx = warp_sum(xhx); gx = warp_sum(gx);
    if (lane == 0) { s_red[wid] = mx; s_red[NW + wid] = sq; s_red[2 * NW + wid] = xhx; s_red[3 * NW + wid] = gx; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0, c = 0.0, e = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) { a = nanmax(a, s_red[i]); b += s_red[NW + i]; c += s_red[2 * NW + i]; e += s_red[3 * NW + i]; }
        partials[blockIdx.x] = a;
        partials[pstride + blockIdx.x] = b;
        partials[2 * pstride + blockIdx.x] = c;
        partials[3 * pstride + blockIdx.x] = e;
    }
}

// Camera side of the update: x_c = -dx_c, varnext[c] = update(variables[c], x_c), plus the camera terms of the step
// statistics.  One thread per camera; per CTA partials (n = gridDim.x): [b] max|x_c|, [n + b] sum x_c^2,
// [2n + b] sum x_c' U_c x_c, [3n + b] g_c . x_c.
template <class R>
__global__ void __launch_bounds__(128) cam_update_kernel(DevProblem p, const double* __restrict__ dxc, const double* __restrict__ cams,
                                                         double* __restrict__ cams_next, double* __restrict__ x, double* __restrict__ partials) {
    constexpr int DC = R::DC;
    __shared__ double s_red[16];
    double mx = 0.0, sq = 0.0, xhx = 0.0, gx = 0.0;
    const int cam = blockIdx.x * 128 + threadIdx.x;
    if (cam < p.nA) {
        double xc[DC];
#pragma unroll
        for (int a = 0; a < DC; ++a) {
            xc[a] = -dxc[(size_t)cam * DC + a];
            x[(size_t)cam * DC + a] = xc[a];
            mx = nanmax(mx, fabs(xc[a]));
            sq += xc[a] * xc[a];
            gx += p.g[(size_t)cam * DC + a] * xc[a];
        }
        const double* U = p.H + (size_t)DC * DC * cam;
#pragma unroll
        for (int b = 0; b < DC; ++b) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DC; ++a) s += U[a + DC * b] * xc[a];
            xhx += xc[b] * s;
        }
        R::update_cam(cams + (size_t)cam * R::CS, xc, cams_next + (size_t)cam * R::CS);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    mx = warp_nanmax(mx); sq = warp_sum(sq); xhx = warp_sum(xhx); gx = warp_sum(gx);
    if (lane == 0) { s_red[wid] = mx; s_red[4 + wid] = sq; s_red[8 + wid] = xhx; s_red[12 + wid] = gx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0, e = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a = nanmax(a, s_red[i]); b += s_red[4 + i]; c += s_red[8 + i]; e += s_red[12 + i]; }
        const int G = gridDim.x;
        partials[blockIdx.x] = a; partials[G + blockIdx.x] = b; partials[2 * G + blockIdx.x] = c; partials[3 * G + blockIdx.x] = e;
    }
}

// gathered[r * 6 + i], r < nranks: the six per-try scalars of every rank -> out[0] = sum of the costs, out[1] = NaN-propagating max
// of max|x_p|, out[2..4] = sums, all in rank order (lane i owns scalar i); out[5] = OR of the ranks' "maxtime reached" flags
__global__ void combine_scalars_kernel(const double* __restrict__ gathered, int nranks, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i >= 6) return;
    double v = gathered[i];
    if (i < 5) for (int r = 1; r < nranks; ++r) v = (i == 1) ? nanmax(v, gathered[r * 6 + i]) : v + gathered[r * 6 + i];
    else for (int r = 1; r < nranks; ++r) v = (gathered[r * 6 + i] != 0.0) ? 1.0 : v;   // any rank's clock stops every rank
    out[i] = v;
}

// CTA k reduces partials[k * n .. (k+1) * n) in a fixed order into out[k]; k == 0 is a NaN-propagating max, the others sums
// (the {max|x|, sum x^2, x'Hx, g.x} quadruple of the step statistics).
__global__ void __launch_bounds__(256) reduce_stats_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double s_red[8];
    const double* src = partials + (size_t)blockIdx.x * n;
    double v = 0.0;
    if (blockIdx.x == 0) { for (int i = threadIdx.x; i < n; i += 256) v = nanmax(v, src[i]); v = block_nanmax(v, s_red); }
    else { for (int i = threadIdx.x; i < n; i += 256) v += src[i]; v = block_sum(v, s_red); }
    if (threadIdx.x == 0) out[blockIdx.x] = v;
}

// ---------------------------------------------------------------------------------------------------
// Irregular points: tracks longer than a tile holds and points with several costs on the same camera (the reference accepts both:
// updatesymA! / updateb! just accumulate, src/linearsystem.jl:132-175).  They stay in place in the layout but belong to no tile
// of any tile kernel / Schur plan; one CTA per point does the point pass and the back-substitution straight from global memory,
// and schur_outlier_kernel (schur5.cuh) eliminates them.  Per-point sums keep observation order.
// ---------------------------------------------------------------------------------------------------
constexpr int LONG_THREADS = 256;

template <class R, bool MS = false>
__global__ void __launch_bounds__(LONG_THREADS) lin_point_long_kernel(DevProblem p, const int* __restrict__ long_pts, const double* __restrict__ cams,
                                                                      const double* __restrict__ pts, double* __restrict__ cost_partials) {
    constexpr int DC = R::DC, WB = 3 * DC, NW = LONG_THREADS / 32;
    __shared__ double s_pc[9 * LONG_THREADS];
    __shared__ double s_red[NW];
    __shared__ double s_acc[9];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int pt = long_pts[blockIdx.x];
    const int ob0 = p.obs_start[pt], k = p.obs_start[pt + 1] - ob0;
    double* hrow = p.H + (size_t)p.hB + (size_t)WB * ob0 + (size_t)9 * pt;
    const double X[3] = {pts[(size_t)3 * pt], pts[(size_t)3 * pt + 1], pts[(size_t)3 * pt + 2]};
    const bool fpt = p.fixB != nullptr && p.fixB[pt];
    if (tid < 9) s_acc[tid] = 0.0;
    double c = 0.0;
    for (int base = 0; base < k; base += LONG_THREADS) {
        const int i = base + tid;
        if (i < k) {
            const int j = ob0 + i;
            const int cam = p.obs_cam[j];
            const double2 z = p.obs_z[j];
            double cv[R::NC];
            R::load_cam(cams, cam, cv);
            double r[2], Jc[2][DC], Jp[2][3];
            R::resjac(cv, X, z.x, z.y, r, Jc, Jp);
            const double s = r[0] * r[0] + r[1] * r[1];
            double rho, d1, d2;
            robustifydcost(rk_point<MS>(p, j), s, rho, d1, d2);
            c += 0.5 * rho;
            const bool fcross = fpt || (p.fixA != nullptr && p.fixA[cam]);
            double gc[DC], gp[3];
#pragma unroll
            for (int a = 0; a < DC; ++a) gc[a] = jtr<R>(Jc, r, a);
#pragma unroll
            for (int b = 0; b < 3; ++b) gp[b] = fma(Jp[1][b], r[1], Jp[0][b] * r[0]);
            const double td2 = 2 * d2;
            double* wd = hrow + (size_t)WB * i;
#pragma unroll
            for (int a = 0; a < DC; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    double h = jtj_pc<R>(Jp, Jc, b, a);
                    if (d1 != 1.0) h *= d1;
                    if (d2 != 0.0) h = fma(td2 * gp[b], gc[a], h);
                    wd[b + 3 * a] = fcross ? 0.0 : h;
                }
            int q = 0;
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2)
#pragma unroll
                for (int b = b2; b < 3; ++b) {
                    double h = fma(Jp[1][b], Jp[1][b2], Jp[0][b] * Jp[0][b2]);
                    if (d1 != 1.0) h *= d1;
                    if (d2 != 0.0) h = fma(td2 * gp[b], gp[b2], h);
                    s_pc[9 * tid + (q++)] = fpt ? 0.0 : h;
                }
#pragma unroll
            for (int b = 0; b < 3; ++b) s_pc[9 * tid + 6 + b] = fpt ? 0.0 : ((d1 != 1.0) ? gp[b] * d1 : gp[b]);
        }
        __syncthreads();
        if (tid < 9) {   // element tid of (V_p lower triangle, g_p): this chunk's observations in order
            double v = s_acc[tid];
            const int n = min(LONG_THREADS, k - base);
            for (int j = 0; j < n; ++j) v += s_pc[9 * j + tid];
            s_acc[tid] = v;
        }
        __syncthreads();
    }
    if (tid < 9) {
        double v = s_acc[tid];
        if (fpt) v = (tid == 0 || tid == 3 || tid == 5) ? 1.0 : 0.0;
        if (tid < 6) {
            double* V = hrow + (size_t)WB * k;
            const int pa = (0x854210 >> (4 * tid)) & 15, pb = (0x874630 >> (4 * tid)) & 15;
            V[pa] = v;
            if (pb != pa) V[pb] = v;
        } else {
            p.g[p.gB + (size_t)3 * pt + (tid - 6)] = v;
        }
    }
    c = warp_sum(c);
    if (lane == 0) s_red[wid] = c;
    __syncthreads();
    if (tid == 0) {
        double tsum = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) tsum += s_red[i];
        cost_partials[blockIdx.x] = tsum;
    }
}

// x_p = -A_p^-1 (g_p - sum_c W_pc dx_c), varnext, step statistics: one CTA per irregular point; partials[k * pstride + slot0 + blockIdx.x]
template <int DC>
__global__ void __launch_bounds__(LONG_THREADS) backsub_long_kernel(DevProblem p, const int* __restrict__ long_pts, const double* __restrict__ dxc,
                                                                    const double* __restrict__ Ainv, const double* __restrict__ pts, double* __restrict__ pts_next,
                                                                    double* __restrict__ x, double* __restrict__ partials, int pstride, int slot0) {
    constexpr int WB = 3 * DC;
    __shared__ double s_u[3 * LONG_THREADS];
    __shared__ double s_acc[3];
    const int tid = threadIdx.x;
    const int pt = long_pts[blockIdx.x];
    const int ob0 = p.obs_start[pt], k = p.obs_start[pt + 1] - ob0;
    const double* hrow = p.H + (size_t)p.hB + (size_t)WB * ob0 + (size_t)9 * pt;
    if (tid < 3) s_acc[tid] = 0.0;
    for (int base = 0; base < k; base += LONG_THREADS) {
        const int i = base + tid;
        if (i < k) {
            const int cam = p.obs_cam[ob0 + i];
            const double* w = hrow + (size_t)WB * i;
            double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
            for (int a = 0; a < DC; ++a) { const double d = dxc[(size_t)cam * DC + a]; u0 += w[3 * a] * d; u1 += w[3 * a + 1] * d; u2 += w[3 * a + 2] * d; }
            s_u[tid] = u0; s_u[LONG_THREADS + tid] = u1; s_u[2 * LONG_THREADS + tid] = u2;
        }
        __syncthreads();
        if (tid < 3) {
            double v = s_acc[tid];
            const int n = min(LONG_THREADS, k - base);
            for (int j = 0; j < n; ++j) v += s_u[tid * LONG_THREADS + j];
            s_acc[tid] = v;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double u0 = s_acc[0], u1 = s_acc[1], u2 = s_acc[2];
        const double* pg = p.g + p.gB + (size_t)3 * pt;
        const double* pa = Ainv + (size_t)6 * pt;
        const double r0 = pg[0] - u0, r1 = pg[1] - u1, r2 = pg[2] - u2;
        const double x0 = -(pa[0] * r0 + pa[1] * r1 + pa[2] * r2);
        const double x1 = -(pa[1] * r0 + pa[3] * r1 + pa[4] * r2);
        const double x2 = -(pa[2] * r0 + pa[4] * r1 + pa[5] * r2);
        x[p.gB + (size_t)3 * pt] = x0; x[p.gB + (size_t)3 * pt + 1] = x1; x[p.gB + (size_t)3 * pt + 2] = x2;
        pts_next[(size_t)3 * pt] = pts[(size_t)3 * pt] + x0;
        pts_next[(size_t)3 * pt + 1] = pts[(size_t)3 * pt + 1] + x1;
        pts_next[(size_t)3 * pt + 2] = pts[(size_t)3 * pt + 2] + x2;
        const double* V = hrow + (size_t)WB * k;
        const double v0 = V[0] * x0 + V[3] * x1 + V[6] * x2;
        const double v1 = V[1] * x0 + V[4] * x1 + V[7] * x2;
        const double v2 = V[2] * x0 + V[5] * x1 + V[8] * x2;
        const int o = slot0 + blockIdx.x;
        partials[o] = nanmax(nanmax(fabs(x0), fabs(x1)), fabs(x2));
        partials[pstride + o] = x0 * x0 + x1 * x1 + x2 * x2;
        partials[2 * pstride + o] = (x0 * v0 + x1 * v1 + x2 * v2) - 2.0 * (x0 * u0 + x1 * u1 + x2 * u2);
        partials[3 * pstride + o] = pg[0] * x0 + pg[1] * x1 + pg[2] * x2;
    }
}

}  // namespace nlls

#include "schur5.cuh"
#include "singles.cuh"
#include "iterators.cuh"
