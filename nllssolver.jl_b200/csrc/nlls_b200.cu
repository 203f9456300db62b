// libnlls_b200 — C ABI implementation (see include/nlls_b200.h for the reference call each symbol replaces).
// Host logic mirrors src/optimize.jl:109-180 and src/iterators.jl:139-172; every numeric operation runs in the
// sm_100a kernels of kernels.cuh.  There is no CPU compute path: without a CUDA device nlls_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nlls_b200.h"
#include "kernels.cuh"
#include "adaptive.cuh"
#include "reduced_plan.hpp"

using namespace nlls;

// ---------------------------------------------------------------------------------------------------
// NCCL, resolved lazily with dlopen so that single-GPU use has no NCCL dependency.
// ---------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSum = 0, ncclMax = 2 };
enum { ncclFloat64 = 8 };
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return false;
        GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
        CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
        Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
        AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
        GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && Broadcast && AllGather && GroupStart && GroupEnd;
    }
};
NcclApi g_nccl;

inline uint64_t now_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct VarSet {
    int vartype = 0;
    int nstore = 0;
    std::vector<int64_t> gidx;  // 1-based positions in problem.variables, ascending
    std::vector<double> vals;   // n x nstore
    bool stale = false;         // device copy is newer than `vals` (values were refreshed straight from the caller's buffer)
};

// scalar slots in the device/pinned scalar buffer
enum Scal {
    SC_COST_LIN = 0, SC_COST_TRY = 1,
    SC_P_MAX = 2, SC_P_SQ = 3, SC_P_XHX = 4, SC_P_GX = 5,      // point-row terms (summed / maxed over ranks)
    SC_TIMEUP = 6,                                             // multi-rank: OR of the ranks' "maxtime reached" flags, rides on the try's all-gather
    SC_C_MAX = 7, SC_C_SQ = 8, SC_C_XHX = 9, SC_C_GX = 10,     // camera terms (replicated)
    SC_MAXDIAG = 11, SC_INFO = 12, SC_EXCH = 13 /* .. 16: host values exchanged between ranks */, SC_COUNT = 17,
    SC_TRY_N = 6                                               // scalars per rank in the try's all-gather: SC_COST_TRY .. SC_TIMEUP
};
}  // namespace

struct nlls_ctx {
    int device = 0;
    cudaStream_t st = nullptr, st2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_b0 = nullptr, ev_b1 = nullptr, ev_c0 = nullptr, ev_c1 = nullptr, ev_g0 = nullptr, ev_g1 = nullptr;
    bool grad_pending = false;       // a re-linearisation was enqueued without a host sync: its time (ev_g0 .. ev_g1) is booked at the next sync
    std::string err;
    int64_t launches = 0;

    // ---- definition
    std::map<int, VarSet> vars;
    int restype = 0;
    int robust = 0;
    double kparams[2] = {0, 1};
    std::vector<int64_t> h_cam_g, h_pt_g;  // global 1-based indices per cost, storage order
    std::vector<double> h_z;               // 2 per cost
    // further cost sets of the same residual struct with their own robust kernel (nlls_add_costs): set id per cost, kernels per set
    struct SetSpec { int robust; double kp[2]; };
    std::vector<SetSpec> sets;
    std::vector<unsigned char> h_set;
    unsigned char *d_obs_set = nullptr, *d_cm_set = nullptr;
    RobustParams* d_rk_tab = nullptr;
    bool prepared = false;
    std::vector<unsigned char> h_unfixed;   // optimize!(…, unfixed): by 0-based variable position; empty = all unfixed
    bool masked = false;                    // some variable of this problem is fixed
    std::vector<unsigned char> h_fixA, h_fixB;
    unsigned char *d_fixA = nullptr, *d_fixB = nullptr;

    // ---- layout (host copies kept for read-back)
    int vtA = 0, vtB = 0, DC = 0, NC = 0, CS = 0, BS = 3;   // BS: stored doubles per variable of the second class (3: points, 1: scalar means)
    // adaptive-kernel problems (NLLS_RES_ADAPTIVE_OFFSET): one ContaminatedGaussian variable + scalar means, small dense system
    bool adaptive = false;
    int64_t kernel_var = 0;
    int ad_nchunks = 0;
    std::vector<int> h_ad_moff;
    int ad_koff = 0;
    double* d_ad_data = nullptr;
    int4* d_ad_chunks = nullptr;
    int* d_ad_moff = nullptr;
    double* d_ad_part = nullptr;
    unsigned char* d_fixdof = nullptr;   // adaptive problems under an unfixed mask: per DoF 1 = fixed
    double* d_em = nullptr;              // EM refit: [0..3] state (old parameters, done flag), [4..7] sums, then 4 partials per chunk
    int64_t nA = 0, nB = 0, nobs = 0, dof = 0, hlen = 0, nred = 0;
    bool cams_first = true;
    std::vector<int> h_obs_cam, h_obs_pt, h_obs_start, h_tile_pt, h_cam_start, h_cm_obs;
    int ntiles = 0, nitems = 0;

    // ---- device
    int *d_obs_cam = nullptr, *d_obs_pt = nullptr, *d_obs_start = nullptr, *d_tile_pt = nullptr;
    double2* d_obs_z = nullptr;
    int *d_cm_pt = nullptr, *d_item_cam = nullptr, *d_item_beg = nullptr, *d_item_end = nullptr, *d_cam_item_start = nullptr;
    double2* d_cm_z = nullptr;
    double *d_A[3] = {nullptr, nullptr, nullptr}, *d_B[3] = {nullptr, nullptr, nullptr};
    int cur = 0, nxt = 1, bst = 2;
    double *d_H = nullptr, *d_g = nullptr, *d_x = nullptr, *d_Ainv = nullptr, *d_S = nullptr, *d_rhs = nullptr;
    double *d_cost_part = nullptr, *d_step_part = nullptr, *d_cam_part = nullptr, *d_cam_part2 = nullptr, *d_camstat_part = nullptr;
    int cam_part_vars = -1;          // variable buffer (index into d_A / d_B) whose camera-pass partials are in d_cam_part, -1: none
    int cost_pointmajor = 0;
    double *d_scal = nullptr, *h_scal = nullptr, *d_gather = nullptr;
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    int use_tma = 1;
    int tile_obs = 0;               // observations per point tile = threads per CTA of the tile kernels: the smallest of 64 / 128 / 256 that
                                    // holds the longest track (NLLS_B200_TILE overrides)
    int tile_env = 0;
    int4* d_tiles = nullptr;        // (pt0, npt, ob0, nob) per point tile
    int nsm = 148, lin_grid = 0, cost_grid = 0, bs_grid = 0;
    int schur_stride = 296;
    // reduced camera system: tile-sparse level-scheduled LDL' (the only storage; a dense reduced system is the same code with every tile present)
    const int s_tiled = 1;
    int NT = 0;
    int64_t ntiles_alloc = 0;
    // multi-rank exchange of the reduced system by tile ownership (nlls_prepare): tiles only this rank's points touch ("exclusive") are
    // stored rank by rank in blocks of xg_block tiles and travel in ONE all-gather; tiles several ranks touch are summed with an
    // all-reduce; fill-only tiles are zero everywhere and are not exchanged
    int64_t xg_block = 0, xg_shared0 = 0, xg_nshared = 0;
    int* d_add_u = nullptr;          // [NT] natural camera tile: this rank adds U_c + lambda I to its diagonal tile
    int red_levels = 0;
    double red_flops = 0.0;          // algorithmic FP64 operations of one reduced solve (tile LDL' + sweeps), counted at prepare
    struct RedLaunch { int kind, off, cnt; };          // kind 0: diagonal tiles of a level, 1: its off-diagonal tiles, 2: its updates
    std::vector<RedLaunch> fact_launches;
    RedTask* d_red_tasks = nullptr;
    RedUpd* d_red_upds = nullptr;
    RedTarget* d_red_targets = nullptr;
    std::vector<std::pair<int, int>> lvl_cols;           // (offset, count) into d_lvl_cols per level
    int *d_tile_id = nullptr, *d_pos = nullptr, *d_diag_tile = nullptr, *d_diag_tile_nat = nullptr, *d_lvl_cols = nullptr;
    int *d_colptr = nullptr, *d_col_tile = nullptr, *d_col_row = nullptr;
    double *d_Linv = nullptr, *d_xp = nullptr;
    int bwd_flow = 2;                // backward sweep: 2 = one dataflow launch, everything column-independent staged before the first wait (default);
                                     // NLLS_B200_BWD=flow: first dataflow version (measured equal to per-level launches: 0.401 vs 0.396 ms per reduced
                                     // solve on the Venice shape — the dependent chain inside a column is the cost, not the launches); =levels: one launch per level
    int *d_bwd_order = nullptr, *d_bwd_flags = nullptr, *d_nat_of_pos = nullptr;
    // Schur v2 plan (per-tile sorted contribution lists)
    int schur_v2 = 1;
    int nstiles = 0;
    int *d_stile_pt = nullptr, *d_chunk_off = nullptr, *d_ent_off = nullptr;
    // Schur v4 plan (super-tiles: tensor-core accumulation of whole S blocks across consecutive tiles); 2 = forced
    cudaGraphExec_t red_graph_exec = nullptr;   // the reduced solve's launch sequence (tile-sparse path)
    int use_graph = 1, red_graph_launches = 0;
    int use_pdl = 1;                 // programmatic dependent launch along the reduced solve's kernel chain (NLLS_B200_PDL=0: plain stream order)
    int schur_v4 = 1, nsuper = 0;   // nsuper: CTAs of the v4 kernel (0: v2 path)
    // Schur v5 plan (window-aligned register accumulation, schur5.cuh); 1: automatic, 2: forced, 0: off.  n5cta: CTAs (0: not in use)
    int schur_v5 = 1, n5cta = 0, nout_pts = 0, s5_ncons = S5_CONSUMERS;
    int* d5_cta_item = nullptr;
    Schur5Item* d5_items = nullptr;
    unsigned int* d5_blob = nullptr;
    long long* d5_ftab = nullptr;
    int* d_out_pts = nullptr;       // points outside the Schur plans (schur_outlier_kernel): the v5 plan's own outliers, then from first_irr on the irregular points
    // irregular points (tracks longer than a tile holds, several costs on one camera): no tile of any kernel / plan holds them
    std::vector<unsigned char> h_irr;
    int nlong = 0, first_irr = 0;
    bool has_dups = false;
    int* d_long_pts = nullptr;
    int* d_cta_item = nullptr;
    SchurItem* d_items = nullptr;
    SchurUnit* d_units = nullptr;
    unsigned int *d_wtab = nullptr, *d_blob = nullptr;
    SchurChunk* d_chunks = nullptr;
    unsigned int* d_ents = nullptr;

    // ---- multi-GPU
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;

    // ---- LM state (src/iterators.jl:120-126, src/optimize.jl:109-121)
    nlls_options opts{};
    double lambda = 0.0, bestcost = 0.0, cost = 0.0, startcost = 0.0, maxstep = 0.0;
    int64_t fails = 0, iternum = 0, converged = 0;
    int64_t costcomputations = 0, gradientcomputations = 0, linearsolvers = 0;
    uint64_t starttime = 0, stoptime = 0, t_init = 0, t_cost = 0, t_grad = 0, t_solver = 0;
    bool lm_active = false;
    bool have_best = false;
    // Dogleg / gradient descent (src/iterators.jl:30-46,177-185)
    double trustradius = 0.0, stepsize = 1.0;
    double *d_cauchy = nullptr, *d_vec_part = nullptr;
    int lm_phase = 0;   // 0: nlls_lm_iterate is next, 1: nlls_lm_advance is next
};

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                   \
            return NLLS_ERR_CUDA;                                                                             \
        }                                                                                                     \
    } while (0)
#define CKN(call)                                                                                             \
    do {                                                                                                      \
        int r__ = (call);                                                                                     \
        if (r__ != 0) {                                                                                       \
            ctx->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
            return NLLS_ERR_NCCL;                                                                             \
        }                                                                                                     \
    } while (0)
#define FAIL(code, msg)        \
    do {                       \
        ctx->err = (msg);      \
        return (code);         \
    } while (0)
#define TRY(call)                     \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != NLLS_OK) return rc__; \
    } while (0)

namespace {

template <class T>
int upload(nlls_ctx* ctx, T** dptr, const std::vector<T>& h) {
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CK(cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) CK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return NLLS_OK;
}
template <class T>
int dalloc(nlls_ctx* ctx, T** dptr, size_t n) {
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    CK(cudaMalloc((void**)dptr, std::max<size_t>(n, 1) * sizeof(T)));
    return NLLS_OK;
}

DevProblem devproblem(const nlls_ctx* c) {
    DevProblem p;
    p.obs_cam = c->d_obs_cam; p.obs_pt = c->d_obs_pt; p.obs_z = c->d_obs_z; p.obs_start = c->d_obs_start;
    p.tile_pt = c->d_tile_pt; p.ntiles = c->ntiles; p.nA = (int)c->nA; p.nB = (int)c->nB; p.nobs = (int)c->nobs;
    p.cm_pt = c->d_cm_pt; p.cm_z = c->d_cm_z; p.item_cam = c->d_item_cam; p.item_beg = c->d_item_beg; p.item_end = c->d_item_end;
    p.cam_item_start = c->d_cam_item_start; p.nitems = c->nitems;
    p.H = c->d_H; p.g = c->d_g; p.hB = (long long)c->DC * c->DC * c->nA; p.gB = (long long)c->DC * c->nA;
    p.rk.kind = c->robust & 15; p.rk.scaled = (c->robust & NLLS_ROBUST_SCALED) ? 1 : 0;
    p.rk.width = c->kparams[0]; p.rk.width2 = c->kparams[0] * c->kparams[0]; p.rk.height = c->kparams[1];
    p.use_tma = c->use_tma;
    const bool multi = c->sets.size() > 1;
    p.obs_set = multi ? c->d_obs_set : nullptr; p.cm_set = multi ? c->d_cm_set : nullptr; p.rk_tab = multi ? c->d_rk_tab : nullptr;
    p.schur_stride = c->schur_stride;
    p.tile_id = c->d_tile_id; p.tile_pos = c->d_pos; p.NT = c->NT; p.s_tiled = c->s_tiled;
    p.fixA = c->masked ? c->d_fixA : nullptr; p.fixB = c->masked ? c->d_fixB : nullptr;
    return p;
}

int vartype_nstore(int vt) {
    switch (vt) {
        case NLLS_VAR_SCALAR: return 1;
        case NLLS_VAR_EUCLID3: return 3;
        case NLLS_VAR_EUCLID6: return 6;
        case NLLS_VAR_CONTAMGAUSS: return 3;
        case NLLS_VAR_PINHOLE: return 15;
    }
    return 0;
}

// pack host variables (nstore per variable) into the device stride and upload to buffer `which`
int upload_vars(nlls_ctx* ctx, const VarSet& vs, int dstride, double* dst) {
    const int64_t n = (int64_t)vs.gidx.size();
    if (dstride == vs.nstore) {
        CK(cudaMemcpyAsync(dst, vs.vals.data(), sizeof(double) * n * dstride, cudaMemcpyHostToDevice, ctx->st));
    } else {
        std::vector<double> tmp((size_t)n * dstride, 0.0);
        for (int64_t i = 0; i < n; ++i) std::memcpy(&tmp[(size_t)i * dstride], &vs.vals[(size_t)i * vs.nstore], sizeof(double) * vs.nstore);
        CK(cudaMemcpyAsync(dst, tmp.data(), sizeof(double) * n * dstride, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    return NLLS_OK;
}

// ---- kernel launch helpers, dispatched on the registered residual type --------------------------------
template <class R, int TO>
int set_tile_attrs(nlls_ctx* ctx) {
    constexpr int TP = TO / 2;
    CK(cudaFuncSetAttribute(lin_point_kernel<R, TO, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LinSmem<R, TO, TP>::bytes));
    CK(cudaFuncSetAttribute(lin_point_kernel<R, TO, TP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LinSmem<R, TO, TP>::bytes));
    CK(cudaFuncSetAttribute(backsub_kernel<R::DC, TO, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BacksubSmem<R::DC, TO, TP>::bytes));
    int occ = 1;
    auto grid_for = [&](int o, const char* env) {
        int g = std::max(1, std::min(ctx->ntiles, ctx->nsm * std::max(1, o)));
        if (const char* e = getenv(env)) g = std::max(1, std::min(ctx->ntiles, atoi(e)));
        return g;
    };
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lin_point_kernel<R, TO, TP>, TO, LinSmem<R, TO, TP>::bytes));
    ctx->lin_grid = grid_for(occ, "NLLS_B200_LIN_GRID");
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cost_kernel<R, TO>, TO, 0));
    ctx->cost_grid = grid_for(std::min(occ, 1024 / TO), "NLLS_B200_COST_GRID");
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, backsub_kernel<R::DC, TO, TP>, TO, BacksubSmem<R::DC, TO, TP>::bytes));
    ctx->bs_grid = grid_for(occ, "NLLS_B200_BS_GRID");
    return NLLS_OK;
}

template <class R>
int set_smem_attrs(nlls_ctx* ctx) {
    if (ctx->tile_obs == 64) TRY((set_tile_attrs<R, 64>(ctx)));
    else if (ctx->tile_obs == 128) TRY((set_tile_attrs<R, 128>(ctx)));
    else TRY((set_tile_attrs<R, 256>(ctx)));
    CK(cudaFuncSetAttribute(schur2_kernel<R::DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Schur2Smem<R::DC>::bytes));
    CK(cudaFuncSetAttribute(schur4_kernel<R::DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Schur4Cfg<R::DC>::bytes));
    CK(cudaFuncSetAttribute(schur5_kernel<R::DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Schur5Smem<R::DC>::bytes));
    CK(cudaFuncSetAttribute(ldl_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    CK(cudaFuncSetAttribute(ldl_off_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OFF_SMEM));
    CK(cudaFuncSetAttribute(ldl_upd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM));
    CK(cudaFuncSetAttribute(ldl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    CK(cudaFuncSetAttribute(ldl_bwd_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    CK(cudaFuncSetAttribute(ldl_bwd_flow2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD2_SMEM));
    return NLLS_OK;
}

int allreduce(nlls_ctx* ctx, double* buf, size_t count, int op) {
    if (ctx->nranks <= 1) return NLLS_OK;
    CKN(g_nccl.AllReduce(buf, buf, count, ncclFloat64, op, ctx->comm, ctx->st));
    return NLLS_OK;
}

// do_cam: 1 = camera pass (lin_cam on the second stream) + finalize, 2 = finalize only: the work-item partials in d_cam_part were
// produced by the cost evaluation of the accepted try at exactly these variables (launch_cost)
template <class R>
int launch_linearize(nlls_ctx* ctx, bool do_point = true, int do_cam = 1) {
    DevProblem p = devproblem(ctx);
    constexpr int NU = R::DC * (R::DC + 1) / 2 + R::DC;
    const bool ms = ctx->sets.size() > 1;   // several cost sets: the kernels look the robust kernel up per observation
    if (do_cam) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->st));
        CK(cudaStreamWaitEvent(ctx->st2, ctx->ev_fork, 0));
        if (do_cam == 1 && ctx->nitems > 0) {
            if (ms) lin_cam_kernel<R, true><<<ctx->nitems, 256, 0, ctx->st2>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_cam_part);
            else lin_cam_kernel<R><<<ctx->nitems, 256, 0, ctx->st2>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_cam_part);
            ctx->launches++;
        }
        const int tot = (int)ctx->nA * NU;
        cam_finalize_kernel<R::DC><<<(tot + 255) / 256, 256, 0, ctx->st2>>>(p, ctx->d_cam_part); ctx->launches++;
        CK(cudaEventRecord(ctx->ev_join, ctx->st2));
    }
    if (do_point && ctx->ntiles > 0) {
        const double *cA = ctx->d_A[ctx->cur], *cB = ctx->d_B[ctx->cur];
        if (ctx->tile_obs == 64) {
            if (ms) lin_point_kernel<R, 64, 32, true><<<ctx->lin_grid, 64, LinSmem<R, 64, 32>::bytes, ctx->st>>>(p, ctx->d_tiles, cA, cB, ctx->d_cost_part);
            else lin_point_kernel<R, 64, 32><<<ctx->lin_grid, 64, LinSmem<R, 64, 32>::bytes, ctx->st>>>(p, ctx->d_tiles, cA, cB, ctx->d_cost_part);
        } else if (ctx->tile_obs == 128) {
            if (ms) lin_point_kernel<R, 128, 64, true><<<ctx->lin_grid, 128, LinSmem<R, 128, 64>::bytes, ctx->st>>>(p, ctx->d_tiles, cA, cB, ctx->d_cost_part);
            else lin_point_kernel<R, 128, 64><<<ctx->lin_grid, 128, LinSmem<R, 128, 64>::bytes, ctx->st>>>(p, ctx->d_tiles, cA, cB, ctx->d_cost_part);
        } else {
            if (ms) lin_point_kernel<R, 256, 128, true><<<ctx->lin_grid, 256, LinSmem<R, 256, 128>::bytes, ctx->st>>>(p, ctx->d_tiles, cA, cB, ctx->d_cost_part);
            else lin_point_kernel<R, 256, 128><<<ctx->lin_grid, 256, LinSmem<R, 256, 128>::bytes, ctx->st>>>(p, ctx->d_tiles, cA, cB, ctx->d_cost_part);
        }
        ctx->launches++;
    }
    if (do_point && ctx->nlong > 0) {   // irregular points: one CTA each, cost partials behind the tiles'
        if (ms) lin_point_long_kernel<R, true><<<ctx->nlong, LONG_THREADS, 0, ctx->st>>>(p, ctx->d_long_pts, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_cost_part + ctx->ntiles);
        else lin_point_long_kernel<R><<<ctx->nlong, LONG_THREADS, 0, ctx->st>>>(p, ctx->d_long_pts, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_cost_part + ctx->ntiles);
        ctx->launches++;
    }
    if (do_cam) CK(cudaStreamWaitEvent(ctx->st, ctx->ev_join, 0));
    CK(cudaGetLastError());
    return NLLS_OK;
}

// cost(vars[which], costs)  (src/cost.jl:11): the camera-major pass (lin_cam_kernel), which also leaves the camera blocks of
// vars[which] in `part` (work-item partials).  part == d_cam_part: the LM try — the partials are reused by the re-linearisation if
// the try is accepted; part == d_cam_part2: nlls_cost(), which must not disturb them.  Same kernel, same reduction tree in both
// cases, so cost(problem) == bestcost exactly (test/optimizeba.jl:67).
template <class R>
int launch_cost(nlls_ctx* ctx, int which, int slot, double* part = nullptr, bool reduce = true) {
    DevProblem p = devproblem(ctx);
    constexpr int NU = R::DC * (R::DC + 1) / 2 + R::DC;
    if (!part) part = ctx->d_cam_part;
    const bool ms = ctx->sets.size() > 1;
    if (ctx->cost_pointmajor && ctx->nlong == 0 && !ms) {   // NLLS_B200_COST=tiles: the round-1 cost kernel (point-major tiles)
        if (ctx->ntiles > 0) {
            if (ctx->tile_obs == 64) cost_kernel<R, 64><<<ctx->cost_grid, 64, 0, ctx->st>>>(p, ctx->d_tiles, ctx->d_A[which], ctx->d_B[which], ctx->d_cost_part);
            else if (ctx->tile_obs == 128) cost_kernel<R, 128><<<ctx->cost_grid, 128, 0, ctx->st>>>(p, ctx->d_tiles, ctx->d_A[which], ctx->d_B[which], ctx->d_cost_part);
            else cost_kernel<R, 256><<<ctx->cost_grid, 256, 0, ctx->st>>>(p, ctx->d_tiles, ctx->d_A[which], ctx->d_B[which], ctx->d_cost_part);
            ctx->launches++;
        }
        reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_cost_part, ctx->ntiles, ctx->d_scal + slot, 0); ctx->launches++;
    } else {
        if (ctx->nitems > 0) {
            if (ms) lin_cam_kernel<R, true><<<ctx->nitems, 256, 0, ctx->st>>>(p, ctx->d_A[which], ctx->d_B[which], part);
            else lin_cam_kernel<R><<<ctx->nitems, 256, 0, ctx->st>>>(p, ctx->d_A[which], ctx->d_B[which], part);
            ctx->launches++;
        }
        cam_cost_reduce_kernel<<<1, 1024, 0, ctx->st>>>(part, ctx->nitems, NU + 1, ctx->d_scal + slot); ctx->launches++;
        if (part == ctx->d_cam_part) ctx->cam_part_vars = which;   // the camera blocks of these variables are now in d_cam_part
    }
    CK(cudaGetLastError());
    if (reduce) TRY(allreduce(ctx, ctx->d_scal + slot, 1, ncclSum));
    return NLLS_OK;
}

RedSolveLists redlists(const nlls_ctx* c) {
    RedSolveLists t;
    t.diag_tile = c->d_diag_tile;
    t.colptr = c->d_colptr; t.col_tile = c->d_col_tile; t.col_row = c->d_col_row;
    return t;
}

template <class R>
int launch_schur(nlls_ctx* ctx, double lambda) {
    constexpr int DC = R::DC;
    DevProblem p = devproblem(ctx);
    const int64_t n = ctx->nred;
    const size_t scount = (size_t)ctx->ntiles_alloc * ST2;
    CK(cudaMemsetAsync(ctx->d_S, 0, sizeof(double) * scount, ctx->st));
    red_init_kernel<DC><<<ctx->NT, 256, 0, ctx->st>>>(ctx->d_S, ctx->d_diag_tile_nat, ctx->d_H, ctx->d_g, ctx->d_rhs, (int)ctx->nA, lambda,
                                                        ctx->d_add_u, ctx->rank == 0 ? 1 : 0);
    ctx->launches++;
    if (ctx->nout_pts > 0) CK(cudaEventRecord(ctx->ev_fork, ctx->st));   // S and rhs are initialised: the outlier kernel (second stream) may start adding
    if (ctx->n5cta > 0) {
        Schur5Dev s5;
        s5.cta_item = ctx->d5_cta_item; s5.items = ctx->d5_items; s5.blob = ctx->d5_blob; s5.ftab = ctx->d5_ftab; s5.dbg = nullptr; s5.ncons = ctx->s5_ncons;
        static const bool s5dbg = getenv("NLLS_B200_S5DBG") != nullptr;
        long long* d_dbg = nullptr;
        if (s5dbg) { CK(cudaMalloc((void**)&d_dbg, sizeof(long long) * 64 * ctx->n5cta)); CK(cudaMemsetAsync(d_dbg, 0, sizeof(long long) * 64 * ctx->n5cta, ctx->st)); s5.dbg = d_dbg; }
        schur5_kernel<DC><<<ctx->n5cta, Schur5Cfg<DC>::THREADS, Schur5Smem<DC>::bytes, ctx->st>>>(p, s5, ctx->d_S, ctx->d_rhs, ctx->d_Ainv, lambda);
        ctx->launches++;
        if (s5dbg) {   // development aid: where the warps of the Schur kernel spend their cycles
            std::vector<long long> h((size_t)64 * ctx->n5cta);
            CK(cudaStreamSynchronize(ctx->st));
            CK(cudaMemcpy(h.data(), d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
            cudaFree(d_dbg);
            double wsum[16] = {0}, tsum[16] = {0}, tmax = 0, idle = 0;
            const int pw = Schur5Cfg<DC>::CONS;   // the producer warp
            for (int c = 0; c < ctx->n5cta; ++c) for (int w = 0; w <= pw; ++w) { wsum[w] += (double)h[((size_t)c * 16 + w) * 4]; tsum[w] += (double)h[((size_t)c * 16 + w) * 4 + 1]; tmax = std::max(tmax, (double)h[((size_t)c * 16 + w) * 4 + 1]); }
            for (int c = 0; c < ctx->n5cta; ++c) idle += (double)h[((size_t)c * 16 + pw) * 4 + 2];
            fprintf(stderr, "[nlls] schur5 cycles: longest warp %.0f; per consumer warp (mean over CTAs) total / waiting:", tmax);
            for (int w = 0; w < pw; ++w) fprintf(stderr, " %d: %.0f/%.0f", w, tsum[w] / ctx->n5cta, wsum[w] / ctx->n5cta);
            fprintf(stderr, "; producer total %.0f point phases %.0f idle polls %.0f\n", tsum[pw] / ctx->n5cta, wsum[pw] / ctx->n5cta, idle / ctx->n5cta);
            // per CTA totals (consumer warp 0) to see the balance between CTAs
            double cmin = 1e30, cmax = 0;
            for (int c = 0; c < ctx->n5cta; ++c) { const double t = (double)h[((size_t)c * 16) * 4 + 1]; cmin = std::min(cmin, t); cmax = std::max(cmax, t); }
            fprintf(stderr, "[nlls] schur5 CTA time (consumer 0): min %.0f max %.0f\n", cmin, cmax);
        }
    } else if (ctx->schur_v4 && ctx->nsuper > 0) {
        SchurPlan4 sp;
        sp.cta_item = ctx->d_cta_item; sp.items = ctx->d_items; sp.units = ctx->d_units; sp.blob = ctx->d_blob; sp.wtab = ctx->d_wtab;
        sp.ld = ctx->s_tiled ? ST : n;
        schur4_kernel<DC><<<ctx->nsuper, SCH4_THREADS, Schur4Cfg<DC>::bytes, ctx->st>>>(p, sp, ctx->d_S, ctx->d_rhs, ctx->d_Ainv, lambda);
        ctx->launches++;
    } else if (ctx->schur_v2 && ctx->nstiles > 0) {
        SchurPlan sp;
        sp.stile_pt = ctx->d_stile_pt; sp.chunk_off = ctx->d_chunk_off; sp.ent_off = ctx->d_ent_off; sp.chunks = ctx->d_chunks; sp.ents = ctx->d_ents; sp.nstiles = ctx->nstiles;
        sp.ld = ctx->s_tiled ? ST : n;
        const int G = std::max(1, ctx->schur_stride);
        const int grid = G * ((ctx->nstiles + G - 1) / G);
        schur2_kernel<DC><<<grid, SCH_THREADS, Schur2Smem<DC>::bytes, ctx->st>>>(p, sp, ctx->d_S, ctx->d_rhs, ctx->d_Ainv, lambda);
        ctx->launches++;
    }
    if (ctx->nout_pts > 0) {   // points outside the plans: gaps in the camera list / wide tracks (v5), and the irregular points (their A_p^-1 too)
        // a few hundred warps of work (29 us alone: latency): on the second stream, behind red_init, so that it fills the tail of the
        // persistent Schur kernel (launched first, it holds every SM's registers until its CTAs finish) instead of following it
        static const bool side = getenv("NLLS_B200_OUTLIER_SERIAL") == nullptr;
        cudaStream_t so = side ? ctx->st2 : ctx->st;
        if (side) CK(cudaStreamWaitEvent(ctx->st2, ctx->ev_fork, 0));
        schur_outlier_kernel<DC><<<(ctx->nout_pts * 32 + 127) / 128, 128, 0, so>>>(p, ctx->d_out_pts, ctx->nout_pts, ctx->d_S, ctx->d_rhs, lambda, ctx->first_irr, ctx->d_Ainv);
        ctx->launches++;
        if (side) { CK(cudaEventRecord(ctx->ev_join, ctx->st2)); CK(cudaStreamWaitEvent(ctx->st, ctx->ev_join, 0)); }
    }
    CK(cudaGetLastError());
    if (ctx->nranks > 1) {
        CKN(g_nccl.GroupStart());
        if (ctx->xg_shared0 > 0 || ctx->xg_nshared > 0) {   // exchange by tile ownership (nlls_prepare)
            const size_t blk = (size_t)ctx->xg_block * ST2;
            if (blk > 0) CKN(g_nccl.AllGather(ctx->d_S + blk * (size_t)ctx->rank, ctx->d_S, blk, ncclFloat64, ctx->comm, ctx->st));
            if (ctx->xg_nshared > 0)
                CKN(g_nccl.AllReduce(ctx->d_S + (size_t)ctx->xg_shared0 * ST2, ctx->d_S + (size_t)ctx->xg_shared0 * ST2, (size_t)ctx->xg_nshared * ST2, ncclFloat64, ncclSum,
                                     ctx->comm, ctx->st));
        } else {
            CKN(g_nccl.AllReduce(ctx->d_S, ctx->d_S, scount, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        }
        CKN(g_nccl.AllReduce(ctx->d_rhs, ctx->d_rhs, (size_t)((int64_t)ctx->NT * ST), ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.GroupEnd());
    }
    return NLLS_OK;
}

int launch_reduced_solve(nlls_ctx* ctx) {
    {
        const RedSolveLists t = redlists(ctx);
        const int nx = ctx->NT * ST;
        // ~40 short dependent launches with constant arguments: captured once into a CUDA graph and replayed (the launch gaps of a
        // plain stream are a quarter of this phase; NLLS_B200_GRAPH=0 keeps plain launches)
        const bool pdl = ctx->use_pdl != 0;
        auto go = [&](auto kernel, int grid, int block, size_t smem, auto... args) -> cudaError_t {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
            return cudaLaunchKernelEx(&cfg, kernel, args...);
        };
        auto enqueue = [&](int& nl) -> int {
            CK(go(red_permute_kernel, (nx + 255) / 256, 256, 0, (const double*)ctx->d_rhs, ctx->d_xp, (const int*)ctx->d_pos, ctx->NT, 1, ctx->bwd_flow == 2 ? ctx->d_bwd_flags : (int*)nullptr)); ++nl;
            for (const auto& l : ctx->fact_launches) {   // factorisation + forward substitution (fused into the diagonal tasks)
                if (l.kind == 0) CK(go(ldl_diag_kernel, l.cnt, DIAG_THREADS, DIAG_SMEM, ctx->d_S, ctx->d_Linv, (const RedTask*)(ctx->d_red_tasks + l.off), ctx->d_xp));
                else if (l.kind == 1) CK(go(ldl_off_kernel, GEMM_CTAS * l.cnt, GEMM_THREADS, OFF_SMEM, ctx->d_S, (const double*)ctx->d_Linv, (const RedTask*)(ctx->d_red_tasks + l.off), ctx->d_xp));
                else CK(go(ldl_upd_kernel, GEMM_CTAS * l.cnt, GEMM_THREADS, UPD_SMEM, ctx->d_S, (const RedUpd*)ctx->d_red_upds, (const RedTarget*)(ctx->d_red_targets + l.off), ctx->d_xp));
                ++nl;
            }
            if (ctx->bwd_flow == 2) {   // one dataflow launch; also writes the solution in natural numbering (no trailing permutation)
                CK(go(ldl_bwd_flow2_kernel, std::min(ctx->NT, ctx->nsm), BW_THREADS, BWD2_SMEM, (const double*)ctx->d_S, (const double*)ctx->d_Linv, t, (const int*)ctx->d_bwd_order,
                      (const int*)ctx->d_nat_of_pos, ctx->NT, ctx->d_bwd_flags, ctx->d_xp, ctx->d_rhs));
                ++nl;
                return NLLS_OK;
            } else if (ctx->bwd_flow) {   // first dataflow version
                CK(cudaMemsetAsync(ctx->d_bwd_flags, 0, sizeof(int) * ctx->NT, ctx->st));
                ldl_bwd_flow_kernel<<<std::min(ctx->NT, ctx->nsm), RED_THREADS, BWD_SMEM, ctx->st>>>(ctx->d_S, ctx->d_Linv, t, ctx->d_bwd_order, ctx->NT, ctx->d_bwd_flags, ctx->d_xp);
                ++nl;
            } else
            for (size_t l = ctx->lvl_cols.size(); l-- > 0;) {
                CK(go(ldl_bwd_kernel, ctx->lvl_cols[l].second, RED_THREADS, BWD_SMEM, (const double*)ctx->d_S, (const double*)ctx->d_Linv, t, (const int*)(ctx->d_lvl_cols + ctx->lvl_cols[l].first), ctx->d_xp));
                ++nl;
            }
            CK(go(red_permute_kernel, (nx + 255) / 256, 256, 0, (const double*)ctx->d_xp, ctx->d_rhs, (const int*)ctx->d_pos, ctx->NT, 0, (int*)nullptr)); ++nl;
            return NLLS_OK;
        };
        if (ctx->use_graph && !ctx->red_graph_exec) {
            cudaGraph_t g = nullptr;
            int nl = 0;
            CK(cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal));
            enqueue(nl);
            CK(cudaStreamEndCapture(ctx->st, &g));
            CK(cudaGraphInstantiate(&ctx->red_graph_exec, g, 0));
            CK(cudaGraphDestroy(g));
            ctx->red_graph_launches = nl;
        }
        if (ctx->red_graph_exec) {
            CK(cudaGraphLaunch(ctx->red_graph_exec, ctx->st));
            ctx->launches += ctx->red_graph_launches;
        } else {
            int nl = 0;
            enqueue(nl);
            ctx->launches += nl;
        }
        CK(cudaGetLastError());
        // every rank factors the same all-reduced system with owner-written updates (no reductions): the camera replicas stay
        // bit-identical without an exchange
        return NLLS_OK;
    }
}

template <class R>
int launch_update(nlls_ctx* ctx) {
    constexpr int DC = R::DC;
    DevProblem p = devproblem(ctx);
    const int bg = ctx->ntiles > 0 ? ctx->bs_grid : 0;
    const int ps = bg + ctx->nlong;      // step-statistics partials per quantity: the tile kernel's CTAs, then one per irregular point
    if (ctx->ntiles > 0) {
        if (ctx->tile_obs == 64)
            backsub_kernel<DC, 64, 32><<<bg, 64, BacksubSmem<DC, 64, 32>::bytes, ctx->st>>>(p, ctx->d_tiles, ctx->d_rhs, ctx->d_Ainv, ctx->d_B[ctx->cur],
                                                                                             ctx->d_B[ctx->nxt], ctx->d_x, ctx->d_step_part, ps);
        else if (ctx->tile_obs == 128)
            backsub_kernel<DC, 128, 64><<<bg, 128, BacksubSmem<DC, 128, 64>::bytes, ctx->st>>>(p, ctx->d_tiles, ctx->d_rhs, ctx->d_Ainv, ctx->d_B[ctx->cur],
                                                                                                ctx->d_B[ctx->nxt], ctx->d_x, ctx->d_step_part, ps);
        else
            backsub_kernel<DC, 256, 128><<<bg, 256, BacksubSmem<DC, 256, 128>::bytes, ctx->st>>>(p, ctx->d_tiles, ctx->d_rhs, ctx->d_Ainv, ctx->d_B[ctx->cur],
                                                                                                  ctx->d_B[ctx->nxt], ctx->d_x, ctx->d_step_part, ps);
        ctx->launches++;
    }
    if (ctx->nlong > 0) {
        backsub_long_kernel<DC><<<ctx->nlong, LONG_THREADS, 0, ctx->st>>>(p, ctx->d_long_pts, ctx->d_rhs, ctx->d_Ainv, ctx->d_B[ctx->cur], ctx->d_B[ctx->nxt], ctx->d_x,
                                                                         ctx->d_step_part, ps, bg);
        ctx->launches++;
    }
    const int cg = (int)((ctx->nA + 127) / 128);
    cam_update_kernel<R><<<cg, 128, 0, ctx->st>>>(p, ctx->d_rhs, ctx->d_A[ctx->cur], ctx->d_A[ctx->nxt], ctx->d_x, ctx->d_camstat_part); ctx->launches++;
    reduce_stats_kernel<<<4, 256, 0, ctx->st>>>(ctx->d_step_part, ps, ctx->d_scal + SC_P_MAX); ctx->launches++;
    reduce_stats_kernel<<<4, 256, 0, ctx->st>>>(ctx->d_camstat_part, cg, ctx->d_scal + SC_C_MAX); ctx->launches++;
    CK(cudaGetLastError());
    return NLLS_OK;
}

// The rank-local scalars of a try — cost(varnext) and the point rows' step statistics {max|x|, sum x^2, x'Hx, g.x}, five
// consecutive slots — are combined over the ranks with ONE all-gather and a tiny kernel that adds them in rank order (the same
// order on every rank, so the replicated accept / reject decision is taken on identical bits).  Round 1 used three all-reduces
// per try here; at 8 ranks their launch latency was a quarter of the step.  A sixth slot carries the rank's "maxtime reached" flag:
// the one termination input that is not replicated (each rank has its own clock) is OR-ed over the ranks, so all of them stop at the
// same iteration, with no collective of its own (a rank that left the loop alone would strand the others in the next all-reduce).
int exchange_try_scalars(nlls_ctx* ctx) {
    if (ctx->nranks <= 1) return NLLS_OK;
    ctx->h_scal[SC_COUNT] = (ctx->lm_active && now_ns() > ctx->stoptime) ? 1.0 : 0.0;   // pinned staging slot outside the read-back range
    CK(cudaMemcpyAsync(ctx->d_scal + SC_TIMEUP, ctx->h_scal + SC_COUNT, sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    CKN(g_nccl.AllGather(ctx->d_scal + SC_COST_TRY, ctx->d_gather, SC_TRY_N, ncclFloat64, ctx->comm, ctx->st));
    combine_scalars_kernel<<<1, 32, 0, ctx->st>>>(ctx->d_gather, ctx->nranks, ctx->d_scal + SC_COST_TRY); ctx->launches++;
    CK(cudaGetLastError());
    return NLLS_OK;
}

template <class R>
int launch_maxdiag(nlls_ctx* ctx) {
    DevProblem p = devproblem(ctx);
    CK(cudaMemsetAsync(ctx->d_scal + SC_MAXDIAG, 0, sizeof(double), ctx->st));
    const long long nv = (long long)ctx->nA + ctx->nB;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nv + 255) / 256, 8LL * ctx->nsm));
    maxdiag_kernel<R::DC><<<grid, 256, 0, ctx->st>>>(p, (unsigned long long*)(ctx->d_scal + SC_MAXDIAG)); ctx->launches++;
    CK(cudaGetLastError());
    TRY(allreduce(ctx, ctx->d_scal + SC_MAXDIAG, 1, ncclMax));
    return NLLS_OK;
}


// ---- adaptive-kernel problems: small dense system, pure reductions (adaptive.cuh) ---------------------------------
AdaptDev adaptdev(const nlls_ctx* c) {
    AdaptDev p;
    p.data = c->d_ad_data; p.chunks = c->d_ad_chunks; p.nchunks = c->ad_nchunks; p.nmeans = (int)c->nB; p.dof = (int)c->dof;
    p.moff = c->d_ad_moff; p.koff = c->ad_koff;
    p.fixdof = c->masked ? c->d_fixdof : nullptr;
    return p;
}
int adapt_linearize(nlls_ctx* ctx) {
    const AdaptDev p = adaptdev(ctx);
    if (ctx->ad_nchunks > 0) { adapt_lin_kernel<<<ctx->ad_nchunks, AD_THREADS, 0, ctx->st>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_ad_part); ctx->launches++; }
    adapt_assemble_kernel<<<1, 64, 0, ctx->st>>>(p, ctx->d_ad_part, ctx->d_H, ctx->d_g, ctx->d_scal + SC_COST_LIN); ctx->launches++;
    CK(cudaGetLastError());
    if (ctx->nranks > 1) {   // residuals are sharded: every entry of the dense system is a sum over all ranks
        CKN(g_nccl.GroupStart());
        CKN(g_nccl.AllReduce(ctx->d_H, ctx->d_H, (size_t)ctx->dof * ctx->dof, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.AllReduce(ctx->d_g, ctx->d_g, (size_t)ctx->dof, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.AllReduce(ctx->d_scal + SC_COST_LIN, ctx->d_scal + SC_COST_LIN, 1, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.GroupEnd());
    }
    return NLLS_OK;
}
int adapt_cost(nlls_ctx* ctx, int which, int slot) {
    const AdaptDev p = adaptdev(ctx);
    if (ctx->ad_nchunks > 0) { adapt_cost_kernel<<<ctx->ad_nchunks, AD_THREADS, 0, ctx->st>>>(p, ctx->d_A[which], ctx->d_B[which], ctx->d_cost_part); ctx->launches++; }
    reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_cost_part, ctx->ad_nchunks, ctx->d_scal + slot, 0); ctx->launches++;
    CK(cudaGetLastError());
    TRY(allreduce(ctx, ctx->d_scal + slot, 1, ncclSum));
    return NLLS_OK;
}
int adapt_solve_update(nlls_ctx* ctx, double lambda) {
    const AdaptDev p = adaptdev(ctx);
    CK(cudaMemsetAsync(ctx->d_scal + SC_P_MAX, 0, 4 * sizeof(double), ctx->st));   // no second variable class on this path
    adapt_solve_kernel<<<1, 32, 0, ctx->st>>>(p, ctx->d_H, ctx->d_g, lambda, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_A[ctx->nxt], ctx->d_B[ctx->nxt],
                                              ctx->d_x, ctx->d_scal + SC_C_MAX);
    ctx->launches++;
    CK(cudaGetLastError());
    return NLLS_OK;
}

#define DISPATCH(ctx, fn, ...)                                                                 \
    ((ctx)->restype == NLLS_RES_AFFINE_BA ? fn<AffineBA>(__VA_ARGS__)                          \
     : (ctx)->restype == NLLS_RES_PINHOLE_BA ? fn<PinholeBA>(__VA_ARGS__)                      \
                                             : (int)NLLS_ERR_NO_KERNEL)

// ---- vector helpers of the Dogleg / gradient-descent iterators (single rank) -----------------------------------------
int fetch_scalars(nlls_ctx* ctx);
constexpr int VEC_GRID = 592;
int vec_dot(nlls_ctx* ctx, const double* a, const double* b, double* out) {   // deterministic: fixed grid, ordered reduction
    dot_kernel<<<VEC_GRID, 256, 0, ctx->st>>>(a, b, (long long)ctx->dof, ctx->d_vec_part); ctx->launches++;
    reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_vec_part, VEC_GRID, ctx->d_scal + SC_EXCH, 0); ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scal + SC_EXCH, ctx->d_scal + SC_EXCH, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    *out = ctx->h_scal[SC_EXCH];
    return NLLS_OK;
}
int vec_axpby(nlls_ctx* ctx, double* x, double sx, const double* y, double sy) {
    axpby_kernel<<<(unsigned)((ctx->dof + 255) / 256), 256, 0, ctx->st>>>(x, sx, y, sy, (long long)ctx->dof); ctx->launches++;
    CK(cudaGetLastError());
    return NLLS_OK;
}
template <class R>
int vec_quadform(nlls_ctx* ctx, const double* v, double* out) {               // v' H v, undamped H  (fast_bAb)
    DevProblem p = devproblem(ctx);
    const int grid = (int)((ctx->nA + ctx->nB + 127) / 128);
    quadform_kernel<R::DC><<<grid, 128, 0, ctx->st>>>(p, v, ctx->d_vec_part); ctx->launches++;
    reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_vec_part, grid, ctx->d_scal + SC_EXCH, 0); ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_scal + SC_EXCH, ctx->d_scal + SC_EXCH, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    *out = ctx->h_scal[SC_EXCH];
    return NLLS_OK;
}
// update!(varnext, variables, linsystem) with the current d_x, then cost(varnext): returns cost, max|x|, |x|
template <class R>
int step_and_cost(nlls_ctx* ctx, double* cost, double* maxstep, double* normx) {
    DevProblem p = devproblem(ctx);
    const int grid = (int)((ctx->nA + ctx->nB + 127) / 128);
    apply_step_kernel<R><<<grid, 128, 0, ctx->st>>>(p, ctx->d_x, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], ctx->d_A[ctx->nxt], ctx->d_B[ctx->nxt], ctx->d_vec_part); ctx->launches++;
    reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_vec_part, grid, ctx->d_scal + SC_EXCH, 1); ctx->launches++;
    reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_vec_part + grid, grid, ctx->d_scal + SC_EXCH + 1, 0); ctx->launches++;
    CK(cudaGetLastError());
    const uint64_t tc = now_ns();
    TRY(launch_cost<R>(ctx, ctx->nxt, SC_COST_TRY));
    CK(cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->t_cost += now_ns() - tc;
    ctx->costcomputations += 1;
    *cost = ctx->h_scal[SC_COST_TRY]; *maxstep = ctx->h_scal[SC_EXCH]; *normx = std::sqrt(ctx->h_scal[SC_EXCH + 1]);
    return NLLS_OK;
}

// iterate!(::DoglegData)                                                     src/iterators.jl:48-113
template <class R>
int iterate_dogleg(nlls_ctx* ctx, nlls_iterinfo* info) {
    const nlls_options& o = ctx->opts;
    double gnorm2 = 0, bAb = 0;
    const uint64_t ts = now_ns();
    TRY(vec_dot(ctx, ctx->d_g, ctx->d_g, &gnorm2));                                             // :52
    TRY(vec_quadform<R>(ctx, ctx->d_g, &bAb));
    const double a = gnorm2 / (bAb + std::numeric_limits<double>::min());                       // :53
    TRY(vec_axpby(ctx, ctx->d_cauchy, 0.0, ctx->d_g, -a));                                      // :54
    const double alpha2 = a * a * gnorm2, alpha = std::sqrt(alpha2);                            // :55-56
    if (ctx->trustradius == 0) ctx->trustradius = alpha;                                        // :57-60
    double beta = 0;
    if (alpha < ctx->trustradius) {                                                             // Newton step :61-66
        TRY(launch_schur<R>(ctx, 0.0));
        TRY(launch_reduced_solve(ctx));
        TRY(launch_update<R>(ctx));
        TRY(fetch_scalars(ctx));
        beta = std::sqrt(ctx->h_scal[SC_P_SQ] + ctx->h_scal[SC_C_SQ]);
        ctx->linearsolvers += 1;
    }
    ctx->t_solver += now_ns() - ts;
    double cost_ = ctx->bestcost, maxstep = 0, normx = 0;                                       // :68
    int64_t ntries = (alpha < ctx->trustradius) ? 1 : 0;   // linear solves of this iteration (what nlls_iterinfo.ntries counts)
    while (true) {
        double linear_approx;
        if (!(alpha < ctx->trustradius)) {                                                      // first leg :71-74
            TRY(vec_axpby(ctx, ctx->d_x, 0.0, ctx->d_cauchy, ctx->trustradius / alpha));
            linear_approx = ctx->trustradius * (2 * alpha - ctx->trustradius) / (2 * a);
        } else if (beta <= ctx->trustradius) {                                                  // full Newton step :77-79
            linear_approx = cost_;
        } else {                                                                                // second leg :80-95
            double sq_leg = 0, c = 0;
            TRY(vec_axpby(ctx, ctx->d_x, 1.0, ctx->d_cauchy, -1.0));
            TRY(vec_dot(ctx, ctx->d_x, ctx->d_x, &sq_leg));
            TRY(vec_dot(ctx, ctx->d_cauchy, ctx->d_x, &c));
            const double trsq = ctx->trustradius * ctx->trustradius - alpha2;
            double step = std::sqrt(c * c + sq_leg * trsq);
            step = (c <= 0) ? (-c + step) / sq_leg : trsq / (c + step);
            TRY(vec_axpby(ctx, ctx->d_x, step, ctx->d_cauchy, 1.0));
            linear_approx = 0.5 * (a * (1 - step) * (1 - step) * gnorm2) + step * (2 - step) * cost_;
        }
        TRY(step_and_cost<R>(ctx, &cost_, &maxstep, &normx));                                   // :98-101
        const double mu_ = (ctx->bestcost - cost_) / linear_approx;                             // :103
        if (mu_ > 0.375) ctx->trustradius = std::max(ctx->trustradius, 3 * normx);              // :104-105
        else if (mu_ < 0.125) ctx->trustradius *= 0.5;                                          // :106-107
        if (!(cost_ > ctx->bestcost) || maxstep < o.dstep) break;                               // :110-113
    }
    ctx->cost = cost_; ctx->maxstep = maxstep;
    if (info) { info->cost = cost_; info->lambda = ctx->trustradius; info->maxstep = maxstep; info->stepnorm = normx; info->ntries = ntries; info->accepted = !(cost_ > ctx->bestcost); }
    return NLLS_OK;
}

// iterate!(::GradientDescentData)                                            src/iterators.jl:187-208
template <class R>
int iterate_gd(nlls_ctx* ctx, nlls_iterinfo* info) {
    double cost_ = 0, maxstep = 0, normx = 0;
    const int64_t ntries = 0;   // no linear solve in this iterator
    TRY(vec_axpby(ctx, ctx->d_x, 0.0, ctx->d_g, -ctx->stepsize));                               // :190
    TRY(step_and_cost<R>(ctx, &cost_, &maxstep, &normx));
    while (cost_ > ctx->bestcost) {                                                             // :195
        double coststep = 0;
        TRY(vec_dot(ctx, ctx->d_x, ctx->d_g, &coststep));                                       // :197
        const double costdiff = ctx->bestcost + coststep - cost_;                               // :198
        ctx->stepsize *= 0.5 * coststep / costdiff;                                             // :200
        TRY(vec_axpby(ctx, ctx->d_x, 0.0, ctx->d_g, -ctx->stepsize));                           // :202
        TRY(step_and_cost<R>(ctx, &cost_, &maxstep, &normx));
    }
    ctx->stepsize *= 2;                                                                         // :207
    ctx->cost = cost_; ctx->maxstep = maxstep;
    if (info) { info->cost = cost_; info->lambda = ctx->stepsize; info->maxstep = maxstep; info->stepnorm = normx; info->ntries = ntries; info->accepted = 1; }
    return NLLS_OK;
}

int fetch_scalars(nlls_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return NLLS_OK;
}

int do_linearize(nlls_ctx* ctx, double* cost) {
    if (ctx->adaptive) {
        TRY(adapt_linearize(ctx));
        TRY(fetch_scalars(ctx));
        if (cost) *cost = ctx->h_scal[SC_COST_LIN];
        return NLLS_OK;
    }
    // the accepted try's cost evaluation already ran the camera pass at these variables: only its finalize is left
    const int cam_mode = (ctx->cam_part_vars == ctx->cur) ? 2 : 1;
    if (!cost) CK(cudaEventRecord(ctx->ev_g0, ctx->st));
    TRY(DISPATCH(ctx, launch_linearize, ctx, true, cam_mode));
    ctx->cam_part_vars = -1;
    reduce_partials_kernel<<<1, 1024, 0, ctx->st>>>(ctx->d_cost_part, ctx->ntiles + ctx->nlong, ctx->d_scal + SC_COST_LIN, 0); ctx->launches++;
    CK(cudaGetLastError());
    if (ctx->nranks > 1) {  // camera blocks, camera gradient and the cost are sums over all ranks' observations: one grouped launch
        CKN(g_nccl.GroupStart());
        CKN(g_nccl.AllReduce(ctx->d_H, ctx->d_H, (size_t)ctx->DC * ctx->DC * ctx->nA, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.AllReduce(ctx->d_g, ctx->d_g, (size_t)ctx->DC * ctx->nA, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.AllReduce(ctx->d_scal + SC_COST_LIN, ctx->d_scal + SC_COST_LIN, 1, ncclFloat64, ncclSum, ctx->comm, ctx->st));
        CKN(g_nccl.GroupEnd());
    }
    // the LM loop discards this cost (src/optimize.jl:169): no read-back, the host goes on to enqueue the next try; the time of this
    // linearisation (events) is booked as timegradient at the next host sync (do_try / nlls_lm_end)
    if (cost) { TRY(fetch_scalars(ctx)); *cost = ctx->h_scal[SC_COST_LIN]; }
    else { CK(cudaEventRecord(ctx->ev_g1, ctx->st)); ctx->grad_pending = true; }
    return NLLS_OK;
}

// one LM try without the host sync: damp + Schur + reduced solve + back-substitution/update + cost(varnext)
int enqueue_try(nlls_ctx* ctx, double lambda) {
    if (ctx->adaptive) {
        TRY(adapt_solve_update(ctx, lambda));
        CK(cudaEventRecord(ctx->ev_c0, ctx->st));
        TRY(adapt_cost(ctx, ctx->nxt, SC_COST_TRY));
        CK(cudaEventRecord(ctx->ev_c1, ctx->st));
        return NLLS_OK;
    }
    TRY(DISPATCH(ctx, launch_schur, ctx, lambda));
    TRY(launch_reduced_solve(ctx));
    TRY(DISPATCH(ctx, launch_update, ctx));
    CK(cudaEventRecord(ctx->ev_c0, ctx->st));
    TRY(DISPATCH(ctx, launch_cost, ctx, ctx->nxt, SC_COST_TRY, nullptr, false));
    CK(cudaEventRecord(ctx->ev_c1, ctx->st));
    return exchange_try_scalars(ctx);
}

// One LM try.  The fused try is booked like the reference's timers (src/iterators.jl:152,157): the cost evaluation (CUDA events around
// its kernels) as timecost, everything else of the try (damp + solve + update, host wall clock minus the cost) as timesolver.
int do_try(nlls_ctx* ctx, double lambda) {
    const uint64_t ts = now_ns();
    TRY(enqueue_try(ctx, lambda));
    TRY(fetch_scalars(ctx));
    uint64_t dt = now_ns() - ts;
    if (ctx->grad_pending) {   // the asynchronous re-linearisation before this try finished inside this wait
        float gms = 0.f;
        CK(cudaEventElapsedTime(&gms, ctx->ev_g0, ctx->ev_g1));
        const uint64_t tg = std::min<uint64_t>(dt, (uint64_t)(gms * 1e6));
        ctx->t_grad += tg; dt -= tg;
        ctx->grad_pending = false;
    }
    float cms = 0.f;
    CK(cudaEventElapsedTime(&cms, ctx->ev_c0, ctx->ev_c1));
    const uint64_t tc = std::min<uint64_t>(dt, (uint64_t)(cms * 1e6));
    ctx->t_cost += tc;
    ctx->t_solver += dt - tc;
    return NLLS_OK;
}

}  // namespace


namespace {
int apply_unfixed(nlls_ctx* ctx);
// makesymmvls for an adaptive-kernel problem: variable -> offset map (block order = variable order, src/linearsystem.jl:93-102),
// residuals grouped by mean variable and cut into single-mean chunks.
int prepare_adaptive(nlls_ctx* ctx) {
    ctx->adaptive = true;
    ctx->vtA = NLLS_VAR_CONTAMGAUSS; ctx->vtB = NLLS_VAR_SCALAR;
    ctx->DC = 3; ctx->NC = 3; ctx->CS = 3; ctx->BS = 1;
    if (!ctx->vars.count(ctx->vtA) || !ctx->vars.count(ctx->vtB)) FAIL(NLLS_ERR_INVALID, "kernel and mean variables must be set before prepare");
    for (auto& kv : ctx->vars)
        if (kv.first != ctx->vtA && kv.first != ctx->vtB && !kv.second.gidx.empty())
            FAIL(NLLS_ERR_UNSUPPORTED, "variables of a type the registered residual does not use");
    for (int vt : {ctx->vtA, ctx->vtB}) {
        VarSet& vs = ctx->vars[vt];
        if (!vs.stale) continue;
        CK(cudaMemcpy(vs.vals.data(), vt == ctx->vtA ? ctx->d_A[ctx->cur] : ctx->d_B[ctx->cur], sizeof(double) * vs.vals.size(), cudaMemcpyDeviceToHost));
        vs.stale = false;
    }
    const VarSet& A = ctx->vars[ctx->vtA];
    const VarSet& B = ctx->vars[ctx->vtB];
    if (A.gidx.size() != 1 || A.gidx[0] != ctx->kernel_var) FAIL(NLLS_ERR_INVALID, "exactly one ContaminatedGaussian variable, at index kernel_var, is expected");
    ctx->nA = 1; ctx->nB = (int64_t)B.gidx.size(); ctx->nobs = (int64_t)ctx->h_pt_g.size();
    if (ctx->nB < 1) FAIL(NLLS_ERR_INVALID, "no mean variable");
    ctx->dof = 3 + ctx->nB;
    if (ctx->dof > AD_MAXDOF) FAIL(NLLS_ERR_UNSUPPORTED, "more than " + std::to_string(AD_MAXDOF - 3) + " mean variables");
    if (ctx->nobs >= (1LL << 31) - 1024) FAIL(NLLS_ERR_UNSUPPORTED, "more than 2^31 costs per rank");
    ctx->hlen = ctx->dof * ctx->dof; ctx->nred = 0; ctx->ntiles = 0; ctx->nitems = 0; ctx->cams_first = true;
    // offsets in variable order
    ctx->h_ad_moff.assign((size_t)ctx->nB, 0);
    {
        int off = 0; size_t ib = 0; bool kdone = false;
        while (ib < B.gidx.size() || !kdone) {
            const bool takeK = !kdone && (ib >= B.gidx.size() || A.gidx[0] < B.gidx[ib]);
            if (takeK) { ctx->ad_koff = off; off += 3; kdone = true; }
            else { if (B.gidx[ib] == A.gidx[0]) FAIL(NLLS_ERR_INVALID, "variable index used by two variable types"); ctx->h_ad_moff[ib++] = off; off += 1; }
        }
    }
    // group residuals by mean variable (stable), chunks of one mean
    const int64_t n = ctx->nobs, M = ctx->nB;
    std::vector<int> ml((size_t)n), cnt((size_t)M + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t gi = ctx->h_pt_g[(size_t)i];
        auto it = std::lower_bound(B.gidx.begin(), B.gidx.end(), gi);
        if (it == B.gidx.end() || *it != gi) FAIL(NLLS_ERR_INVALID, "cost " + std::to_string(i + 1) + ": varind is not a scalar mean variable");
        ml[(size_t)i] = (int)(it - B.gidx.begin());
        cnt[(size_t)ml[(size_t)i] + 1]++;
    }
    for (int64_t m = 0; m < M; ++m) cnt[(size_t)m + 1] += cnt[(size_t)m];
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    std::vector<double> data((size_t)n);
    for (int64_t i = 0; i < n; ++i) data[(size_t)fill[(size_t)ml[(size_t)i]]++] = ctx->h_z[(size_t)i];
    constexpr int AD_CHUNK = 4096;
    std::vector<int4> chunks;
    for (int64_t m = 0; m < M; ++m)
        for (int b = cnt[(size_t)m]; b < cnt[(size_t)m + 1]; b += AD_CHUNK) chunks.push_back(make_int4((int)m, b, std::min(b + AD_CHUNK, cnt[(size_t)m + 1]), 0));
    ctx->ad_nchunks = (int)chunks.size();
    TRY(upload(ctx, &ctx->d_ad_data, data)); TRY(upload(ctx, &ctx->d_ad_chunks, chunks)); TRY(upload(ctx, &ctx->d_ad_moff, ctx->h_ad_moff));
    for (int k = 0; k < 3; ++k) { TRY(dalloc(ctx, &ctx->d_A[k], (size_t)3)); TRY(dalloc(ctx, &ctx->d_B[k], (size_t)M)); }
    ctx->cur = 0; ctx->nxt = 1; ctx->bst = 2;
    TRY(upload_vars(ctx, A, 3, ctx->d_A[0])); TRY(upload_vars(ctx, B, 1, ctx->d_B[0]));
    TRY(dalloc(ctx, &ctx->d_H, (size_t)ctx->hlen)); TRY(dalloc(ctx, &ctx->d_g, (size_t)ctx->dof)); TRY(dalloc(ctx, &ctx->d_x, (size_t)ctx->dof));
    CK(cudaMemsetAsync(ctx->d_H, 0, sizeof(double) * ctx->hlen, ctx->st));
    CK(cudaMemsetAsync(ctx->d_g, 0, sizeof(double) * ctx->dof, ctx->st));
    CK(cudaMemsetAsync(ctx->d_x, 0, sizeof(double) * ctx->dof, ctx->st));
    TRY(dalloc(ctx, &ctx->d_cost_part, (size_t)ctx->ad_nchunks)); TRY(dalloc(ctx, &ctx->d_ad_part, (size_t)ctx->ad_nchunks * AD_NP));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->prepared = true;
    ctx->lm_active = false;
    ctx->masked = false;
    return apply_unfixed(ctx);
}
}  // namespace

namespace {
// (re)build the per-class fixed flags from h_unfixed and upload them (prepared contexts only)
int apply_unfixed(nlls_ctx* ctx) {
    ctx->masked = false;
    if (!ctx->prepared) return NLLS_OK;
    if (ctx->adaptive) {   // dense system: a per-DoF mask (kernel variable: 3 DoF at ad_koff, mean m: 1 DoF at moff[m])
        const VarSet& K = ctx->vars[ctx->vtA];
        const VarSet& M = ctx->vars[ctx->vtB];
        auto fixedv = [&](int64_t gidx) { return (size_t)(gidx - 1) < ctx->h_unfixed.size() && !ctx->h_unfixed[(size_t)(gidx - 1)]; };
        std::vector<unsigned char> fd((size_t)ctx->dof, 0);
        ctx->h_fixA.assign(K.gidx.size(), 0); ctx->h_fixB.assign(M.gidx.size(), 0);
        if (!K.gidx.empty() && fixedv(K.gidx[0])) { ctx->h_fixA[0] = 1; ctx->masked = true; for (int a = 0; a < 3; ++a) fd[(size_t)ctx->ad_koff + a] = 1; }
        for (size_t m = 0; m < M.gidx.size(); ++m) if (fixedv(M.gidx[m])) { ctx->h_fixB[m] = 1; ctx->masked = true; fd[(size_t)ctx->h_ad_moff[m]] = 1; }
        if (ctx->masked) TRY(upload(ctx, &ctx->d_fixdof, fd));
        return NLLS_OK;
    }
    const VarSet& A = ctx->vars[ctx->vtA];
    const VarSet& B = ctx->vars[ctx->vtB];
    auto fixed = [&](int64_t gidx) { return (size_t)(gidx - 1) < ctx->h_unfixed.size() && !ctx->h_unfixed[(size_t)(gidx - 1)]; };
    ctx->h_fixA.assign(A.gidx.size(), 0); ctx->h_fixB.assign(B.gidx.size(), 0);
    for (size_t i = 0; i < A.gidx.size(); ++i) if (fixed(A.gidx[i])) { ctx->h_fixA[i] = 1; ctx->masked = true; }
    for (size_t i = 0; i < B.gidx.size(); ++i) if (fixed(B.gidx[i])) { ctx->h_fixB[i] = 1; ctx->masked = true; }
    if (ctx->masked) { TRY(upload(ctx, &ctx->d_fixA, ctx->h_fixA)); TRY(upload(ctx, &ctx->d_fixB, ctx->h_fixB)); }
    ctx->cam_part_vars = -1;
    return NLLS_OK;
}
}  // namespace

// =====================================================================================================
extern "C" {

int nlls_version(void) { return 200; }

int nlls_set_unfixed(nlls_ctx* ctx, const uint8_t* unfixed, int64_t n) {
    if (!ctx || n < 0 || (n > 0 && !unfixed)) return NLLS_ERR_INVALID;
    if (ctx->lm_active && ctx->lm_phase != 0) FAIL(NLLS_ERR_INVALID, "nlls_set_unfixed inside an LM iteration");
    ctx->h_unfixed.assign(unfixed, unfixed + n);
    CK(cudaSetDevice(ctx->device));
    return apply_unfixed(ctx);
}

int nlls_optimize_singles(nlls_ctx* ctx, int vartype, const nlls_options* opts, int64_t* iterations) {
    if (!ctx || !opts) return NLLS_ERR_INVALID;
    if (opts->iterator != NLLS_ITER_LM) FAIL(NLLS_ERR_UNSUPPORTED, "optimizesingles: only the Levenberg-Marquardt iterator has a batched kernel");
    TRY(nlls_prepare(ctx));
    if (ctx->adaptive || vartype != ctx->vtB) FAIL(NLLS_ERR_NO_KERNEL, "optimizesingles has a registered kernel for the point type of the BA residuals only");
    CK(cudaSetDevice(ctx->device));
    SinglesOpts so;
    so.reldcost = opts->reldcost; so.absdcost = opts->absdcost; so.dstep = opts->dstep; so.maxfails = opts->maxfails; so.maxiters = opts->maxiters;
    so.timeup = opts->maxtime_ns == 0 ? 1 : 0;
    unsigned long long* d_it = reinterpret_cast<unsigned long long*>(ctx->d_scal + SC_INFO);
    CK(cudaMemsetAsync(d_it, 0, sizeof(unsigned long long), ctx->st));
    DevProblem p = devproblem(ctx);
    const int grid = (int)((ctx->nB + 127) / 128);
    const bool ms = ctx->sets.size() > 1;
    if (ctx->restype == NLLS_RES_AFFINE_BA) {
        if (ms) singles_point_kernel<AffineBA, true><<<grid, 128, 0, ctx->st>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], so, d_it);
        else singles_point_kernel<AffineBA><<<grid, 128, 0, ctx->st>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], so, d_it);
    } else {
        if (ms) singles_point_kernel<PinholeBA, true><<<grid, 128, 0, ctx->st>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], so, d_it);
        else singles_point_kernel<PinholeBA><<<grid, 128, 0, ctx->st>>>(p, ctx->d_A[ctx->cur], ctx->d_B[ctx->cur], so, d_it);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    unsigned long long h_it = 0;
    CK(cudaMemcpyAsync(&h_it, d_it, sizeof(h_it), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (iterations) *iterations = (int64_t)h_it;
    for (auto& kv : ctx->vars) kv.second.stale = true;   // the device holds newer variables than the host mirror
    ctx->cam_part_vars = -1;
    ctx->lm_active = false;
    return NLLS_OK;
}

int nlls_create(nlls_ctx** out, int device) {
    if (!out) return NLLS_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return NLLS_ERR_NO_DEVICE;
    nlls_ctx* ctx = new nlls_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return NLLS_ERR_CUDA; }
    cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->st2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreate(&ctx->ev_t0);
    cudaEventCreate(&ctx->ev_t1);
    cudaEventCreate(&ctx->ev_b0);
    cudaEventCreate(&ctx->ev_b1);
    cudaEventCreate(&ctx->ev_c0);
    cudaEventCreate(&ctx->ev_c1);
    cudaEventCreate(&ctx->ev_g0);
    cudaEventCreate(&ctx->ev_g1);
    cudaMalloc((void**)&ctx->d_gather, sizeof(double) * 8 * 64);
    cudaMalloc((void**)&ctx->d_scal, sizeof(double) * SC_COUNT);
    cudaMemset(ctx->d_scal, 0, sizeof(double) * SC_COUNT);
    cudaMallocHost((void**)&ctx->h_scal, sizeof(double) * (SC_COUNT + 2));
    if (ctx->h_scal) std::memset(ctx->h_scal, 0, sizeof(double) * (SC_COUNT + 2));
    const char* e = getenv("NLLS_B200_TMA");
    ctx->use_tma = (e && e[0] == '0') ? 0 : 1;
    if (const char* g = getenv("NLLS_B200_SCHUR_STRIDE")) ctx->schur_stride = std::max(1, atoi(g));
    if (const char* g = getenv("NLLS_B200_GRAPH")) ctx->use_graph = atoi(g) != 0;
    if (const char* g = getenv("NLLS_B200_PDL")) ctx->use_pdl = atoi(g);
    if (const char* g = getenv("NLLS_B200_BWD")) ctx->bwd_flow = std::string(g) == "flow" ? 1 : (std::string(g) == "levels" ? 0 : 2);
    if (const char* g = getenv("NLLS_B200_COST")) ctx->cost_pointmajor = std::string(g) == "tiles";
    if (const char* g = getenv("NLLS_B200_SCHUR")) {
        const std::string m(g);
        ctx->schur_v4 = (m == "v2" || m == "v5") ? 0 : ((m == "v4") ? 2 : 1);
        ctx->schur_v5 = (m == "v2" || m == "v4") ? 0 : ((m == "v5") ? 2 : 1);
    }
    if (const char* g = getenv("NLLS_B200_TILE")) { const int v = atoi(g); ctx->tile_env = (v == 64 || v == 128) ? v : 256; }
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->nsm = v; }
    if (cudaGetLastError() != cudaSuccess || !ctx->st || !ctx->st2 || !ctx->d_scal || !ctx->h_scal) { nlls_destroy(ctx); return NLLS_ERR_CUDA; }
    *out = ctx;
    return NLLS_OK;
}

int nlls_destroy(nlls_ctx* ctx) {
    if (!ctx) return NLLS_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {ctx->d_obs_cam, ctx->d_obs_pt, ctx->d_obs_start, ctx->d_tile_pt, ctx->d_obs_z, ctx->d_cm_pt, ctx->d_item_cam, ctx->d_item_beg,
                    ctx->d_item_end, ctx->d_cam_item_start, ctx->d_cm_z, ctx->d_A[0], ctx->d_A[1], ctx->d_A[2], ctx->d_B[0], ctx->d_B[1], ctx->d_B[2],
                    ctx->d_H, ctx->d_g, ctx->d_x, ctx->d_Ainv, ctx->d_S, ctx->d_rhs, ctx->d_cost_part, ctx->d_step_part, ctx->d_cam_part, ctx->d_cam_part2, ctx->d_scal, ctx->d_gather,
                    ctx->d_flush, ctx->d_tile_id, ctx->d_pos, ctx->d_diag_tile, ctx->d_diag_tile_nat, ctx->d_add_u,
                    ctx->d_lvl_cols, ctx->d_red_tasks, ctx->d_red_upds, ctx->d_red_targets, ctx->d_colptr,
                    ctx->d_col_tile, ctx->d_col_row, ctx->d_Linv, ctx->d_xp, ctx->d_stile_pt, ctx->d_chunk_off, ctx->d_chunks, ctx->d_ents, ctx->d_tiles, ctx->d_camstat_part,
                    ctx->d_ad_data, ctx->d_ad_chunks, ctx->d_ad_moff, ctx->d_ad_part, ctx->d_ent_off, ctx->d_cta_item, ctx->d_items, ctx->d_units, ctx->d_wtab, ctx->d_blob,
                    ctx->d5_cta_item, ctx->d5_items, ctx->d5_blob, ctx->d5_ftab, ctx->d_out_pts, ctx->d_bwd_order, ctx->d_bwd_flags, ctx->d_nat_of_pos, ctx->d_long_pts, ctx->d_obs_set, ctx->d_cm_set, ctx->d_rk_tab, ctx->d_fixdof, ctx->d_em, ctx->d_fixA, ctx->d_fixB, ctx->d_cauchy, ctx->d_vec_part};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->red_graph_exec) cudaGraphExecDestroy(ctx->red_graph_exec);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (cudaEvent_t e : {ctx->ev_fork, ctx->ev_join, ctx->ev_t0, ctx->ev_t1, ctx->ev_b0, ctx->ev_b1, ctx->ev_c0, ctx->ev_c1, ctx->ev_g0, ctx->ev_g1}) if (e) cudaEventDestroy(e);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    if (ctx->st2) cudaStreamDestroy(ctx->st2);
    delete ctx;
    return NLLS_OK;
}

const char* nlls_last_error(nlls_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nlls_comm_unique_id(void* id128) {
    if (!g_nccl.load()) return NLLS_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return NLLS_ERR_NCCL;
    std::memcpy(id128, &id, 128);
    return NLLS_OK;
}

int nlls_comm_init(nlls_ctx* ctx, int rank, int nranks, const void* id128) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return NLLS_ERR_INVALID;
    ctx->rank = rank; ctx->nranks = nranks;
    if (nranks == 1) return NLLS_OK;
    if (nranks > 64) FAIL(NLLS_ERR_UNSUPPORTED, "more than 64 ranks");
    if (!g_nccl.load()) FAIL(NLLS_ERR_NCCL, "libnccl.so.2 not found");
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    CKN(g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
    return NLLS_OK;
}

int nlls_set_variables(nlls_ctx* ctx, int vartype, const double* aos, int64_t n, int64_t stride, int64_t first_index, const int64_t* indices) {
    if (!ctx || !aos || n < 0) return NLLS_ERR_INVALID;
    const int ns = vartype_nstore(vartype);
    if (ns == 0) FAIL(NLLS_ERR_NO_KERNEL, "variable type " + std::to_string(vartype) + " has no registered update kernel");
    if (stride < ns) FAIL(NLLS_ERR_INVALID, "stride smaller than the variable's stored size");
    VarSet& vs = ctx->vars[vartype];
    // same structure as before?  (values-only refresh of problem.variables: one H2D copy straight from the caller's buffer)
    bool same = ctx->prepared && vs.vartype == vartype && (int64_t)vs.gidx.size() == n && n > 0;
    if (same) {
        if (indices) { for (int64_t i = 0; i < n && same; ++i) same = (indices[i] == vs.gidx[(size_t)i]); }
        else same = (vs.gidx.front() == first_index && vs.gidx.back() == first_index + n - 1);
    }
    if (same && (vartype == ctx->vtA || vartype == ctx->vtB)) {
        CK(cudaSetDevice(ctx->device));
        const bool isA = vartype == ctx->vtA;
        const int ds = isA ? ctx->CS : ctx->BS;
        double* dst = isA ? ctx->d_A[ctx->cur] : ctx->d_B[ctx->cur];
        if (ds == stride) {
            CK(cudaMemcpyAsync(dst, aos, sizeof(double) * n * ds, cudaMemcpyHostToDevice, ctx->st));
        } else {
            CK(cudaMemcpy2DAsync(dst, sizeof(double) * ds, aos, sizeof(double) * stride, sizeof(double) * ns, (size_t)n, cudaMemcpyHostToDevice, ctx->st));
        }
        CK(cudaStreamSynchronize(ctx->st));
        vs.stale = true;
        ctx->cam_part_vars = -1;
        return NLLS_OK;
    }
    std::vector<int64_t> gi((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        gi[(size_t)i] = indices ? indices[i] : first_index + i;
        if (gi[(size_t)i] < 1 || (i > 0 && gi[(size_t)i] <= gi[(size_t)i - 1])) FAIL(NLLS_ERR_INVALID, "variable indices must be 1-based and ascending");
    }
    vs.vartype = vartype; vs.nstore = ns; vs.stale = false;
    vs.gidx.swap(gi);
    vs.vals.resize((size_t)n * ns);
    for (int64_t i = 0; i < n; ++i) std::memcpy(&vs.vals[(size_t)i * ns], aos + (size_t)i * stride, sizeof(double) * ns);
    ctx->prepared = false;
    return NLLS_OK;
}

int nlls_set_costs(nlls_ctx* ctx, int restype, const void* aos, int64_t stride_bytes, int64_t n, int robust, const double* kparams, int nkparams,
                   int64_t kernel_var) {
    if (!ctx || (!aos && n > 0) || n < 0) return NLLS_ERR_INVALID;
    if (restype == NLLS_RES_ADAPTIVE_OFFSET) {
        // AbstractAdaptiveResidual: the robust kernel is variable `kernel_var` (src/residual.jl:47, src/problem.jl:97)
        if (robust != NLLS_ROBUST_NONE) FAIL(NLLS_ERR_INVALID, "adaptive residuals take their kernel from a variable, not from robustkernel()");
        if (stride_bytes < 16) FAIL(NLLS_ERR_INVALID, "cost stride must be >= 16 bytes (f64 data + i64 varind)");
        if (kernel_var < 1) FAIL(NLLS_ERR_INVALID, "kernel_var must be the 1-based index of the ContaminatedGaussian variable");
        ctx->restype = restype; ctx->robust = robust; ctx->kernel_var = kernel_var;
        ctx->h_cam_g.clear(); ctx->h_pt_g.resize((size_t)n); ctx->h_z.resize((size_t)n);
        const unsigned char* base = (const unsigned char*)aos;
        for (int64_t i = 0; i < n; ++i) {
            std::memcpy(&ctx->h_z[(size_t)i], base + (size_t)i * stride_bytes, 8);
            std::memcpy(&ctx->h_pt_g[(size_t)i], base + (size_t)i * stride_bytes + 8, 8);
        }
        ctx->prepared = false;
        return NLLS_OK;
    }
    if (restype != NLLS_RES_AFFINE_BA && restype != NLLS_RES_PINHOLE_BA)
        FAIL(NLLS_ERR_NO_KERNEL, "residual type " + std::to_string(restype) + " has no registered sm_100a kernel (no CPU fallback)");
    const int kind = robust & 15;
    if (kind > NLLS_ROBUST_GEMANMCCLURE || (robust & ~(15 | NLLS_ROBUST_SCALED)))
        FAIL(NLLS_ERR_NO_KERNEL, "robust kernel " + std::to_string(robust) + " has no registered device implementation");
    if (stride_bytes < 32) FAIL(NLLS_ERR_INVALID, "cost stride must be >= 32 bytes (2 x f64 measurement + 2 x i64 varind)");
    ctx->restype = restype; ctx->robust = robust;
    ctx->kparams[0] = 0.0; ctx->kparams[1] = 1.0;
    if (kind != NLLS_ROBUST_NONE) {
        if (nkparams < 1 || !kparams) FAIL(NLLS_ERR_INVALID, "robust kernel needs its width in kparams[0]");
        ctx->kparams[0] = kparams[0];
    }
    if (robust & NLLS_ROBUST_SCALED) {
        if (nkparams < 2 || !kparams) FAIL(NLLS_ERR_INVALID, "Scaled kernel needs (width, height) in kparams");
        ctx->kparams[1] = kparams[1];
    }
    ctx->h_cam_g.resize((size_t)n); ctx->h_pt_g.resize((size_t)n); ctx->h_z.resize((size_t)2 * n);
    const unsigned char* base = (const unsigned char*)aos;
    for (int64_t i = 0; i < n; ++i) {
        const unsigned char* e = base + (size_t)i * stride_bytes;
        double z[2]; int64_t vi[2];
        std::memcpy(z, e, 16); std::memcpy(vi, e + 16, 16);
        ctx->h_z[(size_t)2 * i] = z[0]; ctx->h_z[(size_t)2 * i + 1] = z[1];
        ctx->h_cam_g[(size_t)i] = vi[0]; ctx->h_pt_g[(size_t)i] = vi[1];
    }
    ctx->sets.assign(1, nlls_ctx::SetSpec{robust, {ctx->kparams[0], ctx->kparams[1]}});
    ctx->h_set.assign((size_t)n, 0);
    ctx->prepared = false;
    return NLLS_OK;
}

// A further Vector{T} of problem.costs.data (src/VectorRepo.jl:3): costs of the same residual struct whose residual type carries a
// different robustkernel().  The reference sums cost, gradient and Hessian over all cost types (src/cost.jl:54).
int nlls_add_costs(nlls_ctx* ctx, int restype, const void* aos, int64_t stride_bytes, int64_t n, int robust, const double* kparams, int nkparams) {
    if (!ctx || (!aos && n > 0) || n < 0) return NLLS_ERR_INVALID;
    if (ctx->sets.empty() || ctx->restype == 0) FAIL(NLLS_ERR_INVALID, "nlls_add_costs before nlls_set_costs");
    if (restype == NLLS_RES_ADAPTIVE_OFFSET || ctx->restype == NLLS_RES_ADAPTIVE_OFFSET || restype != ctx->restype)
        FAIL(NLLS_ERR_UNSUPPORTED, "cost sets of one problem must share the residual struct (same variables, same kernel family); they may differ in the robust kernel");
    if (ctx->sets.size() >= 8) FAIL(NLLS_ERR_UNSUPPORTED, "more than 8 cost sets");
    const int kind = robust & 15;
    if (kind > NLLS_ROBUST_GEMANMCCLURE || (robust & ~(15 | NLLS_ROBUST_SCALED)))
        FAIL(NLLS_ERR_NO_KERNEL, "robust kernel " + std::to_string(robust) + " has no registered device implementation");
    if (stride_bytes < 32) FAIL(NLLS_ERR_INVALID, "cost stride must be >= 32 bytes (2 x f64 measurement + 2 x i64 varind)");
    nlls_ctx::SetSpec sp{robust, {0.0, 1.0}};
    if (kind != NLLS_ROBUST_NONE) {
        if (nkparams < 1 || !kparams) FAIL(NLLS_ERR_INVALID, "robust kernel needs its width in kparams[0]");
        sp.kp[0] = kparams[0];
    }
    if (robust & NLLS_ROBUST_SCALED) {
        if (nkparams < 2 || !kparams) FAIL(NLLS_ERR_INVALID, "Scaled kernel needs (width, height) in kparams");
        sp.kp[1] = kparams[1];
    }
    const size_t n0 = ctx->h_cam_g.size();
    ctx->h_cam_g.resize(n0 + (size_t)n); ctx->h_pt_g.resize(n0 + (size_t)n); ctx->h_z.resize(2 * (n0 + (size_t)n)); ctx->h_set.resize(n0 + (size_t)n);
    const unsigned char* base = (const unsigned char*)aos;
    for (int64_t i = 0; i < n; ++i) {
        const unsigned char* e = base + (size_t)i * stride_bytes;
        double z[2]; int64_t vi[2];
        std::memcpy(z, e, 16); std::memcpy(vi, e + 16, 16);
        ctx->h_z[2 * (n0 + (size_t)i)] = z[0]; ctx->h_z[2 * (n0 + (size_t)i) + 1] = z[1];
        ctx->h_cam_g[n0 + (size_t)i] = vi[0]; ctx->h_pt_g[n0 + (size_t)i] = vi[1];
        ctx->h_set[n0 + (size_t)i] = (unsigned char)ctx->sets.size();
    }
    ctx->sets.push_back(sp);
    ctx->prepared = false;
    return NLLS_OK;
}

int nlls_prepare(nlls_ctx* ctx) {
    if (!ctx) return NLLS_ERR_INVALID;
    if (ctx->prepared) return NLLS_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->red_graph_exec) { cudaGraphExecDestroy(ctx->red_graph_exec); ctx->red_graph_exec = nullptr; }   // captured pointers go stale
    if (ctx->restype == 0) FAIL(NLLS_ERR_INVALID, "no costs set");
    if (ctx->restype == NLLS_RES_ADAPTIVE_OFFSET) return prepare_adaptive(ctx);
    ctx->adaptive = false; ctx->BS = 3;
    ctx->vtA = (ctx->restype == NLLS_RES_AFFINE_BA) ? NLLS_VAR_EUCLID6 : NLLS_VAR_PINHOLE;
    ctx->vtB = NLLS_VAR_EUCLID3;
    ctx->DC = (ctx->restype == NLLS_RES_AFFINE_BA) ? AffineBA::DC : PinholeBA::DC;
    ctx->NC = (ctx->restype == NLLS_RES_AFFINE_BA) ? AffineBA::NC : PinholeBA::NC;
    ctx->CS = (ctx->restype == NLLS_RES_AFFINE_BA) ? AffineBA::CS : PinholeBA::CS;
    if (!ctx->vars.count(ctx->vtA) || !ctx->vars.count(ctx->vtB)) FAIL(NLLS_ERR_INVALID, "camera and point variables must be set before prepare");
    for (auto& kv : ctx->vars)
        if (kv.first != ctx->vtA && kv.first != ctx->vtB && !kv.second.gidx.empty())
            FAIL(NLLS_ERR_UNSUPPORTED, "variables of a type the registered residual does not use");
    for (int vt : {ctx->vtA, ctx->vtB}) {  // values refreshed on the device since the last full set: pull them back first
        VarSet& vs = ctx->vars[vt];
        if (!vs.stale) continue;
        const bool isA = vt == ctx->vtA;
        const int ds = isA ? ctx->CS : 3;
        CK(cudaMemcpy2D(vs.vals.data(), sizeof(double) * vs.nstore, isA ? ctx->d_A[ctx->cur] : ctx->d_B[ctx->cur], sizeof(double) * ds,
                        sizeof(double) * vs.nstore, vs.gidx.size(), cudaMemcpyDeviceToHost));
        vs.stale = false;
    }
    const VarSet& A = ctx->vars[ctx->vtA];
    const VarSet& B = ctx->vars[ctx->vtB];
    ctx->nA = (int64_t)A.gidx.size(); ctx->nB = (int64_t)B.gidx.size();
    ctx->nobs = (int64_t)ctx->h_cam_g.size();
    if (ctx->nA == 0 || ctx->nB == 0) FAIL(NLLS_ERR_INVALID, "empty variable set");
    if (ctx->nobs >= (1LL << 31) - 1024) FAIL(NLLS_ERR_UNSUPPORTED, "more than 2^31 costs per rank");
    const int DC = ctx->DC, WB = 3 * DC;
    const int64_t maxidx = std::max(A.gidx.back(), B.gidx.back());
    if (maxidx >= (1LL << 30)) FAIL(NLLS_ERR_UNSUPPORTED, "variable index too large");
    ctx->cams_first = A.gidx.back() < B.gidx.front();
    // global -> local lookup
    std::vector<int> g2l((size_t)maxidx + 1, -1);
    std::vector<unsigned char> g2c((size_t)maxidx + 1, 0);
    for (size_t i = 0; i < A.gidx.size(); ++i) { g2l[(size_t)A.gidx[i]] = (int)i; g2c[(size_t)A.gidx[i]] = 1; }
    for (size_t i = 0; i < B.gidx.size(); ++i) {
        if (g2c[(size_t)B.gidx[i]]) FAIL(NLLS_ERR_INVALID, "variable index used by two variable types");
        g2l[(size_t)B.gidx[i]] = (int)i; g2c[(size_t)B.gidx[i]] = 2;
    }
    const int64_t nobs = ctx->nobs, nA = ctx->nA, nB = ctx->nB;
    // counting sort by point (≙ reordercostsforschur!, src/problem.jl:186-196), then by camera inside a point
    std::vector<int> cnt((size_t)nB + 1, 0);
    std::vector<int> caml((size_t)nobs), ptl((size_t)nobs);
    for (int64_t i = 0; i < nobs; ++i) {
        const int64_t cg = ctx->h_cam_g[(size_t)i], pg = ctx->h_pt_g[(size_t)i];
        if (cg < 1 || cg > maxidx || g2c[(size_t)cg] != 1) FAIL(NLLS_ERR_INVALID, "cost " + std::to_string(i + 1) + ": varind[1] is not a camera variable");
        if (pg < 1 || pg > maxidx || g2c[(size_t)pg] != 2) FAIL(NLLS_ERR_INVALID, "cost " + std::to_string(i + 1) + ": varind[2] is not a point variable owned by this rank");
        caml[(size_t)i] = g2l[(size_t)cg]; ptl[(size_t)i] = g2l[(size_t)pg];
        cnt[(size_t)ptl[(size_t)i] + 1]++;
    }
    ctx->h_obs_start.assign((size_t)nB + 1, 0);
    for (int64_t p = 0; p < nB; ++p) ctx->h_obs_start[(size_t)p + 1] = ctx->h_obs_start[(size_t)p] + cnt[(size_t)p + 1];
    // irregular points: more observations than the smallest tile capacity of the kernels / Schur plans in use (KCAP), or two costs
    // on one camera.  The reference accepts both (src/linearsystem.jl:132-175 accumulates); here they are left out of every tile and
    // handled by one-CTA-per-point kernels (lin_point_long / backsub_long) and the Schur fallback kernel.
    const int KCAP = std::min(ctx->tile_env ? ctx->tile_env : 256, (DC <= 6) ? 232 : 128);
    ctx->h_irr.assign((size_t)nB, 0);
    ctx->has_dups = false;
    std::vector<int> fill(ctx->h_obs_start.begin(), ctx->h_obs_start.end() - 1);
    std::vector<int> order((size_t)nobs);
    for (int64_t i = 0; i < nobs; ++i) order[(size_t)fill[(size_t)ptl[(size_t)i]]++] = (int)i;
    {
        int maxk = 0;
        for (int64_t p = 0; p < nB; ++p) {
            int* b = order.data() + ctx->h_obs_start[(size_t)p];
            int* e = order.data() + ctx->h_obs_start[(size_t)p + 1];
            std::stable_sort(b, e, [&](int x, int y) { return caml[(size_t)x] < caml[(size_t)y]; });
            if (e - b > KCAP) ctx->h_irr[(size_t)p] = 1;
            for (int* q = b + 1; q < e; ++q)
                if (caml[(size_t)*q] == caml[(size_t)*(q - 1)]) { ctx->h_irr[(size_t)p] = 1; ctx->has_dups = true; }
            if (!ctx->h_irr[(size_t)p]) maxk = std::max(maxk, (int)(e - b));
        }
        ctx->tile_obs = ctx->tile_env ? ctx->tile_env : (maxk <= 64 ? 64 : (maxk <= 128 ? 128 : 256));
    }
    std::vector<int> long_pts;
    for (int64_t p = 0; p < nB; ++p) if (ctx->h_irr[(size_t)p]) long_pts.push_back((int)p);
    ctx->nlong = (int)long_pts.size();
    TRY(upload(ctx, &ctx->d_long_pts, long_pts));
    if (getenv("NLLS_B200_VERBOSE") && ctx->nlong) fprintf(stderr, "[nlls] %d irregular points (more than %d observations, or several costs on one camera)\n", ctx->nlong, KCAP);
    ctx->h_obs_cam.resize((size_t)nobs); ctx->h_obs_pt.resize((size_t)nobs);
    std::vector<double2> obs_z((size_t)nobs);
    for (int64_t j = 0; j < nobs; ++j) {
        const int i = order[(size_t)j];
        ctx->h_obs_cam[(size_t)j] = caml[(size_t)i]; ctx->h_obs_pt[(size_t)j] = ptl[(size_t)i];
        obs_z[(size_t)j] = make_double2(ctx->h_z[(size_t)2 * i], ctx->h_z[(size_t)2 * i + 1]);
    }
    // tiles: consecutive regular points, <= tile_obs observations and <= tile_obs / 2 points; h_tile_pt holds (first, one past last) pairs
    ctx->h_tile_pt.clear();
    {
        int64_t p0 = 0;
        while (p0 < nB) {
            if (ctx->h_irr[(size_t)p0]) { ++p0; continue; }
            int64_t p1 = p0;
            while (p1 < nB && !ctx->h_irr[(size_t)p1] && (p1 - p0) < ctx->tile_obs / 2 && (ctx->h_obs_start[(size_t)p1 + 1] - ctx->h_obs_start[(size_t)p0]) <= ctx->tile_obs) ++p1;
            ctx->h_tile_pt.push_back((int)p0); ctx->h_tile_pt.push_back((int)p1);
            p0 = p1;
        }
    }
    ctx->ntiles = (int)ctx->h_tile_pt.size() / 2;
    // camera-major copy + work items
    ctx->h_cam_start.assign((size_t)nA + 1, 0);
    for (int64_t j = 0; j < nobs; ++j) ctx->h_cam_start[(size_t)ctx->h_obs_cam[(size_t)j] + 1]++;
    for (int64_t c = 0; c < nA; ++c) ctx->h_cam_start[(size_t)c + 1] += ctx->h_cam_start[(size_t)c];
    std::vector<int> cfill(ctx->h_cam_start.begin(), ctx->h_cam_start.end() - 1);
    ctx->h_cm_obs.resize((size_t)nobs);
    std::vector<int> cm_pt((size_t)nobs);
    std::vector<double2> cm_z((size_t)nobs);
    for (int64_t j = 0; j < nobs; ++j) {
        const int k = cfill[(size_t)ctx->h_obs_cam[(size_t)j]]++;
        ctx->h_cm_obs[(size_t)k] = (int)j; cm_pt[(size_t)k] = ctx->h_obs_pt[(size_t)j]; cm_z[(size_t)k] = obs_z[(size_t)j];
    }
    std::vector<int> item_cam, item_beg, item_end, cam_item_start((size_t)nA + 1, 0);
    for (int64_t c = 0; c < nA; ++c) {
        cam_item_start[(size_t)c] = (int)item_cam.size();
        for (int b = ctx->h_cam_start[(size_t)c]; b < ctx->h_cam_start[(size_t)c + 1]; b += CAM_CHUNK) {
            item_cam.push_back((int)c); item_beg.push_back(b); item_end.push_back(std::min(b + CAM_CHUNK, ctx->h_cam_start[(size_t)c + 1]));
        }
    }
    cam_item_start[(size_t)nA] = (int)item_cam.size();
    ctx->nitems = (int)item_cam.size();
    if (ctx->sets.size() > 1) {   // per-observation cost set in both orders + the sets' kernels
        std::vector<unsigned char> oset((size_t)nobs), cset((size_t)nobs);
        for (int64_t j = 0; j < nobs; ++j) oset[(size_t)j] = ctx->h_set[(size_t)order[(size_t)j]];
        for (int64_t k = 0; k < nobs; ++k) cset[(size_t)k] = oset[(size_t)ctx->h_cm_obs[(size_t)k]];
        std::vector<RobustParams> tab(ctx->sets.size());
        for (size_t q = 0; q < ctx->sets.size(); ++q) {
            tab[q].kind = ctx->sets[q].robust & 15; tab[q].scaled = (ctx->sets[q].robust & NLLS_ROBUST_SCALED) ? 1 : 0;
            tab[q].width = ctx->sets[q].kp[0]; tab[q].width2 = ctx->sets[q].kp[0] * ctx->sets[q].kp[0]; tab[q].height = ctx->sets[q].kp[1];
        }
        TRY(upload(ctx, &ctx->d_obs_set, oset)); TRY(upload(ctx, &ctx->d_cm_set, cset)); TRY(upload(ctx, &ctx->d_rk_tab, tab));
    }

    ctx->dof = (int64_t)DC * nA + 3 * nB;
    ctx->hlen = (int64_t)DC * DC * nA + (int64_t)WB * nobs + 9 * nB;
    ctx->nred = (int64_t)DC * nA;
    // ---- reduced camera system: tile order, tile-level symbolic factorisation, level schedule
    std::vector<int> tile_id, pos, diag_tile, diag_tile_nat, lvl_cols_flat, colptr, col_tile, col_row, add_u_tiles;
    std::vector<RedTask> red_tasks;
    std::vector<RedUpd> red_upds;
    std::vector<RedTarget> red_targets;
    if (ctx->s_tiled) {
        const int TC = ST / DC;
        const int NT = (int)((nA + TC - 1) / TC);
        ctx->NT = NT;
        if ((int64_t)NT * NT > (1LL << 28)) FAIL(NLLS_ERR_UNSUPPORTED, "too many camera tiles for the tile map");
        std::vector<unsigned char> natpat((size_t)NT * NT, 0);   // natural numbering, lower triangle
        for (int I = 0; I < NT; ++I) natpat[(size_t)I * NT + I] = 1;
        std::vector<int> tl;
        for (int64_t pnt = 0; pnt < nB; ++pnt) {
            tl.clear();
            for (int j = ctx->h_obs_start[(size_t)pnt]; j < ctx->h_obs_start[(size_t)pnt + 1]; ++j) {
                const int tI = ctx->h_obs_cam[(size_t)j] / TC;
                if (tl.empty() || tl.back() != tI) tl.push_back(tI);   // cameras ascend within a point
            }
            for (size_t a = 0; a < tl.size(); ++a) for (size_t b = 0; b <= a; ++b) natpat[(size_t)tl[a] * NT + tl[b]] = 2;   // 2: touched by a point of this rank (1: structural diagonal)
        }
        // toucher[tile] (natural numbering, lower triangle): -1 nobody's points touch it, r >= 0 only rank r's do, -2 several ranks'
        std::vector<int> toucher;
        if (ctx->nranks > 1) {  // ranks see different points: every rank needs the union pattern (same tile map, same factorisation) and who touches what
            const size_t np = natpat.size();
            unsigned char* d_pat = nullptr;
            CK(cudaMalloc((void**)&d_pat, np * (size_t)ctx->nranks));
            CK(cudaMemcpy(d_pat + np * (size_t)ctx->rank, natpat.data(), np, cudaMemcpyHostToDevice));
            CKN(g_nccl.AllGather(d_pat + np * (size_t)ctx->rank, d_pat, np, /*ncclUint8*/ 1, ctx->comm, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
            std::vector<unsigned char> all(np * (size_t)ctx->nranks);
            CK(cudaMemcpy(all.data(), d_pat, all.size(), cudaMemcpyDeviceToHost));
            CK(cudaFree(d_pat));
            toucher.assign(np, -1);
            for (int r = 0; r < ctx->nranks; ++r) {
                const unsigned char* pr = all.data() + np * (size_t)r;
                for (size_t e = 0; e < np; ++e) if (pr[e] == 2) { toucher[e] = toucher[e] == -1 ? r : -2; natpat[e] = 2; }
            }
        }
        for (unsigned char& v : natpat) v = v ? 1 : 0;
        // tile order: nested dissection by index when the pattern is banded (half-bandwidth w tiles), identity otherwise
        const int w = red_half_bandwidth(natpat, NT);
        std::vector<int> nat_of_pos;
        const bool use_nd = w >= 1 && NT >= 8 * w && !getenv("NLLS_B200_NO_ND");
        if (use_nd && getenv("NLLS_B200_ND_BAND")) nat_of_pos = red_order_band(NT, w);        // round-1/2a order (Venice shape: 12 levels)
        else if (use_nd) nat_of_pos = red_order_graph(natpat, NT);                              // nested dissection on the actual tile graph (8 levels)
        else for (int i = 0; i < NT; ++i) nat_of_pos.push_back(i);
        const RedSymbolic sym = red_symbolic(natpat, NT, nat_of_pos);                           // reduced_plan.hpp
        pos = sym.pos;
        const std::vector<unsigned char>& pat = sym.pat;
        const std::vector<std::vector<int>>& rows = sym.rows;
        tile_id.assign((size_t)NT * NT, -1);
        int nt = 0;
        for (int J2 = 0; J2 < NT; ++J2) for (int I = J2; I < NT; ++I) if (pat[(size_t)I * NT + J2]) tile_id[(size_t)I * NT + J2] = nt++;
        ctx->ntiles_alloc = nt;
        ctx->xg_block = ctx->xg_shared0 = ctx->xg_nshared = 0;
        std::vector<int> add_u((size_t)NT, (ctx->nranks > 1 && ctx->rank != 0) ? 0 : 1);   // without the ownership exchange: rank 0 adds U_c
        if (ctx->nranks > 1 && !getenv("NLLS_B200_XG_ALLREDUCE")) {
            // Storage order by ownership (a tile id is only a storage slot): [rank 0's exclusive tiles | rank 1's | ... ] in blocks of
            // xg_block tiles, then the tiles several ranks touch, then the fill-only tiles.  A rank's points cover a contiguous camera band,
            // so most tiles have one toucher: the try exchanges S with one all-gather of the blocks (every rank RECEIVES the other ranks'
            // tiles, nothing is summed) + an all-reduce of the few shared tiles, instead of all-reducing all of S (Venice shape, 8 ranks:
            // 18 MB all-reduced = 0.22 ms of a 0.75 ms iteration).  The diagonal tile's U_c + lambda I is added by the tile's owner.
            const RedOwnership own = red_order_by_owner(tile_id, NT, nat_of_pos, toucher, ctx->nranks, ctx->rank);   // reduced_plan.hpp
            add_u = own.add_u;
            ctx->xg_block = own.block; ctx->xg_shared0 = own.shared0; ctx->xg_nshared = own.nshared;
            ctx->ntiles_alloc = own.nslots;
            nt = (int)ctx->ntiles_alloc;
            if (getenv("NLLS_B200_VERBOSE"))
                fprintf(stderr, "[nlls] rank %d: reduced-system exchange by ownership: %lld exclusive tiles (block %lld), %lld shared, %lld fill-only\n", ctx->rank,
                        own.nexcl[(size_t)ctx->rank], own.block, own.nshared, own.nfill);
        }
        add_u_tiles = add_u;
        if ((double)nt * ST2 * 8.0 > 60e9) FAIL(NLLS_ERR_UNSUPPORTED, "reduced camera system needs more than 60 GB");
        diag_tile.resize((size_t)NT); diag_tile_nat.resize((size_t)NT);
        for (int J2 = 0; J2 < NT; ++J2) diag_tile[(size_t)J2] = tile_id[(size_t)J2 * NT + J2];
        for (int o = 0; o < NT; ++o) diag_tile_nat[(size_t)o] = diag_tile[(size_t)pos[(size_t)o]];
        // levels of the elimination tree: column I waits for every column J < I that has I among its rows
        const std::vector<int>& level = sym.level;
        const int nlev = sym.nlev;
        ctx->red_levels = nlev;
        ctx->fact_launches.clear(); ctx->lvl_cols.clear();
        size_t nupd_total = 0;
        for (int lv = 0; lv < nlev; ++lv) {
            const int c0 = (int)lvl_cols_flat.size(), d0 = (int)red_tasks.size();
            for (int J2 = 0; J2 < NT; ++J2) if (level[(size_t)J2] == lv) {
                lvl_cols_flat.push_back(J2);
                RedTask tk; tk.tile = diag_tile[(size_t)J2]; tk.dtile = tk.tile; tk.col = J2; tk.row = J2;
                red_tasks.push_back(tk);
            }
            const int o0 = (int)red_tasks.size(), u0 = (int)red_upds.size(), g0 = (int)red_targets.size();
            std::vector<std::pair<int, RedUpd>> lvl_upds;   // (target tile, update) of this level
            for (int q = c0; q < (int)lvl_cols_flat.size(); ++q) {
                const int J2 = lvl_cols_flat[(size_t)q];
                const std::vector<int>& r = rows[(size_t)J2];
                for (int I : r) { RedTask tk; tk.tile = tile_id[(size_t)I * NT + J2]; tk.dtile = diag_tile[(size_t)J2]; tk.col = J2; tk.row = I; red_tasks.push_back(tk); }
                // eliminating column J couples every pair (a >= b) of its rows: T_{r[a], r[b]} -= L_{r[a],J} D_J L_{r[b],J}'
                for (size_t a = 0; a < r.size(); ++a) for (size_t b = 0; b <= a; ++b) {
                    RedUpd u; u.a = tile_id[(size_t)r[a] * NT + J2]; u.b = tile_id[(size_t)r[b] * NT + J2]; u.dk = diag_tile[(size_t)J2]; u.col = J2;
                    lvl_upds.push_back({tile_id[(size_t)r[a] * NT + r[b]], u});
                }
            }
            // the level's updates grouped by target tile (stable: ascending source column inside a target) — one owner per target
            std::stable_sort(lvl_upds.begin(), lvl_upds.end(), [](const std::pair<int, RedUpd>& x, const std::pair<int, RedUpd>& y) { return x.first < y.first; });
            for (size_t k0 = 0; k0 < lvl_upds.size();) {
                size_t k1 = k0;
                RedTarget tg; tg.target = lvl_upds[k0].first; tg.u0 = (int)red_upds.size(); tg.row = -1;
                while (k1 < lvl_upds.size() && lvl_upds[k1].first == tg.target) { red_upds.push_back(lvl_upds[k1].second); ++k1; }
                tg.u1 = (int)red_upds.size();
                if (lvl_upds[k0].second.a == lvl_upds[k0].second.b) {   // diagonal target (I, I): find I
                    for (int I = 0; I < NT; ++I) if (diag_tile[(size_t)I] == tg.target) { tg.row = I; break; }
                }
                red_targets.push_back(tg);
                k0 = k1;
            }
            ctx->lvl_cols.push_back({c0, (int)lvl_cols_flat.size() - c0});
            ctx->fact_launches.push_back({0, d0, o0 - d0});
            if ((int)red_tasks.size() > o0) ctx->fact_launches.push_back({1, o0, (int)red_tasks.size() - o0});
            if ((int)red_upds.size() > u0) ctx->fact_launches.push_back({2, g0, (int)red_targets.size() - g0});
        }
        nupd_total = red_upds.size();
        {   // algorithmic flops (multiply and add counted separately) of one factorisation + both sweeps, ST^3 = one dense tile GEMM / 2
            const double t3 = (double)ST * ST * ST, t2 = (double)ST * ST;
            const double ndiag = NT, noff = (double)red_tasks.size() - NT;
            double nupd_diag = 0;
            for (const RedTarget& tg : red_targets) if (tg.row >= 0) nupd_diag += tg.u1 - tg.u0;
            ctx->red_flops = ndiag * (t3 / 3 + t3 / 3)            // LDL' of the diagonal tile + its triangular inverse
                             + noff * t3                          // L_IJ = T_IJ L_JJ^-T D^-1 (triangular)
                             + ((double)nupd_total - nupd_diag) * 2 * t3 + nupd_diag * t3   // Schur updates (symmetric targets: half)
                             + (ndiag * 2 + noff * 2) * 2 * t2;   // forward + backward sweeps
        }
        if (getenv("NLLS_B200_VERBOSE"))
            fprintf(stderr, "[nlls] reduced system: NT=%d tiles=%d half-bandwidth=%d nd=%d levels=%d fact_launches=%zu tasks=%zu updates=%zu\n", NT, nt, w,
                    (int)use_nd, nlev, ctx->fact_launches.size(), red_tasks.size(), nupd_total);
        // block-row and block-column lists of L for the sweeps
        colptr.assign((size_t)NT + 1, 0);
        for (int J2 = 0; J2 < NT; ++J2) {
            for (int I : rows[(size_t)J2]) { col_tile.push_back(tile_id[(size_t)I * NT + J2]); col_row.push_back(I); }
            colptr[(size_t)J2 + 1] = (int)col_tile.size();
        }
    }

    // ---- Schur v5 plan (schur5_plan.hpp): window-aligned register accumulation; points that do not fit go to the per-chunk kernel
    ctx->n5cta = 0; ctx->nout_pts = 0;
    std::vector<int> out_pts;
    if (ctx->schur_v5) {
        if (const char* e = getenv("NLLS_B200_S5_COST")) {   // development aid: "dmma,afrag,bfrag,fixed"
            Schur5Cost& cm = schur5_cost();
            sscanf(e, "%lf,%lf,%lf,%lf", &cm.dmma, &cm.afrag, &cm.bfrag, &cm.fixed);
        }
        const int cons_max = (DC == 6) ? Schur5Cfg<6>::CONS : Schur5Cfg<9>::CONS;
        ctx->s5_ncons = cons_max;
        if (const char* e = getenv("NLLS_B200_S5_CONS")) ctx->s5_ncons = std::max(1, std::min(atoi(e), cons_max));
        int maxrun = 48;
        if (const char* e = getenv("NLLS_B200_S5_RUN")) maxrun = std::max(1, atoi(e));
        const std::vector<unsigned char>* irr = ctx->nlong ? &ctx->h_irr : nullptr;
        Schur5Plan P5 = (DC == 6) ? schur5_build_plan<6>(ctx->h_obs_start, ctx->h_obs_cam, nA, ctx->nsm, maxrun, ctx->s5_ncons, irr)
                                  : schur5_build_plan<9>(ctx->h_obs_start, ctx->h_obs_cam, nA, ctx->nsm, maxrun, ctx->s5_ncons, irr);
        ctx->s5_ncons = std::max(ctx->s5_ncons, DC == 6 ? Schur5Cfg<6>::NBANDS : Schur5Cfg<9>::NBANDS);
        // automatic choice: (almost) everything fits the window plan, and every CTA has a few tiles to pipeline (a problem of one tile
        // per SM — Ladybug-shape — is 10 % faster with v4)
        const bool ok = !P5.cta_item.empty() && (ctx->schur_v5 == 2 || (P5.out_frac <= 0.15 && (long long)P5.items.size() >= 4ll * ctx->nsm)) && P5.blob.size() < (1ull << 31);
        if (getenv("NLLS_B200_VERBOSE"))
            fprintf(stderr, "[nlls] schur v5 plan: %s, %zu CTAs, %zu tiles, %lld super-tiles, %lld entries, %lld DMMAs, %lld flushes, %zu outlier points (%.2f %% of the contributions), imbalance %.3f, blob %.1f MB\n",
                    ok ? "in use" : "declined", P5.cta_item.empty() ? 0 : P5.cta_item.size() - 1, P5.items.size(), P5.n_super, P5.n_entries, P5.n_dmma, P5.n_flush, P5.outliers.size(),
                    100.0 * P5.out_frac, P5.imbalance, P5.blob.size() * 4e-6);
        if (ok) {
            ctx->n5cta = (int)P5.cta_item.size() - 1;
            TRY(upload(ctx, &ctx->d5_cta_item, P5.cta_item)); TRY(upload(ctx, &ctx->d5_items, P5.items)); TRY(upload(ctx, &ctx->d5_blob, P5.blob));
            const std::vector<long long> ft = (DC == 6) ? schur5_flush_table<6>(P5.super_base, tile_id, pos, ctx->NT) : schur5_flush_table<9>(P5.super_base, tile_id, pos, ctx->NT);
            TRY(upload(ctx, &ctx->d5_ftab, ft));
            out_pts = P5.outliers;
        }
    }
    // the fallback kernel's list: the plan's own outliers (their A_p^-1 comes from the plan's kernel), then the irregular points
    ctx->first_irr = (int)out_pts.size();
    for (int64_t p2 = 0; p2 < nB; ++p2) if (ctx->h_irr[(size_t)p2]) out_pts.push_back((int)p2);
    ctx->nout_pts = (int)out_pts.size();
    TRY(upload(ctx, &ctx->d_out_pts, out_pts));

    // ---- Schur v2 plan: larger point tiles + per-tile contribution lists sorted by target block
    ctx->nsuper = 0; ctx->nstiles = 0;
    if (ctx->schur_v2 && ctx->n5cta == 0) {
        std::vector<int> stile_pt;     // (first point, one past the last point) per tile; irregular points belong to no tile
        {
            auto aligned = [&](int64_t pt) { return ((WB * (int64_t)ctx->h_obs_start[(size_t)pt] + 9 * pt) & 1) == 0; };
            const int sobs = (DC <= 7) ? 232 : 128, spts = sobs / 2;   // == Schur4Cfg<DC>::OBS / PTS (<= SCH_OBS / SCH_PTS of the v2 kernel)
            int64_t p0 = 0;
            while (p0 < nB) {
                if (ctx->h_irr[(size_t)p0]) { ++p0; continue; }
                int64_t p1 = p0;
                int64_t contrib = 0;   // pairs (i, j <= i) of the tile: the v4 kernel stages the padded list in shared memory
                while (p1 < nB && !ctx->h_irr[(size_t)p1] && (p1 - p0) < spts && (ctx->h_obs_start[(size_t)p1 + 1] - ctx->h_obs_start[(size_t)p0]) <= sobs) {
                    const int64_t kk = ctx->h_obs_start[(size_t)p1 + 1] - ctx->h_obs_start[(size_t)p1];
                    if (p1 > p0 && contrib + kk * (kk + 1) / 2 > Schur4Cfg<6>::MAXENT - 4 * SCH4_WARPS) break;   // (+ quad padding per warp)
                    contrib += kk * (kk + 1) / 2;
                    ++p1;
                }
                if (p1 < nB && !aligned(p1) && p1 - 1 > p0 && aligned(p1 - 1)) --p1;
                if (p1 == p0) FAIL(NLLS_ERR_UNSUPPORTED, "a point with more observations than a Schur tile holds");
                stile_pt.push_back((int)p0); stile_pt.push_back((int)p1);
                p0 = p1;
            }
        }
        const int nst = (int)stile_pt.size() / 2;
        ctx->nstiles = nst;
        const int TC = ST / DC;
        const int NTl = ctx->NT;
        const long long nn = ctx->nred;
        struct TilePlan { std::vector<SchurChunk> chunks; std::vector<unsigned int> ents, ents4; std::vector<unsigned long long> bkey; std::vector<int> bstart; };
        std::vector<TilePlan> plans((size_t)nst);
        auto build = [&](int t0, int t1) {
            std::vector<unsigned long long> keys;
            for (int t = t0; t < t1; ++t) {
                const int pa = stile_pt[(size_t)2 * t], pb = stile_pt[(size_t)2 * t + 1];
                const int ob0 = ctx->h_obs_start[(size_t)pa];
                keys.clear();
                for (int pp = pa; pp < pb; ++pp) {
                    const int b = ctx->h_obs_start[(size_t)pp], e = ctx->h_obs_start[(size_t)pp + 1];
                    for (int i = b; i < e; ++i) for (int j = b; j <= i; ++j) {
                        const unsigned long long key = ((unsigned long long)ctx->h_obs_cam[(size_t)i] << 22) | (unsigned long long)ctx->h_obs_cam[(size_t)j];   // 22 + 22 + 10 + 10 bits
                        keys.push_back((key << 20) | ((unsigned long long)(i - ob0) << 10) | (unsigned long long)(j - ob0));   // SCH_OBS <= 1024
                    }
                }
                std::sort(keys.begin(), keys.end());
                TilePlan& pl = plans[(size_t)t];
                pl.ents.resize(keys.size()); pl.ents4.resize(keys.size());
                size_t k0 = 0;
                while (k0 < keys.size()) {
                    const unsigned long long key = keys[k0] >> 20;
                    size_t k1 = k0;
                    while (k1 < keys.size() && (keys[k1] >> 20) == key) ++k1;
                    pl.bkey.push_back(key); pl.bstart.push_back((int)k0);   // distinct blocks of the tile, in key order, with their first entry
                    const int cam = (int)(key >> 22), camj = (int)(key & 0x3fffffu);
                    long long soff; short flags = (cam == camj) ? 2 : 0;
                    if (ctx->s_tiled) {
                        const int I = cam / TC, Jt = camj / TC, r0 = (cam - I * TC) * DC, c0 = (camj - Jt * TC) * DC;
                        const int pI = pos[(size_t)I], pJ = pos[(size_t)Jt];
                        if (pI >= pJ) soff = (long long)tile_id[(size_t)pI * NTl + pJ] * ST2 + r0 + (long long)ST * c0;
                        else { soff = (long long)tile_id[(size_t)pJ * NTl + pI] * ST2 + c0 + (long long)ST * r0; flags |= 1; }
                    } else {
                        soff = (long long)cam * DC + nn * ((long long)camj * DC);
                    }
                    for (size_t k = k0; k < k1; k += SCH_CHUNK) {
                        SchurChunk ck;
                        ck.soff = soff; ck.ent0 = (int)k; ck.cam = cam; ck.n = (short)std::min<size_t>(SCH_CHUNK, k1 - k); ck.flags = flags;
                        pl.chunks.push_back(ck);
                    }
                    for (size_t k = k0; k < k1; ++k) {
                        const unsigned int il = (unsigned int)((keys[k] >> 10) & 0x3ffu), jl = (unsigned int)(keys[k] & 0x3ffu);
                        pl.ents[k] = (il << 16) | jl;
                        // v4: shared-memory byte offsets of W_i (inside the staged H span) and of Y_j
                        pl.ents4[k] = ((8u * (unsigned int)(WB * il + 9u * (unsigned int)(ctx->h_obs_pt[(size_t)ob0 + il] - pa))) << 16) | (8u * (unsigned int)(4 * (DC + 1) + 1) * jl);   // bytes
                    }
                    k0 = k1;
                }
                pl.bstart.push_back((int)keys.size());
                // long chunks first: the lanes of a warp then run loops of similar length
                std::stable_sort(pl.chunks.begin(), pl.chunks.end(), [](const SchurChunk& x, const SchurChunk& y) { return x.n > y.n; });
            }
        };
        if (nA >= (1 << 22)) FAIL(NLLS_ERR_UNSUPPORTED, "more than 2^22 cameras");
        {
            const int nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            std::vector<std::thread> th;
            for (int k = 0; k < nth; ++k) th.emplace_back(build, (int)((int64_t)nst * k / nth), (int)((int64_t)nst * (k + 1) / nth));
            for (auto& x : th) x.join();
        }
        // ---- Schur v4: one persistent CTA per SM walks a contiguous tile range, cut into super-tiles (runs of consecutive tiles
        //      whose distinct S blocks fit the warps' accumulator slots); see schur4_kernel
        const int NB4 = (DC <= 7) ? 16 : 4;   // == Schur4Cfg<DC>::NB
        const int cap4 = SCH4_WARPS * NB4;
        const int obs4 = (DC <= 7) ? 232 : 128, row4 = WB * obs4 + 9 * (obs4 / 2), ys4 = 4 * (DC + 1) + 1;   // == Schur4Cfg<DC>::OBS, ROW, YS
        std::vector<int> cta_item;
        std::vector<SchurItem> items;
        std::vector<SchurUnit> units;
        std::vector<unsigned int> wtab, blob;
        bool v4ok = ctx->schur_v4 != 0;
        ctx->nsuper = 0;
        if (v4ok) {
            // per tile: contributions and distinct blocks
            std::vector<long long> tile_groups((size_t)nst, 0);
            long long nblk_tiles = 0, ncontrib = 0;
            const unsigned int nullent = ((8u * (unsigned int)(row4 + 2)) << 16) | (8u * (unsigned int)(ys4 * obs4));   // byte offsets of the zero pads
            for (int t = 0; t < nst; ++t) {
                const TilePlan& pl = plans[(size_t)t];
                const long long cnt = pl.bstart.empty() ? 0 : pl.bstart.back();
                tile_groups[(size_t)t] = cnt;
                ncontrib += cnt; nblk_tiles += (long long)pl.bkey.size();
                if (cnt + 4 * SCH4_WARPS > Schur4Cfg<6>::MAXENT) v4ok = false;   // (very long tracks: v2 path)
            }
            const int wstride = cap4 + SCH4_WARPS;   // wtab row: groups per (warp, slot), then the first entry of every warp
            std::vector<int> blk_at_slot((size_t)cap4);
            // contiguous tile ranges of similar weight, one per CTA
            const int ncta = std::max(1, std::min(ctx->nsm, nst));
            std::vector<long long> wsum((size_t)nst + 1, 0);
            for (int t = 0; t < nst; ++t) wsum[(size_t)t + 1] = wsum[(size_t)t] + tile_groups[(size_t)t] + 64;
            std::vector<int> cta_tile((size_t)ncta + 1, nst);
            cta_tile[0] = 0;
            for (int c = 1; c < ncta; ++c)
                cta_tile[(size_t)c] = (int)(std::lower_bound(wsum.begin(), wsum.end(), wsum[(size_t)nst] * c / ncta) - wsum.begin());
            for (int c = 1; c <= ncta; ++c) cta_tile[(size_t)c] = std::max(cta_tile[(size_t)c], cta_tile[(size_t)c - 1]);
            int maxrun = 32;   // tiles per super-tile: longer runs drift across cameras (per-tile warp imbalance), shorter ones flush more often
            if (const char* g = getenv("NLLS_B200_SCHUR_RUN")) maxrun = std::max(1, atoi(g));
            cta_item.assign((size_t)ncta + 1, 0);
            std::vector<unsigned long long> cur, merged, keys_su;
            std::vector<long long> cnt_su;
            std::vector<int> order, slot_of, su_first;
            int nsu_total = 0;
            long long stat_tot = 0, stat_maxw = 0, stat_maxq = 0;
            for (int c = 0; c < ncta; ++c) {
                // super-tiles of this CTA's range
                su_first.clear();
                const int ta0 = cta_tile[(size_t)c], tb0 = cta_tile[(size_t)c + 1];
                cur.clear();
                bool wide = false;   // the open super-tile is a single oversize tile (it must stand alone)
                for (int t = ta0; t < tb0; ++t) {
                    const std::vector<unsigned long long>& bk = plans[(size_t)t].bkey;
                    const bool big = (int)bk.size() > cap4;
                    if (t == ta0 || wide || big) { su_first.push_back(t); cur = bk; wide = big; continue; }
                    merged.clear();
                    std::set_union(cur.begin(), cur.end(), bk.begin(), bk.end(), std::back_inserter(merged));
                    if ((int)merged.size() > cap4 || t - su_first.back() >= maxrun) { su_first.push_back(t); cur = bk; }
                    else cur.swap(merged);
                }
                su_first.push_back(tb0);
                for (size_t q = 0; q + 1 < su_first.size(); ++q) {
                    const int ta = su_first[q], tb = su_first[q + 1];
                    if (ta >= tb) continue;
                    ++nsu_total;
                    keys_su.clear();
                    for (int t = ta; t < tb; ++t) keys_su.insert(keys_su.end(), plans[(size_t)t].bkey.begin(), plans[(size_t)t].bkey.end());
                    std::sort(keys_su.begin(), keys_su.end());
                    keys_su.erase(std::unique(keys_su.begin(), keys_su.end()), keys_su.end());
                    const int nblk = (int)keys_su.size();
                    cnt_su.assign((size_t)nblk, 0);
                    for (int t = ta; t < tb; ++t) {
                        const TilePlan& pl = plans[(size_t)t];
                        for (size_t bb = 0; bb < pl.bkey.size(); ++bb) {
                            const int idx = (int)(std::lower_bound(keys_su.begin(), keys_su.end(), pl.bkey[bb]) - keys_su.begin());
                            cnt_su[(size_t)idx] += (pl.bstart[bb + 1] - pl.bstart[bb]) + 2;   // DMMAs + the slot's own overhead
                        }
                    }
                    // heaviest blocks first, each to the least-loaded warp of its round that still has a free slot
                    order.resize((size_t)nblk);
                    for (int k = 0; k < nblk; ++k) order[(size_t)k] = k;
                    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cnt_su[(size_t)x] > cnt_su[(size_t)y]; });
                    const int nrounds = std::max(1, (nblk + cap4 - 1) / cap4);
                    slot_of.assign((size_t)nblk, 0);
                    const int urow0 = (int)(units.size() / (size_t)cap4);
                    SchurUnit none; none.soff = 0; none.cam = 0; none.flags = 0;
                    units.resize(units.size() + (size_t)nrounds * cap4, none);
                    for (int r = 0; r < nrounds; ++r) {
                        long long load[SCH4_WARPS] = {0};
                        int used[SCH4_WARPS] = {0};
                        for (int k = r; k < nblk; k += nrounds) {   // round r takes every nrounds-th block of the sorted order
                            const int idx = order[(size_t)k];
                            int w = -1;
                            for (int q2 = 0; q2 < SCH4_WARPS; ++q2) if (used[q2] < NB4 && (w < 0 || load[q2] < load[w])) w = q2;
                            const int slot = r * cap4 + w * NB4 + used[w]++;
                            load[w] += cnt_su[(size_t)idx];
                            slot_of[(size_t)idx] = slot;
                            const unsigned long long key = keys_su[(size_t)idx];
                            const int cam = (int)(key >> 22), camj = (int)(key & 0x3fffffu);
                            SchurUnit u; u.cam = cam; u.flags = 8 | ((cam == camj) ? 2 : 0);
                            if (ctx->s_tiled) {
                                const int I = cam / TC, Jt = camj / TC, r0 = (cam - I * TC) * DC, c0 = (camj - Jt * TC) * DC;
                                const int pI = pos[(size_t)I], pJ = pos[(size_t)Jt];
                                if (pI >= pJ) u.soff = (long long)tile_id[(size_t)pI * NTl + pJ] * ST2 + r0 + (long long)ST * c0;
                                else { u.soff = (long long)tile_id[(size_t)pJ * NTl + pI] * ST2 + c0 + (long long)ST * r0; u.flags |= 1; }
                            } else {
                                u.soff = (long long)cam * DC + nn * ((long long)camj * DC);
                            }
                            units[(size_t)urow0 * cap4 + (size_t)slot] = u;
                        }
                    }
                    for (int r = 0; r < nrounds; ++r)
                        for (int t = ta; t < tb; ++t) {
                            const TilePlan& pl = plans[(size_t)t];
                            SchurItem it;
                            it.pt0 = stile_pt[(size_t)2 * t]; it.npt = stile_pt[(size_t)2 * t + 1] - it.pt0;
                            it.ob0 = ctx->h_obs_start[(size_t)it.pt0]; it.nob = ctx->h_obs_start[(size_t)stile_pt[(size_t)2 * t + 1]] - it.ob0;
                            it.wrow = (int)(wtab.size() / (size_t)wstride);
                            it.urow = (t == tb - 1) ? urow0 + r : -1;
                            it.flags = ((t == ta) ? 1 : 0) | ((r == 0) ? 2 : 0) |
                                       ((((long long)DC * DC * nA + (long long)WB * it.ob0 + 9ll * it.pt0) & 1) ? 4 : 0);   // bit2: the H span starts 8 bytes off a 16-byte boundary
                            it.pad0 = it.pad1 = it.pad2 = 0;
                            // the item's blob: [per-observation table | contribution entries in (warp, slot) order, groups of four]
                            it.blob0 = (int)blob.size();
                            for (int i = it.ob0; i < it.ob0 + it.nob; ++i) {
                                const int pp = ctx->h_obs_pt[(size_t)i];
                                const unsigned int first = (i == ctx->h_obs_start[(size_t)pp]) ? 0x80000000u : 0u;
                                blob.push_back(first | ((unsigned int)(ctx->h_obs_start[(size_t)pp + 1] - it.ob0) << 16) | (unsigned int)(pp - it.pt0));
                            }
                            while (blob.size() & 3) blob.push_back(0u);
                            const size_t eb = blob.size();
                            std::fill(blk_at_slot.begin(), blk_at_slot.end(), -1);
                            for (size_t bb = 0; bb < pl.bkey.size(); ++bb) {
                                const int idx = (int)(std::lower_bound(keys_su.begin(), keys_su.end(), pl.bkey[bb]) - keys_su.begin());
                                const int slot = slot_of[(size_t)idx];
                                if (slot / cap4 == r) blk_at_slot[(size_t)(slot - r * cap4)] = (int)bb;
                            }
                            const size_t wb = wtab.size();
                            wtab.resize(wb + (size_t)wstride, 0u);
                            long long wl[SCH4_WARPS] = {0}, ql[4] = {0}, tot = 0, mx = 0, mq = 0;
                            for (int w = 0; w < SCH4_WARPS; ++w) {
                                while ((blob.size() - eb) & 3) blob.push_back(nullent);   // every warp's run starts on a quad of entries
                                wtab[wb + (size_t)cap4 + (size_t)w] = (unsigned int)(blob.size() - eb);
                                for (int b2 = 0; b2 < NB4; ++b2) {
                                    const int bb = blk_at_slot[(size_t)(w * NB4 + b2)];
                                    if (bb < 0) continue;
                                    const long long nc = pl.bstart[(size_t)bb + 1] - pl.bstart[(size_t)bb];
                                    for (int k = pl.bstart[(size_t)bb]; k < pl.bstart[(size_t)bb + 1]; ++k) blob.push_back(pl.ents4[(size_t)k]);
                                    wtab[wb + (size_t)(w * NB4 + b2)] = (unsigned int)nc;
                                    wl[w] += nc; ql[w & 3] += nc; tot += nc;
                                }
                            }
                            while ((blob.size() - eb) & 3) blob.push_back(nullent);
                            it.ne4 = (int)(blob.size() - eb);
                            for (int k = 0; k < 8; ++k) blob.push_back(nullent);   // the kernel's prefetch runs up to two quads past a warp's run
                            if (blob.size() >= (1ull << 31)) FAIL(NLLS_ERR_UNSUPPORTED, "too many Schur contributions per rank");
                            items.push_back(it);
                            // balance statistics: the slowest warp (and the slowest scheduler: warp % 4) sets the pace of a tile
                            for (int w = 0; w < SCH4_WARPS; ++w) mx = std::max(mx, wl[w]);
                            for (int q2 = 0; q2 < 4; ++q2) mq = std::max(mq, ql[q2]);
                            stat_tot += tot; stat_maxw += mx * SCH4_WARPS; stat_maxq += mq * 4;
                        }
                }
                cta_item[(size_t)c + 1] = (int)items.size();
            }
            // blocks that recur too rarely within a tile (unsorted points) make the per-slot bookkeeping and the final reductions
            // dominate: keep the v2 path then
            const double reuse = (double)ncontrib / std::max(1.0, (double)nblk_tiles);
            if (getenv("NLLS_B200_VERBOSE"))
                fprintf(stderr, "[nlls] schur v4 plan: %d CTAs, %d super-tiles over %d tiles, %zu items, %.2f contributions per (tile, block), per-tile imbalance: warp %.2f scheduler %.2f\n", ncta, nsu_total, nst,
                        items.size(), reuse, (double)stat_maxw / std::max(1ll, stat_tot), (double)stat_maxq / std::max(1ll, stat_tot));
            if (reuse < 1.5 && ctx->schur_v4 != 2) v4ok = false;
            if (v4ok) ctx->nsuper = ncta;
        }
        if (ctx->nsuper == 0) {   // v2 structures
            std::vector<int> chunk_off((size_t)nst + 1, 0), ent_off((size_t)nst + 1, 0);
            size_t nent = 0, nch = 0;
            for (int t = 0; t < nst; ++t) { nent += plans[(size_t)t].ents.size(); nch += plans[(size_t)t].chunks.size(); }
            if (nent >= (1ull << 31)) FAIL(NLLS_ERR_UNSUPPORTED, "too many Schur contributions per rank");
            std::vector<SchurChunk> chunks; chunks.reserve(nch);
            std::vector<unsigned int> ents; ents.reserve(nent);
            for (int t = 0; t < nst; ++t) {
                const int base = (int)ents.size();
                for (SchurChunk ck : plans[(size_t)t].chunks) { ck.ent0 += base; chunks.push_back(ck); }
                ents.insert(ents.end(), plans[(size_t)t].ents.begin(), plans[(size_t)t].ents.end());
                chunk_off[(size_t)t + 1] = (int)chunks.size();
                ent_off[(size_t)t + 1] = (int)ents.size();
                plans[(size_t)t] = TilePlan();
            }
            if (getenv("NLLS_B200_VERBOSE"))
                fprintf(stderr, "[nlls] schur v2 plan: %d tiles, %zu contributions, %zu chunks (%.2f contributions per chunk)\n", nst, nent, nch, (double)nent / std::max<size_t>(nch, 1));
            TRY(upload(ctx, &ctx->d_stile_pt, stile_pt)); TRY(upload(ctx, &ctx->d_chunk_off, chunk_off)); TRY(upload(ctx, &ctx->d_ent_off, ent_off));
            TRY(upload(ctx, &ctx->d_chunks, chunks)); TRY(upload(ctx, &ctx->d_ents, ents));
        }
        if (ctx->nsuper > 0) {
            TRY(upload(ctx, &ctx->d_cta_item, cta_item)); TRY(upload(ctx, &ctx->d_items, items));
            TRY(upload(ctx, &ctx->d_units, units)); TRY(upload(ctx, &ctx->d_wtab, wtab)); TRY(upload(ctx, &ctx->d_blob, blob));
        }
    }

    TRY(upload(ctx, &ctx->d_obs_cam, ctx->h_obs_cam)); TRY(upload(ctx, &ctx->d_obs_pt, ctx->h_obs_pt)); TRY(upload(ctx, &ctx->d_obs_z, obs_z));
    TRY(upload(ctx, &ctx->d_obs_start, ctx->h_obs_start)); TRY(upload(ctx, &ctx->d_tile_pt, ctx->h_tile_pt));
    {
        std::vector<int4> tiles((size_t)ctx->ntiles);
        for (int t = 0; t < ctx->ntiles; ++t) {
            const int a = ctx->h_tile_pt[(size_t)2 * t], b = ctx->h_tile_pt[(size_t)2 * t + 1];
            tiles[(size_t)t] = make_int4(a, b - a, ctx->h_obs_start[(size_t)a], ctx->h_obs_start[(size_t)b] - ctx->h_obs_start[(size_t)a]);
        }
        TRY(upload(ctx, &ctx->d_tiles, tiles));
    }
    TRY(upload(ctx, &ctx->d_cm_pt, cm_pt)); TRY(upload(ctx, &ctx->d_cm_z, cm_z));
    TRY(upload(ctx, &ctx->d_item_cam, item_cam)); TRY(upload(ctx, &ctx->d_item_beg, item_beg)); TRY(upload(ctx, &ctx->d_item_end, item_end));
    TRY(upload(ctx, &ctx->d_cam_item_start, cam_item_start));
    for (int k = 0; k < 3; ++k) { TRY(dalloc(ctx, &ctx->d_A[k], (size_t)nA * ctx->CS)); TRY(dalloc(ctx, &ctx->d_B[k], (size_t)nB * 3)); }
    ctx->cur = 0; ctx->nxt = 1; ctx->bst = 2;
    TRY(upload_vars(ctx, A, ctx->CS, ctx->d_A[0])); TRY(upload_vars(ctx, B, 3, ctx->d_B[0]));
    TRY(dalloc(ctx, &ctx->d_H, (size_t)ctx->hlen + 2));   // + slack: the Schur kernel's 16-byte-aligned bulk loads may read one element past a span
    TRY(dalloc(ctx, &ctx->d_g, (size_t)ctx->dof + 2)); TRY(dalloc(ctx, &ctx->d_x, (size_t)ctx->dof));   // + slack: 16-byte-aligned bulk loads of g_p
    CK(cudaMemsetAsync(ctx->d_H, 0, sizeof(double) * (ctx->hlen + 2), ctx->st));
    CK(cudaMemsetAsync(ctx->d_g, 0, sizeof(double) * (ctx->dof + 2), ctx->st));
    CK(cudaMemsetAsync(ctx->d_x, 0, sizeof(double) * ctx->dof, ctx->st));
    TRY(dalloc(ctx, &ctx->d_Ainv, (size_t)6 * nB));
    if (ctx->s_tiled) {
        TRY(dalloc(ctx, &ctx->d_S, (size_t)ctx->ntiles_alloc * ST2)); TRY(dalloc(ctx, &ctx->d_rhs, (size_t)ctx->NT * ST));
        TRY(dalloc(ctx, &ctx->d_Linv, (size_t)ctx->NT * ST2)); TRY(dalloc(ctx, &ctx->d_xp, (size_t)ctx->NT * ST));
        TRY(upload(ctx, &ctx->d_tile_id, tile_id)); TRY(upload(ctx, &ctx->d_pos, pos));
        TRY(upload(ctx, &ctx->d_diag_tile, diag_tile)); TRY(upload(ctx, &ctx->d_diag_tile_nat, diag_tile_nat));
        TRY(upload(ctx, &ctx->d_add_u, add_u_tiles));
        TRY(upload(ctx, &ctx->d_red_tasks, red_tasks)); TRY(upload(ctx, &ctx->d_red_upds, red_upds)); TRY(upload(ctx, &ctx->d_red_targets, red_targets));
        CK(cudaMemsetAsync(ctx->d_Linv, 0, sizeof(double) * (size_t)ctx->NT * ST2, ctx->st));
        TRY(upload(ctx, &ctx->d_lvl_cols, lvl_cols_flat));
        {
            std::vector<int> bwd_order;   // columns by level, last level first
            for (size_t l = ctx->lvl_cols.size(); l-- > 0;)
                for (int q = 0; q < ctx->lvl_cols[l].second; ++q) bwd_order.push_back(lvl_cols_flat[(size_t)(ctx->lvl_cols[l].first + q)]);
            TRY(upload(ctx, &ctx->d_bwd_order, bwd_order));
            TRY(dalloc(ctx, &ctx->d_bwd_flags, (size_t)ctx->NT));
            std::vector<int> nat_of((size_t)ctx->NT);
            for (int o = 0; o < ctx->NT; ++o) nat_of[(size_t)pos[(size_t)o]] = o;
            TRY(upload(ctx, &ctx->d_nat_of_pos, nat_of));
        }
        TRY(upload(ctx, &ctx->d_colptr, colptr)); TRY(upload(ctx, &ctx->d_col_tile, col_tile)); TRY(upload(ctx, &ctx->d_col_row, col_row));
    }
    TRY(dalloc(ctx, &ctx->d_cost_part, (size_t)ctx->ntiles + ctx->nlong)); TRY(dalloc(ctx, &ctx->d_step_part, (size_t)4 * (ctx->ntiles + ctx->nlong)));
    const int NU = DC * (DC + 1) / 2 + DC;
    TRY(dalloc(ctx, &ctx->d_cam_part, (size_t)ctx->nitems * (NU + 1))); TRY(dalloc(ctx, &ctx->d_cam_part2, (size_t)ctx->nitems * (NU + 1)));
    ctx->cam_part_vars = -1;
    TRY(dalloc(ctx, &ctx->d_camstat_part, (size_t)4 * ((nA + 127) / 128)));
    TRY(DISPATCH(ctx, set_smem_attrs, ctx));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->prepared = true;
    ctx->lm_active = false;
    return apply_unfixed(ctx);
}

int nlls_linearize(nlls_ctx* ctx, double* cost) {
    if (!ctx) return NLLS_ERR_INVALID;
    TRY(nlls_prepare(ctx));
    CK(cudaSetDevice(ctx->device));
    return do_linearize(ctx, cost);
}

// optimize(kernel, squarederrors, maxiters) of src/robustadaptive.jl:48-73 on the device: refits the adaptive kernel variable of
// buffer `which` (0 variables, 1 varnext, 2 varbest) in place from the squared residuals at that buffer's means.
int nlls_adaptive_em(nlls_ctx* ctx, int which, int maxiters) {
    if (!ctx || which < 0 || which > 2 || maxiters < 0) return NLLS_ERR_INVALID;
    TRY(nlls_prepare(ctx));
    if (!ctx->adaptive) FAIL(NLLS_ERR_INVALID, "nlls_adaptive_em: the problem has no adaptive robust kernel");
    CK(cudaSetDevice(ctx->device));
    const int buf = which == 0 ? ctx->cur : (which == 1 ? ctx->nxt : ctx->bst);
    const AdaptDev p = adaptdev(ctx);
    if (!ctx->d_em) TRY(dalloc(ctx, &ctx->d_em, (size_t)8 + 4 * (size_t)std::max(ctx->ad_nchunks, 1)));
    double *state = ctx->d_em, *sums = ctx->d_em + 4, *part = ctx->d_em + 8;
    adapt_em_init_kernel<<<1, 32, 0, ctx->st>>>(ctx->d_A[buf], state); ctx->launches++;
    for (int it = 0; it < maxiters; ++it) {
        if (ctx->ad_nchunks > 0) { adapt_em_partial_kernel<<<ctx->ad_nchunks, AD_THREADS, 0, ctx->st>>>(p, ctx->d_A[buf], ctx->d_B[buf], state, part); ctx->launches++; }
        adapt_em_sums_kernel<<<1, 32, 0, ctx->st>>>(part, ctx->ad_nchunks, state, sums); ctx->launches++;
        TRY(allreduce(ctx, sums, 4, ncclSum));   // residuals are sharded over the ranks; `done` is replicated, so every rank takes part
        adapt_em_update_kernel<<<1, 32, 0, ctx->st>>>(sums, ctx->d_A[buf], state); ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->st));
    for (auto& kv : ctx->vars) kv.second.stale = true;
    return NLLS_OK;
}

int nlls_cost(nlls_ctx* ctx, int which, double* cost) {
    if (!ctx || which < 0 || which > 2) return NLLS_ERR_INVALID;
    TRY(nlls_prepare(ctx));
    CK(cudaSetDevice(ctx->device));
    const int buf = which == 0 ? ctx->cur : (which == 1 ? ctx->nxt : ctx->bst);
    if (ctx->adaptive) TRY(adapt_cost(ctx, buf, SC_COST_TRY));
    else TRY(DISPATCH(ctx, launch_cost, ctx, buf, SC_COST_TRY, ctx->d_cam_part2));
    TRY(fetch_scalars(ctx));
    if (cost) *cost = ctx->h_scal[SC_COST_TRY];
    if (ctx->lm_active && ctx->lm_phase == 1) ctx->costcomputations += 1;   // a callback re-evaluating the cost (test/adaptivecost.jl:21-22)
    return NLLS_OK;
}

int nlls_solve(nlls_ctx* ctx, double lambda) {
    if (!ctx || !ctx->prepared) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->adaptive) { TRY(adapt_solve_update(ctx, lambda)); return fetch_scalars(ctx); }
    ctx->cam_part_vars = -1;   // varnext is rewritten without a cost evaluation
    TRY(DISPATCH(ctx, launch_schur, ctx, lambda));
    TRY(launch_reduced_solve(ctx));
    // back-substitution also writes x and varnext; nlls_update is then a no-op kept for API symmetry
    TRY(DISPATCH(ctx, launch_update, ctx));
    CK(cudaMemsetAsync(ctx->d_scal + SC_COST_TRY, 0, sizeof(double), ctx->st));
    TRY(exchange_try_scalars(ctx));
    TRY(fetch_scalars(ctx));
    return NLLS_OK;
}

int nlls_update(nlls_ctx* ctx) {
    if (!ctx || !ctx->prepared) return NLLS_ERR_INVALID;
    return NLLS_OK;  // varnext is produced by the fused back-substitution in nlls_solve
}

// ---- LM -----------------------------------------------------------------------------------------------
int nlls_lm_begin(nlls_ctx* ctx, const nlls_options* opts) {
    if (!ctx || !opts) return NLLS_ERR_INVALID;
    const uint64_t t0 = now_ns();
    if (opts->iterator < NLLS_ITER_NEWTON || opts->iterator > NLLS_ITER_GD) FAIL(NLLS_ERR_INVALID, "unknown iterator");
    TRY(nlls_prepare(ctx));
    CK(cudaSetDevice(ctx->device));
    if (opts->iterator == NLLS_ITER_DOGLEG || opts->iterator == NLLS_ITER_GD) {
        if (ctx->adaptive) FAIL(NLLS_ERR_UNSUPPORTED, "Dogleg / gradient descent are implemented for the bundle-adjustment residuals");
        if (ctx->nranks > 1) FAIL(NLLS_ERR_UNSUPPORTED, "Dogleg / gradient descent run on one rank");
        const size_t npart = std::max<size_t>(2 * ((size_t)(ctx->nA + ctx->nB + 127) / 128), VEC_GRID);
        TRY(dalloc(ctx, &ctx->d_cauchy, (size_t)ctx->dof)); TRY(dalloc(ctx, &ctx->d_vec_part, npart));
    }
    ctx->trustradius = 0.0; ctx->stepsize = 1.0;                 // DoglegData / GradientDescentData  src/iterators.jl:35,184
    ctx->opts = *opts;
    ctx->starttime = t0;
    ctx->stoptime = t0 + opts->maxtime_ns;                       // src/optimize.jl:115
    ctx->lambda = 0.0;                                           // LevMarData(0.0)  src/iterators.jl:124
    ctx->fails = 0; ctx->iternum = 0; ctx->converged = 0;
    ctx->costcomputations = ctx->gradientcomputations = ctx->linearsolvers = 0;
    ctx->t_init = ctx->t_cost = ctx->t_grad = ctx->t_solver = 0;
    ctx->have_best = false; ctx->lm_phase = 0;
    if (ctx->adaptive && ctx->masked) {   // varnext = deepcopy(variables) (src/optimize.jl:80-82): update! never writes its fixed entries
        CK(cudaMemcpyAsync(ctx->d_A[ctx->nxt], ctx->d_A[ctx->cur], sizeof(double) * ctx->nA * ctx->CS, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->d_B[ctx->nxt], ctx->d_B[ctx->cur], sizeof(double) * ctx->nB * ctx->BS, cudaMemcpyDeviceToDevice, ctx->st));
    }
    ctx->t_init += now_ns() - t0;                                // :116
    const uint64_t tg = now_ns();
    double c = 0.0;
    TRY(do_linearize(ctx, &c));                                  // :118
    ctx->t_grad += now_ns() - tg;
    ctx->gradientcomputations += 1;
    ctx->cost = c; ctx->bestcost = c; ctx->startcost = c;        // :120-121
    ctx->lm_active = true;
    return NLLS_OK;
}

int nlls_lm_iterate(nlls_ctx* ctx, nlls_iterinfo* info) {
    if (!ctx || !ctx->lm_active) return NLLS_ERR_INVALID;
    // call order is begin -> (iterate -> advance)* -> end: after a non-zero termination word the linear system belongs to the
    // previous point (advance skipped the re-linearisation, src/optimize.jl:165-171), so another iteration would be silently wrong
    if (ctx->converged != 0) FAIL(NLLS_ERR_INVALID, "nlls_lm_iterate after termination (converged != 0): call nlls_lm_end");
    if (ctx->lm_phase != 0) FAIL(NLLS_ERR_INVALID, "nlls_lm_iterate called twice without nlls_lm_advance");
    ctx->lm_phase = 1;
    CK(cudaSetDevice(ctx->device));
    const nlls_options& o = ctx->opts;
    ctx->iternum += 1;                                           // src/optimize.jl:124
    if (o.iterator == NLLS_ITER_NEWTON) {
        // ---- iterate!(::NewtonData)                               src/iterators.jl:17-27: undamped solve, update, cost — no
        // acceptance test; the outer loop's best / fails bookkeeping (nlls_lm_advance) handles a cost increase
        TRY(do_try(ctx, 0.0));
        ctx->linearsolvers += 1; ctx->costcomputations += 1;
        const double* s = ctx->h_scal;
        const double cost_ = s[SC_COST_TRY];
        const double maxstep = (std::isnan(s[SC_P_MAX]) || std::isnan(s[SC_C_MAX])) ? std::numeric_limits<double>::quiet_NaN() : std::max(s[SC_P_MAX], s[SC_C_MAX]);
        ctx->cost = cost_;
        ctx->maxstep = maxstep;
        if (info) { info->cost = cost_; info->lambda = 0.0; info->maxstep = maxstep; info->stepnorm = std::sqrt(s[SC_P_SQ] + s[SC_C_SQ]); info->ntries = 1; info->accepted = !(cost_ > ctx->bestcost); }
        return NLLS_OK;
    }
    if (o.iterator == NLLS_ITER_DOGLEG) return DISPATCH(ctx, iterate_dogleg, ctx, info);
    if (o.iterator == NLLS_ITER_GD) return DISPATCH(ctx, iterate_gd, ctx, info);
    // ---- iterate!(::LevMarData)                                 src/iterators.jl:139-172
    if (ctx->lambda == 0) {                                      // initlambda  :131-137,142-144
        if (ctx->adaptive) { adapt_maxdiag_kernel<<<1, 32, 0, ctx->st>>>(ctx->d_H, (int)ctx->dof, ctx->d_scal + SC_MAXDIAG, ctx->masked ? ctx->d_fixdof : nullptr); ctx->launches++; CK(cudaGetLastError()); }
        else TRY(DISPATCH(ctx, launch_maxdiag, ctx));
        TRY(fetch_scalars(ctx));
        ctx->lambda = ctx->h_scal[SC_MAXDIAG] * 1e-6;
    }
    double mu = 2.0, cost_ = 0.0, maxstep = 0.0, sq = 0.0;
    int64_t ntries = 0, accepted = 0;
    while (true) {
        TRY(do_try(ctx, ctx->lambda));                           // :149-157 (damp, solve, negate, update, cost); books timesolver / timecost
        ctx->linearsolvers += 1; ctx->costcomputations += 1; ntries += 1;
        const double* s = ctx->h_scal;
        cost_ = s[SC_COST_TRY];
        maxstep = (std::isnan(s[SC_P_MAX]) || std::isnan(s[SC_C_MAX])) ? std::numeric_limits<double>::quiet_NaN() : std::max(s[SC_P_MAX], s[SC_C_MAX]);
        sq = s[SC_P_SQ] + s[SC_C_SQ];
        if (!(cost_ > ctx->bestcost) || maxstep < o.dstep) {     // :160
            accepted = !(cost_ > ctx->bestcost);
            const double xhx = s[SC_P_XHX] + s[SC_C_XHX], gx = s[SC_P_GX] + s[SC_C_GX];
            const double q = (cost_ - ctx->bestcost) / (0.5 * xhx + gx);   // :163
            const double t = 2 * q - 1;
            ctx->lambda *= q < 0.983 ? 1 - t * t * t : 0.1;     // :164
            break;
        }
        ctx->lambda *= mu;                                       // :169-170
        mu *= 2.0;
    }
    ctx->cost = cost_;
    ctx->maxstep = maxstep;
    if (info) { info->cost = cost_; info->lambda = ctx->lambda; info->maxstep = maxstep; info->stepnorm = std::sqrt(sq); info->ntries = ntries; info->accepted = accepted; }
    return NLLS_OK;
}

static int lm_advance(nlls_ctx* ctx, double cost, int64_t terminate, int64_t* converged, bool collective_terminate) {
    if (!ctx || !ctx->lm_active) return NLLS_ERR_INVALID;
    if (ctx->lm_phase != 1) FAIL(NLLS_ERR_INVALID, "nlls_lm_advance without a preceding nlls_lm_iterate");
    ctx->lm_phase = 0;
    CK(cudaSetDevice(ctx->device));
    if (collective_terminate && ctx->nranks > 1) {
        // The callback's flag is the caller's: one rank's callback may ask to stop while the others' do not.  A rank that left the loop
        // alone would strand the others in the next all-reduce, so the flag is combined (maximum) over the ranks: one 8-byte
        // all-reduce per iteration, only on this callback path (nlls_optimize / nlls_lm_step pass the null callback's 0).
        ctx->h_scal[SC_COUNT + 1] = (double)terminate;
        CK(cudaMemcpyAsync(ctx->d_scal + SC_EXCH + 2, ctx->h_scal + SC_COUNT + 1, sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        TRY(allreduce(ctx, ctx->d_scal + SC_EXCH + 2, 1, ncclMax));
        CK(cudaMemcpyAsync(ctx->h_scal + SC_COUNT + 1, ctx->d_scal + SC_EXCH + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        terminate = (int64_t)ctx->h_scal[SC_COUNT + 1];
    }
    const nlls_options& o = ctx->opts;
    const double maxstep = ctx->maxstep;
    // ---- optimizeinternal! after the callback                    src/optimize.jl:130-171
    double dcost = ctx->bestcost - cost;
    if (dcost >= 0) { ctx->bestcost = cost; ctx->fails = 0; }
    else {
        dcost = cost;                                            // sic :135
        ctx->fails += 1;
        if (ctx->fails == 1) {                                   // :137-144: keep the current (best) variables
            CK(cudaMemcpyAsync(ctx->d_A[ctx->bst], ctx->d_A[ctx->cur], sizeof(double) * ctx->nA * ctx->CS, cudaMemcpyDeviceToDevice, ctx->st));
            CK(cudaMemcpyAsync(ctx->d_B[ctx->bst], ctx->d_B[ctx->cur], sizeof(double) * ctx->nB * ctx->BS, cudaMemcpyDeviceToDevice, ctx->st));
            ctx->have_best = true;
        }
    }
    std::swap(ctx->cur, ctx->nxt);                               // updatefromnext!  :147,207-209
    for (auto& kv : ctx->vars) kv.second.stale = true;           // the device now holds newer variables than the host mirror
    ctx->cost = cost;
    int64_t conv = 0;                                            // :150-161
    conv |= (int64_t)std::isinf(cost) << 0;
    conv |= (int64_t)std::isnan(cost) << 1;
    conv |= (int64_t)(dcost < ctx->bestcost * o.reldcost) << 2;
    conv |= (int64_t)(dcost < o.absdcost) << 3;
    conv |= (int64_t)std::isinf(maxstep) << 4;
    conv |= (int64_t)std::isnan(maxstep) << 5;
    conv |= (int64_t)(maxstep < o.dstep) << 6;
    conv |= (int64_t)(ctx->fails > o.maxfails) << 7;
    conv |= (int64_t)(ctx->iternum >= o.maxiters) << 8;
    // Multi-rank: every rank must take the same decision (a rank that leaves alone strands the others in the next all-reduce).  All
    // inputs of the termination word are replicated except the clock — the OR of the ranks' readings came with the try's scalars —
    // and the callback's flag (combined over the ranks above).
    const int64_t timeup = (ctx->nranks > 1) ? (ctx->h_scal[SC_TIMEUP] != 0.0) : (now_ns() > ctx->stoptime);
    conv |= timeup << 9;
    conv |= terminate << 16;
    ctx->converged = conv;
    if (converged) *converged = conv;
    if (conv == 0) {                                             // :167-171
        const uint64_t tg = now_ns();
        TRY(do_linearize(ctx, nullptr));
        ctx->t_grad += now_ns() - tg;
        ctx->gradientcomputations += 1;
    }
    return NLLS_OK;
}

int nlls_lm_advance(nlls_ctx* ctx, double cost, int64_t terminate, int64_t* converged) { return lm_advance(ctx, cost, terminate, converged, true); }

// One outer iteration with the null callback (src/callbacks.jl:20): nlls_lm_iterate + nlls_lm_advance(info.cost, 0) — what nlls_optimize
// runs per iteration, for callers that want the per-iteration record without a callback's flag to exchange.
int nlls_lm_step(nlls_ctx* ctx, nlls_iterinfo* info, int64_t* converged) {
    nlls_iterinfo local;
    if (!info) info = &local;
    TRY(nlls_lm_iterate(ctx, info));
    return lm_advance(ctx, info->cost, 0, converged, false);
}

int nlls_lm_end(nlls_ctx* ctx, nlls_result* r) {
    if (!ctx || !ctx->lm_active) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (!(ctx->bestcost >= ctx->cost) && ctx->have_best) std::swap(ctx->cur, ctx->bst);   // src/optimize.jl:173-176
    for (auto& kv : ctx->vars) kv.second.stale = true;
    CK(cudaStreamSynchronize(ctx->st));
    if (ctx->grad_pending) {
        float gms = 0.f;
        CK(cudaEventElapsedTime(&gms, ctx->ev_g0, ctx->ev_g1));
        ctx->t_grad += (uint64_t)(gms * 1e6);
        ctx->grad_pending = false;
    }
    ctx->lm_active = false;
    if (r) {
        r->startcost = ctx->startcost; r->bestcost = ctx->bestcost;
        r->timetotal = (now_ns() - ctx->starttime) * 1e-9;       // :178
        r->timeinit = ctx->t_init * 1e-9; r->timecost = ctx->t_cost * 1e-9; r->timegradient = ctx->t_grad * 1e-9; r->timesolver = ctx->t_solver * 1e-9;
        r->termination = ctx->converged; r->niterations = ctx->iternum;
        r->costcomputations = ctx->costcomputations; r->gradientcomputations = ctx->gradientcomputations; r->linearsolvers = ctx->linearsolvers;
    }
    return NLLS_OK;
}

int nlls_optimize(nlls_ctx* ctx, const nlls_options* opts, nlls_result* result) {
    TRY(nlls_lm_begin(ctx, opts));
    int64_t conv = 0;
    while (conv == 0) {
        nlls_iterinfo info;
        TRY(nlls_lm_iterate(ctx, &info));
        TRY(lm_advance(ctx, info.cost, 0, &conv, false));        // nullcallback returns (cost, 0)  src/callbacks.jl:20
    }
    return nlls_lm_end(ctx, result);
}

// ---- read-back -----------------------------------------------------------------------------------------
int nlls_get_variables(nlls_ctx* ctx, int vartype, int which, double* aos, int64_t n, int64_t stride) {
    if (!ctx || !ctx->prepared || !aos || which < 0 || which > 2) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int buf = which == 0 ? ctx->cur : (which == 1 ? ctx->nxt : ctx->bst);
    const bool isA = vartype == ctx->vtA;
    if (!isA && vartype != ctx->vtB) return NLLS_ERR_INVALID;
    const int64_t cnt = isA ? ctx->nA : ctx->nB;
    const int ds = isA ? ctx->CS : ctx->BS, ns = isA ? ctx->NC : ctx->BS;
    if (n != cnt || stride < ns) FAIL(NLLS_ERR_INVALID, "size mismatch in nlls_get_variables");
    if (ds == stride) {
        CK(cudaMemcpyAsync(aos, isA ? ctx->d_A[buf] : ctx->d_B[buf], sizeof(double) * cnt * ds, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    } else {
        std::vector<double> tmp((size_t)cnt * ds);
        CK(cudaMemcpyAsync(tmp.data(), isA ? ctx->d_A[buf] : ctx->d_B[buf], sizeof(double) * cnt * ds, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        for (int64_t i = 0; i < cnt; ++i) std::memcpy(aos + (size_t)i * stride, &tmp[(size_t)i * ds], sizeof(double) * ns);
    }
    return NLLS_OK;
}

int64_t nlls_dof(nlls_ctx* ctx) {   // length of linsystem.b / x: the unfixed variables' DoF
    if (!ctx || !ctx->prepared) return -1;
    if (!ctx->masked || ctx->adaptive) return ctx->dof;   // (adaptive problems keep the full dense system: fixed entries of b / x are zero)
    int64_t d = 0;
    for (unsigned char f : ctx->h_fixA) if (!f) d += ctx->DC;
    for (unsigned char f : ctx->h_fixB) if (!f) d += 3;
    return d;
}
int64_t nlls_hessian_len(nlls_ctx* ctx) { return (ctx && ctx->prepared) ? ctx->hlen : -1; }
int64_t nlls_hessian_nblocks(nlls_ctx* ctx) { return (ctx && ctx->prepared && !ctx->adaptive) ? ctx->nA + ctx->nB + ctx->nobs : -1; }

}  // extern "C"

namespace {
// vector in reference (variable-index) order from the internal [cameras | points] order
int reorder_vec(nlls_ctx* ctx, const double* dsrc, double* out) {
    std::vector<double> tmp((size_t)ctx->dof);
    CK(cudaMemcpyAsync(tmp.data(), dsrc, sizeof(double) * ctx->dof, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (ctx->cams_first && !ctx->masked) { std::memcpy(out, tmp.data(), sizeof(double) * ctx->dof); return NLLS_OK; }
    const VarSet& A = ctx->vars[ctx->vtA];
    const VarSet& B = ctx->vars[ctx->vtB];
    size_t ia = 0, ib = 0, o = 0;
    const int DC = ctx->DC;
    while (ia < A.gidx.size() || ib < B.gidx.size()) {
        const bool takeA = ib >= B.gidx.size() || (ia < A.gidx.size() && A.gidx[ia] < B.gidx[ib]);
        // (fixed variables have no block in the reference's linear system: they are skipped)
        if (takeA) { if (!(ctx->masked && ctx->h_fixA[ia])) { std::memcpy(out + o, &tmp[(size_t)DC * ia], sizeof(double) * DC); o += DC; } ++ia; }
        else { if (!(ctx->masked && ctx->h_fixB[ib])) { std::memcpy(out + o, &tmp[(size_t)DC * ctx->nA + 3 * ib], sizeof(double) * 3); o += 3; } ++ib; }
    }
    return NLLS_OK;
}

// Walk the reference's BlockSparseMatrix in storage order (block rows by variable index, blocks by ascending block column,
// src/BlockSparseMatrix.jl:37-44, src/linearsystem.jl:108-115) and hand every block to `emit`.
template <class F>
void walk_reference_blocks(nlls_ctx* ctx, F emit) {
    const VarSet& A = ctx->vars[ctx->vtA];
    const VarSet& B = ctx->vars[ctx->vtB];
    const int DC = ctx->DC, WB = 3 * DC;
    const int64_t hB = (int64_t)DC * DC * ctx->nA;
    // block index (1-based rank among all variables) of every camera / point
    std::vector<int64_t> blkA(A.gidx.size()), blkB(B.gidx.size());
    {
        size_t ia = 0, ib = 0; int64_t r = 1;
        while (ia < A.gidx.size() || ib < B.gidx.size()) {
            const bool takeA = ib >= B.gidx.size() || (ia < A.gidx.size() && A.gidx[ia] < B.gidx[ib]);
            if (takeA) blkA[ia++] = r++; else blkB[ib++] = r++;
        }
    }
    size_t ia = 0, ib = 0;
    while (ia < A.gidx.size() || ib < B.gidx.size()) {
        const bool takeA = ib >= B.gidx.size() || (ia < A.gidx.size() && A.gidx[ia] < B.gidx[ib]);
        if (takeA) {
            const int c = (int)ia;
            for (int k = ctx->h_cam_start[(size_t)c]; k < ctx->h_cam_start[(size_t)c + 1]; ++k) {   // points ascending
                const int j = ctx->h_cm_obs[(size_t)k];
                const int p = ctx->h_obs_pt[(size_t)j];
                if (B.gidx[(size_t)p] < A.gidx[ia])  // (camera row, point column): transpose of W
                    emit(blkA[ia], blkB[(size_t)p], DC, 3, hB + (int64_t)WB * j + 9 * (int64_t)p, true);
            }
            emit(blkA[ia], blkA[ia], DC, DC, (int64_t)DC * DC * c, false);
            ++ia;
        } else {
            const int p = (int)ib;
            for (int j = ctx->h_obs_start[(size_t)p]; j < ctx->h_obs_start[(size_t)p + 1]; ++j) {     // cameras ascending
                const int c = ctx->h_obs_cam[(size_t)j];
                if (A.gidx[(size_t)c] < B.gidx[ib]) emit(blkB[ib], blkA[(size_t)c], 3, DC, hB + (int64_t)WB * j + 9 * (int64_t)p, false);
            }
            emit(blkB[ib], blkB[ib], 3, 3, hB + (int64_t)WB * ctx->h_obs_start[(size_t)p + 1] + 9 * (int64_t)p, false);
            ++ib;
        }
    }
}
}  // namespace

extern "C" {

int nlls_get_gradient(nlls_ctx* ctx, double* b) {
    if (!ctx || !ctx->prepared || !b) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    return reorder_vec(ctx, ctx->d_g, b);
}
int nlls_get_step(nlls_ctx* ctx, double* x) {
    if (!ctx || !ctx->prepared || !x) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    return reorder_vec(ctx, ctx->d_x, x);
}

int nlls_get_hessian_blocks(nlls_ctx* ctx, double* data) {
    if (!ctx || !ctx->prepared || !data) return NLLS_ERR_INVALID;
    if (ctx->masked) FAIL(NLLS_ERR_UNSUPPORTED, "Hessian read-back under an unfixed mask (fixed variables are frozen in place, not removed)");
    if (ctx->has_dups) FAIL(NLLS_ERR_UNSUPPORTED, "Hessian read-back with several costs on one (camera, point) pair (the device keeps one block per cost)");
    CK(cudaSetDevice(ctx->device));
    if (ctx->cams_first) {  // internal layout == reference layout
        CK(cudaMemcpyAsync(data, ctx->d_H, sizeof(double) * ctx->hlen, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        return NLLS_OK;
    }
    std::vector<double> tmp((size_t)ctx->hlen);
    CK(cudaMemcpyAsync(tmp.data(), ctx->d_H, sizeof(double) * ctx->hlen, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    int64_t o = 0;
    walk_reference_blocks(ctx, [&](int64_t, int64_t, int rows, int cols, int64_t src, bool transposed) {
        if (!transposed) { std::memcpy(data + o, &tmp[(size_t)src], sizeof(double) * rows * cols); }
        else {  // stored block is cols x rows (column-major); emit its transpose
            for (int c = 0; c < cols; ++c) for (int r = 0; r < rows; ++r) data[o + r + (int64_t)rows * c] = tmp[(size_t)(src + c + (int64_t)cols * r)];
        }
        o += (int64_t)rows * cols;
    });
    return NLLS_OK;
}

int nlls_get_hessian_index(nlls_ctx* ctx, int64_t* rowblock, int64_t* colblock, int64_t* start) {
    if (!ctx || !ctx->prepared || !rowblock || !colblock || !start) return NLLS_ERR_INVALID;
    if (ctx->adaptive) FAIL(NLLS_ERR_UNSUPPORTED, "adaptive problems use the dense layout: nlls_get_hessian_blocks returns the dof x dof matrix");
    int64_t o = 1, k = 0;
    walk_reference_blocks(ctx, [&](int64_t rb, int64_t cb, int rows, int cols, int64_t, bool) {
        rowblock[k] = rb; colblock[k] = cb; start[k] = o; ++k;
        o += (int64_t)rows * cols;
    });
    return NLLS_OK;
}

// ---- measurement hooks -----------------------------------------------------------------------------------
int64_t nlls_kernel_launches(nlls_ctx* ctx) { return ctx ? ctx->launches : -1; }

int nlls_timer_start(nlls_ctx* ctx) {
    if (!ctx) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaEventRecord(ctx->ev_b0, ctx->st));
    return NLLS_OK;
}
int nlls_timer_stop(nlls_ctx* ctx, double* ms) {
    if (!ctx || !ms) return NLLS_ERR_INVALID;
    CK(cudaEventRecord(ctx->ev_b1, ctx->st));
    CK(cudaEventSynchronize(ctx->ev_b1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, ctx->ev_b0, ctx->ev_b1));
    *ms = f;
    return NLLS_OK;
}

int nlls_algorithmic_bytes(nlls_ctx* ctx, int which, double* bytes) {
    if (!ctx || !ctx->prepared || !bytes) return NLLS_ERR_INVALID;
    if (ctx->adaptive) { *bytes = 8.0 * (double)ctx->nobs; return (which == NLLS_TIME_LINEARIZE || which == NLLS_TIME_COST) ? NLLS_OK : NLLS_ERR_INVALID; }
    const double nobs = (double)ctx->nobs, nA = (double)ctx->nA, nB = (double)ctx->nB, DC = ctx->DC;
    const double rd = nobs * (8 * 2 + 8) + 8 * (nA * ctx->NC + nB * 3);   // measurement + 2 x int32 index, each variable once
    const double wr = 8 * (nobs * DC * 3 + nA * DC * DC + nB * 9 + nA * DC + nB * 3);
    switch (which) {
        case NLLS_TIME_LINEARIZE: *bytes = rd + wr; break;                                  // B_lin
        case NLLS_TIME_LIN_POINT: *bytes = rd + 8 * (nobs * DC * 3 + nB * 9 + nB * 3); break;
        case NLLS_TIME_LIN_CAM: *bytes = rd + 8 * (nA * DC * DC + nA * DC); break;
        case NLLS_TIME_COST: *bytes = rd; break;
        // Schur elimination: H's point rows (W, V) and g_p read once, A_p^-1 written; the reduced system itself is small
        case NLLS_TIME_SCHUR: *bytes = 8 * (nobs * DC * 3 + nB * 9 + nB * 3 + nB * 6); break;
        // back-substitution: H's point rows, A_p^-1, g_p, camera step read once; points read and written, point step written
        case NLLS_TIME_BACKSUB: *bytes = 8 * (nobs * DC * 3 + nB * 9 + nB * 6 + nB * 3 + nA * DC + 3 * nB * 3); break;
        default: *bytes = 0; return NLLS_ERR_INVALID;
    }
    return NLLS_OK;
}

// Algorithmic FP64 operations per launch (multiply and add counted separately) of the two phases north_star wants against the FP64
// peak: the Schur elimination (per point with k observations: 3 x 3 inverse, Y = A_p^-1 W (k blocks of 3 x 3 x DC), k (k + 1) / 2
// products W_i' Y_j (DC x 3 x DC) and k right-hand side terms) and the reduced solve.
int nlls_algorithmic_flops(nlls_ctx* ctx, int which, double* flops) {
    if (!ctx || !ctx->prepared || !flops || ctx->adaptive) return NLLS_ERR_INVALID;
    const double DC = ctx->DC;
    if (which == NLLS_TIME_SCHUR) {
        double f = 0;
        for (int64_t p = 0; p < ctx->nB; ++p) {
            const double k = ctx->h_obs_start[(size_t)p + 1] - ctx->h_obs_start[(size_t)p];
            f += 60.0 + k * (2 * 9 * DC) + 0.5 * k * (k + 1) * (2 * 3 * DC * DC) + k * (2 * 3 * DC) + 18.0;
        }
        *flops = f;
        return NLLS_OK;
    }
    if (which == NLLS_TIME_SOLVE_REDUCED) { *flops = ctx->red_flops; return NLLS_OK; }
    *flops = 0;
    return NLLS_ERR_INVALID;
}

int nlls_time_kernels(nlls_ctx* ctx, int which, int reps, int flush_l2, double* ms_per_call) {
    if (!ctx || !ctx->prepared || reps < 1 || !ms_per_call) return NLLS_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (flush_l2 && !ctx->d_flush) {
        ctx->flush_bytes = (size_t)256 << 20;  // larger than the 126 MB L2
        CK(cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    }
    double lambda = ctx->lambda;
    if (lambda == 0) lambda = 1e-3;
    double total = 0.0;
    const int keep_cam = ctx->cam_part_vars;
    if (which != NLLS_TIME_LIN_LOOP) ctx->cam_part_vars = -1;
    for (int r = 0; r < reps; ++r) {
        if (flush_l2) CK(cudaMemsetAsync(ctx->d_flush, r & 0xff, ctx->flush_bytes, ctx->st));
        CK(cudaEventRecord(ctx->ev_t0, ctx->st));
        if (ctx->adaptive) {
            if (which == NLLS_TIME_LINEARIZE) TRY(adapt_linearize(ctx));
            else if (which == NLLS_TIME_COST) TRY(adapt_cost(ctx, ctx->cur, SC_COST_TRY));
            else if (which == NLLS_TIME_TRY) TRY(enqueue_try(ctx, lambda));
            else return NLLS_ERR_INVALID;
        } else
        switch (which) {
            case NLLS_TIME_LINEARIZE: TRY(DISPATCH(ctx, launch_linearize, ctx, true, 1)); break;
            case NLLS_TIME_LIN_POINT: TRY(DISPATCH(ctx, launch_linearize, ctx, true, 0)); break;
            case NLLS_TIME_LIN_CAM: TRY(DISPATCH(ctx, launch_linearize, ctx, false, 1)); break;
            case NLLS_TIME_LIN_LOOP: TRY(DISPATCH(ctx, launch_linearize, ctx, true, 2)); break;
            case NLLS_TIME_COST: TRY(DISPATCH(ctx, launch_cost, ctx, ctx->cur, SC_COST_TRY, ctx->d_cam_part2, false)); break;   // kernels only: the LM try combines its scalars in exchange_try_scalars
            case NLLS_TIME_SCHUR: TRY(DISPATCH(ctx, launch_schur, ctx, lambda)); break;
            case NLLS_TIME_SOLVE_REDUCED: TRY(DISPATCH(ctx, launch_schur, ctx, lambda)); CK(cudaEventRecord(ctx->ev_t0, ctx->st)); TRY(launch_reduced_solve(ctx)); break;
            case NLLS_TIME_BACKSUB: TRY(DISPATCH(ctx, launch_update, ctx)); break;
            case NLLS_TIME_TRY: TRY(enqueue_try(ctx, lambda)); break;
            case NLLS_TIME_MEMSET_H: CK(cudaMemsetAsync(ctx->d_H + (size_t)ctx->DC * ctx->DC * ctx->nA, 0, sizeof(double) * (size_t)(ctx->hlen - (int64_t)ctx->DC * ctx->DC * ctx->nA), ctx->st)); break;
            default: return NLLS_ERR_INVALID;
        }
        CK(cudaEventRecord(ctx->ev_t1, ctx->st));
        CK(cudaEventSynchronize(ctx->ev_t1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
        total += ms;
    }
    *ms_per_call = total / reps;
    ctx->cam_part_vars = (which == NLLS_TIME_LIN_LOOP || which == NLLS_TIME_COST || which == NLLS_TIME_LIN_POINT) ? keep_cam : -1;
    return NLLS_OK;
}

}  // extern "C"
