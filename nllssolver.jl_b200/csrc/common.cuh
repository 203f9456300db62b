// Shared device helpers for the sm_100a kernels of libnlls_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nlls {

constexpr int TILE_OBS = 256;   // observations per point-major tile (one thread each)
constexpr int TILE_PTS = 128;   // points per tile (upper bound)
constexpr int LIN_THREADS = 256;
constexpr int CAM_CHUNK = 4096; // observations per camera-pass work item
constexpr int DP = 3;           // point DoF

// ---------------------------------------------------------------------------------------------------
// Robust kernels (src/robust.jl:7-77).  kind is warp-uniform, so the switch does not diverge.
// ---------------------------------------------------------------------------------------------------
struct RobustParams {
    int kind;      // nlls_robust without the SCALED bit
    int scaled;
    double width, width2, height;
};

// robustify(kernel, s)                                             src/robust.jl:11,26,47,72
__device__ __forceinline__ double robustify(const RobustParams& k, double s) {
    double r;
    switch (k.kind) {
        default: r = s; break;
        case 1:
        case 2: r = s < k.width2 ? s : sqrt(s) * (k.width * 2) - k.width2; break;
        case 3: r = s * k.width2 / (s + k.width2); break;
    }
    return k.scaled ? r * k.height : r;
}
// robustifydcost(kernel, s) -> (rho, rho', rho'')                   src/robust.jl:12,28-31,48-55,73-77
__device__ __forceinline__ void robustifydcost(const RobustParams& k, double s, double& rho, double& d1, double& d2) {
    switch (k.kind) {
        default: rho = s; d1 = 1.0; d2 = 0.0; break;
        case 1:
        case 2:
            if (s < k.width2) { rho = s; d1 = 1.0; d2 = 0.0; }
            else {
                double sq = sqrt(s);
                rho = sq * (k.width * 2) - k.width2;
                d1 = k.width / sq;
                d2 = (k.kind == 2) ? (-0.5 * k.width) / (s * sq) : 0.0;
            }
            break;
        case 3: {
            double r = 1.0 / (s + k.width2);
            double w = k.width2 * r;
            double ww = w * w;
            rho = s * w; d1 = ww; d2 = -2 * ww * r;
            break;
        }
    }
    if (k.scaled) { rho *= k.height; d1 *= k.height; d2 *= k.height; }
}

// ---------------------------------------------------------------------------------------------------
// Deterministic block reductions (fixed tree: xor-shuffle inside a warp, then sequential over warps).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// NaN-propagating max like Julia's maximum(abs, x)
__device__ __forceinline__ double nanmax(double a, double b) { return (isnan(a) || isnan(b)) ? nan("") : fmax(a, b); }
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Sum over the block; result valid in thread 0.  s_red must hold blockDim.x/32 doubles.  Ends with no barrier:
// callers that reuse s_red must __syncthreads() first.
__device__ __forceinline__ double block_sum(double v, double* s_red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) t += s_red[i];
    }
    return t;
}
__device__ __forceinline__ double block_nanmax(double v, double* s_red) {
    v = warp_nanmax(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) t = nanmax(t, s_red[i]);
    }
    return t;
}

// ---------------------------------------------------------------------------------------------------
// TMA bulk copies (1-D cp.async.bulk; SASS: UBLKCP).  Addresses and sizes must be multiples of 16 bytes.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// a global load that stays where it is written (the compiler sinks plain loads into the branch that consumes them)
__device__ __forceinline__ double ldg_f64_here(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared -> global; the calling thread must wait (bulk_store_wait) before the smem is reused / the CTA exits
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, completion signalled on the mbarrier
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 3x3 symmetric inverse by the adjugate. in: a00,a10,a20,a11,a21,a22 (lower, column order). out same order.
__device__ __forceinline__ void inv_sym3(const double a[6], double inv[6]) {
    const double a00 = a[0], a10 = a[1], a20 = a[2], a11 = a[3], a21 = a[4], a22 = a[5];
    const double c00 = a11 * a22 - a21 * a21;
    const double c10 = a20 * a21 - a10 * a22;
    const double c20 = a10 * a21 - a20 * a11;
    const double det = a00 * c00 + a10 * c10 + a20 * c20;
    const double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = c10 * id;
    inv[2] = c20 * id;
    inv[3] = (a00 * a22 - a20 * a20) * id;
    inv[4] = (a10 * a20 - a00 * a21) * id;
    inv[5] = (a00 * a11 - a10 * a10) * id;
}

}  // namespace nlls
