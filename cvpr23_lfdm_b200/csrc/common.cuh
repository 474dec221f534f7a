// common.cuh — shared device helpers for the LFDM sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdlib>
#include <cstring>
#include <utility>
#include "../../include/lfdm_b200.h"

typedef __nv_bfloat16 bf16;

#define LFDM_CHECK_LAUNCH()                                   \
    do {                                                      \
        cudaError_t e__ = cudaPeekAtLastError();              \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// One sampling step is ~180 back-to-back launches of short kernels on one stream.  Kernels launched through
// lfdm_launch_pdl() carry the programmatic-stream-serialization attribute (also inside a captured CUDA graph): their
// CTAs may become resident and run their prologue (barrier init, TMEM allocation, index math) while the previous kernel
// drains.  Contract: such a kernel calls pdl_prologue_done() (or pdl_trigger() ... pdl_wait()) before its first global
// memory access - griddepcontrol.wait returns only when the preceding grid has completed and its writes are visible.
// LFDM_NO_PDL=1 turns the attribute off (the device-side instructions are then no-ops).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue_done() { pdl_trigger(); pdl_wait(); }

inline bool lfdm_pdl_enabled() {
    static const bool on = (getenv("LFDM_NO_PDL") == nullptr);
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t lfdm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    memset(at, 0, sizeof(at));
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = lfdm_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
// same, launched as thread-block clusters of `cluster_x` CTAs (cluster_x <= 1: plain launch)
template <typename... KArgs, typename... Args>
inline cudaError_t lfdm_launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                           Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    memset(at, 0, sizeof(at));
    int n = 0;
    if (lfdm_pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
#define LFDM_LAUNCH_PDL(...)                                   \
    do {                                                      \
        cudaError_t e__ = lfdm_launch_pdl(__VA_ARGS__);       \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

// The > 48 KiB dynamic-shared-memory opt-in (cudaFuncSetAttribute) is PER DEVICE: one flag per device ordinal, not per process.
struct PerDeviceOnce {
    bool done[64] = {};
    // -> true when the caller still has to run the one-time set-up for the current device
    bool need() const { int d = 0; cudaGetDevice(&d); return d < 0 || d >= 64 || !done[d]; }
    void mark() { int d = 0; cudaGetDevice(&d); if (d >= 0 && d < 64) done[d] = true; }
};

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(bf16 hi, bf16 lo) { return __bfloat162float(hi) + __bfloat162float(lo); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == LFDM_ACT_RELU) return fmaxf(v, 0.f);
    if (act == LFDM_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}
__device__ __forceinline__ float silu_f(float v) { return v / (1.f + expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// store 4 consecutive channels of one row as split-bf16 (8-byte stores into each plane)
__device__ __forceinline__ void store_sb4(bf16* hi_plane, int64_t plane, int64_t idx, float4 v) {
    bf16 h0, l0, h1, l1, h2, l2, h3, l3;
    split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
    __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    __nv_bfloat162 c = __halves2bfloat162(l0, l1), d = __halves2bfloat162(l2, l3);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<uint32_t*>(&a); hv.y = *reinterpret_cast<uint32_t*>(&b);
    lv.x = *reinterpret_cast<uint32_t*>(&c); lv.y = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint2*>(hi_plane + idx) = hv;
    *reinterpret_cast<uint2*>(hi_plane + plane + idx) = lv;
}
__device__ __forceinline__ void store_sb1(bf16* hi_plane, int64_t plane, int64_t idx, float v) {
    bf16 h, l; split_bf16(v, h, l);
    hi_plane[idx] = h; hi_plane[plane + idx] = l;
}
