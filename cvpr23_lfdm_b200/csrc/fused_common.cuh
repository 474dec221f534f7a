// fused_common.cuh — small device helpers shared by the fused tcgen05 attention kernels (explicit 32-bit shared-memory
// addressing, split-bf16 packing, TMEM loads / stores, non-blocking mbarrier tests, bulk copies).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace fz {

__device__ __forceinline__ void bulk_copy_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(ptx::smem_u32(bar))
                 : "memory");
}
// non-blocking phase test (an issuer polls several independent streams)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) { ptx::tmem_ld16(taddr, r); }
// 32 lanes x 8 consecutive 32-bit columns <- 8 registers per thread
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (x, y) -> packed bf16 pairs hi = bf16(.), lo = bf16(. - hi)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of 16-byte chunk `c` of row `r` in a SW128 K-major tile (rows of 128 B, 8-row groups of 1024 B)
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// one arrival per warp: all lanes have finished their part (and fenced it) before lane 0 signals
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar);
}
// 64 fp32 values of one operand row (K = 64) -> split-bf16 row of a SW128 K-major tile: hi plane at `tile`, lo plane at
// `tile + plane_bytes`
__device__ __forceinline__ void store_row64_hilo(uint32_t tile, uint32_t plane_bytes, int r, const float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split2(v[8 * c], v[8 * c + 1], h0, l0);
        split2(v[8 * c + 2], v[8 * c + 3], h1, l1);
        split2(v[8 * c + 4], v[8 * c + 5], h2, l2);
        split2(v[8 * c + 6], v[8 * c + 7], h3, l3);
        const uint32_t off = tile + sw_off(r, c);
        sts128(off, h0, h1, h2, h3);
        sts128(off + plane_bytes, l0, l1, l2, l3);
    }
}

}  // namespace fz
