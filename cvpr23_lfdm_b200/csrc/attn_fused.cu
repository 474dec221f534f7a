// attn_fused.cu — the whole temporal-attention block of the LFDM UNet as ONE persistent tcgen05 kernel (sm_100a):
//
//     out = x + to_out( softmax( rot(q*scale) rot(k)^T + pos_bias ) v ),   q|k|v = to_qkv( LayerNorm(x) )
//
// Replaces Residual(PreNorm(EinopsToAndFrom(Attention))) of the reference (DM/modules/video_flow_diffusion.py:132-138,
// 170-190, 270-283, 286-363) for C = 64 channels and 40 frames (the 32x32 / 16x16 levels: init_temporal_attn, downs.0,
// ups.2, ups.3).  q|k|v, the scores and the per-head outputs never leave the SM: the four GEMM-shaped stages run on
// tcgen05 with fp32 accumulators in TMEM, everything between them is done by three warp-groups straight out of TMEM.
//
// One tile = 3 pixel columns x 40 frames = 120 rows of the row matrix (padded to the 128-row UMMA tile); per head:
//               qkv  D[128x96]   = Xn . W_h^T          3 split-bf16 products x 4 K-steps, double-buffered in TMEM
//               QK   S[128x128]  = Q . K^T             all 3 pixels at once; a row only uses its own 40-column block
//               PV   D[128x96]   = P . V^T^T           compact K = 48 positions; column block 32*pixel(row) is the result
//               out  OUT[128x64] += O_h . Wout_h^T     accumulated over heads in TMEM
// The kernel is a DATAFLOW pipeline over the flat (tile, head) sequence g.  Measured history (tools/attn_trace.py stamps every
// barrier crossing of CTA 0 with clock64()): one blocking in-order issue thread + single-buffered V^T: 5.8 k clk per head and
// 29 k clk per tile boundary (0.77 ms per 32x32 block at batch 8); MMAs issued by lanes of the compute groups: the spinning
// sibling lanes steal the issue slots (1 k clk per 6 MMAs); thread-per-row epilogue: 32 lines per store instruction, every
// st.shared of the SM stalls behind it for 17 k clk per tile.  Current roles (18 warps):
//   warp 0      MMA issuer (one elected thread) POLLING four independent streams -- qkv(g) into T_QKV[g & 1], QK(g), PV(g),
//               out(g) -- with non-blocking mbarrier tests, downstream first: no stage waits behind an unrelated one
//   warp 1      weight producer: per head one [W_hi|W_lo] slice of to_qkv (24 KiB) and of to_out (8 KiB, two heads later),
//               pre-swizzled on the host, cp.async.bulk from L2 into 2-stage rings
//   warps 2-5   LayerNorm producers + tile epilogue: the NEXT tile's x rows are loaded and normalised in registers while the
//               current tile is being processed; once the last qkv GEMM of the current tile has drained the operand buffer
//               they split and store the A operand (128x64, SW128 K-major, hi / lo planes); then OUT (+bias) + x of the
//               previous tile -> F32 and split-bf16 rows straight from TMEM, coalesced by a 4x4 chunk transpose inside
//               every lane quad (shuffles)
//   warps 6-9   WG-A: rotary(q) (scale folded into W_q), rotary(k) TMEM -> [hi|lo] operand rows
//   warps 10-13 WG-B: PV(g) result -> [hi|lo] operand rows of the output projection (downstream first), then v(g+2) -> the
//               MN-major B operand V[(g+2) & 1][position][(pixel, d)]: 64 contiguous bytes per thread instead of 64 two-byte
//               scatter stores of the transposed K-major form (kept behind LFDM_ATTN_VT_KMAJOR=1 as a cross-check)
//   warps 14-17 WG-C: +bias, softmax over the row's 40 columns, P -> compact A operand [128 x 48] (hi / lo planes)
// Synchronisation: mbarriers only (tcgen05.commit for MMA completion, one elected arrival per warp).  Shared memory is
// addressed with explicit st.shared / ld.shared on 32-bit addresses.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int FL = 40;       // frames (sequence length)
constexpr int FG = 3;        // pixel columns per tile
constexpr int FC = 64;       // channels
constexpr int NTHREADS = 576;       // 18 warps (register cap 112)

constexpr int OFF_XN = 0;            // 2 planes x 16 KiB : LayerNorm output, A operand of the qkv GEMM
constexpr int OFF_Q = 32768;         // 16 KiB : rows [q_hi(32) | q_lo(32)]
constexpr int OFF_K = 49152;         // 16 KiB : rows [k_hi | k_lo]
constexpr int OFF_VT = 65536;        // 2 buffers x 2 planes x 12 KiB : V^T, 96 rows (pixel, d) x 64 positions (48 used)
constexpr int VT_PLANE = 12288;
constexpr int VT_BUF = 24576;
constexpr int OFF_P = 114688;        // 2 planes x 16 KiB : P rows x 64 positions (48 used, 40..47 zero); fp32 staging tile of the epilogue
constexpr int OFF_O = 147456;        // 16 KiB : rows [o_hi(32) | o_lo(32)]
constexpr int OFF_WQ = 163840;       // 2 stages x 24 KiB : [W_hi (96 x 64) | W_lo (96 x 64)] of one head
constexpr int WQ_STAGE = 24576;
constexpr int OFF_WO = 212992;       // 2 stages x 8 KiB : 64 rows [w_hi(32) | w_lo(32)] of one head
constexpr int WO_STAGE = 8192;
constexpr int OFF_BAR = 229376;      // mbarriers, TMEM pointer
constexpr int OFF_INV = OFF_BAR + 1024;   // 2 x 128 floats: 1 / (softmax denominator) of every tile row, per head parity
constexpr int SMEM_BYTES = OFF_INV + 1024 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KiB per CTA)");

// TMEM column map (512 columns)
constexpr uint32_t T_QKV = 0;        // 2 x 96 (double-buffered q | k | v of one head)
constexpr uint32_t T_S = 192;        // 128
constexpr uint32_t T_PVD = 320;      // 128 (96 used with the K-major V^T operand, 128 = two 64-wide atoms with the MN-major V operand)
constexpr uint32_t T_OUT = 448;      // 64
constexpr int VMN_ATOM = 6144;       // MN-major V operand: one 64-column atom = 48 positions x 128 B

enum {
    B_XN_FULL = 0, B_XN_EMPTY = 1, B_WQ_FULL = 2 /* 2 */, B_WQ_EMPTY = 4 /* 2 */, B_WO_FULL = 6 /* 2 */, B_WO_EMPTY = 8 /* 2 */,
    B_QKV_FULL = 10 /* 2 */, B_QKV_EMPTY = 12 /* 2 */, B_S_FULL = 14, B_S_EMPTY = 15, B_VT_READY = 16 /* 2 */, B_PVD_FULL = 18,
    B_PVD_EMPTY = 19, B_O_FREE = 20, B_OUT_FULL = 21, B_OUT_EMPTY = 22, B_QK_READY = 23, B_P_READY = 24, B_O_READY = 25
};

struct FusedArgs {
    const float* x;
    const float* gamma;
    const uint8_t* wq;           // [heads][WQ_STAGE] pre-swizzled smem images
    const uint8_t* wo;           // [heads][WO_STAGE]
    const float* out_bias;       // [64] or null
    const float* rot_cos;        // [40][16]
    const float* rot_sin;
    const float* pos_bias;       // [heads][40][40] or null
    float* out_f32;
    bf16* out_sb;
    int64_t out_plane;
    float* dbg;                  // diagnostics (tests): [M][3*hid + heads*40 + hid] or null
    long long* trace;            // LFDM_ATTN_TRACE=1: [32 event slots][64 heads] clock64() stamps of CTA 0 (pipeline timeline)
    int32_t heads, n_pc, pix, n_tiles;
    float eps;
};

// timeline stamp of event `slot` for flat head index g (CTA 0 only; `who` = the one thread of the role that stamps)
#define TR(slot, g, who)                                                                             \
    do {                                                                                             \
        if (a.trace && blockIdx.x == 0 && (who) && (g) < 64u) a.trace[(slot) * 64 + (g)] = clock64(); \
    } while (0)

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ptx::smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(ptx::smem_u32(bar))
                 : "memory");
}
// non-blocking phase test (the issuer polls several independent streams)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint64_t* bar_at(uint8_t* smem, int i) { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + i; }
// byte offset of 16-byte chunk `c` of row `r` in a SW128 K-major tile (rows of 128 B, 8-row groups of 1024 B)
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t a) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(a) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// 16 fp32 values (dims 16*half .. 16*half+15) of one row -> [hi(32) | lo(32)] operand row (chunks 0-3 hi, 4-7 lo)
__device__ __forceinline__ void store_hilo_half(uint32_t tile, int r, int half, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split2(v[8 * c], v[8 * c + 1], h0, l0);
        split2(v[8 * c + 2], v[8 * c + 3], h1, l1);
        split2(v[8 * c + 4], v[8 * c + 5], h2, l2);
        split2(v[8 * c + 6], v[8 * c + 7], h3, l3);
        sts128(tile + sw_off(r, 2 * half + c), h0, h1, h2, h3);
        sts128(tile + sw_off(r, 2 * half + c + 4), l0, l1, l2, l3);
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// one arrival per warp: all lanes have finished their part (and fenced it) before lane 0 signals
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar);
}

// named barrier of one compute warp-group (ids 1..3), 128 threads
__device__ __forceinline__ void wg_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// VMN: v is stored as an MN-major B operand [position][(pixel, d)] (64 contiguous bytes per thread) instead of the
// K-major transposed V^T[(pixel, d)][position] (64 two-byte scatter stores per thread)
template <bool DBG, bool VMN>
__global__ void __launch_bounds__(NTHREADS, 1) attn_temporal_fused_kernel(const __grid_constant__ FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 512);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int heads = a.heads;
    const int hid = heads * 32;

    pdl_trigger();
    if (warp == 1 && ptx::elect_one()) {
        ptx::mbar_init(bar_at(smem, B_XN_FULL), 4);
        ptx::mbar_init(bar_at(smem, B_XN_EMPTY), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_at(smem, B_WQ_FULL + i), 1);
            ptx::mbar_init(bar_at(smem, B_WQ_EMPTY + i), 1);
            ptx::mbar_init(bar_at(smem, B_WO_FULL + i), 1);
            ptx::mbar_init(bar_at(smem, B_WO_EMPTY + i), 1);
            ptx::mbar_init(bar_at(smem, B_QKV_FULL + i), 1);
            ptx::mbar_init(bar_at(smem, B_QKV_EMPTY + i), 8);
            ptx::mbar_init(bar_at(smem, B_VT_READY + i), 4);
        }
        ptx::mbar_init(bar_at(smem, B_S_FULL), 1);
        ptx::mbar_init(bar_at(smem, B_S_EMPTY), 4);
        ptx::mbar_init(bar_at(smem, B_PVD_FULL), 1);
        ptx::mbar_init(bar_at(smem, B_PVD_EMPTY), 4);
        ptx::mbar_init(bar_at(smem, B_O_FREE), 1);
        ptx::mbar_init(bar_at(smem, B_OUT_FULL), 1);
        ptx::mbar_init(bar_at(smem, B_OUT_EMPTY), 4);
        ptx::mbar_init(bar_at(smem, B_QK_READY), 4);
        ptx::mbar_init(bar_at(smem, B_P_READY), 4);
        ptx::mbar_init(bar_at(smem, B_O_READY), 4);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr, 512);
        ptx::tmem_relinquish();
    }
    // V^T pad positions (40..47 of every row) and P positions 40..63 are never written afterwards: zero both operands once
    for (int i = threadIdx.x; i < (OFF_O - OFF_VT) / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(smem + OFF_VT)[i] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();

    const int n_tiles = a.n_tiles;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t NG = (uint32_t)(my_tiles * heads);        // flat (tile, head) sequence of this CTA
    const uint32_t sb = ptx::smem_u32(smem);

    if (warp == 1) {
        // ===================== weight producer =====================
        // to_out slices trail the to_qkv slices by two heads: out(g) runs ~3 head periods after qkv(g), and waiting here for
        // out(g-2) before fetching W_qkv(g+1) would tie the head of the pipeline to its tail
        if (ptx::elect_one()) {
            int h = 0, ho = 0;
            for (uint32_t g = 0; g < NG + 2; ++g) {
                if (g < NG) {
                    const int s = g & 1;
                    ptx::mbar_wait(bar_at(smem, B_WQ_EMPTY + s), ((g >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(bar_at(smem, B_WQ_FULL + s), WQ_STAGE);
                    bulk_copy_g2s(smem + OFF_WQ + s * WQ_STAGE, a.wq + (size_t)h * WQ_STAGE, WQ_STAGE, bar_at(smem, B_WQ_FULL + s));
                    if (++h == heads) h = 0;
                }
                if (g >= 2) {
                    const uint32_t go = g - 2;
                    const int s = go & 1;
                    ptx::mbar_wait(bar_at(smem, B_WO_EMPTY + s), ((go >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(bar_at(smem, B_WO_FULL + s), WO_STAGE);
                    bulk_copy_g2s(smem + OFF_WO + s * WO_STAGE, a.wo + (size_t)ho * WO_STAGE, WO_STAGE, bar_at(smem, B_WO_FULL + s));
                    if (++ho == heads) ho = 0;
                }
            }
        }
    } else if (warp == 0) {
        // ===================== MMA issuer (one elected thread) =====================
        // Four independent streams (qkv, QK, PV, out), each with its own cursor; the thread POLLS their input barriers and
        // issues whichever is ready, downstream first: a stage never waits behind an unrelated one (a blocking in-order issue
        // loop measured 5.8 k clk per head; issuing from lanes of the compute groups steals their issue slots).
        if (ptx::elect_one()) {
            const uint64_t d_xn = ptx::make_sw128_kmajor_desc(sb + OFF_XN);
            const uint64_t d_wq = ptx::make_sw128_kmajor_desc(sb + OFF_WQ);
            const uint64_t d_q = ptx::make_sw128_kmajor_desc(sb + OFF_Q);
            const uint64_t d_k = ptx::make_sw128_kmajor_desc(sb + OFF_K);
            const uint64_t d_p = ptx::make_sw128_kmajor_desc(sb + OFF_P);
            const uint64_t d_vt = ptx::make_sw128_kmajor_desc(sb + OFF_VT);
            const uint64_t d_o = ptx::make_sw128_kmajor_desc(sb + OFF_O);
            const uint64_t d_wo = ptx::make_sw128_kmajor_desc(sb + OFF_WO);
            // MN-major, 128-byte-swizzled B operand: 64 columns (128 B) contiguous, 8 positions per 1024-byte group (SBO), the
            // second 64-column atom VMN_ATOM bytes further (LBO)
            const uint64_t d_vmn = (uint64_t)(((sb + OFF_VT) & 0x3FFFFu) >> 4) | ((uint64_t)(VMN_ATOM >> 4) << 16) |
                                   ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            constexpr uint32_t ID96 = ptx::make_idesc_bf16(128, 96);
            constexpr uint32_t ID128 = ptx::make_idesc_bf16(128, 128);
            constexpr uint32_t ID128T = ptx::make_idesc_bf16(128, 128) | (1u << 16);       // B operand MN-major
            constexpr uint32_t ID64 = ptx::make_idesc_bf16(128, 64);
            uint32_t gq = 0, gk = 0, gv = 0, go = 0;     // next head of each stream
            uint32_t it_q = 0, it_o = 0;
            int h_q = 0, h_o = 0;
            while (go < NG) {
                bool issued = false;
                // ---- OUT += O_h Wout_h^T
                {
                    const uint32_t so = go & 1u;
                    if (mbar_test(bar_at(smem, B_O_READY), go & 1) && mbar_test(bar_at(smem, B_WO_FULL + so), (go >> 1) & 1) &&
                        (h_o != 0 || it_o == 0 || mbar_test(bar_at(smem, B_OUT_EMPTY), (it_o - 1) & 1))) {
                        ptx::tc_fence_after();
                        TR(7, go, true);
                        const uint32_t td = tmem_base + T_OUT;
                        const uint64_t w = d_wo + (uint64_t)((so * WO_STAGE) >> 4);
                        ptx::umma_bf16(td, d_o + 4, w + 0, ID64, h_o > 0 ? 1u : 0u);   // o_lo . w_hi
                        ptx::umma_bf16(td, d_o + 6, w + 2, ID64, 1u);
                        ptx::umma_bf16(td, d_o + 0, w + 4, ID64, 1u);               // o_hi . w_lo
                        ptx::umma_bf16(td, d_o + 2, w + 6, ID64, 1u);
                        ptx::umma_bf16(td, d_o + 0, w + 0, ID64, 1u);               // o_hi . w_hi
                        ptx::umma_bf16(td, d_o + 2, w + 2, ID64, 1u);
                        ptx::umma_commit(bar_at(smem, B_WO_EMPTY + so));
                        ptx::umma_commit(bar_at(smem, B_O_FREE));
                        if (++h_o == heads) { ptx::umma_commit(bar_at(smem, B_OUT_FULL)); h_o = 0; ++it_o; }
                        ++go;
                        issued = true;
                    }
                }
                // ---- D = P V : compact K = 48 positions; D column block 32*pixel
                if (gv < NG) {
                    const uint32_t s = gv & 1u;
                    if (mbar_test(bar_at(smem, B_P_READY), gv & 1) && mbar_test(bar_at(smem, B_VT_READY + s), (gv >> 1) & 1) &&
                        (gv == 0 || mbar_test(bar_at(smem, B_PVD_EMPTY), (gv - 1) & 1))) {
                        ptx::tc_fence_after();
                        TR(5, gv, true);
                        const uint32_t td = tmem_base + T_PVD;
                        const uint64_t p_hi = d_p, p_lo = d_p + (uint64_t)(16384 >> 4);
                        if (VMN) {
                            const uint64_t v_hi = d_vmn + (uint64_t)((s * VT_BUF) >> 4), v_lo = v_hi + (uint64_t)(VT_PLANE >> 4);
#pragma unroll
                            for (int ks = 0; ks < 3; ++ks) {
                                const uint64_t o = (uint64_t)(ks * 2), ov = (uint64_t)(ks * (2048 >> 4));   // 16 positions = 2 row groups
                                ptx::umma_bf16(td, p_lo + o, v_hi + ov, ID128T, ks > 0 ? 1u : 0u);
                                ptx::umma_bf16(td, p_hi + o, v_lo + ov, ID128T, 1u);
                                ptx::umma_bf16(td, p_hi + o, v_hi + ov, ID128T, 1u);
                            }
                        } else {
                            const uint64_t v_hi = d_vt + (uint64_t)((s * VT_BUF) >> 4), v_lo = v_hi + (uint64_t)(VT_PLANE >> 4);
#pragma unroll
                            for (int ks = 0; ks < 3; ++ks) {
                                const uint64_t o = (uint64_t)(ks * 2);
                                ptx::umma_bf16(td, p_lo + o, v_hi + o, ID96, ks > 0 ? 1u : 0u);
                                ptx::umma_bf16(td, p_hi + o, v_lo + o, ID96, 1u);
                                ptx::umma_bf16(td, p_hi + o, v_hi + o, ID96, 1u);
                            }
                        }
                        ptx::umma_commit(bar_at(smem, B_PVD_FULL));
                        ++gv;
                        issued = true;
                    }
                }
                // ---- S = Q K^T : rows [hi | lo]: K-steps 0,1 = hi dims, 2,3 = lo dims
                if (gk < NG) {
                    if (mbar_test(bar_at(smem, B_QK_READY), gk & 1) && (gk == 0 || mbar_test(bar_at(smem, B_S_EMPTY), (gk - 1) & 1))) {
                        ptx::tc_fence_after();
                        TR(3, gk, true);
                        const uint32_t td = tmem_base + T_S;
                        ptx::umma_bf16(td, d_q + 4, d_k + 0, ID128, 0u);      // q_lo . k_hi
                        ptx::umma_bf16(td, d_q + 6, d_k + 2, ID128, 1u);
                        ptx::umma_bf16(td, d_q + 0, d_k + 4, ID128, 1u);      // q_hi . k_lo
                        ptx::umma_bf16(td, d_q + 2, d_k + 6, ID128, 1u);
                        ptx::umma_bf16(td, d_q + 0, d_k + 0, ID128, 1u);      // q_hi . k_hi
                        ptx::umma_bf16(td, d_q + 2, d_k + 2, ID128, 1u);
                        ptx::umma_commit(bar_at(smem, B_S_FULL));
                        ++gk;
                        issued = true;
                    }
                }
                // ---- q | k | v of one head: Xn . [W_hi | W_lo]^T into T_QKV[gq & 1]
                if (gq < NG) {
                    const uint32_t s = gq & 1u, par = (gq >> 1) & 1u;
                    if ((h_q != 0 || mbar_test(bar_at(smem, B_XN_FULL), it_q & 1)) && mbar_test(bar_at(smem, B_WQ_FULL + s), par) &&
                        mbar_test(bar_at(smem, B_QKV_EMPTY + s), par ^ 1)) {
                        ptx::tc_fence_after();
                        TR(1, gq, true);
                        const uint32_t td = tmem_base + T_QKV + 96u * s;
                        const uint64_t w_hi = d_wq + (uint64_t)((s * WQ_STAGE) >> 4), w_lo = w_hi + (uint64_t)(12288 >> 4);
                        const uint64_t x_hi = d_xn, x_lo = d_xn + (uint64_t)(16384 >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t o = (uint64_t)(ks * 2);
                            ptx::umma_bf16(td, x_lo + o, w_hi + o, ID96, ks > 0 ? 1u : 0u);
                            ptx::umma_bf16(td, x_hi + o, w_lo + o, ID96, 1u);
                            ptx::umma_bf16(td, x_hi + o, w_hi + o, ID96, 1u);
                        }
                        ptx::umma_commit(bar_at(smem, B_WQ_EMPTY + s));
                        ptx::umma_commit(bar_at(smem, B_QKV_FULL + s));
                        if (++h_q == heads) {                     // last qkv GEMM of the tile: the operand buffer may be refilled
                            ptx::umma_commit(bar_at(smem, B_XN_EMPTY));
                            h_q = 0; ++it_q;
                        }
                        ++gq;
                        issued = true;
                    }
                }
                if (!issued) __nanosleep(40);          // idle pass: leave the issue slots of this scheduler to the compute warps
            }
        }
    } else if (warp < 6) {
        // ===================== LayerNorm producers (warps 2-5): next tile's x rows -> normalised in registers -> XN ============
        const int t128 = (int)threadIdx.x - 64;         // 0..127
        const int l16 = t128 & 15, rg = t128 >> 4;      // 16 lanes per row, 8 rows per pass, 16 passes
        const float4 gam = *reinterpret_cast<const float4*>(a.gamma + l16 * 4);
        const uint32_t sb_ln = sb + OFF_XN;
        // global row of tile row `row` of tile `tile`: one division per tile (first pixel column), the other two columns follow
        auto tile_base = [&](int tile, int& b0, int& rem0) { const int pc0 = tile * FG; b0 = pc0 / a.pix; rem0 = pc0 - b0 * a.pix; };
        auto row_addr = [&](int tile, int b0, int rem0, int row) -> int64_t {       // -1: padding row / beyond the last pixel column
            const int tpx = (row * 205) >> 13;                                       // row / 40 for row < 128
            const int tf = row - tpx * FL;
            if (row >= FG * FL || tile * FG + tpx >= a.n_pc) return -1;
            int rem = rem0 + tpx, b = b0;
            if (rem >= a.pix) { rem -= a.pix; ++b; }
            return ((int64_t)(b * FL + tf) * a.pix + rem) * FC;
        };
        // tile epilogue (also this group's job: it idles between two LayerNorms): OUT (+bias) + x -> F32 and split-bf16 rows.
        // thread = TMEM lane = tile row out of the accumulator; a 4x4 transpose of 16-byte chunks inside every lane quad (two
        // shuffle butterflies) turns that into "4 lanes = 64 contiguous bytes of one row" for the global loads / stores (the
        // thread-per-row form touches 32 lines per instruction and was measured to stall every shared-memory store of the SM)
        const int eq = warp & 3, er = eq * 32 + lane;                  // TMEM lane quarter / tile row of this thread
        const int ec = lane & 3;                                        // 16-byte chunk of each 64-byte group this lane ends up with
        auto epilogue = [&](int tile_e, uint32_t it_e) {
            ptx::mbar_wait(bar_at(smem, B_OUT_FULL), it_e & 1);
            ptx::tc_fence_after();
            uint32_t u[4][16];
#pragma unroll
            for (int part = 0; part < 4; ++part) tmem_ld16(tmem_base + ((uint32_t)(eq * 32) << 16) + T_OUT + 16u * (uint32_t)part, u[part]);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            warp_arrive(bar_at(smem, B_OUT_EMPTY), lane);             // out(head 0) of the next tile may overwrite the accumulator
            int64_t goff[4];                                            // element offsets of the 4 tile rows of this lane's quad
            {
                int b0, rem0;
                tile_base(tile_e, b0, rem0);
#pragma unroll
                for (int j = 0; j < 4; ++j) goff[j] = row_addr(tile_e, b0, rem0, (er & ~3) + j);
            }
            const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[part][i]);
                // stage 1 (lanes a, a^2): slot s <- (row (s&2)|(a&1), chunk (a&2)|(s&1));  stage 2 (lanes a, a^1): slot s <- (row s, chunk a)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float send = up2 ? v[4 * c + e] : v[4 * (c + 2) + e];
                        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
                        if (up2) v[4 * c + e] = recv; else v[4 * (c + 2) + e] = recv;
                    }
#pragma unroll
                for (int sp = 0; sp < 4; sp += 2)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float send = up1 ? v[4 * sp + e] : v[4 * (sp + 1) + e];
                        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                        if (up1) v[4 * sp + e] = recv; else v[4 * (sp + 1) + e] = recv;
                    }
                const int col = part * 16 + ec * 4;
                float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.out_bias) ob = __ldg(reinterpret_cast<const float4*>(a.out_bias + col));
                float4 xres[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    xres[j] = goff[j] >= 0 ? *reinterpret_cast<const float4*>(a.x + goff[j] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (goff[j] >= 0) {
                        const float4 t4 = make_float4(v[4 * j] + xres[j].x + ob.x, v[4 * j + 1] + xres[j].y + ob.y,
                                                      v[4 * j + 2] + xres[j].z + ob.z, v[4 * j + 3] + xres[j].w + ob.w);
                        if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + goff[j] + col) = t4;
                        if (a.out_sb) store_sb4(a.out_sb, a.out_plane, goff[j] + col, t4);
                    }
                }
            }
            TR(23, it_e, t128 == 0);
        };
        uint32_t it = 0;
        int tile_prev = -1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            float4 v[16];
            int b0, rem0;
            tile_base(tile, b0, rem0);
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int64_t off = row_addr(tile, b0, rem0, p * 8 + rg);
                v[p] = off >= 0 ? *reinterpret_cast<const float4*>(a.x + off + l16 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                float s = (v[p].x + v[p].y) + (v[p].z + v[p].w);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s * (1.f / (float)FC);
                const float d0 = v[p].x - mean, d1 = v[p].y - mean, d2 = v[p].z - mean, d3 = v[p].w - mean;
                float sq = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                const float rstd = rsqrtf(sq * (1.f / (float)FC) + a.eps);
                v[p] = make_float4(d0 * rstd * gam.x, d1 * rstd * gam.y, d2 * rstd * gam.z, d3 * rstd * gam.w);
            }
            TR(24, it, t128 == 0);
            ptx::mbar_wait(bar_at(smem, B_XN_EMPTY), (it & 1) ^ 1);     // previous tile's qkv GEMMs are done with XN
            TR(25, it, t128 == 0);
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int row = p * 8 + rg;
                uint2 hv, lv;
                split2(v[p].x, v[p].y, hv.x, lv.x);
                split2(v[p].z, v[p].w, hv.y, lv.y);
                const uint32_t off = sb_ln + sw_off(row, l16 >> 1) + ((l16 & 1) << 3);
                sts64(off, hv.x, hv.y);
                sts64(off + 16384, lv.x, lv.y);
            }
            ptx::fence_proxy_async();
            warp_arrive(bar_at(smem, B_XN_FULL), lane);
            TR(26, it, t128 == 0);
            {   // next tile's rows towards L2 while this group is busy with the epilogue
                const int tn = tile + (int)gridDim.x;
                if (tn < n_tiles) {
                    int bn, remn;
                    tile_base(tn, bn, remn);
                    const int64_t off = row_addr(tn, bn, remn, t128);
                    if (off >= 0) { ptx::prefetch_l2(a.x + off); ptx::prefetch_l2(a.x + off + 32); }
                }
            }
            if (tile_prev >= 0) epilogue(tile_prev, it - 1);           // the previous tile finishes ~3 heads after its last qkv GEMM
            tile_prev = tile;
        }
        if (tile_prev >= 0) epilogue(tile_prev, it - 1);
    } else {
        // ===================== three compute warp-groups: A = warps 6-9, B = warps 10-13, C = warps 14-17 =====================
        const int tc = (int)threadIdx.x - 192;          // 0..383
        const int wg = tc >> 7;                          // 0: A (q, k, QK), 1: B (v, O, out, epilogue), 2: C (softmax, PV)
        const bool first = (tc & 127) == 0;              // the group's MMA-issuing thread
        const int q = warp & 3;                          // TMEM lane quarter
        const int r = q * 32 + lane;                     // tile row of this thread
        const int rpx = r / FL;                          // 0..3 (3 = pad rows)
        const int px = rpx < FG ? rpx : FG - 1;
        const int fr = r - rpx * FL;                     // frame of this row (pad rows: 0..7)
        const bool row_real = r < FG * FL;
        const int pxlo = (q * 32) / FL;
        const int pxhi_raw = (q * 32 + 31) / FL;
        const int pxhi = pxhi_raw < FG ? pxhi_raw : FG - 1;
        const bool straddle = pxhi != pxlo;              // warp-uniform
        const bool use_hi = straddle && px != pxlo;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int dbg_ld = 3 * hid + heads * FL + hid;
        auto grow_of = [&](int tile) -> int64_t {
            const int pc = tile * FG + rpx;
            return (row_real && pc < a.n_pc) ? ((int64_t)(pc / a.pix) * FL + fr) * a.pix + (pc % a.pix) : -1;
        };

        if (wg == 0) {
            // ---------- WG-A: q*scale, rotary(q), rotary(k) -> operand rows; QK(g) ----------
            float rc[16], rs[16];                        // rotary table row of this thread's frame
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(a.rot_cos + fr * 16 + i);
                const float4 s4 = *reinterpret_cast<const float4*>(a.rot_sin + fr * 16 + i);
                rc[i] = c4.x; rc[i + 1] = c4.y; rc[i + 2] = c4.z; rc[i + 3] = c4.w;
                rs[i] = s4.x; rs[i + 1] = s4.y; rs[i + 2] = s4.z; rs[i + 3] = s4.w;
            }
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int64_t my_grow = DBG ? grow_of(tile) : -1;
                for (int h = 0; h < heads; ++h, ++g) {
                    const uint32_t s = g & 1u;
                    const uint32_t t_qkv = lane_base + T_QKV + 96u * s;
                    TR(8, g, first);
                    ptx::mbar_wait(bar_at(smem, B_QKV_FULL + s), (g >> 1) & 1);
                    TR(9, g, first);
                    if (g > 0) ptx::mbar_wait(bar_at(smem, B_S_FULL), (g - 1) & 1);     // QK(g-1) is done with Q / K
                    ptx::tc_fence_after();
                    TR(10, g, first);
#pragma unroll
                    for (int part = 0; part < 4; ++part) {       // q dims 0-15, 16-31, k dims 0-15, 16-31
                        uint32_t u[16];
                        float v[16];
                        tmem_ld16(t_qkv + 16u * (uint32_t)part, u);
                        ptx::tmem_ld_wait();
                        // (q * 32^-0.5 of the reference, :325, is folded into the packed q rows of W_qkv on the host)
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float x0 = __uint_as_float(u[2 * i]), y0 = __uint_as_float(u[2 * i + 1]);
                            const float cc = rc[8 * (part & 1) + i], sn = rs[8 * (part & 1) + i];
                            v[2 * i] = x0 * cc - y0 * sn;
                            v[2 * i + 1] = y0 * cc + x0 * sn;
                        }
                        store_hilo_half(sb + (part < 2 ? OFF_Q : OFF_K), r, part & 1, v);
                        if (DBG) {
                            if (a.dbg && my_grow >= 0) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) a.dbg[my_grow * dbg_ld + (part < 2 ? 0 : hid) + h * 32 + 16 * (part & 1) + i] = v[i];
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::mbar_arrive(bar_at(smem, B_QK_READY));
                        ptx::mbar_arrive(bar_at(smem, B_QKV_EMPTY + s));
                    }
                    TR(11, g, first);
                }
            }
        } else if (wg == 2) {
            // ---------- WG-C: softmax rows; PV(g) ----------
            constexpr float L2E = 1.4426950408889634f;
            uint32_t g = 0, it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int64_t my_grow = DBG ? grow_of(tile) : -1;
                for (int h = 0; h < heads; ++h, ++g) {
                    TR(12, g, first);
                    ptx::mbar_wait(bar_at(smem, B_S_FULL), g & 1);
                    ptx::tc_fence_after();
                    TR(13, g, first);
                    float sv[FL];
                    {
                        // this row's 40 scores (column block of its pixel): all loads in flight, one wait per batch
                        const uint32_t t_s = lane_base + T_S;
                        if (!straddle) {
                            uint32_t u[FL / 8][8];
#pragma unroll
                            for (int c = 0; c < FL / 8; ++c) tmem_ld8(t_s + (uint32_t)(FL * pxlo + 8 * c), u[c]);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < FL / 8; ++c)
#pragma unroll
                                for (int j = 0; j < 8; ++j) sv[8 * c + j] = __uint_as_float(u[c][j]);
                        } else {
                            {
                                uint32_t u0[3][8], u1[3][8];
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    tmem_ld8(t_s + (uint32_t)(FL * pxlo + 8 * c), u0[c]);
                                    tmem_ld8(t_s + (uint32_t)(FL * pxhi + 8 * c), u1[c]);
                                }
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int c = 0; c < 3; ++c)
#pragma unroll
                                    for (int j = 0; j < 8; ++j) sv[8 * c + j] = __uint_as_float(use_hi ? u1[c][j] : u0[c][j]);
                            }
                            {
                                uint32_t u0[2][8], u1[2][8];
#pragma unroll
                                for (int c = 0; c < 2; ++c) {
                                    tmem_ld8(t_s + (uint32_t)(FL * pxlo + 24 + 8 * c), u0[c]);
                                    tmem_ld8(t_s + (uint32_t)(FL * pxhi + 24 + 8 * c), u1[c]);
                                }
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int c = 0; c < 2; ++c)
#pragma unroll
                                    for (int j = 0; j < 8; ++j) sv[24 + 8 * c + j] = __uint_as_float(use_hi ? u1[c][j] : u0[c][j]);
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    warp_arrive(bar_at(smem, B_S_EMPTY), lane);          // QK(g+1) may overwrite the scores
                    if (a.pos_bias) {      // T5 relative-position bias row (h, frame): L1 / L2 resident
                        const float4* bp = reinterpret_cast<const float4*>(a.pos_bias + ((int64_t)h * FL + fr) * FL);
#pragma unroll
                        for (int i = 0; i < FL / 4; ++i) {
                            const float4 t4 = __ldg(bp + i);
                            sv[4 * i] += t4.x; sv[4 * i + 1] += t4.y; sv[4 * i + 2] += t4.z; sv[4 * i + 3] += t4.w;
                        }
                    }
                    float m0 = fmaxf(sv[0], sv[1]), m1 = fmaxf(sv[2], sv[3]), m2 = fmaxf(sv[4], sv[5]), m3 = fmaxf(sv[6], sv[7]);
#pragma unroll
                    for (int j = 8; j < FL; j += 4) {
                        m0 = fmaxf(m0, sv[j]); m1 = fmaxf(m1, sv[j + 1]); m2 = fmaxf(m2, sv[j + 2]); m3 = fmaxf(m3, sv[j + 3]);
                    }
                    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    const float nb = -mx * L2E;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int j = 0; j < FL; j += 4) {
                        sv[j] = ex2(fmaf(sv[j], L2E, nb)); sv[j + 1] = ex2(fmaf(sv[j + 1], L2E, nb));
                        sv[j + 2] = ex2(fmaf(sv[j + 2], L2E, nb)); sv[j + 3] = ex2(fmaf(sv[j + 3], L2E, nb));
                        s0 += sv[j]; s1 += sv[j + 1]; s2 += sv[j + 2]; s3 += sv[j + 3];
                    }
                    // P stays un-normalised (exp(s - max) in (0, 1]); WG-B scales the PV result by 1 / sum instead (32 products
                    // per row there instead of 40 here: this group is the slowest stage of the pipeline)
                    const float inv = 1.f / ((s0 + s1) + (s2 + s3));
                    if (DBG) {
                        if (a.dbg && my_grow >= 0) {
#pragma unroll
                            for (int j = 0; j < FL; ++j) a.dbg[my_grow * dbg_ld + 3 * hid + h * FL + j] = sv[j] * inv;
                        }
                    }
                    TR(14, g, first);
                    // P is single-buffered: PV of the previous head must have read it
                    if (g > 0) ptx::mbar_wait(bar_at(smem, B_PVD_FULL), (g - 1) & 1);
                    TR(15, g, first);
                    // slot g & 1 was last read for head g-2, before WG-B released the accumulator that PV(g-1) -- just seen complete --
                    // had to wait for; WG-B sees this store through P_READY -> PV(g) -> PVD_FULL
                    reinterpret_cast<float*>(smem + OFF_INV)[(g & 1u) * 128u + (uint32_t)r] = inv;
#pragma unroll
                    for (int c = 0; c < FL / 8; ++c) {
                        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                        split2(sv[8 * c], sv[8 * c + 1], h0, l0);
                        split2(sv[8 * c + 2], sv[8 * c + 3], h1, l1);
                        split2(sv[8 * c + 4], sv[8 * c + 5], h2, l2);
                        split2(sv[8 * c + 6], sv[8 * c + 7], h3, l3);
                        sts128(sb + OFF_P + sw_off(r, c), h0, h1, h2, h3);
                        sts128(sb + OFF_P + 16384 + sw_off(r, c), l0, l1, l2, l3);
                    }
                    ptx::fence_proxy_async();
                    warp_arrive(bar_at(smem, B_P_READY), lane);
                    TR(16, g, first);
                }
            }
        } else {
            // ---------- WG-B: v(g) -> transposed compact operand (one head ahead); PV(g-1) result -> O operand rows, out(g-1);
            // ---------- tile epilogue ----------
            const int t128 = tc & 127;
            const int l16 = t128 & 15, rg = t128 >> 4;   // store mapping of the epilogue: 16 lanes per row, 8 rows per pass
            int h_v = 0, tile_v = blockIdx.x;            // cursor of the v stage
            int h_o = 0, tile_o = blockIdx.x;            // cursor of the O / out stage
            int64_t grow_v = DBG ? grow_of(tile_v) : -1, grow_o = grow_v;
            // v of head g -> B operand buffer g & 1 (last read by PV(g-2), whose completion this group has waited for)
            auto do_v = [&](uint32_t g) {
                const uint32_t s = g & 1u;
                const uint32_t t_qkv = lane_base + T_QKV + 96u * s;
                TR(17, g, first);
                ptx::mbar_wait(bar_at(smem, B_QKV_FULL + s), (g >> 1) & 1);
                ptx::tc_fence_after();
                TR(18, g, first);
                const uint32_t vt = sb + OFF_VT + s * VT_BUF;
                if (VMN) {
                    uint32_t u0[16], u1[16];
                    tmem_ld16(t_qkv + 64u, u0);
                    tmem_ld16(t_qkv + 80u, u1);
                    ptx::tmem_ld_wait();
                    if (row_real) {
                        // row k = frame of the [position][column] tile; this row's 32 columns = 64 contiguous bytes
                        const uint32_t base = vt + (uint32_t)((px >> 1) * VMN_ATOM + (fr >> 3) * 1024 + (fr & 7) * 128);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t* u = c < 2 ? u0 : u1;
                            const int o = (c & 1) * 8;
                            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                            split2(__uint_as_float(u[o]), __uint_as_float(u[o + 1]), h0, l0);
                            split2(__uint_as_float(u[o + 2]), __uint_as_float(u[o + 3]), h1, l1);
                            split2(__uint_as_float(u[o + 4]), __uint_as_float(u[o + 5]), h2, l2);
                            split2(__uint_as_float(u[o + 6]), __uint_as_float(u[o + 7]), h3, l3);
                            const uint32_t off = base + (uint32_t)(((((px & 1) << 2) + c) ^ (fr & 7)) << 4);
                            sts128(off, h0, h1, h2, h3);
                            sts128(off + VT_PLANE, l0, l1, l2, l3);
                        }
                        if (DBG) {
                            if (a.dbg && grow_v >= 0) {
#pragma unroll
                                for (int dd = 0; dd < 16; ++dd) {
                                    a.dbg[grow_v * dbg_ld + 2 * hid + h_v * 32 + dd] = __uint_as_float(u0[dd]);
                                    a.dbg[grow_v * dbg_ld + 2 * hid + h_v * 32 + 16 + dd] = __uint_as_float(u1[dd]);
                                }
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        uint32_t u[16];
                        tmem_ld16(t_qkv + 64u + 16u * (uint32_t)part, u);
                        ptx::tmem_ld_wait();
                        if (row_real) {
#pragma unroll
                            for (int dd = 0; dd < 16; ++dd) {
                                const int d = 16 * part + dd;
                                const float x0 = __uint_as_float(u[dd]);
                                const bf16 hi = __float2bfloat16_rn(x0);
                                const bf16 lo = __float2bfloat16_rn(x0 - __bfloat162float(hi));
                                const int n = px * 32 + d;
                                const uint32_t off = vt + sw_off(n, fr >> 3) + (uint32_t)((fr & 7) * 2);
                                sts16(off, *reinterpret_cast<const uint16_t*>(&hi));
                                sts16(off + VT_PLANE, *reinterpret_cast<const uint16_t*>(&lo));
                            }
                            if (DBG) {
                                if (a.dbg && grow_v >= 0) {
#pragma unroll
                                    for (int dd = 0; dd < 16; ++dd) a.dbg[grow_v * dbg_ld + 2 * hid + h_v * 32 + 16 * part + dd] = __uint_as_float(u[dd]);
                                }
                            }
                        }
                    }
                }
                ptx::tc_fence_before();
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(bar_at(smem, B_VT_READY + s));
                    ptx::mbar_arrive(bar_at(smem, B_QKV_EMPTY + s));
                }
                TR(19, g, first);
                if (++h_v == heads) { h_v = 0; tile_v += gridDim.x; if (DBG) grow_v = grow_of(tile_v); }
            };
            do_v(0);
            if (NG > 1) do_v(1);
            // downstream first: O(gp) / out(gp) as soon as PV(gp) lands (PV(gp+1) waits for this group to drain the accumulator),
            // then v two heads ahead while the softmax of the next head is still running
            for (uint32_t gp = 0; gp < NG; ++gp) {
                // ---- this row's 32 output columns of PV(gp) -> O operand row
                ptx::mbar_wait(bar_at(smem, B_PVD_FULL), gp & 1);
                ptx::tc_fence_after();
                TR(20, gp, first);
                float ov[32];
                {
                    uint32_t u0[4][8], u1[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        tmem_ld8(lane_base + T_PVD + (uint32_t)(32 * pxlo + 8 * c), u0[c]);
                        if (straddle) tmem_ld8(lane_base + T_PVD + (uint32_t)(32 * pxhi + 8 * c), u1[c]);
                    }
                    ptx::tmem_ld_wait();
#pragma unroll
                    const float inv = reinterpret_cast<const float*>(smem + OFF_INV)[(gp & 1u) * 128u + (uint32_t)r];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 8; ++j) ov[8 * c + j] = __uint_as_float(use_hi ? u1[c][j] : u0[c][j]) * inv;
                }
                ptx::tc_fence_before();
                warp_arrive(bar_at(smem, B_PVD_EMPTY), lane);            // PV(gp+1) may overwrite the accumulator
                if (gp > 0) ptx::mbar_wait(bar_at(smem, B_O_FREE), (gp - 1) & 1);       // out(gp-1) has read the O operand
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = ov[16 * half + i];
                    store_hilo_half(sb + OFF_O, r, half, v);
                }
                if (DBG) {
                    if (a.dbg && grow_o >= 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) a.dbg[grow_o * dbg_ld + 3 * hid + heads * FL + h_o * 32 + i] = ov[i];
                    }
                }
                ptx::fence_proxy_async();
                warp_arrive(bar_at(smem, B_O_READY), lane);
                TR(21, gp, first);
                if (gp + 2 < NG) do_v(gp + 2);
                if (++h_o == heads) { h_o = 0; tile_o += gridDim.x; if (DBG) grow_o = grow_of(tile_o); }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int lfdm_attn_temporal_fused(const float* x, const float* gamma, const void* wq_packed, const void* wo_packed,
                             const float* out_bias, const float* rot_cos, const float* rot_sin, const float* pos_bias,
                             float* out_f32, void* out_sb, int64_t out_plane, int n_b, int frames, int pixels, int c,
                             int heads, float eps, float* debug, void* stream) {
    if (!x || !gamma || !wq_packed || !wo_packed || !rot_cos || !rot_sin || (!out_f32 && !out_sb)) return LFDM_E_BADARG;
    if (frames != FL || c != FC || heads < 1 || heads > 16 || n_b < 1 || pixels < 1) return LFDM_E_UNSUPP;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq_packed) | reinterpret_cast<uintptr_t>(wo_packed) |
         reinterpret_cast<uintptr_t>(rot_cos) | reinterpret_cast<uintptr_t>(rot_sin) | reinterpret_cast<uintptr_t>(pos_bias) |
         reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(out_bias) | reinterpret_cast<uintptr_t>(out_f32) |
         reinterpret_cast<uintptr_t>(out_sb)) & 15)
        return LFDM_E_UNSUPP;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_done[64] = {};
    static int sms[64] = {};
    if (dev < 0 || dev >= 64) return LFDM_E_UNSUPP;
    static const bool vmn = (getenv("LFDM_ATTN_VT_KMAJOR") == nullptr);     // A/B switch: K-major transposed V^T operand
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_temporal_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_temporal_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_temporal_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_temporal_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
        attr_done[dev] = true;
    }
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.gamma = gamma;
    a.wq = reinterpret_cast<const uint8_t*>(wq_packed); a.wo = reinterpret_cast<const uint8_t*>(wo_packed);
    a.out_bias = out_bias; a.rot_cos = rot_cos; a.rot_sin = rot_sin; a.pos_bias = pos_bias;
    a.out_f32 = out_f32; a.out_sb = reinterpret_cast<bf16*>(out_sb); a.out_plane = out_plane; a.dbg = debug;
    static const bool trace_mode = (getenv("LFDM_ATTN_TRACE") != nullptr);     // `debug` is then a [32][64] int64 stamp buffer
    if (trace_mode && debug) { a.trace = reinterpret_cast<long long*>(debug); a.dbg = nullptr; debug = nullptr; }
    a.heads = heads; a.n_pc = n_b * pixels; a.pix = pixels; a.n_tiles = (a.n_pc + FG - 1) / FG; a.eps = eps;
    const int grid = a.n_tiles < sms[dev] ? a.n_tiles : sms[dev];
    if (debug && vmn)
        LFDM_LAUNCH_PDL((attn_temporal_fused_kernel<true, true>), dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a);
    else if (debug)
        LFDM_LAUNCH_PDL((attn_temporal_fused_kernel<true, false>), dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a);
    else if (vmn)
        LFDM_LAUNCH_PDL((attn_temporal_fused_kernel<false, true>), dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a);
    else
        LFDM_LAUNCH_PDL((attn_temporal_fused_kernel<false, false>), dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a);
    return 0;
}
