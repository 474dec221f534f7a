// attn_fused.cu — the whole temporal-attention block of the LFDM UNet as ONE persistent tcgen05 kernel (sm_100a):
//
//     out = x + to_out( softmax( rot(q*scale) rot(k)^T + pos_bias ) v ),   q|k|v = to_qkv( LayerNorm(x) )
//
// Replaces Residual(PreNorm(EinopsToAndFrom(Attention))) of the reference (DM/modules/video_flow_diffusion.py:132-138,
// 170-190, 270-283, 286-363) for C = 64 channels and 40 frames (the 32x32 / 16x16 levels: init_temporal_attn, downs.0,
// ups.2, ups.3).  q|k|v, the scores and the per-head outputs never leave the SM: the four GEMM-shaped stages run on
// tcgen05 with fp32 accumulators in TMEM, everything between them is done by two warp-groups straight out of TMEM.
//
// One tile = 3 pixel columns x 40 frames = 120 rows of the row matrix (padded to the 128-row UMMA tile).  Per tile:
//   LN      : 256 threads, 16 lanes per row: x -> LayerNorm -> split-bf16 A operand (128x64, SW128 K-major) + fp32 copy
//   per head h (weights streamed from L2 with cp.async.bulk into a 2-stage ring, pre-swizzled on the host):
//     qkv   : D[128x96]  = Xn . W_h^T            (3 split-bf16 products x 4 K-steps, double-buffered TMEM)
//     WG-A  : q*scale, rotary(q), rotary(k) from TMEM -> [hi|lo] operand rows in smem
//     WG-B  : v from TMEM -> transposed "compact" B operand  V^T[(pixel, d)][j]
//     QK    : S[128x128] = Q . K^T               (all 3 pixels at once; each row only uses its own 40-column block)
//     WG-A  : +bias, softmax over the row's 40 columns, P -> compact A operand [128 x 48] (hi / lo planes)
//     PV    : D[128x96]  = P . V^T^T             (column block 32*pixel(row) is the row's result)
//     WG-B  : own 32 columns -> split-bf16 [hi|lo] operand rows
//     out   : OUT[128x64] += O_h . Wout_h^T      (accumulated over heads in TMEM)
//   epilogue: OUT + bias + x (kept in smem) -> F32 and split-bf16 rows, coalesced.
// Synchronisation is mbarrier-only (tcgen05.commit for MMA completion, counted arrivals for the warp-groups).
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int FL = 40;       // frames (sequence length)
constexpr int FG = 3;        // pixel columns per tile
constexpr int FC = 64;       // channels
constexpr int NTHREADS = 384;

constexpr int OFF_XN = 0;            // 2 planes x 16 KiB : LayerNorm output, A operand of the qkv GEMM
constexpr int OFF_Q = 32768;         // 16 KiB : rows [q_hi(32) | q_lo(32)]
constexpr int OFF_K = 49152;         // 16 KiB : rows [k_hi | k_lo]
constexpr int OFF_VT = 65536;        // 2 planes x 12 KiB : V^T, 96 rows (pixel, d) x 64 positions (48 used)
constexpr int VT_PLANE = 12288;
constexpr int OFF_P = 90112;         // 2 planes x 16 KiB : P rows x 64 positions (48 used); hi plane doubles as the O tile
constexpr int OFF_WQ = 122880;       // 2 stages x 24 KiB : [W_hi (96 x 64) | W_lo (96 x 64)] of one head
constexpr int WQ_STAGE = 24576;
constexpr int OFF_WO = 172032;       // 2 stages x 8 KiB : 64 rows [w_hi(32) | w_lo(32)] of one head
constexpr int WO_STAGE = 8192;
constexpr int OFF_XR = 188416;       // 32 KiB : fp32 copy of the x tile (residual + coalescing stage of the output)
constexpr int OFF_BAR = 221184;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;

// TMEM column map (512 columns allocated)
constexpr uint32_t T_QKV = 0;        // 2 x 96
constexpr uint32_t T_S = 192;        // 128
constexpr uint32_t T_PVD = 320;      // 96
constexpr uint32_t T_OUT = 416;      // 64

enum {
    B_XN_FULL = 0, B_WQ_FULL = 1, B_WQ_EMPTY = 3, B_WO_FULL = 5, B_WO_EMPTY = 7, B_QKV_FULL = 9, B_QKV_EMPTY = 11,
    B_QK_READY = 13, B_S_FULL = 14, B_P_READY = 15, B_VT_READY = 16, B_PVD_FULL = 17, B_O_READY = 18, B_OUT_FULL = 19,
    B_OUT_EMPTY = 20, B_COUNT = 21
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
    int32_t heads, n_pc, pix, n_tiles;
    float eps;
};

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ptx::smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(ptx::smem_u32(bar))
                 : "memory");
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
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of 16-byte chunk `c` of row `r` in a SW128 K-major tile (rows of 128 B, 8-row groups of 1024 B)
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// 32 fp32 values of one row -> [hi(32) | lo(32)] operand row (chunks 0-3 hi, 4-7 lo)
__device__ __forceinline__ void store_hilo_row(uint8_t* tile, int r, const float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 h, l;
        split2(v[8 * c], v[8 * c + 1], h.x, l.x);
        split2(v[8 * c + 2], v[8 * c + 3], h.y, l.y);
        split2(v[8 * c + 4], v[8 * c + 5], h.z, l.z);
        split2(v[8 * c + 6], v[8 * c + 7], h.w, l.w);
        *reinterpret_cast<uint4*>(tile + sw_off(r, c)) = h;
        *reinterpret_cast<uint4*>(tile + sw_off(r, c + 4)) = l;
    }
}

__device__ __forceinline__ uint64_t* bar_at(uint8_t* smem, int i) { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + i; }

__global__ void __launch_bounds__(NTHREADS, 1) attn_temporal_fused_kernel(const __grid_constant__ FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 256);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int heads = a.heads;
    const int hid = heads * 32;

    pdl_trigger();
    if (warp == 1 && ptx::elect_one()) {
        ptx::mbar_init(bar_at(smem, B_XN_FULL), 256);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_at(smem, B_WQ_FULL + i), 1);
            ptx::mbar_init(bar_at(smem, B_WQ_EMPTY + i), 1);
            ptx::mbar_init(bar_at(smem, B_WO_FULL + i), 1);
            ptx::mbar_init(bar_at(smem, B_WO_EMPTY + i), 1);
            ptx::mbar_init(bar_at(smem, B_QKV_FULL + i), 1);
            ptx::mbar_init(bar_at(smem, B_QKV_EMPTY + i), 256);
        }
        ptx::mbar_init(bar_at(smem, B_QK_READY), 128);
        ptx::mbar_init(bar_at(smem, B_S_FULL), 1);
        ptx::mbar_init(bar_at(smem, B_P_READY), 128);
        ptx::mbar_init(bar_at(smem, B_VT_READY), 128);
        ptx::mbar_init(bar_at(smem, B_PVD_FULL), 1);
        ptx::mbar_init(bar_at(smem, B_O_READY), 128);
        ptx::mbar_init(bar_at(smem, B_OUT_FULL), 1);
        ptx::mbar_init(bar_at(smem, B_OUT_EMPTY), 256);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr, 512);
        ptx::tmem_relinquish();
    }
    // V^T pad positions (40..47 of every row) are never written afterwards: zero the operand once
    for (int i = threadIdx.x; i < (2 * VT_PLANE) / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(smem + OFF_VT)[i] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();

    const int n_tiles = a.n_tiles;

    if (warp == 1) {
        // ===================== weight producer: one head's W_qkv / W_out slices per ring stage =====================
        if (ptx::elect_one()) {
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int h = 0; h < heads; ++h, ++g) {
                    const int s = g & 1;
                    const uint32_t par = ((g >> 1) & 1) ^ 1;
                    ptx::mbar_wait(bar_at(smem, B_WQ_EMPTY + s), par);
                    ptx::mbar_arrive_expect_tx(bar_at(smem, B_WQ_FULL + s), WQ_STAGE);
                    bulk_copy_g2s(smem + OFF_WQ + s * WQ_STAGE, a.wq + (size_t)h * WQ_STAGE, WQ_STAGE, bar_at(smem, B_WQ_FULL + s));
                    ptx::mbar_wait(bar_at(smem, B_WO_EMPTY + s), par);
                    ptx::mbar_arrive_expect_tx(bar_at(smem, B_WO_FULL + s), WO_STAGE);
                    bulk_copy_g2s(smem + OFF_WO + s * WO_STAGE, a.wo + (size_t)h * WO_STAGE, WO_STAGE, bar_at(smem, B_WO_FULL + s));
                }
            }
        }
    } else if (warp == 0) {
        // ===================== MMA issuer (one elected thread) =====================
        if (ptx::elect_one()) {
            const uint32_t sb = ptx::smem_u32(smem);
            const uint64_t d_xn = ptx::make_sw128_kmajor_desc(sb + OFF_XN);
            const uint64_t d_q = ptx::make_sw128_kmajor_desc(sb + OFF_Q);
            const uint64_t d_k = ptx::make_sw128_kmajor_desc(sb + OFF_K);
            const uint64_t d_vt = ptx::make_sw128_kmajor_desc(sb + OFF_VT);
            const uint64_t d_p = ptx::make_sw128_kmajor_desc(sb + OFF_P);
            const uint64_t d_wq = ptx::make_sw128_kmajor_desc(sb + OFF_WQ);
            const uint64_t d_wo = ptx::make_sw128_kmajor_desc(sb + OFF_WO);
            constexpr uint32_t ID96 = ptx::make_idesc_bf16(128, 96);
            constexpr uint32_t ID128 = ptx::make_idesc_bf16(128, 128);
            constexpr uint32_t ID64 = ptx::make_idesc_bf16(128, 64);
            auto issue_qkv = [&](uint32_t gq) {
                const int s = gq & 1;
                const uint32_t par = (gq >> 1) & 1;
                ptx::mbar_wait(bar_at(smem, B_WQ_FULL + s), par);
                ptx::mbar_wait(bar_at(smem, B_QKV_EMPTY + s), par ^ 1);
                ptx::tc_fence_after();
                const uint32_t td = tmem_base + T_QKV + 96u * (uint32_t)s;
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
            };
            uint32_t g = 0, it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                ptx::mbar_wait(bar_at(smem, B_XN_FULL), it & 1);
                ptx::tc_fence_after();
                issue_qkv(g);
                for (int h = 0; h < heads; ++h) {
                    const uint32_t gh = g + (uint32_t)h;
                    if (h + 1 < heads) issue_qkv(gh + 1);
                    // ---- S = Q K^T : rows [hi | lo]: K-steps 0,1 = hi dims, 2,3 = lo dims
                    ptx::mbar_wait(bar_at(smem, B_QK_READY), gh & 1);
                    ptx::tc_fence_after();
                    {
                        const uint32_t td = tmem_base + T_S;
                        ptx::umma_bf16(td, d_q + 4, d_k + 0, ID128, 0u);      // q_lo . k_hi
                        ptx::umma_bf16(td, d_q + 6, d_k + 2, ID128, 1u);
                        ptx::umma_bf16(td, d_q + 0, d_k + 4, ID128, 1u);      // q_hi . k_lo
                        ptx::umma_bf16(td, d_q + 2, d_k + 6, ID128, 1u);
                        ptx::umma_bf16(td, d_q + 0, d_k + 0, ID128, 1u);      // q_hi . k_hi
                        ptx::umma_bf16(td, d_q + 2, d_k + 2, ID128, 1u);
                    }
                    ptx::umma_commit(bar_at(smem, B_S_FULL));
                    // ---- D = P V : compact K = 48 positions; D column block 32*pixel
                    ptx::mbar_wait(bar_at(smem, B_P_READY), gh & 1);
                    ptx::mbar_wait(bar_at(smem, B_VT_READY), gh & 1);
                    ptx::tc_fence_after();
                    {
                        const uint32_t td = tmem_base + T_PVD;
                        const uint64_t p_hi = d_p, p_lo = d_p + (uint64_t)(16384 >> 4);
                        const uint64_t v_hi = d_vt, v_lo = d_vt + (uint64_t)(VT_PLANE >> 4);
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) {
                            const uint64_t o = (uint64_t)(ks * 2);
                            ptx::umma_bf16(td, p_lo + o, v_hi + o, ID96, ks > 0 ? 1u : 0u);
                            ptx::umma_bf16(td, p_hi + o, v_lo + o, ID96, 1u);
                            ptx::umma_bf16(td, p_hi + o, v_hi + o, ID96, 1u);
                        }
                    }
                    ptx::umma_commit(bar_at(smem, B_PVD_FULL));
                    // ---- OUT += O_h Wout_h^T
                    const int so = gh & 1;
                    ptx::mbar_wait(bar_at(smem, B_O_READY), gh & 1);
                    ptx::mbar_wait(bar_at(smem, B_WO_FULL + so), (gh >> 1) & 1);
                    if (h == 0) ptx::mbar_wait(bar_at(smem, B_OUT_EMPTY), (it & 1) ^ 1);
                    ptx::tc_fence_after();
                    {
                        const uint32_t td = tmem_base + T_OUT;
                        const uint64_t w = d_wo + (uint64_t)((so * WO_STAGE) >> 4);
                        const uint64_t o_t = d_p;      // O tile lives in the P hi plane
                        ptx::umma_bf16(td, o_t + 4, w + 0, ID64, h > 0 ? 1u : 0u);   // o_lo . w_hi
                        ptx::umma_bf16(td, o_t + 6, w + 2, ID64, 1u);
                        ptx::umma_bf16(td, o_t + 0, w + 4, ID64, 1u);               // o_hi . w_lo
                        ptx::umma_bf16(td, o_t + 2, w + 6, ID64, 1u);
                        ptx::umma_bf16(td, o_t + 0, w + 0, ID64, 1u);               // o_hi . w_hi
                        ptx::umma_bf16(td, o_t + 2, w + 2, ID64, 1u);
                    }
                    ptx::umma_commit(bar_at(smem, B_WO_EMPTY + so));
                    if (h == heads - 1) ptx::umma_commit(bar_at(smem, B_OUT_FULL));
                }
                g += (uint32_t)heads;
            }
        }
    } else if (warp >= 4) {
        // ===================== two compute warp-groups: A = warps 4-7, B = warps 8-11 =====================
        const int tc = (int)threadIdx.x - 128;          // 0..255
        const int wg = tc >> 7;                          // 0: A, 1: B
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
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        float* xr = reinterpret_cast<float*>(smem + OFF_XR);
        const int dbg_ld = 3 * hid + heads * FL + hid;

        // rotary table row of this thread's frame (WG-A only uses it)
        float rc[16], rs[16];
        if (wg == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(a.rot_cos + fr * 16 + i);
                const float4 s4 = *reinterpret_cast<const float4*>(a.rot_sin + fr * 16 + i);
                rc[i] = c4.x; rc[i + 1] = c4.y; rc[i + 2] = c4.z; rc[i + 3] = c4.w;
                rs[i] = s4.x; rs[i + 1] = s4.y; rs[i + 2] = s4.z; rs[i + 3] = s4.w;
            }
        }
        const int l16 = tc & 15, rg = tc >> 4;           // LayerNorm / store mapping: 16 lanes per row, 16 rows per pass
        const float4 gam = *reinterpret_cast<const float4*>(a.gamma + l16 * 4);
        float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.out_bias) ob = *reinterpret_cast<const float4*>(a.out_bias + l16 * 4);

        uint32_t g = 0, it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            // ---------------- LayerNorm of the x tile -> XN (split-bf16 operand) + XR (fp32) ----------------
            {
                float4 v[8];
                int64_t grow[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int row = p * 16 + rg;
                    const int tpx = row / FL, tf = row - tpx * FL;
                    const int pc = tile * FG + tpx;
                    const bool ok = row < FG * FL && pc < a.n_pc;
                    grow[p] = ok ? ((int64_t)(pc / a.pix) * FL + tf) * a.pix + (pc % a.pix) : -1;
                    v[p] = ok ? *reinterpret_cast<const float4*>(a.x + grow[p] * FC + l16 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int row = p * 16 + rg;
                    float s = (v[p].x + v[p].y) + (v[p].z + v[p].w);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    const float mean = s / (float)FC;
                    const float d0 = v[p].x - mean, d1 = v[p].y - mean, d2 = v[p].z - mean, d3 = v[p].w - mean;
                    float sq = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    const float rstd = 1.f / sqrtf(sq / (float)FC + a.eps);
                    const float n0 = d0 * rstd * gam.x, n1 = d1 * rstd * gam.y, n2 = d2 * rstd * gam.z, n3 = d3 * rstd * gam.w;
                    uint2 hv, lv;
                    split2(n0, n1, hv.x, lv.x);
                    split2(n2, n3, hv.y, lv.y);
                    const uint32_t off = sw_off(row, l16 >> 1) + ((l16 & 1) << 3);
                    *reinterpret_cast<uint2*>(smem + OFF_XN + off) = hv;
                    *reinterpret_cast<uint2*>(smem + OFF_XN + 16384 + off) = lv;
                    reinterpret_cast<float4*>(xr)[row * 16 + (l16 ^ (row & 7))] = v[p];
                }
                ptx::fence_proxy_async();
                ptx::mbar_arrive(bar_at(smem, B_XN_FULL));
            }
            // global row of this thread's tile row (diagnostics only)
            int64_t my_grow = -1;
            {
                const int pc = tile * FG + rpx;
                if (row_real && pc < a.n_pc) my_grow = ((int64_t)(pc / a.pix) * FL + fr) * a.pix + (pc % a.pix);
            }

            for (int h = 0; h < heads; ++h) {
                const uint32_t gh = g + (uint32_t)h;
                const int b = gh & 1;
                const uint32_t t_qkv = lane_base + T_QKV + 96u * (uint32_t)b;
                if (wg == 0) {
                    // ---------- WG-A: q, k -> operand rows ----------
                    ptx::mbar_wait(bar_at(smem, B_QKV_FULL + b), (gh >> 1) & 1);
                    ptx::tc_fence_after();
                    {
                        uint32_t u[32];
                        float v[32];
                        tmem_ld32(t_qkv, u);
                        ptx::tmem_ld_wait();
                        const float scale = 0.17677669529663687f;        // 32^-0.5, reference :325
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float x0 = __uint_as_float(u[2 * i]) * scale, y0 = __uint_as_float(u[2 * i + 1]) * scale;
                            v[2 * i] = x0 * rc[i] - y0 * rs[i];
                            v[2 * i + 1] = y0 * rc[i] + x0 * rs[i];
                        }
                        store_hilo_row(smem + OFF_Q, r, v);
                        if (a.dbg && my_grow >= 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) a.dbg[my_grow * dbg_ld + h * 32 + i] = v[i];
                        }
                        tmem_ld32(t_qkv + 32, u);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float x0 = __uint_as_float(u[2 * i]), y0 = __uint_as_float(u[2 * i + 1]);
                            v[2 * i] = x0 * rc[i] - y0 * rs[i];
                            v[2 * i + 1] = y0 * rc[i] + x0 * rs[i];
                        }
                        store_hilo_row(smem + OFF_K, r, v);
                        if (a.dbg && my_grow >= 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) a.dbg[my_grow * dbg_ld + hid + h * 32 + i] = v[i];
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar_at(smem, B_QK_READY));
                    ptx::mbar_arrive(bar_at(smem, B_QKV_EMPTY + b));
                    // ---------- WG-A: softmax of this row's 40-column block ----------
                    ptx::mbar_wait(bar_at(smem, B_S_FULL), gh & 1);
                    ptx::tc_fence_after();
                    float sv[FL];
                    {
                        const uint32_t t_s = lane_base + T_S;
                        const bool use_hi = straddle && px != pxlo;
#pragma unroll
                        for (int c = 0; c < FL / 8; ++c) {       // 8 columns at a time keeps the live register set small
                            uint32_t u0[8], u1[8];
                            tmem_ld8(t_s + (uint32_t)(FL * pxlo + 8 * c), u0);
                            if (straddle) tmem_ld8(t_s + (uint32_t)(FL * pxhi + 8 * c), u1);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j) sv[8 * c + j] = __uint_as_float(use_hi ? u1[j] : u0[j]);
                        }
                    }
                    if (a.pos_bias) {      // T5 relative-position bias row (h, frame): 160 B, L1 / L2 resident
                        const float4* bp = reinterpret_cast<const float4*>(a.pos_bias + ((int64_t)h * FL + fr) * FL);
#pragma unroll
                        for (int i = 0; i < FL / 4; ++i) {
                            const float4 t4 = __ldg(bp + i);
                            sv[4 * i] += t4.x; sv[4 * i + 1] += t4.y; sv[4 * i + 2] += t4.z; sv[4 * i + 3] += t4.w;
                        }
                    }
                    float mx = sv[0];
#pragma unroll
                    for (int j = 1; j < FL; ++j) mx = fmaxf(mx, sv[j]);
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < FL; ++j) { sv[j] = __expf(sv[j] - mx); sum += sv[j]; }
                    const float inv = 1.f / sum;
#pragma unroll
                    for (int j = 0; j < FL; ++j) sv[j] *= inv;
                    if (a.dbg && my_grow >= 0) {
#pragma unroll
                        for (int j = 0; j < FL; ++j) a.dbg[my_grow * dbg_ld + 3 * hid + h * FL + j] = sv[j];
                    }
#pragma unroll
                    for (int c = 0; c < FL / 8; ++c) {
                        uint4 hh, ll;
                        split2(sv[8 * c], sv[8 * c + 1], hh.x, ll.x);
                        split2(sv[8 * c + 2], sv[8 * c + 3], hh.y, ll.y);
                        split2(sv[8 * c + 4], sv[8 * c + 5], hh.z, ll.z);
                        split2(sv[8 * c + 6], sv[8 * c + 7], hh.w, ll.w);
                        *reinterpret_cast<uint4*>(smem + OFF_P + sw_off(r, c)) = hh;
                        *reinterpret_cast<uint4*>(smem + OFF_P + 16384 + sw_off(r, c)) = ll;
                    }
                    {   // positions 40..47 (third K-step reads them): zero; the hi-plane chunk was overwritten by the O tile
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(smem + OFF_P + sw_off(r, FL / 8)) = z;
                        *reinterpret_cast<uint4*>(smem + OFF_P + 16384 + sw_off(r, FL / 8)) = z;
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar_at(smem, B_P_READY));
                } else {
                    // ---------- WG-B: v -> transposed compact operand ----------
                    ptx::mbar_wait(bar_at(smem, B_QKV_FULL + b), (gh >> 1) & 1);
                    ptx::tc_fence_after();
                    {
                        uint32_t u[32];
                        tmem_ld32(t_qkv + 64, u);
                        ptx::tmem_ld_wait();
                        if (row_real) {
                            uint8_t* vt = smem + OFF_VT;
#pragma unroll
                            for (int d = 0; d < 32; ++d) {
                                const float x0 = __uint_as_float(u[d]);
                                const bf16 hi = __float2bfloat16_rn(x0);
                                const bf16 lo = __float2bfloat16_rn(x0 - __bfloat162float(hi));
                                const int n = px * 32 + d;
                                const uint32_t off = sw_off(n, fr >> 3) + (uint32_t)((fr & 7) * 2);
                                *reinterpret_cast<bf16*>(vt + off) = hi;
                                *reinterpret_cast<bf16*>(vt + VT_PLANE + off) = lo;
                            }
                            if (a.dbg && my_grow >= 0) {
#pragma unroll
                                for (int d = 0; d < 32; ++d) a.dbg[my_grow * dbg_ld + 2 * hid + h * 32 + d] = __uint_as_float(u[d]);
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar_at(smem, B_VT_READY));
                    ptx::mbar_arrive(bar_at(smem, B_QKV_EMPTY + b));
                    // ---------- WG-B: this row's 32 output columns -> O operand row ----------
                    ptx::mbar_wait(bar_at(smem, B_PVD_FULL), gh & 1);
                    ptx::tc_fence_after();
                    {
                        float v[32];
                        const bool use_hi = straddle && px != pxlo;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t u0[8], u1[8];
                            tmem_ld8(lane_base + T_PVD + (uint32_t)(32 * pxlo + 8 * c), u0);
                            if (straddle) tmem_ld8(lane_base + T_PVD + (uint32_t)(32 * pxhi + 8 * c), u1);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[8 * c + j] = __uint_as_float(use_hi ? u1[j] : u0[j]);
                        }
                        store_hilo_row(smem + OFF_P, r, v);
                        if (a.dbg && my_grow >= 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) a.dbg[my_grow * dbg_ld + 3 * hid + heads * FL + h * 32 + i] = v[i];
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar_at(smem, B_O_READY));
                }
            }
            g += (uint32_t)heads;

            // ---------------- epilogue: OUT (+bias) + x -> global, coalesced through XR ----------------
            ptx::mbar_wait(bar_at(smem, B_OUT_FULL), it & 1);
            ptx::tc_fence_after();
            {
                uint32_t u[32];
                tmem_ld32(lane_base + T_OUT + 32u * (uint32_t)wg, u);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(bar_at(smem, B_OUT_EMPTY));
                float4* row4 = reinterpret_cast<float4*>(xr) + r * 16;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int lc = wg * 8 + c;
                    float4 t4 = row4[lc ^ (r & 7)];
                    t4.x += __uint_as_float(u[4 * c]); t4.y += __uint_as_float(u[4 * c + 1]);
                    t4.z += __uint_as_float(u[4 * c + 2]); t4.w += __uint_as_float(u[4 * c + 3]);
                    row4[lc ^ (r & 7)] = t4;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int row = p * 16 + rg;
                const int tpx = row / FL, tf = row - tpx * FL;
                const int pc = tile * FG + tpx;
                if (row < FG * FL && pc < a.n_pc) {
                    const int64_t gr = ((int64_t)(pc / a.pix) * FL + tf) * a.pix + (pc % a.pix);
                    float4 t4 = reinterpret_cast<const float4*>(xr)[row * 16 + (l16 ^ (row & 7))];
                    t4.x += ob.x; t4.y += ob.y; t4.z += ob.z; t4.w += ob.w;
                    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + gr * FC + l16 * 4) = t4;
                    if (a.out_sb) store_sb4(a.out_sb, a.out_plane, gr * FC + l16 * 4, t4);
                }
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
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_temporal_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
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
    a.heads = heads; a.n_pc = n_b * pixels; a.pix = pixels; a.n_tiles = (a.n_pc + FG - 1) / FG; a.eps = eps;
    const int grid = a.n_tiles < sms[dev] ? a.n_tiles : sms[dev];
    LFDM_LAUNCH_PDL(attn_temporal_fused_kernel, dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, a);
    return 0;
}
