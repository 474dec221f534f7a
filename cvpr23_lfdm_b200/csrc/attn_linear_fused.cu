// attn_linear_fused.cu — the whole spatial linear-attention block of the LFDM UNet on tcgen05 (sm_100a):
//
//     out = x + to_out( ctx^T (softmax_d(q) * scale) ) + bias,   ctx = softmax_n(k)^T v,   q|k|v = to_qkv( LayerNorm(x) )
//
// Replaces Residual(PreNorm(SpatialLinearAttention)) of the reference (DM/modules/video_flow_diffusion.py:132-138, 170-190,
// 240-265) for C = 64 channels, 8 heads x 32 (the 32x32 level: downs.0.2, ups.3.2).  The composed path runs LayerNorm, a
// qkv projection that writes 3 KiB of fp32 q|k|v per row to HBM, the attention core that reads them back, and an output
// projection; here q|k|v never leave the SM.  Frames are independent; one frame = `pos` consecutive rows.
//
// Three launches (the softmax over the POSITIONS of a frame needs all of k before any q can be used):
//   1. linattn_ctx_kernel   (persistent, 64-row tiles, contiguous tile range per CTA)
//        K^T[hd][pos] = W_k Xn^T and V^T[he][pos] = W_v Xn^T: the weights are the A operand (128 rows = 4 heads x 32), the
//        normalised rows the B operand, so a thread owns one (head, d) ROW and sees every position of the tile: the softmax
//        over positions is thread-local (online maximum kept in registers across the tiles of a frame; the TMEM context
//        accumulator is rescaled only when a row's maximum grows by more than 2^8).  P^T = exp2(K^T - max) and V^T go back
//        to shared memory as K-major split-bf16 operands and ctx[hd][he] += P^T V^T^T accumulates in TMEM (4 heads per
//        128x128 accumulator; only the 4 diagonal 32x32 blocks are read).  At the end of a frame segment every thread writes
//        its row (32 context values, maximum, denominator) to the partials buffer.
//   2. linattn_combine_kernel (one small block per (frame, 2-head chunk)): merges the <= LFDM_LINATTN_MAXP partials of a
//        frame, folds the q scale and 1 / denominator, and multiplies by W_out:  G[hd][c] = sum_e ctx[hd][e] W_out[c][h*32 + e],
//        written as the split-bf16 B-operand image of launch 3.
//   3. linattn_apply_kernel (persistent, 128-row tiles; Xn comes back as the operand images launch 1 stored -- one 32 KiB bulk
//        copy per tile instead of a second LayerNorm, which was this kernel's critical path): Q = Xn W_q^T in 64-column chunks (2 heads) through a 4-deep TMEM ring,
//        softmax over d per (row, head) in registers, Qs -> split-bf16 A operand, OUT[128x64] += Qs_chunk G_chunk, epilogue
//        OUT + bias + x -> F32 (and optional split-bf16) rows, coalesced by the quad transpose of attn_fused.cu.
// Both persistent kernels use the role layout measured on the temporal block (attn_fused.cu): one POLLING issuer thread per
// CTA (independent MMA streams, non-blocking mbarrier tests), a LayerNorm (+ epilogue) group that works one tile ahead, and
// compute warp-groups that only ever wait on mbarriers.
#include <cstdlib>
#include <cstring>
#include "fused_common.cuh"

namespace {
using namespace fz;

constexpr int FC = 64;               // channels
constexpr int HEADS = 8;
constexpr int HID = 256;             // heads * 32
constexpr int MAXP = LFDM_LINATTN_MAXP;
constexpr int PART_LD = 34;          // floats per partial row: 32 context values, maximum (log2 domain), denominator
constexpr int NTHREADS1 = 576;       // launch 1: 18 warps
constexpr int NTHREADS = 448;        // launch 3: 14 warps
constexpr float RESCALE_THRESHOLD = 8.f;       // log2 units: rows are re-based only when their maximum grows by > 2^8

struct LinArgs {
    const float* x;
    const float* gamma;
    const uint8_t* wk;           // [2 halves][2 planes][128 rows x 128 B]   SW128 images, rows (head, d)
    const uint8_t* wv;           // same, rows (head, e)
    const uint8_t* wq;           // [4 chunks][2 planes][64 rows x 128 B]    rows (head, d) of 2 heads
    const float* wout;           // [64][256] fp32
    const float* out_bias;       // [64] or null
    float* part;                 // [frames][MAXP][256][PART_LD]
    uint8_t* gimg;               // [frames][4 chunks][2 planes][64 rows x 128 B]  G^T operand images
    uint8_t* ximg;               // [rows / 128][2 planes][128 rows x 128 B]       LayerNorm(x) operand images (written by launch 1)
    float* out_f32;
    bf16* out_sb;
    int64_t out_plane;
    int32_t frames, pos, tiles, tpc, tpf;      // tiles / tiles per CTA / tiles per frame of the kernel at hand
    float eps;
};

__device__ __forceinline__ uint64_t* bar_at(uint8_t* bars, int i) { return reinterpret_cast<uint64_t*>(bars) + i; }

// LayerNorm of 16-lane row groups: v[p] <- (v[p] - mean) * rstd * gamma   (reference :170-180: biased variance, eps inside)
template <int NP>
__device__ __forceinline__ void ln_rows(float4 (&v)[NP], const float4 gam, float eps) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float s = (v[p].x + v[p].y) + (v[p].z + v[p].w);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.f / (float)FC);
        const float d0 = v[p].x - mean, d1 = v[p].y - mean, d2 = v[p].z - mean, d3 = v[p].w - mean;
        float sq = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * (1.f / (float)FC) + eps);
        v[p] = make_float4(d0 * rstd * gam.x, d1 * rstd * gam.y, d2 * rstd * gam.z, d3 * rstd * gam.w);
    }
}
// normalised rows (16 lanes per row, ROWS_PER_PASS rows per pass) -> split-bf16 SW128 K-major operand (hi at `tile`, lo at +plane)
template <int NP, int ROWS_PER_PASS>
__device__ __forceinline__ void ln_store(const float4 (&v)[NP], uint32_t tile, uint32_t plane, int rg, int l16) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int row = p * ROWS_PER_PASS + rg;
        uint2 hv, lv;
        split2(v[p].x, v[p].y, hv.x, lv.x);
        split2(v[p].z, v[p].w, hv.y, lv.y);
        const uint32_t off = tile + sw_off(row, l16 >> 1) + ((l16 & 1) << 3);
        sts64(off, hv.x, hv.y);
        sts64(off + plane, lv.x, lv.y);
    }
}

// ============================================================================================================================
// launch 1: context partials
// ============================================================================================================================
namespace l1 {
constexpr int OFF_WK = 0;            // 2 halves x (hi 16 KiB | lo 16 KiB)
constexpr int OFF_WV = 65536;
constexpr int OFF_XN = 131072;       // 2 buffers x (hi 8 KiB | lo 8 KiB): 64 rows x 128 B
constexpr int OFF_P = 163840;        // hi 16 KiB | lo 16 KiB: P^T, 128 rows (head, d) x 64 positions
constexpr int OFF_V = 196608;        // V^T, 128 rows (head, e) x 64 positions
constexpr int OFF_BAR = 229376;
constexpr int SMEM_BYTES = OFF_BAR + 1024 + 1024;
constexpr uint32_t T_KT = 0;         // 2 halves x 64 columns
constexpr uint32_t T_VT = 128;       // 2 halves x 64
constexpr uint32_t T_CTX = 256;      // 2 halves x 128
enum { B_W_FULL = 0, B_XN_FULL = 1 /* 2 */, B_XN_EMPTY = 3 /* 2 */, B_KT_FULL = 5 /* 2 */, B_KT_EMPTY = 7 /* 2 */, B_VT_FULL = 9 /* 2 */,
       B_VT_EMPTY = 11 /* 2 */, B_P_READY = 13, B_V_READY = 15, B_V_FREE = 16, B_CTX_DONE = 17 /* 2 */, B_CTX_EMPTY = 19 /* 2 */,
       B_P_FREE = 21 /* 2: [h] = the operand buffer is free for the group of half h (a parity wait must see EVERY phase of its barrier) */ };
}  // namespace l1

__global__ void __launch_bounds__(NTHREADS1, 1) linattn_ctx_kernel(const __grid_constant__ LinArgs a) {
    using namespace l1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bars = smem + OFF_BAR;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 512);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t sb = ptx::smem_u32(smem);

    pdl_trigger();
    if (warp == 1 && ptx::elect_one()) {
        ptx::mbar_init(bar_at(bars, B_W_FULL), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_at(bars, B_XN_FULL + i), 4);
            ptx::mbar_init(bar_at(bars, B_XN_EMPTY + i), 1);
            ptx::mbar_init(bar_at(bars, B_KT_FULL + i), 1);
            ptx::mbar_init(bar_at(bars, B_KT_EMPTY + i), 4);
            ptx::mbar_init(bar_at(bars, B_VT_FULL + i), 1);
            ptx::mbar_init(bar_at(bars, B_VT_EMPTY + i), 4);
            ptx::mbar_init(bar_at(bars, B_CTX_DONE + i), 1);
            ptx::mbar_init(bar_at(bars, B_CTX_EMPTY + i), 4);
        }
        ptx::mbar_init(bar_at(bars, B_P_READY), 4);
        ptx::mbar_init(bar_at(bars, B_P_FREE), 1);
        ptx::mbar_init(bar_at(bars, B_P_FREE + 1), 1);
        ptx::mbar_init(bar_at(bars, B_V_READY), 4);
        ptx::mbar_init(bar_at(bars, B_V_FREE), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // contiguous tile range of this CTA (64-row tiles); a frame segment = the tiles of one frame inside the range
    const int t_begin = (int)blockIdx.x * a.tpc;
    const int t_end = min(t_begin + a.tpc, a.tiles);
    const int my_tiles = max(t_end - t_begin, 0);
    const uint32_t NG = (uint32_t)(2 * my_tiles);                  // flat (tile, half) sequence
    auto seg_first = [&](int tl) { return tl == 0 || ((t_begin + tl) % a.tpf) == 0; };
    auto seg_last = [&](int tl) { return tl == my_tiles - 1 || ((t_begin + tl + 1) % a.tpf) == 0; };

    if (warp == 1) {
        // ===================== weights: resident for the whole kernel (constants: may be fetched before the PDL wait) ==========
        if (ptx::elect_one() && my_tiles > 0) {
            ptx::mbar_arrive_expect_tx(bar_at(bars, B_W_FULL), 131072);
            for (int i = 0; i < 4; ++i) {
                bulk_copy_g2s(sb + OFF_WK + i * 16384, a.wk + (size_t)i * 16384, 16384, bar_at(bars, B_W_FULL));
                bulk_copy_g2s(sb + OFF_WV + i * 16384, a.wv + (size_t)i * 16384, 16384, bar_at(bars, B_W_FULL));
            }
        }
    } else if (warp == 0) {
        // ===================== MMA issuer: polls the projection stream (kv) and the context stream (ctx) =====================
        if (ptx::elect_one() && my_tiles > 0) {
            constexpr uint32_t ID64 = ptx::make_idesc_bf16(128, 64);
            constexpr uint32_t ID128 = ptx::make_idesc_bf16(128, 128);
            const uint64_t d_wk = ptx::make_sw128_kmajor_desc(sb + OFF_WK), d_wv = ptx::make_sw128_kmajor_desc(sb + OFF_WV);
            const uint64_t d_xn = ptx::make_sw128_kmajor_desc(sb + OFF_XN);
            const uint64_t d_p = ptx::make_sw128_kmajor_desc(sb + OFF_P), d_v = ptx::make_sw128_kmajor_desc(sb + OFF_V);
            ptx::mbar_wait(bar_at(bars, B_W_FULL), 0);
            uint32_t gk = 0, gc = 0;
            while (gc < NG) {
                bool issued = false;
                // ---- ctx[half] += P^T V^T^T   (K = 64 positions)
                {
                    const uint32_t hf = gc & 1u;
                    const int tl = (int)(gc >> 1);
                    const bool first = seg_first(tl);
                    // a new segment overwrites the accumulator: the previous segment's rows must have been flushed
                    bool ok = mbar_test(bar_at(bars, B_P_READY), gc & 1) && mbar_test(bar_at(bars, B_V_READY), gc & 1);
                    if (ok && first && tl > 0) {
                        // number of flushes of this half so far = segments completed before tile tl
                        int nseg = 0;
                        for (int q = 1; q <= tl; ++q) nseg += seg_first(q) ? 1 : 0;
                        ok = mbar_test(bar_at(bars, B_CTX_EMPTY + hf), (uint32_t)(nseg - 1) & 1u);
                    }
                    if (ok) {
                        ptx::tc_fence_after();
                        const uint32_t td = tmem_base + T_CTX + 128u * hf;
                        const uint64_t p_hi = d_p, p_lo = d_p + (uint64_t)(16384 >> 4), v_hi = d_v, v_lo = d_v + (uint64_t)(16384 >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t o = (uint64_t)(ks * 2);
                            ptx::umma_bf16(td, p_lo + o, v_hi + o, ID128, (ks > 0 || !first) ? 1u : 0u);
                            ptx::umma_bf16(td, p_hi + o, v_lo + o, ID128, 1u);
                            ptx::umma_bf16(td, p_hi + o, v_hi + o, ID128, 1u);
                        }
                        ptx::umma_commit(bar_at(bars, B_P_FREE + (hf ^ 1u)));       // the next writer of P^T is the other half's group
                        ptx::umma_commit(bar_at(bars, B_V_FREE));
                        ptx::umma_commit(bar_at(bars, B_CTX_DONE + hf));
                        ++gc;
                        issued = true;
                    }
                }
                // ---- K^T[half] = W_k[half] Xn^T,  V^T[half] = W_v[half] Xn^T   (N = 64 positions)
                if (gk < NG) {
                    const uint32_t hf = gk & 1u;
                    const uint32_t t = gk >> 1, xb = t & 1u;
                    if ((hf != 0 || mbar_test(bar_at(bars, B_XN_FULL + xb), (t >> 1) & 1)) &&
                        mbar_test(bar_at(bars, B_KT_EMPTY + hf), (t & 1u) ^ 1u) && mbar_test(bar_at(bars, B_VT_EMPTY + hf), (t & 1u) ^ 1u)) {
                        ptx::tc_fence_after();
                        const uint64_t x_hi = d_xn + (uint64_t)((xb * 16384) >> 4), x_lo = x_hi + (uint64_t)(8192 >> 4);
                        for (int kv = 0; kv < 2; ++kv) {
                            const uint64_t w_hi = (kv ? d_wv : d_wk) + (uint64_t)((hf * 32768) >> 4), w_lo = w_hi + (uint64_t)(16384 >> 4);
                            const uint32_t td = tmem_base + (kv ? T_VT : T_KT) + 64u * hf;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t o = (uint64_t)(ks * 2);
                                ptx::umma_bf16(td, w_lo + o, x_hi + o, ID64, ks > 0 ? 1u : 0u);
                                ptx::umma_bf16(td, w_hi + o, x_lo + o, ID64, 1u);
                                ptx::umma_bf16(td, w_hi + o, x_hi + o, ID64, 1u);
                            }
                            ptx::umma_commit(bar_at(bars, (kv ? B_VT_FULL : B_KT_FULL) + hf));
                        }
                        if (hf == 1) ptx::umma_commit(bar_at(bars, B_XN_EMPTY + xb));
                        ++gk;
                        issued = true;
                    }
                }
                if (!issued) __nanosleep(40);
            }
        }
    } else if (warp < 6) {
        // ===================== LayerNorm producers (warps 2-5): 64 rows per tile, one tile ahead ===========================
        pdl_wait();
        const int t128 = (int)threadIdx.x - 64;
        const int l16 = t128 & 15, rg = t128 >> 4;      // 16 lanes per row, 8 rows per pass, 8 passes
        const float4 gam = *reinterpret_cast<const float4*>(a.gamma + l16 * 4);
        for (int tl = 0; tl < my_tiles; ++tl) {
            const int64_t row0 = (int64_t)(t_begin + tl) * 64;
            float4 v[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) v[p] = *reinterpret_cast<const float4*>(a.x + (row0 + p * 8 + rg) * FC + l16 * 4);
            ln_rows<8>(v, gam, a.eps);
            const uint32_t xb = (uint32_t)tl & 1u;
            ptx::mbar_wait(bar_at(bars, B_XN_EMPTY + xb), (((uint32_t)tl >> 1) & 1u) ^ 1u);
            ln_store<8, 8>(v, sb + OFF_XN + xb * 16384, 8192, rg, l16);
            ptx::fence_proxy_async();
            warp_arrive(bar_at(bars, B_XN_FULL + xb), lane);
            // the same operand image goes to global memory: launch 3 bulk-copies it instead of normalising the rows again
            {
                const int tg = t_begin + tl;
                uint8_t* img = a.ximg + (size_t)(tg >> 1) * 32768 + (size_t)(tg & 1) * 8192;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int row = p * 8 + rg;
                    uint2 hv, lv;
                    split2(v[p].x, v[p].y, hv.x, lv.x);
                    split2(v[p].z, v[p].w, hv.y, lv.y);
                    const uint32_t off = sw_off(row, l16 >> 1) + ((l16 & 1) << 3);
                    *reinterpret_cast<uint2*>(img + off) = hv;
                    *reinterpret_cast<uint2*>(img + 16384 + off) = lv;
                }
            }
        }
    } else {
        // ===================== WG-K0 / WG-K1 (warps 6-9 / 10-13): rows (head, d) of K^T of half 0 / 1; WG-V (warps 14-17) ======
        const int tc = (int)threadIdx.x - 192;
        const int wg = tc >> 7;
        const int q = warp & 3;
        const int r = q * 32 + lane;                     // operand row = TMEM lane: (head % 4) * 32 + d
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        if (wg < 2) {
            // the exponentials are the slowest stage: the two halves (4 heads each) run on their own warp-groups, each keeping the
            // running maximum / denominator of its rows in registers; they alternate on the single P^T operand buffer
            const uint32_t hf = (uint32_t)wg;
            float mx = -INFINITY, zs = 0.f;               // running maximum (log2 domain) / denominator of this row
            uint32_t nctx = 0u;                           // ctx MMAs of this half requested so far (CTX_DONE phases)
            const uint32_t t_kt = lane_base + T_KT + 64u * hf;
            const uint32_t tcx = lane_base + T_CTX + 128u * hf + 32u * (uint32_t)q;      // this row's own 32 x 32 context block
            for (uint32_t g = hf; g < NG; g += 2) {
                const int tl = (int)(g >> 1);
                const bool first = seg_first(tl), last = seg_last(tl);
                ptx::mbar_wait(bar_at(bars, B_KT_FULL + hf), (uint32_t)tl & 1u);
                ptx::tc_fence_after();
                // ---- pass A: maximum of this row over the 64 positions of the tile (W_k carries log2(e): log2 domain)
                float tm = -INFINITY;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t u[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld8(t_kt + 32u * (uint32_t)hh + 8u * (uint32_t)c, u[c]);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 8; ++j) tm = fmaxf(tm, __uint_as_float(u[c][j]));
                }
                if (first) { mx = -INFINITY; zs = 0.f; }
                // re-base this row when its maximum grew by more than the threshold (always on the first tile of a segment, where
                // nothing has been accumulated yet); the TMEM loads / stores are warp-collective: any lane -> whole warp
                const bool grow = tm > mx + RESCALE_THRESHOLD;
                if (first) {
                    mx = tm;
                } else if (__any_sync(0xffffffffu, grow)) {
                    const float nm = grow ? tm : mx;
                    const float sc = ex2(mx - nm);                           // 1 for the lanes that keep their maximum
                    ptx::mbar_wait(bar_at(bars, B_CTX_DONE + hf), (nctx - 1u) & 1u);         // every context MMA of this half has landed
                    ptx::tc_fence_after();
                    uint32_t cu[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld8(tcx + 8u * (uint32_t)c, cu[c]);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) cu[c][j] = __float_as_uint(__uint_as_float(cu[c][j]) * sc);
                        tmem_st8(tcx + 8u * (uint32_t)c, cu[c]);
                    }
                    tmem_st_wait();
                    ptx::tc_fence_before();
                    zs *= sc;
                    mx = nm;
                }
                // ctx(g-1) (the other half's) has read the operand: P_FREE[hf] completes once per step of the other group
                if (g > 0) ptx::mbar_wait(bar_at(bars, B_P_FREE + hf), (hf ? (uint32_t)tl : (uint32_t)tl - 1u) & 1u);
                // ---- pass B: exponentials, 32 positions at a time -> 4 operand chunks of the hi and of the lo plane
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t u[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld8(t_kt + 32u * (uint32_t)hh + 8u * (uint32_t)c, u[c]);
                    ptx::tmem_ld_wait();
                    if (hh == 1) {
                        ptx::tc_fence_before();
                        warp_arrive(bar_at(bars, B_KT_EMPTY + hf), lane);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float e[8];
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            e[j] = ex2(__uint_as_float(u[c][j]) - mx); e[j + 1] = ex2(__uint_as_float(u[c][j + 1]) - mx);
                            s0 += e[j]; s1 += e[j + 1];
                        }
                        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                        split2(e[0], e[1], h0, l0);
                        split2(e[2], e[3], h1, l1);
                        split2(e[4], e[5], h2, l2);
                        split2(e[6], e[7], h3, l3);
                        const uint32_t off = sb + OFF_P + sw_off(r, 4 * hh + c);
                        sts128(off, h0, h1, h2, h3);
                        sts128(off + 16384, l0, l1, l2, l3);
                    }
                }
                zs += s0 + s1;
                ptx::fence_proxy_async();
                warp_arrive(bar_at(bars, B_P_READY), lane);
                ++nctx;
                if (last) {
                    // flush this row of the segment: 32 context values (un-normalised, relative to the row maximum), maximum, denominator
                    ptx::mbar_wait(bar_at(bars, B_CTX_DONE + hf), (nctx - 1u) & 1u);
                    ptx::tc_fence_after();
                    uint32_t cu[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld8(tcx + 8u * (uint32_t)c, cu[c]);
                    ptx::tmem_ld_wait();
                    ptx::tc_fence_before();
                    warp_arrive(bar_at(bars, B_CTX_EMPTY + hf), lane);
                    const int tg = t_begin + tl;
                    const int frame = tg / a.tpf;
                    const int part = (int)blockIdx.x - (frame * a.tpf) / a.tpc;
                    float* dst = a.part + (((int64_t)frame * MAXP + part) * HID + (int64_t)hf * 128 + r) * PART_LD;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 8; j += 2)
                            *reinterpret_cast<float2*>(dst + 8 * c + j) = make_float2(__uint_as_float(cu[c][j]), __uint_as_float(cu[c][j + 1]));
                    *reinterpret_cast<float2*>(dst + 32) = make_float2(mx, zs);
                }
            }
        } else {
            for (uint32_t g = 0; g < NG; ++g) {
                const uint32_t hf = g & 1u;
                const uint32_t tl = g >> 1;
                ptx::mbar_wait(bar_at(bars, B_VT_FULL + hf), tl & 1u);
                ptx::tc_fence_after();
                if (g > 0) ptx::mbar_wait(bar_at(bars, B_V_FREE), (g - 1) & 1);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t u[4][8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld8(lane_base + T_VT + 64u * hf + 32u * (uint32_t)hh + 8u * (uint32_t)c, u[c]);
                    ptx::tmem_ld_wait();
                    if (hh == 1) {
                        ptx::tc_fence_before();
                        warp_arrive(bar_at(bars, B_VT_EMPTY + hf), lane);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                        split2(__uint_as_float(u[c][0]), __uint_as_float(u[c][1]), h0, l0);
                        split2(__uint_as_float(u[c][2]), __uint_as_float(u[c][3]), h1, l1);
                        split2(__uint_as_float(u[c][4]), __uint_as_float(u[c][5]), h2, l2);
                        split2(__uint_as_float(u[c][6]), __uint_as_float(u[c][7]), h3, l3);
                        const uint32_t off = sb + OFF_V + sw_off(r, 4 * hh + c);
                        sts128(off, h0, h1, h2, h3);
                        sts128(off + 16384, l0, l1, l2, l3);
                    }
                }
                ptx::fence_proxy_async();
                warp_arrive(bar_at(bars, B_V_READY), lane);
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

// ============================================================================================================================
// launch 2: merge the partials of a frame, G = (scale * ctx / Z) W_out^T, as the B-operand image of launch 3
// ============================================================================================================================
// grid (4 chunks, frames): block (j, frame) merges the 64 rows (2 heads) of chunk j and writes its 16 KiB operand image
__global__ void __launch_bounds__(256) linattn_combine_kernel(const __grid_constant__ LinArgs a, int tpc1, int tpf1) {
    __shared__ float ctx[64][33];                   // rows (hh, d) of the chunk's 2 heads x 32 e, scaled
    __shared__ float gsm[64][65];                   // G^T[c][hd] of the chunk
    pdl_trigger();
    pdl_wait();
    const int j = blockIdx.x, frame = blockIdx.y;
    const int t = threadIdx.x;
    const int c0 = (frame * tpf1) / tpc1, c1 = ((frame + 1) * tpf1 - 1) / tpc1;
    const int nparts = c1 - c0 + 1;
    {
        // thread (row = t / 4, e-octet = t % 4)
        const int row = t >> 2, e0 = (t & 3) * 8;
        const float* base = a.part + ((int64_t)frame * MAXP * HID + j * 64 + row) * PART_LD;
        float m = -INFINITY;
        for (int p = 0; p < nparts; ++p) m = fmaxf(m, base[(int64_t)p * HID * PART_LD + 32]);
        float z = 0.f, acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int p = 0; p < nparts; ++p) {
            const float* rowp = base + (int64_t)p * HID * PART_LD;
            const float sc = exp2f(rowp[32] - m);
            z += sc * rowp[33];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const float2 v2 = *reinterpret_cast<const float2*>(rowp + e0 + e);
                acc[e] += sc * v2.x; acc[e + 1] += sc * v2.y;
            }
        }
        const float f = 0.17677669529663687f / z;       // q * 32^-0.5 (reference :258) folded here
#pragma unroll
        for (int e = 0; e < 8; ++e) ctx[row][e0 + e] = acc[e] * f;
    }
    __syncthreads();
    {
        // thread (c = t % 64, d-octet = t / 64): G[hd = hh * 32 + d][c] = sum_e ctx[hd][e] W_out[c][(2 j + hh) * 32 + e]
        const int c = t & 63, dq = t >> 6;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float w[32];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(a.wout + (int64_t)c * HID + (2 * j + hh) * 32 + e);
                w[e] = w4.x; w[e + 1] = w4.y; w[e + 2] = w4.z; w[e + 3] = w4.w;
            }
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) {
                const int hd = hh * 32 + dq * 8 + dd;
                const float* cr = ctx[hd];
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int e = 0; e < 32; e += 2) { s0 = fmaf(cr[e], w[e], s0); s1 = fmaf(cr[e + 1], w[e + 1], s1); }
                gsm[c][hd] = s0 + s1;
            }
        }
    }
    __syncthreads();
    {
        // thread (row c = t / 4, 16-element quarter = t % 4): two 16-byte chunks of the hi and of the lo plane
        const int c = t >> 2, qd = t & 3;
        uint8_t* dst = a.gimg + ((int64_t)frame * 4 + j) * 16384;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int ch = qd * 2 + cc;
            const float* gv = &gsm[c][8 * ch];
            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
            split2(gv[0], gv[1], h0, l0);
            split2(gv[2], gv[3], h1, l1);
            split2(gv[4], gv[5], h2, l2);
            split2(gv[6], gv[7], h3, l3);
            const uint32_t off = sw_off(c, ch);
            *reinterpret_cast<uint4*>(dst + off) = make_uint4(h0, h1, h2, h3);
            *reinterpret_cast<uint4*>(dst + 8192 + off) = make_uint4(l0, l1, l2, l3);
        }
    }
}

// ============================================================================================================================
// launch 3: out = x + bias + softmax_d(q) G
// ============================================================================================================================
namespace l3 {
constexpr int OFF_WQ = 0;            // 4 chunks x (hi 8 KiB | lo 8 KiB): 64 rows (2 heads x 32 d) x 128 B
constexpr int OFF_XN = 65536;        // 2 buffers x (hi 16 KiB | lo 16 KiB): 128 rows
constexpr int OFF_QS = 131072;       // 2 buffers x (hi 16 KiB | lo 16 KiB): 128 rows x 64 (2 heads x 32 d)
constexpr int OFF_G = 196608;        // 2 stages x (hi 8 KiB | lo 8 KiB): G^T chunk, 64 rows (c) x 64 (hd)
constexpr int OFF_BAR = 229376;
constexpr int SMEM_BYTES = OFF_BAR + 1024 + 1024;
constexpr uint32_t T_Q = 0;          // 4 x 64 columns (chunk j -> buffer j)
constexpr uint32_t T_OUT = 256;      // 2 x 64
enum { B_W_FULL = 0, B_XN_FULL = 1 /* 2 */, B_XN_EMPTY = 3 /* 2 */, B_Q_FULL = 5 /* 4 */, B_Q_EMPTY = 9 /* 4 */, B_QS_READY = 13 /* 2 */,
       B_QS_FREE = 15 /* 2 */, B_G_FULL = 17 /* 2 */, B_G_EMPTY = 19 /* 2 */, B_OUT_FULL = 21 /* 2 */, B_OUT_EMPTY = 23 /* 2 */ };
}  // namespace l3

__global__ void __launch_bounds__(NTHREADS, 1) linattn_apply_kernel(const __grid_constant__ LinArgs a) {
    using namespace l3;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bars = smem + OFF_BAR;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 512);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t sb = ptx::smem_u32(smem);

    pdl_trigger();
    if (warp == 1 && ptx::elect_one()) {
        ptx::mbar_init(bar_at(bars, B_W_FULL), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar_at(bars, B_XN_FULL + i), 1);
            ptx::mbar_init(bar_at(bars, B_XN_EMPTY + i), 1);
            ptx::mbar_init(bar_at(bars, B_QS_READY + i), 4);
            ptx::mbar_init(bar_at(bars, B_QS_FREE + i), 1);
            ptx::mbar_init(bar_at(bars, B_G_FULL + i), 1);
            ptx::mbar_init(bar_at(bars, B_G_EMPTY + i), 1);
            ptx::mbar_init(bar_at(bars, B_OUT_FULL + i), 1);
            ptx::mbar_init(bar_at(bars, B_OUT_EMPTY + i), 4);
        }
        for (int i = 0; i < 4; ++i) {
            ptx::mbar_init(bar_at(bars, B_Q_FULL + i), 1);
            ptx::mbar_init(bar_at(bars, B_Q_EMPTY + i), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int my_tiles = (a.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // 128-row tiles, round-robin
    const uint32_t NG = (uint32_t)(4 * my_tiles);                                                  // flat (tile, chunk) sequence
    auto tile_of = [&](uint32_t tl) { return (int)blockIdx.x + (int)tl * (int)gridDim.x; };

    if (warp == 1) {
        // ===================== producer: W_q once, then the G^T chunk of every (tile, chunk) through a 2-stage ring ==========
        if (ptx::elect_one() && my_tiles > 0) {
            ptx::mbar_arrive_expect_tx(bar_at(bars, B_W_FULL), 65536);
            for (int i = 0; i < 4; ++i) bulk_copy_g2s(sb + OFF_WQ + i * 16384, a.wq + (size_t)i * 16384, 16384, bar_at(bars, B_W_FULL));
            pdl_wait();                                   // G comes from the combine kernel, the Xn images from launch 1
            auto load_xn = [&](uint32_t tl) {            // normalised rows of tile tl: one 32 KiB image (hi | lo plane)
                const uint32_t xb = tl & 1u;
                ptx::mbar_wait(bar_at(bars, B_XN_EMPTY + xb), ((tl >> 1) & 1u) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar_at(bars, B_XN_FULL + xb), 32768);
                bulk_copy_g2s(sb + OFF_XN + xb * 32768, a.ximg + (size_t)tile_of(tl) * 32768, 32768, bar_at(bars, B_XN_FULL + xb));
            };
            load_xn(0);
            for (uint32_t g = 0; g < NG; ++g) {
                const uint32_t s = g & 1u;
                if ((g & 3u) == 0 && (g >> 2) + 1 < (uint32_t)my_tiles) load_xn((g >> 2) + 1);
                const int frame = tile_of(g >> 2) / a.tpf;
                ptx::mbar_wait(bar_at(bars, B_G_EMPTY + s), ((g >> 1) & 1u) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar_at(bars, B_G_FULL + s), 16384);
                bulk_copy_g2s(sb + OFF_G + s * 16384, a.gimg + ((size_t)frame * 4 + (g & 3u)) * 16384, 16384, bar_at(bars, B_G_FULL + s));
            }
        }
    } else if (warp == 0) {
        // ===================== MMA issuer: polls the Q stream and the OUT stream =====================
        if (ptx::elect_one() && my_tiles > 0) {
            constexpr uint32_t ID64 = ptx::make_idesc_bf16(128, 64);
            const uint64_t d_wq = ptx::make_sw128_kmajor_desc(sb + OFF_WQ), d_xn = ptx::make_sw128_kmajor_desc(sb + OFF_XN);
            const uint64_t d_qs = ptx::make_sw128_kmajor_desc(sb + OFF_QS), d_g = ptx::make_sw128_kmajor_desc(sb + OFF_G);
            ptx::mbar_wait(bar_at(bars, B_W_FULL), 0);
            uint32_t gq = 0, go = 0;
            while (go < NG) {
                bool issued = false;
                // ---- OUT[tile] += Qs_chunk G_chunk
                {
                    const uint32_t s = go & 1u, j = go & 3u, t = go >> 2, ob = t & 1u;
                    if (mbar_test(bar_at(bars, B_QS_READY + s), (go >> 1) & 1) && mbar_test(bar_at(bars, B_G_FULL + s), (go >> 1) & 1) &&
                        (j != 0 || mbar_test(bar_at(bars, B_OUT_EMPTY + ob), ((t >> 1) & 1u) ^ 1u))) {
                        ptx::tc_fence_after();
                        const uint32_t td = tmem_base + T_OUT + 64u * ob;
                        const uint64_t q_hi = d_qs + (uint64_t)((s * 32768) >> 4), q_lo = q_hi + (uint64_t)(16384 >> 4);
                        const uint64_t g_hi = d_g + (uint64_t)((s * 16384) >> 4), g_lo = g_hi + (uint64_t)(8192 >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t o = (uint64_t)(ks * 2);
                            ptx::umma_bf16(td, q_lo + o, g_hi + o, ID64, (ks > 0 || j > 0) ? 1u : 0u);
                            ptx::umma_bf16(td, q_hi + o, g_lo + o, ID64, 1u);
                            ptx::umma_bf16(td, q_hi + o, g_hi + o, ID64, 1u);
                        }
                        ptx::umma_commit(bar_at(bars, B_QS_FREE + s));
                        ptx::umma_commit(bar_at(bars, B_G_EMPTY + s));
                        if (j == 3) ptx::umma_commit(bar_at(bars, B_OUT_FULL + ob));
                        ++go;
                        issued = true;
                    }
                }
                // ---- Q chunk j = Xn W_q[j]^T
                if (gq < NG) {
                    const uint32_t j = gq & 3u, t = gq >> 2, xb = t & 1u;
                    if ((j != 0 || mbar_test(bar_at(bars, B_XN_FULL + xb), (t >> 1) & 1)) && mbar_test(bar_at(bars, B_Q_EMPTY + j), (t & 1u) ^ 1u)) {
                        ptx::tc_fence_after();
                        const uint32_t td = tmem_base + T_Q + 64u * j;
                        const uint64_t x_hi = d_xn + (uint64_t)((xb * 32768) >> 4), x_lo = x_hi + (uint64_t)(16384 >> 4);
                        const uint64_t w_hi = d_wq + (uint64_t)((j * 16384) >> 4), w_lo = w_hi + (uint64_t)(8192 >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t o = (uint64_t)(ks * 2);
                            ptx::umma_bf16(td, x_lo + o, w_hi + o, ID64, ks > 0 ? 1u : 0u);
                            ptx::umma_bf16(td, x_hi + o, w_lo + o, ID64, 1u);
                            ptx::umma_bf16(td, x_hi + o, w_hi + o, ID64, 1u);
                        }
                        ptx::umma_commit(bar_at(bars, B_Q_FULL + j));
                        if (j == 3) ptx::umma_commit(bar_at(bars, B_XN_EMPTY + xb));
                        ++gq;
                        issued = true;
                    }
                }
                if (!issued) __nanosleep(40);
            }
        }
    } else if (warp < 6) {
        // ===================== tile epilogue (warps 2-5) =====================
        pdl_wait();
        const int eq = warp & 3, er = eq * 32 + lane;   // TMEM lane quarter / tile row of this thread
        const int ec = lane & 3;
        auto epilogue = [&](uint32_t tl) {
            const uint32_t ob = tl & 1u;
            ptx::mbar_wait(bar_at(bars, B_OUT_FULL + ob), (tl >> 1) & 1u);
            ptx::tc_fence_after();
            uint32_t u[4][16];
#pragma unroll
            for (int part = 0; part < 4; ++part) tmem_ld16(tmem_base + ((uint32_t)(eq * 32) << 16) + T_OUT + 64u * ob + 16u * (uint32_t)part, u[part]);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            warp_arrive(bar_at(bars, B_OUT_EMPTY + ob), lane);
            const int64_t row0 = (int64_t)tile_of(tl) * 128 + (er & ~3);        // first row of this lane's quad
            const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[part][i]);
                // 4x4 transpose of 16-byte chunks inside the lane quad: slot s <- (row s of the quad, chunk lane & 3)
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
                float4 ob4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.out_bias) ob4 = __ldg(reinterpret_cast<const float4*>(a.out_bias + col));
                float4 xres[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) xres[jj] = *reinterpret_cast<const float4*>(a.x + (row0 + jj) * FC + col);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 t4 = make_float4(v[4 * jj] + xres[jj].x + ob4.x, v[4 * jj + 1] + xres[jj].y + ob4.y,
                                                  v[4 * jj + 2] + xres[jj].z + ob4.z, v[4 * jj + 3] + xres[jj].w + ob4.w);
                    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + (row0 + jj) * FC + col) = t4;
                    if (a.out_sb) store_sb4(a.out_sb, a.out_plane, (row0 + jj) * FC + col, t4);
                }
            }
        };
        for (uint32_t tl = 0; tl < (uint32_t)my_tiles; ++tl) epilogue(tl);
    } else {
        // ===================== two softmax warp-groups: WG w handles the chunks with (chunk & 1) == w =====================
        const int tc = (int)threadIdx.x - 192;
        const uint32_t w = (uint32_t)(tc >> 7);
        const int q = warp & 3;
        const int r = q * 32 + lane;                     // tile row
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t use = 0;                                // uses of this group's Qs buffer so far
        for (uint32_t g = w; g < NG; g += 2, ++use) {
            const uint32_t j = g & 3u, t = g >> 2;
            ptx::mbar_wait(bar_at(bars, B_Q_FULL + j), t & 1u);
            ptx::tc_fence_after();
            float qv[64];
            {
                uint32_t u[8][8];
#pragma unroll
                for (int c = 0; c < 8; ++c) tmem_ld8(lane_base + T_Q + 64u * j + 8u * (uint32_t)c, u[c]);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int i = 0; i < 8; ++i) qv[8 * c + i] = __uint_as_float(u[c][i]);       // log2 domain: W_q carries log2(e) (host packer)
            }
            ptx::tc_fence_before();
            warp_arrive(bar_at(bars, B_Q_EMPTY + j), lane);
            // softmax over the 32 d of each of the 2 heads of the chunk (reference :256; the scale lives in G)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float m = qv[32 * hh];
#pragma unroll
                for (int i = 1; i < 32; ++i) m = fmaxf(m, qv[32 * hh + i]);
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    qv[32 * hh + i] = ex2(qv[32 * hh + i] - m); qv[32 * hh + i + 1] = ex2(qv[32 * hh + i + 1] - m);
                    s0 += qv[32 * hh + i]; s1 += qv[32 * hh + i + 1];
                }
                const float inv = 1.f / (s0 + s1);
#pragma unroll
                for (int i = 0; i < 32; ++i) qv[32 * hh + i] *= inv;
            }
            if (use > 0) ptx::mbar_wait(bar_at(bars, B_QS_FREE + w), (use - 1) & 1u);     // OUT MMA of the previous use has read the buffer
            store_row64_hilo(sb + OFF_QS + w * 32768, 16384, r, qv);
            ptx::fence_proxy_async();
            warp_arrive(bar_at(bars, B_QS_READY + w), lane);
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

int lfdm_attn_linear_fused(const float* x, const float* gamma, const void* wk_packed, const void* wv_packed, const void* wq_packed,
                           const float* wout, const float* out_bias, float* partials, void* g_images, void* xn_images,
                           float* out_f32, void* out_sb, int64_t out_plane, int frames, int pos, int c, int heads, float eps,
                           void* stream) {
    if (!x || !gamma || !wk_packed || !wv_packed || !wq_packed || !wout || !partials || !g_images || !xn_images || (!out_f32 && !out_sb))
        return LFDM_E_BADARG;
    if (c != FC || heads != HEADS || frames < 1 || pos < 128 || (pos % 128) != 0) return LFDM_E_UNSUPP;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wk_packed) | reinterpret_cast<uintptr_t>(wv_packed) |
         reinterpret_cast<uintptr_t>(wq_packed) | reinterpret_cast<uintptr_t>(wout) | reinterpret_cast<uintptr_t>(gamma) |
         reinterpret_cast<uintptr_t>(out_bias) | reinterpret_cast<uintptr_t>(out_f32) | reinterpret_cast<uintptr_t>(out_sb) |
         reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(g_images) | reinterpret_cast<uintptr_t>(xn_images)) & 15)
        return LFDM_E_UNSUPP;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static PerDeviceOnce once;
    static int sms = 0;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(linattn_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, l1::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(linattn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, l3::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        once.mark();
    }
    LinArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.gamma = gamma;
    a.wk = reinterpret_cast<const uint8_t*>(wk_packed); a.wv = reinterpret_cast<const uint8_t*>(wv_packed);
    a.wq = reinterpret_cast<const uint8_t*>(wq_packed); a.wout = wout; a.out_bias = out_bias;
    a.part = partials; a.gimg = reinterpret_cast<uint8_t*>(g_images); a.ximg = reinterpret_cast<uint8_t*>(xn_images);
    a.out_f32 = out_f32; a.out_sb = reinterpret_cast<bf16*>(out_sb); a.out_plane = out_plane;
    a.frames = frames; a.pos = pos; a.eps = eps;
    // ---- launch 1: 64-row tiles, contiguous ranges
    const int tiles1 = frames * (pos / 64), tpf1 = pos / 64;
    int grid1 = tiles1 < sms ? tiles1 : sms;
    int tpc1 = (tiles1 + grid1 - 1) / grid1;
    // a frame may be cut into at most MAXP segments: widen the per-CTA range if the frame is long compared with it
    while ((tpf1 + tpc1 - 1) / tpc1 + 1 > MAXP) ++tpc1;
    grid1 = (tiles1 + tpc1 - 1) / tpc1;
    a.tiles = tiles1; a.tpc = tpc1; a.tpf = tpf1;
    LFDM_LAUNCH_PDL(linattn_ctx_kernel, dim3(grid1), dim3(NTHREADS1), (size_t)l1::SMEM_BYTES, st, a);
    // ---- launch 2
    LFDM_LAUNCH_PDL(linattn_combine_kernel, dim3(4, frames), dim3(256), (size_t)0, st, a, tpc1, tpf1);
    // ---- launch 3: 128-row tiles, round-robin
    const int tiles3 = frames * (pos / 128);
    a.tiles = tiles3; a.tpc = 0; a.tpf = pos / 128;
    const int grid3 = tiles3 < sms ? tiles3 : sms;
    LFDM_LAUNCH_PDL(linattn_apply_kernel, dim3(grid3), dim3(NTHREADS), (size_t)l3::SMEM_BYTES, st, a);
    return 0;
}
