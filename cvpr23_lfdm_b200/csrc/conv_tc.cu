// conv_tc.cu — LFDM_ENGINE_TC: persistent, warp-specialised tcgen05 implicit-GEMM convolution for sm_100a.
//
//   D[128 x BN] (fp32, TMEM accumulator ring) += A[128 x 64] * B[BN x 64]^T     per (filter tap, 64-channel chunk)
//
// * A tiles are gathered straight from the channels-last split-bf16 activation tensor by TMA: one 5-D box
//   {64 ch, BW, BH, BNF, plane} per tap, shifted by the tap offset; out-of-bounds coordinates are zero-filled by
//   the TMA unit, which *is* the zero padding of the convolution.  A "virtual concat" (torch.cat of the UNet skip
//   connections) is two tensor maps walked back to back.  Stride-2 convs read four parity views of the input;
//   transposed / up-sampling convs run as four 2x2 phase launches.
// * 3x3 "halo mode" (tile inside one image, BW <= 32): one (BH+2)-row halo copy per horizontal tap offset; the three
//   vertical taps are descriptor start offsets into it.  Separate A (2 x 48 KiB) and B (single-tap) rings, one
//   producer warp each.
// * B tiles (weights, pre-split to bf16 hi/lo and packed [plane][tap][Cout][Cin]) come from a 4-D map; W_hi / W_lo
//   of a tap sit back to back so one N = 2*BN MMA covers both.
// * fp32-class accuracy on bf16 tensor cores: x*w ~ hi*hi + hi*lo + lo*hi, fp32 accumulate (2 MMAs per K-step in
//   the wide form, 3 in the classic form used for the short-K projections).
// * warp roles: warp0 = TMA producer (A in halo mode), warp1 = MMA issuer (one elected thread runs the whole loop),
//   warp2 = TMEM allocator, warp3 = B producer (halo mode), warps 4-11 = two epilogue groups draining alternate
//   tiles (tcgen05.ld -> bias / residual / activation / GroupNorm partial sums -> swizzled smem stage -> TMA store,
//   or coalesced global stores when a residual / split-bf16 output is involved).
//   smem ring: full/empty mbarriers; TMEM ring: 2-4 accumulator stages, tmem_full/tmem_empty mbarriers.
// * launched with the programmatic-dependent-launch attribute: everything above pdl_wait() overlaps the previous
//   kernel's tail.
//
// Replaces cuDNN/cuBLAS behind every Conv3d(1,k,k)/ConvTranspose3d/Conv2d/Linear on the path (include/lfdm_b200.h).
#include <cuda.h>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;               // bf16 elements per K block = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * BK * 2; // 16 KiB per plane
constexpr int MAX_TAPS = 52;
constexpr int NUM_THREADS = 384;     // warp0 TMA (A), warp1 MMA, warp2 TMEM alloc, warp3 TMA (B, halo mode), warps 4-11 epilogue

struct TcArgs {
    CUtensorMap tmA[8];              // [source*4 + parity view]
    CUtensorMap tmB;
    CUtensorMap tmBh;                // weight tile cut in two (box rows = BN / 2): pair mode, each CTA holds half of the rows of B
    int32_t pair;                    // 1: clusters of 2 CTAs = one cta_group::2 MMA pair: M = 256 (two M-adjacent tiles), B split in two
    CUtensorMap tmOut;               // fp32 [M][c_out], box {32 cols, 32 rows}, 128B swizzle (TMA-store epilogue)
    int32_t tma_store;
    int32_t dbg;                     // LFDM_CONV_DBG ablation bits (timing experiments only): 1 no A loads, 2 no B loads, 8 no TMA stores, 128 no epilogue body
    int32_t na2;                     // 1 (default): two halo stages + six weight stages at BN = 64; LFDM_CONV_NA3=1 -> three + three
    int32_t halo, halo_plane;        // 3x3 halo mode: tmA[src*4+1] = box {64, bw, bh+2}; bytes of one halo plane
    int32_t n_taps, tap_base;
    int8_t tap_map[MAX_TAPS], tap_dy[MAX_TAPS], tap_dx[MAX_TAPS];
    int32_t chunks[2];               // 64-channel chunks per source
    int32_t bw, bh, bnf;             // A box (bw*bh*bnf == 128)
    int32_t tiles_w, tiles_h, m_tiles, n_tiles;
    int32_t ho_full, wo_full, mul, off_h, off_w;   // output row = ((nf*ho_full + h*mul+off_h)*wo_full + w*mul+off_w)
    int32_t c_out;
    const float* bias;
    const float* residual;
    int32_t res_bcast_f;
    int64_t p_out;
    float* out_f32;
    int32_t f32_act;
    bf16* out_sb;
    int64_t out_plane;
    int32_t sb_act;
    const float* sb_scale;
    const float* sb_shift;
    double* gn_stats;
    int32_t gn_cpg, gn_groups;
    int64_t rows_per_sample;
    // stream-K (generic path): the flat (tile, K-block) space is cut into equal contiguous ranges, one per CTA; a tile cut in two
    // is finished by the CTA that owns its end, the other one hands over its partial accumulator through sk_ws / sk_flag
    int32_t sk;                      // K-blocks per CTA (0 = off: whole tiles, round-robin)
    float* sk_ws;                    // [grid][128][bn] fp32 partial accumulators
    int32_t* sk_flag;                // [grid] 0 / 1, left at 0 by the consumer
    const float* rot_cos;            // fused rotary of a temporal qkv projection (TMA-store epilogue only), see lfdm_conv_desc
    const float* rot_sin;
    int32_t rot_frames, rot_rows_per_frame, rot_cols, rot_scale_cols;
    float rot_scale;
};

// epilogue of 16 accumulator columns [nb, nb+16) of one output row
__device__ __forceinline__ void epi_chunk16(const TcArgs& a, const uint32_t (&rr)[16], int nb, int64_t orow, int64_t rrow,
                                            float* gn_acc, int lane, bool vec_ok) {
    if (nb >= a.c_out) return;       // warp-uniform (padded N)
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
        const bool full = vec_ok && (nb + 16 <= a.c_out);
        if (full) {
            if (a.bias) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 b4 = *reinterpret_cast<const float4*>(a.bias + nb + j);
                    v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
            }
            if (a.residual) {
                const float* rp = a.residual + rrow * a.c_out + nb;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                    v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
                }
            }
            if (a.gn_stats) {
                // 16 columns = two 8-column halves; cpg is a multiple of 8
                float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { s0 += v[j]; q0 = fmaf(v[j], v[j], q0); }
#pragma unroll
                for (int j = 8; j < 16; ++j) { s1 += v[j]; q1 = fmaf(v[j], v[j], q1); }
                s0 = warp_sum(s0); q0 = warp_sum(q0); s1 = warp_sum(s1); q1 = warp_sum(q1);
                if (lane == 0) {
                    const int g0 = nb / a.gn_cpg, g1 = (nb + 8) / a.gn_cpg;
                    gn_acc[g0 * 2] += s0; gn_acc[g0 * 2 + 1] += q0;
                    gn_acc[g1 * 2] += s1; gn_acc[g1 * 2 + 1] += q1;
                }
            }
            const int64_t o = orow * a.c_out + nb;
            if (a.out_f32) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 o4 = make_float4(apply_act(v[j], a.f32_act), apply_act(v[j + 1], a.f32_act),
                                            apply_act(v[j + 2], a.f32_act), apply_act(v[j + 3], a.f32_act));
                    *reinterpret_cast<float4*>(a.out_f32 + o + j) = o4;
                }
            }
            if (a.out_sb) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float u[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = v[j + e];
                        if (a.sb_scale) t *= a.sb_scale[nb + j + e];
                        if (a.sb_shift) t += a.sb_shift[nb + j + e];
                        u[e] = apply_act(t, a.sb_act);
                    }
                    store_sb4(a.out_sb, a.out_plane, o + j, make_float4(u[0], u[1], u[2], u[3]));
                }
            }
        } else {
            // ragged N (e.g. Cout = 3): scalar, masked
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = nb + j;
                if (n < a.c_out) {
                    float t = v[j];
                    if (a.bias) t += a.bias[n];
                    if (a.residual) t += a.residual[rrow * a.c_out + n];
                    const int64_t o = orow * a.c_out + n;
                    if (a.out_f32) a.out_f32[o] = apply_act(t, a.f32_act);
                    if (a.out_sb) {
                        float u = t;
                        if (a.sb_scale) u *= a.sb_scale[n];
                        if (a.sb_shift) u += a.sb_shift[n];
                        store_sb1(a.out_sb, a.out_plane, o, apply_act(u, a.sb_act));
                    }
                }
            }
        }
}

// Coalesced epilogue of a 32-row x 32-column accumulator block of one warp.  The thread-per-row TMEM layout is
// transposed through a 4 KiB XOR-swizzled smem stage so that 8 consecutive lanes cover one 128-byte row segment:
// every global store / residual load instruction then touches full 128 B lines (the thread-per-row form issues 32
// scattered 16 B accesses per instruction).  `orow_l[i]` / `rrow_l[i]` are the output / residual rows of tile row
// (lane >> 3) + 4 i of this warp's 32-row slab.
__device__ __forceinline__ void epi_block32(const TcArgs& a, float* stage, const uint32_t (&r0)[16], const uint32_t (&r1)[16],
                                            int nb, const int32_t (&orow_l)[8], const int32_t (&rrow_l)[8], float* gn_acc,
                                            int lane) {
    // ---- write: thread = row `lane`, logical 16-byte chunk q -> physical chunk q ^ (lane & 7)
    float4* st4 = reinterpret_cast<float4*>(stage) + lane * 8;
    const int sw = lane & 7;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        st4[q ^ sw] = make_float4(__uint_as_float(r0[4 * q]), __uint_as_float(r0[4 * q + 1]), __uint_as_float(r0[4 * q + 2]),
                                  __uint_as_float(r0[4 * q + 3]));
        st4[(q + 4) ^ sw] = make_float4(__uint_as_float(r1[4 * q]), __uint_as_float(r1[4 * q + 1]),
                                        __uint_as_float(r1[4 * q + 2]), __uint_as_float(r1[4 * q + 3]));
    }
    __syncwarp();
    // ---- read transposed: lane -> (row sub-index lane >> 3, column chunk lane & 7)
    const int cq = lane & 7, rsub = lane >> 3;
    const int n = nb + 4 * cq;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias4 = *reinterpret_cast<const float4*>(a.bias + n);
    float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.out_sb && a.sb_scale) sc4 = *reinterpret_cast<const float4*>(a.sb_scale + n);
    if (a.out_sb && a.sb_shift) sh4 = *reinterpret_cast<const float4*>(a.sb_shift + n);
    float gs = 0.f, gq = 0.f;
    // all eight residual loads are issued before the first store: the output may alias the residual (in-place add), so
    // the compiler cannot hoist them itself and the loop would otherwise pay eight serial global round trips
    float4 res4[8];
    if (a.residual) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            res4[i] = *reinterpret_cast<const float4*>(a.residual + (int64_t)rrow_l[i] * a.c_out + n);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = rsub + 4 * i;
        float4 v = reinterpret_cast<const float4*>(stage)[rr * 8 + (cq ^ (rr & 7))];
        v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
        if (a.residual) {
            const float4 r4 = res4[i];
            v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
        }
        if (a.gn_stats) {
            gs += (v.x + v.y) + (v.z + v.w);
            gq = fmaf(v.x, v.x, gq); gq = fmaf(v.y, v.y, gq); gq = fmaf(v.z, v.z, gq); gq = fmaf(v.w, v.w, gq);
        }
        const int64_t o = (int64_t)orow_l[i] * a.c_out + n;
        if (a.out_f32)
            *reinterpret_cast<float4*>(a.out_f32 + o) = make_float4(apply_act(v.x, a.f32_act), apply_act(v.y, a.f32_act),
                                                                   apply_act(v.z, a.f32_act), apply_act(v.w, a.f32_act));
        if (a.out_sb)
            store_sb4(a.out_sb, a.out_plane, o,
                      make_float4(apply_act(v.x * sc4.x + sh4.x, a.sb_act), apply_act(v.y * sc4.y + sh4.y, a.sb_act),
                                  apply_act(v.z * sc4.z + sh4.z, a.sb_act), apply_act(v.w * sc4.w + sh4.w, a.sb_act)));
    }
    if (a.gn_stats) {
        // rows: lanes differing in bits 3,4; columns of one group: 4*cq .. -> lanes differing in the low bits
        gs += __shfl_xor_sync(0xffffffffu, gs, 8);  gq += __shfl_xor_sync(0xffffffffu, gq, 8);
        gs += __shfl_xor_sync(0xffffffffu, gs, 16); gq += __shfl_xor_sync(0xffffffffu, gq, 16);
        gs += __shfl_xor_sync(0xffffffffu, gs, 1);  gq += __shfl_xor_sync(0xffffffffu, gq, 1);      // cpg % 8 == 0
        const int lanes_per_group = a.gn_cpg >= 32 ? 8 : a.gn_cpg / 4;                             // 2, 4 or 8
        if (lanes_per_group >= 4) { gs += __shfl_xor_sync(0xffffffffu, gs, 2); gq += __shfl_xor_sync(0xffffffffu, gq, 2); }
        if (lanes_per_group >= 8) { gs += __shfl_xor_sync(0xffffffffu, gs, 4); gq += __shfl_xor_sync(0xffffffffu, gq, 4); }
        if (rsub == 0 && (cq % lanes_per_group) == 0) {
            gn_acc[(n / a.gn_cpg) * 2] += gs;                    // this warp's own shared-memory slot: one lane per (group, k)
            gn_acc[(n / a.gn_cpg) * 2 + 1] += gq;
        }
    }
    __syncwarp();      // stage is reused by the next block
}

// ---- stream-K work list of one CTA.  Range [s, e) of the flat K-block space; natural segments are the pieces of consecutive
// tiles inside it.  The piece at the END of the range that stops short of its tile's end (a HEAD: the next CTA finishes that
// tile) is processed FIRST and the piece at the START that begins inside a tile (a TAIL: finishes the tile with the previous
// CTA's partial) LAST, with the whole tiles in between: the partial a TAIL needs was published after at most one tile's worth
// of K-blocks, while its consumer gets there after a whole range (>= one tile: host) -- nobody ever waits.  (HEAD first,
// TAIL second measured +25 us: a CTA with a short HEAD reaches its TAIL long before a neighbour with a long HEAD publishes.)
// Every range holds at least one whole tile's worth of K-blocks, so a tile is never cut into more than two pieces.
struct SkSeg { int tile, kb0, kb1; };
struct SkList {
    int s, e, n_kb, t0, nseg, has_head, has_tail, n_mid;
    __device__ SkList(int cta, int q, int n_kb_, int total_tiles) {
        n_kb = n_kb_;
        const long long tot = (long long)total_tiles * n_kb;
        long long s64 = (long long)cta * q, e64 = s64 + q;
        if (s64 > tot) s64 = tot;
        if (e64 > tot) e64 = tot;
        s = (int)s64; e = (int)e64;
        t0 = s / n_kb;
        nseg = e > s ? (e - 1) / n_kb - t0 + 1 : 0;
        has_head = (nseg > 1 && (e % n_kb) != 0) ? 1 : 0;
        has_tail = (nseg > 0 && (s % n_kb) != 0) ? 1 : 0;
        n_mid = nseg - has_head - has_tail;
    }
    __device__ SkSeg at(int i) const {           // i-th piece in processing order: [HEAD] whole tiles ... [TAIL]
        int j;
        if (has_head && i == 0) {
            j = nseg - 1;
        } else {
            const int k = i - has_head;
            j = k < n_mid ? k + has_tail : 0;
        }
        const int tile = t0 + j;
        const int lo = tile * n_kb, hi = lo + n_kb;
        SkSeg g;
        g.tile = tile;
        g.kb0 = (s > lo ? s : lo) - lo;
        g.kb1 = (e < hi ? e : hi) - lo;
        return g;
    }
};

// PAIR: the cta_group::2 build (must be launched as clusters of 2: a kernel that contains 2-CTA instructions is refused by a plain launch)
template <int BN, int STAGES, bool WIDE, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    // Each accumulator stage holds TWO column ranges: D1 = A_hi.W_hi + A_lo.W_hi and D2 = A_hi.W_lo.  W_hi / W_lo tiles are
    // adjacent in smem, so ONE N = 2*BN MMA with A_hi produces (A_hi.W_hi | A_hi.W_lo) while sweeping A_hi once; a second
    // N = BN MMA adds A_lo.W_hi.  2 MMAs / 14 KiB of operand reads per K-step instead of 3 MMAs / 18 KiB (BN = 64): the
    // kernel is bound by the 128 B/clk shared-memory port, not by the tensor pipe.  The epilogue sums D1 + D2.
    // WIDE = false (K = 1-2 blocks, epilogue-bound GEMMs such as the qkv projections): classic 3 MMAs into one D and a
    // 4-deep accumulator ring even at BN = 128.
    constexpr int DCOLS = WIDE ? 2 * BN : BN;             // TMEM columns per accumulator stage
    constexpr int ACC = (4 * DCOLS <= 512) ? 4 : 2;       // TMEM accumulator ring depth
    constexpr uint32_t TMEM_COLS = (ACC * DCOLS <= 32) ? 32 : (ACC * DCOLS <= 64) ? 64 : (ACC * DCOLS <= 128) ? 128 : (ACC * DCOLS <= 256) ? 256 : 512;
    static_assert(ACC * DCOLS <= 512, "accumulator ring exceeds TMEM");
    constexpr uint32_t IDESC_WIDE = ptx::make_idesc_bf16(BM, 2 * BN);
    constexpr uint32_t IDESC = ptx::make_idesc_bf16(BM, BN);
    // CTA pair (cta_group::2): M = 256 = this CTA's 128 rows + the peer's, every CTA holds HALF of the rows of each B operand
    // at the same shared-memory offset: rank 0 [W_hi[0:BN/2] ; W_lo[BN/2:BN]], rank 1 [W_hi[BN/2:BN] ; W_lo[0:BN/2]].
    //   wide MMA   (A_hi, N = 2 BN): accumulator column blocks (BN/2 wide)  [hi0 | lo1 | hi1 | lo0]
    //   narrow MMA (A_lo, N = BN)  : the first BN/2 rows of both CTAs = W_hi[0:BN/2], W_hi[BN/2:BN] -> added to blocks 0 and 1
    // so blocks 0 + 3 sum to output columns [0, BN/2) and blocks 1 + 2 to [BN/2, BN).  Per K-step and CTA the shared-memory port
    // carries A 8 KiB + B 3/4 BN x 32 B instead of A 8 KiB + B 3/2 BN x 32 B, and the weight fills are halved.
    constexpr uint32_t IDESC_WIDE2 = ptx::make_idesc_bf16(2 * BM, 2 * BN);
    constexpr uint32_t IDESC2 = ptx::make_idesc_bf16(2 * BM, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [operand ring][8 x 4 KiB epilogue staging (1024-aligned: TMA-store / swizzle atoms)][barriers][GN accumulators]
    // barrier block: operand ring (generic: STAGES stages; halo mode: the B ring, up to 6 stages), halo A ring (2), TMEM ring
    constexpr int MAXB = 6;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 8 * 4096);
    uint64_t* empty_bar = full_bar + MAXB;
    uint64_t* fullA_bar = empty_bar + MAXB;
    uint64_t* emptyA_bar = fullA_bar + 3;
    uint64_t* tfull_bar = emptyA_bar + 3;
    uint64_t* tempty_bar = tfull_bar + ACC;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + ACC);
    // 3x3 halo mode (see the host side): the operand ring is re-cut into 2 A stages of 48 KiB -- one (bh+2)-row halo copy
    // per horizontal tap offset dx, shared by the three vertical taps through the descriptor start address -- and a ring
    // of NB single-tap weight stages.  A bytes per tile: 3 (bh+2)/bh tiles instead of 9.
    constexpr int A_HALO_STAGE = 49152;
    // ring split: NA halo stages + NB weight stages out of the same bytes: 2 + 6 (BN = 64) / 2 + 3 (BN = 128).  LFDM_CONV_NA3=1 tries
    // 3 + 3 at BN = 64 (measured: no difference, 71.7 vs 73.6 us -- the 56 % tensor-pipe activity is not an A-ring depth problem).
    const int NA = (BN == 64 && !a.na2) ? 3 : 2;
    const int NB = (STAGES * STAGE_BYTES - NA * A_HALO_STAGE) / (2 * B_BYTES);
    constexpr bool HALO_OK = WIDE && BN >= 64;
    static_assert(!HALO_OK || ((STAGES * STAGE_BYTES - 2 * A_HALO_STAGE) / (2 * B_BYTES) <= MAXB && (STAGES * STAGE_BYTES - (BN == 64 ? 3 : 2) * A_HALO_STAGE) / (2 * B_BYTES) >= 3), "halo-mode rings");
    const bool halo = HALO_OK && a.halo;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    pdl_trigger();       // PDL: the next kernel's CTAs may become resident as SMs free up; everything up to pdl_wait() below
                         // (barrier init, TMEM allocation, tensor-map prefetch) overlaps the previous kernel's tail
    if (warp == 0 && ptx::elect_one()) {
        for (int i = 0; i < 8; ++i) ptx::prefetch_tensormap(&a.tmA[i]);
        ptx::prefetch_tensormap(&a.tmB);
        if (a.tma_store) ptx::prefetch_tensormap(&a.tmOut);
    }
    if (warp == 1 && ptx::elect_one()) {
        // pair mode: the "full" / "tmem empty" barriers of the EVEN CTA collect both CTAs (TMA bytes, epilogue threads); "empty" /
        // "tmem full" are signalled in both CTAs by the multicast commits of the even CTA's MMA thread
        for (int i = 0; i < MAXB; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 3; ++i) { ptx::mbar_init(&fullA_bar[i], 1); ptx::mbar_init(&emptyA_bar[i], 1); }
        for (int i = 0; i < ACC; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], PAIR ? 256 : 128); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { ptx::tmem_alloc_2sm(tmem_ptr, TMEM_COLS); ptx::tmem_relinquish_2sm(); }
        else { ptx::tmem_alloc(tmem_ptr, TMEM_COLS); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) ptx::cluster_sync_all();       // the peer's barriers exist before anything is signalled on them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();          // the producing kernel has completed: activations / residual / GroupNorm accumulators are safe to touch

    const int kb_per_tap = a.chunks[0] + a.chunks[1];
    const int n_kb = a.n_taps * kb_per_tap;
    const int total_tiles = a.m_tiles * a.n_tiles;
    // tiles are dealt round-robin: at any instant the CTAs sweep ~gridDim consecutive tiles, i.e. one contiguous window of the
    // activation tensor (measured ~8 % faster than giving every CTA its own contiguous range: DRAM/L2 locality across CTAs)
    // work units: tiles, or (pair mode) pairs of M-adjacent tiles with the same n_tile, one per CTA of the cluster
    const int crank = PAIR ? (int)ptx::cluster_ctarank() : 0;
    const int ncta = PAIR ? 2 : 1;
    const int wid = (int)blockIdx.x / ncta, nw = (int)gridDim.x / ncta;
    const int total_units = total_tiles / ncta;
    auto unit_tile = [&](int u) { return PAIR ? ((2 * (u / a.n_tiles) + crank) * a.n_tiles + (u % a.n_tiles)) : u; };
    const SkList skl(wid, a.sk, n_kb > 0 ? n_kb : 1, total_units);
    const int my_count = a.sk ? skl.nseg : (total_units - wid + nw - 1) / nw;
    const uint16_t pair_mask = 3;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (halo) {
            if constexpr (HALO_OK) {
                if (ptx::elect_one()) {      // A producer: one halo copy per (dx, source, 64-channel chunk)
                    int sa = 0; uint32_t pa = 0;
                    for (int it = 0; it < my_count; ++it) {
                        const int tile = unit_tile(wid + it * nw);
                        const int m_tile = tile / a.n_tiles;
                        const int w_t = m_tile % a.tiles_w;
                        const int h_t = (m_tile / a.tiles_w) % a.tiles_h;
                        const int nf0 = m_tile / (a.tiles_w * a.tiles_h);
                        const int w0 = w_t * a.bw, h0 = h_t * a.bh;
                        for (int dxi = 0; dxi < 3; ++dxi)
                            for (int src = 0; src < 2; ++src) {
                                const CUtensorMap* tm = &a.tmA[src * 4 + 1];
                                for (int ch = 0; ch < a.chunks[src]; ++ch) {
                                    ptx::mbar_wait(&emptyA_bar[sa], pa ^ 1);
                                    uint8_t* s = smem + sa * A_HALO_STAGE;
                                    if constexpr (PAIR) {        // both CTAs' halo copies are counted on the even CTA's barrier
                                        if (crank == 0) ptx::mbar_arrive_expect_tx(&fullA_bar[sa], 4 * a.halo_plane);
                                        ptx::tma_load_5d_2sm(s, tm, &fullA_bar[sa], ch * BK, w0 + dxi - 1, h0 - 1, nf0, 0);
                                        ptx::tma_load_5d_2sm(s + a.halo_plane, tm, &fullA_bar[sa], ch * BK, w0 + dxi - 1, h0 - 1, nf0, 1);
                                    } else
                                    if (a.dbg & 1) { ptx::mbar_arrive_expect_tx(&fullA_bar[sa], 0); } else {
                                    ptx::mbar_arrive_expect_tx(&fullA_bar[sa], 2 * a.halo_plane);
                                    ptx::tma_load_5d(s, tm, &fullA_bar[sa], ch * BK, w0 + dxi - 1, h0 - 1, nf0, 0);
                                    ptx::tma_load_5d(s + a.halo_plane, tm, &fullA_bar[sa], ch * BK, w0 + dxi - 1, h0 - 1, nf0, 1); }
                                    if (++sa == NA) { sa = 0; pa ^= 1; }
                                }
                            }
                    }
                }
            }
        } else if (ptx::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < my_count; ++it) {
                int tile, kb0 = 0, kb1 = n_kb;
                if (a.sk) { const SkSeg sg = skl.at(it); tile = unit_tile(sg.tile); kb0 = sg.kb0; kb1 = sg.kb1; }
                else tile = unit_tile(wid + it * nw);
                const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
                const int w_t = m_tile % a.tiles_w;
                const int h_t = (m_tile / a.tiles_w) % a.tiles_h;
                const int nf_t = m_tile / (a.tiles_w * a.tiles_h);
                const int w0 = w_t * a.bw, h0 = h_t * a.bh, nf0 = nf_t * a.bnf, n0 = n_tile * BN;
                for (int tap = kb0 / kb_per_tap; tap < a.n_taps && tap * kb_per_tap < kb1; ++tap) {
                    const int dy = a.tap_dy[tap], dx = a.tap_dx[tap], mp = a.tap_map[tap];
                    for (int src = 0; src < 2; ++src) {
                        const CUtensorMap* tm = &a.tmA[src * 4 + mp];
                        const int kbase = src ? a.chunks[0] * BK : 0;
                        for (int ch = 0; ch < a.chunks[src]; ++ch) {
                            const int kb = tap * kb_per_tap + (src ? a.chunks[0] : 0) + ch;
                            if (kb < kb0 || kb >= kb1) continue;
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* s = smem + stage * STAGE_BYTES;
                            if constexpr (PAIR) {
                                // own A tile + own half of B ([W_hi rows of this rank ; W_lo rows of the other half]); bytes of both CTAs -> even CTA
                                if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (2 * A_BYTES + B_BYTES));
                                ptx::tma_load_5d_2sm(s, tm, &full_bar[stage], ch * BK, w0 + dx, h0 + dy, nf0, 0);
                                ptx::tma_load_5d_2sm(s + A_BYTES, tm, &full_bar[stage], ch * BK, w0 + dx, h0 + dy, nf0, 1);
                                ptx::tma_load_4d_2sm(s + 2 * A_BYTES, &a.tmBh, &full_bar[stage], kbase + ch * BK, n0 + crank * (BN / 2),
                                                     a.tap_base + tap, 0);
                                ptx::tma_load_4d_2sm(s + 2 * A_BYTES + B_BYTES / 2, &a.tmBh, &full_bar[stage], kbase + ch * BK,
                                                     n0 + (1 - crank) * (BN / 2), a.tap_base + tap, 1);
                            } else {
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], ((a.dbg & 1) ? 0 : 2 * A_BYTES) + ((a.dbg & 2) ? 0 : 2 * B_BYTES));
                            if (!(a.dbg & 1)) {
                            ptx::tma_load_5d(s, tm, &full_bar[stage], ch * BK, w0 + dx, h0 + dy, nf0, 0);
                            ptx::tma_load_5d(s + A_BYTES, tm, &full_bar[stage], ch * BK, w0 + dx, h0 + dy, nf0, 1); }
                            if (!(a.dbg & 2)) {
                            ptx::tma_load_4d(s + 2 * A_BYTES, &a.tmB, &full_bar[stage], kbase + ch * BK, n0,
                                             a.tap_base + tap, 0);
                            ptx::tma_load_4d(s + 2 * A_BYTES + B_BYTES, &a.tmB, &full_bar[stage], kbase + ch * BK, n0,
                                             a.tap_base + tap, 1); } }
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== halo mode: weight (B) producer =====================
        if constexpr (HALO_OK) {
            if (halo && ptx::elect_one()) {
                int sb = 0; uint32_t pb = 0;
                for (int it = 0; it < my_count; ++it) {
                    const int tile = unit_tile(wid + it * nw);
                    const int n0 = (tile % a.n_tiles) * BN;
                    for (int dxi = 0; dxi < 3; ++dxi)
                        for (int src = 0; src < 2; ++src) {
                            const int kbase = src ? a.chunks[0] * BK : 0;
                            for (int ch = 0; ch < a.chunks[src]; ++ch)
                                for (int dyi = 0; dyi < 3; ++dyi) {
                                    ptx::mbar_wait(&empty_bar[sb], pb ^ 1);
                                    uint8_t* s = smem + NA * A_HALO_STAGE + sb * (2 * B_BYTES);
                                    if constexpr (PAIR) {
                                        if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[sb], 2 * B_BYTES);
                                        ptx::tma_load_4d_2sm(s, &a.tmBh, &full_bar[sb], kbase + ch * BK, n0 + crank * (BN / 2),
                                                             a.tap_base + dyi * 3 + dxi, 0);
                                        ptx::tma_load_4d_2sm(s + B_BYTES / 2, &a.tmBh, &full_bar[sb], kbase + ch * BK, n0 + (1 - crank) * (BN / 2),
                                                             a.tap_base + dyi * 3 + dxi, 1);
                                    } else
                                    if (a.dbg & 2) { ptx::mbar_arrive_expect_tx(&full_bar[sb], 0); } else {
                                    ptx::mbar_arrive_expect_tx(&full_bar[sb], 2 * B_BYTES);
                                    ptx::tma_load_4d(s, &a.tmB, &full_bar[sb], kbase + ch * BK, n0, a.tap_base + dyi * 3 + dxi, 0);
                                    ptx::tma_load_4d(s + B_BYTES, &a.tmB, &full_bar[sb], kbase + ch * BK, n0, a.tap_base + dyi * 3 + dxi, 1); }
                                    if (++sb == NB) { sb = 0; pb ^= 1; }
                                }
                        }
                }
            }
        }
    } else if (warp == 1 && halo) {
        // ===================== halo mode: MMA issuer (one elected thread runs the whole loop; pair mode: of the even CTA) ========
        if constexpr (HALO_OK) {
            if (crank == 0 && ptx::elect_one()) {
                constexpr uint64_t DB_STRIDE = (2 * B_BYTES) >> 4;
                const uint64_t dA0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
                const uint64_t dB0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + NA * A_HALO_STAGE));
                const uint64_t dy_step = (uint64_t)((a.bw * 128) >> 4), lo_step = (uint64_t)(a.halo_plane >> 4);
                int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
                uint64_t dA = dA0, dB = dB0;
                int it = 0;
                const int groups = 3 * kb_per_tap;
                for (; it < my_count; ++it) {
                    const int as = it % ACC;
                    const uint32_t aphase = (it / ACC) & 1;
                    ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * DCOLS;
                    uint32_t acc = 0u;
                    for (int g = 0; g < groups; ++g) {
                        ptx::mbar_wait(&fullA_bar[sa], pa);
                        uint64_t da_hi = dA;
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            ptx::mbar_wait(&full_bar[sb], pb);
                            ptx::tc_fence_after();
                            const uint64_t da_lo = da_hi + lo_step;
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks) {
                                if constexpr (PAIR) {
                                    ptx::umma_bf16_2sm(tmem_d, da_hi + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC_WIDE2, acc);
                                    ptx::umma_bf16_2sm(tmem_d, da_lo + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC2, 1u);
                                } else {
                                ptx::umma_bf16(tmem_d, da_hi + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC_WIDE, acc);   // A_hi.[W_hi;W_lo]
                                ptx::umma_bf16(tmem_d, da_lo + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC, 1u);          // + A_lo.W_hi
                                }
                                acc = 1u;
                            }
                            if constexpr (PAIR) ptx::umma_commit_2sm(&empty_bar[sb], pair_mask); else ptx::umma_commit(&empty_bar[sb]);
                            da_hi += dy_step;
                            dB += DB_STRIDE;
                            if (++sb == NB) { sb = 0; pb ^= 1; dB = dB0; }
                        }
                        if constexpr (PAIR) ptx::umma_commit_2sm(&emptyA_bar[sa], pair_mask); else ptx::umma_commit(&emptyA_bar[sa]);
                        dA += (uint64_t)(A_HALO_STAGE >> 4);
                        if (++sa == NA) { sa = 0; pa ^= 1; dA = dA0; }
                    }
                    if constexpr (PAIR) ptx::umma_commit_2sm(&tfull_bar[as], pair_mask); else ptx::umma_commit(&tfull_bar[as]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected thread runs the whole loop) =====================
        // The tensor-core queue is short: whatever the issuing thread executes between the last MMA of one stage and the
        // first of the next is exposed as tensor-pipe idle time (measured ~330 clk / stage with a per-stage elect + descriptor
        // rebuild, i.e. as long as the 8 MMAs of a BN = 64 stage).  Descriptors are therefore carried incrementally.
        if (crank == 0 && ptx::elect_one()) {
            constexpr uint64_t D_STRIDE = STAGE_BYTES >> 4, D_ALO = A_BYTES >> 4, D_B = (2 * A_BYTES) >> 4;
            const uint64_t d0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
            int stage = 0; uint32_t phase = 0;
            uint64_t da_hi = d0;
            int it = 0;
            for (; it < my_count; ++it) {
                int nkb_seg = n_kb;
                if (a.sk) { const SkSeg sg = skl.at(it); nkb_seg = sg.kb1 - sg.kb0; }
                const int as = it % ACC;
                const uint32_t aphase = (it / ACC) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * DCOLS;
                uint32_t acc = 0u;
                for (int kb = 0; kb < nkb_seg; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da_lo = da_hi + D_ALO, db_hi = da_hi + D_B;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t off = (uint64_t)(ks * 2);   // 16 bf16 = 32 B, encoded >> 4
                        if constexpr (WIDE) {
                            if constexpr (PAIR) {
                                ptx::umma_bf16_2sm(tmem_d, da_hi + off, db_hi + off, IDESC_WIDE2, acc);
                                ptx::umma_bf16_2sm(tmem_d, da_lo + off, db_hi + off, IDESC2, 1u);
                            } else {
                            ptx::umma_bf16(tmem_d, da_hi + off, db_hi + off, IDESC_WIDE, acc);      // A_hi.[W_hi;W_lo]
                            ptx::umma_bf16(tmem_d, da_lo + off, db_hi + off, IDESC, 1u);            // + A_lo.W_hi
                            }
                        } else {
                            const uint64_t db_lo = db_hi + (uint64_t)(B_BYTES >> 4);
                            ptx::umma_bf16(tmem_d, da_lo + off, db_hi + off, IDESC, acc);
                            ptx::umma_bf16(tmem_d, da_hi + off, db_lo + off, IDESC, 1u);
                            ptx::umma_bf16(tmem_d, da_hi + off, db_hi + off, IDESC, 1u);
                        }
                        acc = 1u;
                    }
                    if constexpr (PAIR) ptx::umma_commit_2sm(&empty_bar[stage], pair_mask); else ptx::umma_commit(&empty_bar[stage]);
                    da_hi += D_STRIDE;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; da_hi = d0; }
                }
                if constexpr (PAIR) ptx::umma_commit_2sm(&tfull_bar[as], pair_mask); else ptx::umma_commit(&tfull_bar[as]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;              // TMEM lane quarter this warp may access (warp % 4)
        const int grp = (warp - 4) >> 2;     // two epilogue groups of 4 warps drain alternate tiles concurrently
        const int r = q * 32 + lane;         // tile row handled by this thread
        const bool vec_ok = (a.c_out % 16) == 0;
        float* stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + (warp - 4) * 1024;          // 4 KiB per warp
        // GroupNorm sums without float atomics (bit-reproducible): every warp owns a slot [16 groups][sum, sumsq], one lane per address
        float* gn_base = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 8 * 4096 + 256);
        float* gn_acc = gn_base + (warp - 4) * 32;
        const int tig = threadIdx.x & 127;                     // thread index inside the epilogue group
        int cur_sample = -1;
        if (a.gn_stats) {
            gn_acc[lane] = 0.f;
            __syncwarp();
        }
        auto gn_flush = [&](int next_sample) {
            // all 4 warps of the group are between tiles here (named barrier), so the slots are quiescent.  Inside the CTA tiles
            // and warps are added in a fixed order; ACROSS CTAs the double atomics commute exactly, because every contribution is
            // first rounded to a multiple of 2^-24 (a no-op for |v| >= 0.5: a float has no finer bits) and sums of such multiples
            // are exact in a double while they stay below 2^29 - the order CTAs arrive in cannot change the result.
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if (tig < 2 * a.gn_groups) {
                float* s4 = gn_base + grp * 128 + tig;
                const float v = ((s4[0] + s4[32]) + s4[64]) + s4[96];
                s4[0] = 0.f; s4[32] = 0.f; s4[64] = 0.f; s4[96] = 0.f;
                if (cur_sample >= 0 && v != 0.f)
                    atomicAdd(a.gn_stats + (int64_t)cur_sample * a.gn_groups * 2 + tig, rint((double)v * 16777216.0) * (1.0 / 16777216.0));
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            cur_sample = next_sample;
        };
        for (int it = grp; it < my_count; it += 2) {
            int tile;
            bool sk_head = false, sk_tail = false;       // this piece hands its partial over / finishes a tile with the previous CTA's
            if (a.sk) {
                const SkSeg sg = skl.at(it);
                tile = unit_tile(sg.tile);
                sk_head = sg.kb1 < n_kb;
                sk_tail = sg.kb0 > 0;
            } else {
                tile = unit_tile(wid + it * nw);
            }
            const int as = it % ACC;
            const uint32_t aphase = (it / ACC) & 1;
            if (a.dbg & 128) {
                ptx::mbar_wait(&tfull_bar[as], aphase);
                ptx::tc_fence_after();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[as]);
                continue;
            }
            const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
            const int w_t = m_tile % a.tiles_w;
            const int h_t = (m_tile / a.tiles_w) % a.tiles_h;
            const int nf_t = m_tile / (a.tiles_w * a.tiles_h);
            const int n0 = n_tile * BN;
            const int nf = nf_t * a.bnf + r / (a.bh * a.bw);
            const int h = h_t * a.bh + (r / a.bw) % a.bh;
            const int w = w_t * a.bw + r % a.bw;
            const int64_t orow = ((int64_t)nf * a.ho_full + (h * a.mul + a.off_h)) * a.wo_full + (w * a.mul + a.off_w);
            int64_t rrow = orow;
            if (a.res_bcast_f > 0) rrow = (orow / ((int64_t)a.res_bcast_f * a.p_out)) * a.p_out + (orow % a.p_out);
            if (a.gn_stats && !sk_head) {
                const int bsample = (int)(__shfl_sync(0xffffffffu, orow, 0) / a.rows_per_sample);   // tile-uniform (128 | rows_per_sample)
                if (bsample != cur_sample) gn_flush(bsample);
            }
            // rows this lane handles in the transposed (coalesced) epilogue: tile rows q*32 + (lane>>3) + 4i
            int32_t orow_l[8], rrow_l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int src_lane = (lane >> 3) + 4 * i;
                const int64_t oo = __shfl_sync(0xffffffffu, orow, src_lane);
                const int64_t rro = __shfl_sync(0xffffffffu, rrow, src_lane);
                orow_l[i] = (int32_t)oo; rrow_l[i] = (int32_t)rro;
            }

            if (a.residual && vec_ok) {
                // pull this warp's residual block (32 rows x BN columns) into L2 while the MMAs of the tile are in flight
                const int pc = n0 + 32 * (lane & 7);
                if (32 * (lane & 7) < BN && pc < a.c_out) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) ptx::prefetch_l2(a.residual + (int64_t)rrow_l[i] * a.c_out + pc);
                }
            }
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * DCOLS);
            // stream-K hand-over: row r of the [128][BN] fp32 slot of the producing CTA
            float* sk_row = nullptr;
            if (sk_head) sk_row = a.sk_ws + ((size_t)blockIdx.x * BM + r) * BN;
            if (sk_tail) {
                sk_row = a.sk_ws + ((size_t)(blockIdx.x - ncta) * BM + r) * BN;       // the same rank of the previous work unit
                if (tig == 0) {
                    int v;
                    do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a.sk_flag + (blockIdx.x - ncta)) : "memory"); } while (v == 0);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            }
            if constexpr (BN >= 32) {
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r0[16], r1[16], s0[16], s1[16];
                    ptx::tmem_ld16(taddr + c0, r0);
                    ptx::tmem_ld16(taddr + c0 + 16, r1);
                    if constexpr (WIDE) {
                        // partner column block of this chunk: [x | x + BN], or (pair mode, blocks [hi0 | lo1 | hi1 | lo0]) 0 <-> 3, 1 <-> 2
                        const int sec = PAIR ? (c0 < BN / 2 ? c0 + BN + BN / 2 : c0 + BN / 2) : BN + c0;
                        ptx::tmem_ld16(taddr + sec, s0);
                        ptx::tmem_ld16(taddr + sec + 16, s1);
                    }
                    ptx::tmem_ld_wait();
                    if constexpr (WIDE) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __uint_as_float(s0[j]));
                            r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __uint_as_float(s1[j]));
                        }
                    }
                    if (sk_head) {           // partial accumulator of this row -> the slot (fp32, 128 contiguous bytes per chunk)
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            *reinterpret_cast<uint4*>(sk_row + c0 + j) = make_uint4(r0[j], r0[j + 1], r0[j + 2], r0[j + 3]);
                            *reinterpret_cast<uint4*>(sk_row + c0 + 16 + j) = make_uint4(r1[j], r1[j + 1], r1[j + 2], r1[j + 3]);
                        }
                        continue;
                    }
                    if (sk_tail) {           // + the first part of the K range, computed by the previous CTA (L2, not L1)
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 p0 = __ldcg(reinterpret_cast<const float4*>(sk_row + c0 + j));
                            const float4 p1 = __ldcg(reinterpret_cast<const float4*>(sk_row + c0 + 16 + j));
                            r0[j] = __float_as_uint(__uint_as_float(r0[j]) + p0.x); r0[j + 1] = __float_as_uint(__uint_as_float(r0[j + 1]) + p0.y);
                            r0[j + 2] = __float_as_uint(__uint_as_float(r0[j + 2]) + p0.z); r0[j + 3] = __float_as_uint(__uint_as_float(r0[j + 3]) + p0.w);
                            r1[j] = __float_as_uint(__uint_as_float(r1[j]) + p1.x); r1[j + 1] = __float_as_uint(__uint_as_float(r1[j + 1]) + p1.y);
                            r1[j + 2] = __float_as_uint(__uint_as_float(r1[j + 2]) + p1.z); r1[j + 3] = __float_as_uint(__uint_as_float(r1[j + 3]) + p1.w);
                        }
                    }
                    if (a.tma_store) {
                        // plain F32 output (+bias, +GroupNorm sums): thread-per-row -> swizzled smem -> one TMA store per
                        // 32x32 block; the warp never waits for its global stores
                        const int nb = n0 + c0;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(r0[j]); v[16 + j] = __uint_as_float(r1[j]); }
                        if (a.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(a.bias + nb + j);
                                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                            }
                        }
                        if (a.rot_cos && nb < a.rot_cols) {
                            // q (scaled) / k block of one head: rotate the 16 (2i, 2i+1) pairs by this row's frame angle
                            const int fr = (int)((orow / a.rot_rows_per_frame) % a.rot_frames);
                            const float sc = nb < a.rot_scale_cols ? a.rot_scale : 1.f;
                            const float4* c4p = reinterpret_cast<const float4*>(a.rot_cos + fr * 16);
                            const float4* s4p = reinterpret_cast<const float4*>(a.rot_sin + fr * 16);
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) {
                                const float4 c4 = c4p[qd], s4 = s4p[qd];
                                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float x = v[8 * qd + 2 * e] * sc, y = v[8 * qd + 2 * e + 1] * sc;
                                    v[8 * qd + 2 * e] = x * cc[e] - y * ss[e];
                                    v[8 * qd + 2 * e + 1] = y * cc[e] + x * ss[e];
                                }
                            }
                        }
                        if (a.gn_stats) {
#pragma unroll
                            for (int oct = 0; oct < 4; ++oct) {        // cpg is a multiple of 8: one (sum, sumsq) per 8 columns
                                float s8 = 0.f, q8 = 0.f;
#pragma unroll
                                for (int j = 0; j < 8; ++j) { s8 += v[8 * oct + j]; q8 = fmaf(v[8 * oct + j], v[8 * oct + j], q8); }
                                s8 = warp_sum(s8); q8 = warp_sum(q8);
                                if (lane == 0) {
                                    const int gi = (nb + 8 * oct) / a.gn_cpg;
                                    gn_acc[gi * 2] += s8; gn_acc[gi * 2 + 1] += q8;
                                }
                            }
                        }
                        if (lane == 0) ptx::tma_store_wait_read0();      // previous store has finished reading this stage
                        __syncwarp();
                        float4* st4 = reinterpret_cast<float4*>(stage) + lane * 8;
                        const int sw = lane & 7;
#pragma unroll
                        for (int qd = 0; qd < 8; ++qd) st4[qd ^ sw] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && !(a.dbg & 8)) {
                            ptx::tma_store_2d(&a.tmOut, stage, nb, orow_l[0]);     // lane 0: orow_l[0] = first row of this warp's slab
                            ptx::tma_store_commit();
                        }
                    } else if (vec_ok && (n0 + c0 + 32 <= a.c_out)) {
                        epi_block32(a, stage, r0, r1, n0 + c0, orow_l, rrow_l, gn_acc, lane);
                    } else {
                        epi_chunk16(a, r0, n0 + c0, orow, rrow, gn_acc, lane, vec_ok);
                        epi_chunk16(a, r1, n0 + c0 + 16, orow, rrow, gn_acc, lane, vec_ok);
                    }
                }
            } else {
                uint32_t r0[16], s0[16];
                ptx::tmem_ld16(taddr, r0);
                if constexpr (WIDE) ptx::tmem_ld16(taddr + BN, s0);
                ptx::tmem_ld_wait();
                if constexpr (WIDE) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __uint_as_float(s0[j]));
                }
                epi_chunk16(a, r0, n0, orow, rrow, gn_acc, lane, vec_ok);
            }
            ptx::tc_fence_before();
            if constexpr (PAIR) ptx::mbar_arrive_leader(&tempty_bar[as]); else ptx::mbar_arrive(&tempty_bar[as]);
            if (sk_head) {
                // every row of the slot is written and fenced before one thread publishes it
                __threadfence();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if (tig == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.sk_flag + blockIdx.x), "r"(1) : "memory");
            }
            if (sk_tail) {
                // all 128 threads have read the slot: re-arm the flag for the next launch
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if (tig == 0) a.sk_flag[blockIdx.x - ncta] = 0;
            }
        }
        if (a.gn_stats) gn_flush(-1);
        if (a.tma_store && lane == 0) ptx::tma_store_wait0();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) ptx::cluster_sync_all();       // the peer may still multicast into this CTA's ring / arrive on its barriers
    if (warp == 2) {
        ptx::tc_fence_after();
        if constexpr (PAIR) ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS); else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int make_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return LFDM_E_NODRIVER;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims,
                     strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (1000 + (int)r);
}

template <int BN, int STAGES, bool WIDE, bool PAIR>
int launch(const TcArgs& a_in, cudaStream_t st) {
    TcArgs a = a_in;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BN * BK * 2;
    constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 8 * 4096 + 256 + 1024;  // + 8 staging tiles + barriers + GN slots (8 warps x 128 B)
    static_assert(SMEM <= 232448, "shared memory budget");
    static PerDeviceOnce once;
    static int num_sms = 0, max_clusters = 0;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, WIDE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return (int)e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
        max_clusters = num_sms / 2;
        once.mark();
    }
    int total = a.m_tiles * a.n_tiles;
    int grid = total < num_sms ? total : num_sms;
    int units = total, workers = grid;          // work units (tiles / tile pairs) and the CTAs / clusters that walk them
    if (PAIR) {                                 // clusters of 2 CTAs, one pair of M-adjacent tiles per cluster and step
        static PerDeviceOnce once_c;
        if (once_c.need()) {
            // clusters of 2 that can be resident at a time (GPCs with an odd number of usable SMs leave one SM out)
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)(num_sms & ~1)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM;
            cudaLaunchAttribute at[1];
            memset(at, 0, sizeof(at));
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int mc = 0;
            if (cudaOccupancyMaxActiveClusters(&mc, conv_tc_kernel<BN, STAGES, WIDE, PAIR>, &cfg) == cudaSuccess && mc > 0 && mc < max_clusters)
                max_clusters = mc;
            (void)cudaGetLastError();
            once_c.mark();
        }
        units = total / 2;
        workers = units < max_clusters ? units : max_clusters;
        grid = 2 * workers;
    }
    if (a.sk) {
        // stream-K pays when the last wave is mostly empty; K-blocks per worker >= K-blocks per tile because units > workers
        const int n_kb = a.sk;
        const int waves = (units + workers - 1) / workers;
        const double ideal = (double)units / workers;
        a.sk = (units > workers && (waves - ideal) / waves > 0.08) ? (int)(((long long)units * n_kb + workers - 1) / workers) : 0;
    }
    cudaError_t e = lfdm_launch_pdl_cluster(conv_tc_kernel<BN, STAGES, WIDE, PAIR>, dim3(grid), dim3(NUM_THREADS), (size_t)SMEM, st,
                                            PAIR ? 2 : 1, a);
    if (e != cudaSuccess) return (int)e;
    return 0;
}

}  // namespace

int lfdm_conv_tc(const lfdm_conv_desc* d, cudaStream_t st) {
    if (!d || !d->w_sb) return LFDM_E_BADARG;
    // ---- operand checks
    int nsrc = d->a_c[1] > 0 ? 2 : 1;
    for (int s = 0; s < nsrc; ++s) {
        if (!d->a_sb[s] || (d->a_c[s] % BK) != 0) return LFDM_E_UNSUPP;
        if (reinterpret_cast<uintptr_t>(d->a_sb[s]) & 15) return LFDM_E_UNSUPP;
    }
    // reflect padding (NATOPS up-sampling conv, reference :156-163): reflecting the x2 nearest up-sampled map by one pixel equals
    // CLAMPING on the low-resolution map, so the caller hands in the replicate-padded input [(h_in+2) x (w_in+2)]
    // (lfdm_pad_replicate_rows) and every tap coordinate shifts by +1: no out-of-bounds access remains.
    if (d->reflect && d->mode != LFDM_CONV_UPNEAREST) return LFDM_E_UNSUPP;
    const int rpad = d->reflect ? 1 : 0;
    const int cin_total = d->a_c[0] + d->a_c[1];

    // ---- iteration geometry (the grid the M tiles walk) and per-launch tap lists
    int ht, wt, mul = 1, n_launch = 1, taps_per_launch, total_taps;
    bool parity_views = false;
    if (d->mode == LFDM_CONV_DIRECT && d->stride == 1) {
        if (d->h_out != d->h_in || d->w_out != d->w_in || d->kh * d->kw > MAX_TAPS) return LFDM_E_UNSUPP;
        ht = d->h_out; wt = d->w_out; taps_per_launch = d->kh * d->kw; total_taps = taps_per_launch;
    } else if (d->mode == LFDM_CONV_DIRECT && d->stride == 2) {
        if (d->kh != 4 || d->kw != 4 || d->pad != 1 || d->h_out * 2 != d->h_in || d->w_out * 2 != d->w_in) return LFDM_E_UNSUPP;
        ht = d->h_out; wt = d->w_out; taps_per_launch = 16; total_taps = 16; parity_views = true;
    } else if (d->mode == LFDM_CONV_TRANSPOSED) {
        if (d->kh != 4 || d->kw != 4 || d->pad != 1 || d->stride != 2 || d->h_out != 2 * d->h_in || d->w_out != 2 * d->w_in) return LFDM_E_UNSUPP;
        ht = d->h_in; wt = d->w_in; mul = 2; n_launch = 4; taps_per_launch = 4; total_taps = 16;
    } else if (d->mode == LFDM_CONV_UPNEAREST) {
        if (d->kh != 3 || d->kw != 3 || d->pad != 1 || d->h_out != 2 * d->h_in || d->w_out != 2 * d->w_in) return LFDM_E_UNSUPP;
        ht = d->h_in; wt = d->w_in; mul = 2; n_launch = 4; taps_per_launch = 4; total_taps = 16;
    } else {
        return LFDM_E_UNSUPP;
    }
    int bw = wt < 128 ? wt : 128;
    if (wt % bw) return LFDM_E_UNSUPP;
    int bh = ht < 128 / bw ? ht : 128 / bw;
    if (bh < 1 || ht % bh) return LFDM_E_UNSUPP;
    if (128 % (bw * bh)) return LFDM_E_UNSUPP;
    int bnf = 128 / (bw * bh);
    if (d->nf % bnf) return LFDM_E_UNSUPP;
    if (bw > 256 || bh > 256 || bnf > 256) return LFDM_E_UNSUPP;

    // ---- N tile
    int bn;
    if (d->c_out % 128 == 0) bn = 128;
    else if (d->c_out % 64 == 0) bn = 64;
    else if (d->c_out % 32 == 0) bn = 32;
    else if (d->c_out <= 16) bn = 16;
    else return LFDM_E_UNSUPP;
    const int c_out_pad = ((d->c_out + bn - 1) / bn) * bn;
    if (d->gn_stats && ((d->gn_cpg % 8) != 0 || d->rows_per_sample % 128 != 0 || mul != 1 || d->c_out / d->gn_cpg > 16)) return LFDM_E_UNSUPP;

    TcArgs a;
    memset(&a, 0, sizeof(a));
    // ---- tensor maps: activations
    for (int s = 0; s < 2; ++s) {
        int src = s < nsrc ? s : 0;   // unused maps alias source 0 (never dereferenced: chunks == 0)
        const bf16* base = reinterpret_cast<const bf16*>(d->a_sb[src]);
        const cuuint64_t C = (cuuint64_t)d->a_c[src], W = (cuuint64_t)d->w_in, H = (cuuint64_t)d->h_in;
        for (int v = 0; v < 4; ++v) {
            int rc;
            if (parity_views) {
                const int ph = v >> 1, pw = v & 1;
                cuuint64_t dims[5] = {C, W / 2, H / 2, (cuuint64_t)d->nf, 2};
                cuuint64_t strides[4] = {2 * C * 2, 2 * W * C * 2, H * W * C * 2, (cuuint64_t)d->a_plane[src] * 2};
                cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bnf, 1};
                rc = make_map(&a.tmA[s * 4 + v], base + ((int64_t)ph * W + pw) * C, 5, dims, strides, box);
            } else {
                const cuuint64_t Wp = W + 2 * rpad, Hp = H + 2 * rpad;
                cuuint64_t dims[5] = {C, Wp, Hp, (cuuint64_t)d->nf, 2};
                cuuint64_t strides[4] = {C * 2, Wp * C * 2, Hp * Wp * C * 2, (cuuint64_t)d->a_plane[src] * 2};
                cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bnf, 1};
                rc = make_map(&a.tmA[s * 4 + v], base, 5, dims, strides, box);
            }
            if (rc) return rc;
        }
    }
    // ---- weights: [plane][tap][c_out_pad][cin_total]
    {
        cuuint64_t dims[4] = {(cuuint64_t)cin_total, (cuuint64_t)c_out_pad, (cuuint64_t)total_taps, 2};
        cuuint64_t strides[3] = {(cuuint64_t)cin_total * 2, (cuuint64_t)cin_total * c_out_pad * 2, (cuuint64_t)d->w_plane * 2};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bn, 1, 1};
        int rc = make_map(&a.tmB, d->w_sb, 4, dims, strides, box);
        if (rc) return rc;
        cuuint32_t boxh[4] = {(cuuint32_t)BK, (cuuint32_t)(bn >= 16 ? bn / 2 : bn), 1, 1};
        rc = make_map(&a.tmBh, d->w_sb, 4, dims, strides, boxh);
        if (rc) return rc;
    }
    // ---- TMA-store epilogue: plain F32 output (+bias, +GroupNorm sums), consecutive output rows per tile
    a.tma_store = 0;
    {
        static const bool allow = (getenv("LFDM_CONV_NO_TMA_STORE") == nullptr);     // A/B switch
        const int64_t m_out = (int64_t)d->nf * d->h_out * d->w_out;
        if (allow && d->out_f32 && !d->out_sb && !d->residual && d->f32_act == LFDM_ACT_NONE && mul == 1 && bn >= 32 &&
            d->c_out % 32 == 0 && (reinterpret_cast<uintptr_t>(d->out_f32) & 15) == 0) {
            cuuint64_t dims[2] = {(cuuint64_t)d->c_out, (cuuint64_t)m_out};
            cuuint64_t strides[1] = {(cuuint64_t)d->c_out * 4};
            cuuint32_t box[2] = {32, 32};
            int rc = make_map(&a.tmOut, d->out_f32, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
            if (rc) return rc;
            a.tma_store = 1;
        }
    }
    // ---- 3x3 halo mode: tile inside one image, (bh+2) x bw halo copies per dx (A re-reads 9x -> 3(bh+2)/bh x)
    a.halo = 0;
    {
        static const bool na2 = (getenv("LFDM_CONV_NA3") == nullptr);
        a.na2 = na2 ? 1 : 0;
    }
    {
        static const bool allow = (getenv("LFDM_CONV_NO_HALO") == nullptr);           // A/B switch
        if (allow && d->mode == LFDM_CONV_DIRECT && d->stride == 1 && d->kh == 3 && d->kw == 3 && d->pad == 1 && bnf == 1 &&
            bw >= 8 && bw <= 32 && (bn == 64 || bn == 128)) {
            for (int s = 0; s < nsrc; ++s) {
                const bf16* base = reinterpret_cast<const bf16*>(d->a_sb[s]);
                const cuuint64_t C = (cuuint64_t)d->a_c[s], W = (cuuint64_t)d->w_in, H = (cuuint64_t)d->h_in;
                cuuint64_t dims[5] = {C, W, H, (cuuint64_t)d->nf, 2};
                cuuint64_t strides[4] = {C * 2, W * C * 2, H * W * C * 2, (cuuint64_t)d->a_plane[s] * 2};
                cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)(bh + 2), 1, 1};
                int rc = make_map(&a.tmA[s * 4 + 1], base, 5, dims, strides, box);
                if (rc) return rc;
            }
            a.halo = 1;
            a.halo_plane = (bh + 2) * bw * 128;
        }
    }
    {
        static const int dbg = getenv("LFDM_CONV_DBG") ? atoi(getenv("LFDM_CONV_DBG")) : 0;
        a.dbg = dbg;
    }
    if (d->rot_cos) {
        if (!a.tma_store || !d->rot_sin || d->rot_frames <= 0 || d->rot_rows_per_frame <= 0 || (d->rot_cols % 32) || (d->rot_scale_cols % 32) ||
            d->rot_cols > d->c_out || ((reinterpret_cast<uintptr_t>(d->rot_cos) | reinterpret_cast<uintptr_t>(d->rot_sin)) & 15))
            return LFDM_E_UNSUPP;
        a.rot_cos = d->rot_cos; a.rot_sin = d->rot_sin; a.rot_frames = d->rot_frames; a.rot_rows_per_frame = d->rot_rows_per_frame;
        a.rot_cols = d->rot_cols; a.rot_scale_cols = d->rot_scale_cols; a.rot_scale = d->rot_scale;
    }
    a.chunks[0] = d->a_c[0] / BK;
    a.chunks[1] = nsrc > 1 ? d->a_c[1] / BK : 0;
    a.bw = bw; a.bh = bh; a.bnf = bnf;
    a.tiles_w = wt / bw; a.tiles_h = ht / bh;
    a.m_tiles = a.tiles_w * a.tiles_h * (d->nf / bnf);
    a.n_tiles = c_out_pad / bn;
    a.ho_full = d->h_out; a.wo_full = d->w_out; a.mul = mul;
    a.c_out = d->c_out;
    a.bias = d->bias; a.residual = d->residual; a.res_bcast_f = d->res_bcast_f;
    a.p_out = (int64_t)d->h_out * d->w_out;
    a.out_f32 = d->out_f32; a.f32_act = d->f32_act;
    a.out_sb = reinterpret_cast<bf16*>(d->out_sb); a.out_plane = d->out_plane; a.sb_act = d->sb_act;
    a.sb_scale = d->sb_scale; a.sb_shift = d->sb_shift;
    a.gn_stats = d->gn_stats; a.gn_cpg = d->gn_cpg > 0 ? d->gn_cpg : 8;
    a.gn_groups = d->c_out / a.gn_cpg; a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
    a.n_taps = taps_per_launch;

    for (int L = 0; L < n_launch; ++L) {
        a.tap_base = L * taps_per_launch;
        a.off_h = 0; a.off_w = 0;
        if (d->mode == LFDM_CONV_DIRECT && d->stride == 1) {
            for (int t = 0; t < taps_per_launch; ++t) {
                a.tap_map[t] = 0; a.tap_dy[t] = (int8_t)(t / d->kw - d->pad); a.tap_dx[t] = (int8_t)(t % d->kw - d->pad);
            }
        } else if (parity_views) {
            // input row = 2*ho - 1 + kh = 2*(ho + a) + ph  with (kh - 1) = 2a + ph
            for (int t = 0; t < 16; ++t) {
                int kh = t / 4, kw = t % 4;
                int eh = kh - 1, ew = kw - 1;
                int ph = eh & 1, pw = ew & 1;              // two's complement: (-1 & 1) == 1
                int ah = (eh - ph) / 2, aw = (ew - pw) / 2;
                a.tap_map[t] = (int8_t)(ph * 2 + pw); a.tap_dy[t] = (int8_t)ah; a.tap_dx[t] = (int8_t)aw;
            }
        } else {
            // 2x2 sub-kernels per output parity (p, q); weights packed [phase][tap = a*2+b] by the host packer
            const int p = L >> 1, qq = L & 1;
            a.off_h = p; a.off_w = qq;
            int dyl[2], dxl[2];
            if (d->mode == LFDM_CONV_TRANSPOSED) {          // p=0: kh in {1,3} -> dy {0,-1};  p=1: kh in {0,2} -> dy {+1,0}
                dyl[0] = p ? 1 : 0; dyl[1] = p ? 0 : -1;
                dxl[0] = qq ? 1 : 0; dxl[1] = qq ? 0 : -1;
            } else {                                        // up-nearest: p=0: dy {-1,0};  p=1: dy {0,+1}
                dyl[0] = p ? 0 : -1; dyl[1] = p ? 1 : 0;
                dxl[0] = qq ? 0 : -1; dxl[1] = qq ? 1 : 0;
            }
            for (int t = 0; t < 4; ++t) { a.tap_map[t] = 0; a.tap_dy[t] = (int8_t)(dyl[t >> 1] + rpad); a.tap_dx[t] = (int8_t)(dxl[t & 1] + rpad); }
        }
        int rc;
        const int n_kb = taps_per_launch * (a.chunks[0] + a.chunks[1]);
        const bool wide = n_kb >= 3;          // MMA/smem-bound tiles: wide 2-MMA scheme; short-K GEMMs: deeper accumulator ring
        // ---- pair mode (cta_group::2): clusters of 2 CTAs on M-adjacent tiles of the same n_tile run one M = 256 MMA stream; each
        // CTA holds half of the rows of every weight tile.  (Sharing the weight tile by TMA multicast with cta_group::1 MMAs was
        // measured first: no gain -- the bound is the shared-memory port of each SM, not L2 -> SM traffic.)
        a.pair = 0;
        {
            // OFF by default: measured on B200 (profiles/r02_conv_pair_experiments.md) the CTA-pair form is 5-20 % SLOWER than
            // independent CTAs on every 3x3 layer of the UNet, and sharing the weight tile by TMA multicast (cta_group::1 MMAs)
            // changes nothing; LFDM_CONV_PAIR=1 turns it on for experiments.
            static const bool allow = (getenv("LFDM_CONV_PAIR") != nullptr);
            if (allow && wide && (bn == 64 || bn == 128) && n_kb >= 8 && (a.m_tiles % 2) == 0 && a.m_tiles * a.n_tiles >= 64) a.pair = 1;
        }
        // ---- stream-K: few, long tiles whose last wave is mostly empty (4x4 / 8x8 levels: 160 / 320 tiles of 72 / 36 K-blocks on
        // 148 SMs) are cut at K-block granularity into one equal range per CTA
        a.sk = 0;
        {
            static const bool allow = (getenv("LFDM_CONV_NO_STREAMK") == nullptr);       // A/B switch
            int sms = 0, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (sms <= 0) sms = 148;
            // eligibility only: the split itself (K-blocks per worker) is fixed in launch<>(), where the number of resident
            // CTAs / clusters is known
            if (allow && d->sk_workspace && d->sk_flags && !a.halo && n_launch == 1 && wide && bn >= 64 && n_kb >= 8 &&
                (long long)sms * BM * bn * 4 <= d->sk_workspace_bytes && sms <= d->sk_slots) {
                a.sk = n_kb;                                          // marker: K-blocks per tile
                a.sk_ws = reinterpret_cast<float*>(d->sk_workspace);
                a.sk_flag = d->sk_flags;
            }
        }
        switch (bn) {
            case 128: rc = wide ? (a.pair ? launch<128, 3, true, true>(a, st) : launch<128, 3, true, false>(a, st)) : launch<128, 3, false, false>(a, st); break;
            case 64: rc = wide ? (a.pair ? launch<64, 4, true, true>(a, st) : launch<64, 4, true, false>(a, st)) : launch<64, 4, false, false>(a, st); break;
            case 32: rc = launch<32, 4, true, false>(a, st); break;
            default: rc = launch<16, 4, true, false>(a, st); break;
        }
        if (rc) return rc;
    }
    return 0;
}
