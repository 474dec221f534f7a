// attention.cu — attention cores of the denoising UNet, fp32 CUDA-core kernels (dim_head = 32).
//
//  lfdm_attn_softmax : softmax(q k^T + bias) v over short sequences (temporal attention over F=40 frames with rotary
//                      q/k and T5-style relative position bias; mid-block spatial attention over h*w tokens without).
//                      Reference: Attention.forward DM/modules/video_flow_diffusion.py:303-363; the 'b c f h w <->
//                      b (h w) f c' permutes of EinopsToAndFrom (:270-283) are folded into the row gather, so no
//                      layout copy is ever materialised.
//  lfdm_attn_linear  : SpatialLinearAttention core (:253-263): q softmax over d, k softmax over n, q*scale,
//                      ctx = k v^T, out = ctx^T q, per (frame, head).
//
// qkv is the fp32 output of the fused qkv projection GEMM: row-major [M][3*heads*32] (q | k | v, head-major inside).
// Outputs are written in SB (operand of the out-projection GEMM) and/or F32.
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int DH = 32;
constexpr int MAXL = 64;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one warp per (sequence, head).  Register-tiled: lane (a = lane/8, b = lane%8) owns score rows i = a + 4r and
// columns j = b + 8c (RIP x CJ accumulators), so every smem operand fetched with one LDS.128 feeds RIP*CJ*4/(RIP+CJ)
// FMAs instead of one; P is staged through smem once and P.V is tiled the same way (rows i, 4 consecutive d per lane).
constexpr int QP = DH + 4;     // smem row pitch (floats): 16-byte aligned rows, conflict-free LDS.128 over 8 rows

template <int RIP, int CJ, bool ALIAS_P>   // ALIAS_P: single row pass -> P may overwrite the dead Q/K tiles
__global__ void __launch_bounds__(128) attn_softmax_kernel(const float* __restrict__ qkv, bf16* __restrict__ out_sb,
                                                           int64_t out_plane, float* __restrict__ out_f32,
                                                           int64_t n_seq, int L, int heads, int64_t inner,
                                                           int64_t outer_stride, int64_t inner_stride,
                                                           int64_t row_stride, const float* __restrict__ rot_cos,
                                                           const float* __restrict__ rot_sin,
                                                           const float* __restrict__ pos_bias, int warps_per_block) {
    pdl_prologue_done();
    extern __shared__ __align__(16) float s_dyn[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int LP = CJ * 8;                 // padded sequence length (multiple of 8, >= L)
    const int PP = LP + 4;                 // P row pitch
    float* sq = s_dyn + (size_t)w * (3 * LP * QP + (ALIAS_P ? 0 : LP * PP));
    float* sk = sq + LP * QP;
    float* sv = sk + LP * QP;
    float* sp = ALIAS_P ? sq : sv + LP * QP;
    const int64_t unit = (int64_t)blockIdx.x * warps_per_block + w;
    if (unit >= n_seq * heads) return;     // whole warp exits together (only __syncwarp below)
    const int64_t s = unit / heads;
    const int h = (int)(unit - s * heads);
    const int hid = heads * DH;
    const int64_t base = (s / inner) * outer_stride + (s % inner) * inner_stride;
    const float scale = 0.17677669529663687f;  // 32^-0.5

    // gather: 3 tensors x L rows x 8 segments of 16 B, all in flight at once (cp.async), pad rows zeroed
    {
        // (no runtime divisions: t is an outer loop, row = idx >> 3, segment = idx & 7)
        const float* src0 = qkv + base * (3 * hid) + h * DH;
        const int64_t rstep = row_stride * (3 * hid);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            float* dst = sq + t * LP * QP;
            const float* srct = src0 + t * hid;
            for (int idx = lane; idx < L * 8; idx += 32) {
                const int j = idx >> 3, seg = idx & 7;
                cp_async16(dst + j * QP + seg * 4, srct + (int64_t)j * rstep + seg * 4);
            }
        }
        cp_async_commit();
        for (int idx = lane; idx < 3 * (LP - L) * 8; idx += 32) {
            const int t = idx / ((LP - L) * 8);
            const int rem = idx - t * ((LP - L) * 8);
            const int j = L + (rem >> 3), seg = rem & 7;
            *reinterpret_cast<float4*>(sq + t * LP * QP + j * QP + seg * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_wait<0>();
        __syncwarp();
        // q *= scale; rotary on interleaved pairs (2p, 2p+1): x' = x c - y s, y' = y c + x s   (reference :325-331)
        for (int idx = lane; idx < L * (DH / 2); idx += 32) {
            const int j = idx >> 4, pr = idx & 15;
            float2 q2 = *reinterpret_cast<float2*>(sq + j * QP + 2 * pr);
            q2.x *= scale; q2.y *= scale;
            if (rot_cos) {
                const float c = rot_cos[j * (DH / 2) + pr], sn = rot_sin[j * (DH / 2) + pr];
                float2 k2 = *reinterpret_cast<float2*>(sk + j * QP + 2 * pr);
                const float qx = q2.x * c - q2.y * sn, qy = q2.y * c + q2.x * sn;
                const float kx = k2.x * c - k2.y * sn, ky = k2.y * c + k2.x * sn;
                q2 = make_float2(qx, qy);
                *reinterpret_cast<float2*>(sk + j * QP + 2 * pr) = make_float2(kx, ky);
            }
            *reinterpret_cast<float2*>(sq + j * QP + 2 * pr) = q2;
        }
    }
    __syncwarp();

    const int a = lane >> 3, b = lane & 7;
    const float* pb = pos_bias ? pos_bias + (int64_t)h * L * L : nullptr;
    for (int i0 = 0; i0 < L; i0 += 4 * RIP) {
        // ---- S = Q K^T for rows i0 + a + 4r, columns b + 8c
        float acc[RIP][CJ];
#pragma unroll
        for (int r = 0; r < RIP; ++r)
#pragma unroll
            for (int c = 0; c < CJ; ++c) acc[r][c] = 0.f;
        // rows i0 + a + 4r never exceed LP - 1 when the pass covers exactly 4*RIP rows of an LP-row tile
        constexpr bool ROWS_IN_RANGE = (CJ * 8) % (4 * RIP) == 0;
        const float* sq_a = sq + (i0 + a) * QP;
        const float* sk_b = sk + b * QP;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
            float4 kq[CJ];
#pragma unroll
            for (int c = 0; c < CJ; ++c) kq[c] = *reinterpret_cast<const float4*>(sk_b + 8 * c * QP + d);
#pragma unroll
            for (int r = 0; r < RIP; ++r) {
                float4 qv;
                if (ROWS_IN_RANGE) {
                    qv = *reinterpret_cast<const float4*>(sq_a + 4 * r * QP + d);
                } else {
                    int i = i0 + a + 4 * r;
                    i = i < LP ? i : LP - 1;
                    qv = *reinterpret_cast<const float4*>(sq + i * QP + d);
                }
#pragma unroll
                for (int c = 0; c < CJ; ++c) {
                    acc[r][c] = fmaf(qv.x, kq[c].x, acc[r][c]);
                    acc[r][c] = fmaf(qv.y, kq[c].y, acc[r][c]);
                    acc[r][c] = fmaf(qv.z, kq[c].z, acc[r][c]);
                    acc[r][c] = fmaf(qv.w, kq[c].w, acc[r][c]);
                }
            }
        }
        if (ALIAS_P) __syncwarp();      // every lane is done reading Q/K before P overwrites them
        // ---- bias, row softmax (a row lives in the 8 lanes that share `a`), P -> smem
#pragma unroll
        for (int r = 0; r < RIP; ++r) {
            const int i = i0 + a + 4 * r;
            const bool row_ok = i < L;
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CJ; ++c) {
                const int j = b + 8 * c;
                float v = acc[r][c];
                if (pb && row_ok && j < L) v += pb[i * L + j];
                v = (j < L) ? v : -INFINITY;
                acc[r][c] = v;
                mx = fmaxf(mx, v);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < CJ; ++c) {
                float e = expf(acc[r][c] - mx);      // exp(-inf) = 0 for padded columns
                acc[r][c] = e;
                sum += e;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            const float inv = 1.f / sum;
            if (i < LP) {
#pragma unroll
                for (int c = 0; c < CJ; ++c) sp[i * PP + b + 8 * c] = acc[r][c] * inv;
            }
        }
        __syncwarp();
        // ---- O = P V for rows i0 + a + 4r, columns d = 4b .. 4b+3
        float4 o[RIP];
#pragma unroll
        for (int r = 0; r < RIP; ++r) o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* sp_a = sp + (i0 + a) * PP;
        const float* sv_b = sv + 4 * b;
#pragma unroll
        for (int j = 0; j < LP; j += 4) {
            const float4 v0 = *reinterpret_cast<const float4*>(sv_b + (j + 0) * QP);
            const float4 v1 = *reinterpret_cast<const float4*>(sv_b + (j + 1) * QP);
            const float4 v2 = *reinterpret_cast<const float4*>(sv_b + (j + 2) * QP);
            const float4 v3 = *reinterpret_cast<const float4*>(sv_b + (j + 3) * QP);
#pragma unroll
            for (int r = 0; r < RIP; ++r) {
                float4 p;
                if (ROWS_IN_RANGE) {
                    p = *reinterpret_cast<const float4*>(sp_a + 4 * r * PP + j);
                } else {
                    int i = i0 + a + 4 * r;
                    i = i < LP ? i : LP - 1;
                    p = *reinterpret_cast<const float4*>(sp + i * PP + j);
                }
                o[r].x = fmaf(p.x, v0.x, o[r].x); o[r].y = fmaf(p.x, v0.y, o[r].y); o[r].z = fmaf(p.x, v0.z, o[r].z); o[r].w = fmaf(p.x, v0.w, o[r].w);
                o[r].x = fmaf(p.y, v1.x, o[r].x); o[r].y = fmaf(p.y, v1.y, o[r].y); o[r].z = fmaf(p.y, v1.z, o[r].z); o[r].w = fmaf(p.y, v1.w, o[r].w);
                o[r].x = fmaf(p.z, v2.x, o[r].x); o[r].y = fmaf(p.z, v2.y, o[r].y); o[r].z = fmaf(p.z, v2.z, o[r].z); o[r].w = fmaf(p.z, v2.w, o[r].w);
                o[r].x = fmaf(p.w, v3.x, o[r].x); o[r].y = fmaf(p.w, v3.y, o[r].y); o[r].z = fmaf(p.w, v3.z, o[r].z); o[r].w = fmaf(p.w, v3.w, o[r].w);
            }
        }
#pragma unroll
        for (int r = 0; r < RIP; ++r) {
            const int i = i0 + a + 4 * r;
            if (i < L) {
                const int64_t oi = (base + (int64_t)i * row_stride) * hid + h * DH + 4 * b;
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi) = o[r];
                if (out_sb) store_sb4(out_sb, out_plane, oi, o[r]);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core softmax attention for 17 <= L <= 40 (the temporal attention over F = 40 frames at the C >= 128 levels; the
// C = 64 levels run the fused tcgen05 block of attn_fused.cu): one warp per (sequence, head) on mma.sync.m16n8k16 bf16 with
// the split-bf16 x3 scheme of the conv engine (x = hi + lo: hi.hi + hi.lo + lo.hi, fp32 accumulate); the softmaxed scores
// never leave registers: the C fragments of S = Q K^T are exactly the A fragments of O = P V.
// (History: CUDA-core tiling 0.53 ms -> 3xTF32 m16n8k8 -> bf16 m16n8k16 0.47 -> the ldmatrix kernel below 0.42 ms per
// 32x32 call; the two intermediate kernels were removed in round 2, the numbers are in profiles/r01_*.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int ML = 40;              // padded sequence length of this kernel (5 n-tiles of 8; M padded to 48 = 3 m-tiles)

__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - __low2float(h), y - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma3_bf16(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_bf16(d, al, bh);
    mma_bf16(d, ah, bl);
    mma_bf16(d, ah, bh);
}


// ---------------------------------------------------------------------------------------------------------------
// v2 of the same kernel: 16 resident warps per SM instead of 12 and ~40 % fewer instructions per unit.
//  * K and V land as fp32 (cp.async) and are converted ONCE, in place, to split-bf16 planes (hi | lo, 40 rows x 64 B,
//    16-byte chunks XOR-swizzled by (row >> 1) & 3); rotary is applied to K during that pass.  Every B fragment is
//    then an ldmatrix (.trans for V) of ready-made bf16 pairs: no per-fragment splits, no bank conflicts, no padding.
//  * Q never touches shared memory: the A fragments are float2 loads straight from global (each element belongs to
//    exactly one lane; 8 rows x 32 B per instruction = full sectors), scaled, rotated and split in registers, and the
//    next 16-row tile's Q is prefetched while the current one is processed.
//  * the three 16-row query tiles run one after the other (S: 20, O: 16 accumulators) -> <= 128 registers, 50 KiB of
//    smem per CTA -> 4 CTAs per SM.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KV_PLANE = ML * 64;                        // one bf16 plane: 40 rows x 32 dims
constexpr int OST_PITCH = 40;                            // O staging pitch (floats): conflict-free float2 stores
constexpr int V2_WARP_BYTES = 4 * KV_PLANE + 16 * OST_PITCH * 4;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// byte offset of 16-byte chunk `chunk` (8 dims) of row `row` inside a swizzled plane
__device__ __forceinline__ int kv_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

template <bool FULL>      // FULL: seq_len == 40 (the 40-frame production case): every length test folds away
__global__ void __launch_bounds__(128, 4) attn_softmax_mma16v2_kernel(const float* __restrict__ qkv, bf16* __restrict__ out_sb,
                                                                      int64_t out_plane, float* __restrict__ out_f32,
                                                                      int64_t n_seq, int L_arg, int heads, int64_t inner,
                                                                      int64_t outer_stride, int64_t inner_stride,
                                                                      int64_t row_stride, const float* __restrict__ rot_cos,
                                                                      const float* __restrict__ rot_sin,
                                                                      const float* __restrict__ pos_bias, int q_prescaled) {
    pdl_prologue_done();
    const int L = FULL ? ML : L_arg;
    extern __shared__ __align__(16) uint8_t s_dyn2[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t unit = (int64_t)blockIdx.x * 4 + w;
    if (unit >= n_seq * heads) return;
    uint8_t* wb = s_dyn2 + (size_t)w * V2_WARP_BYTES;
    float* kraw = reinterpret_cast<float*>(wb);                       // [40][32] fp32  -> K_hi | K_lo planes
    float* vraw = reinterpret_cast<float*>(wb + 2 * KV_PLANE);        // [40][32] fp32  -> V_hi | V_lo planes
    float* ost = reinterpret_cast<float*>(wb + 4 * KV_PLANE);         // [16][OST_PITCH] fp32 output staging
    const int64_t s = unit / heads;
    const int h = (int)(unit - s * heads);
    const int hid = heads * DH;
    const int64_t base = (s / inner) * outer_stride + (s % inner) * inner_stride;
    const int64_t rstep = row_stride * (3 * hid);
    const float* src0 = qkv + base * (3 * hid) + h * DH;
    const float scale = q_prescaled ? 1.f : 0.17677669529663687f;  // 32^-0.5 (already applied by the qkv epilogue when fused)
    const int g = lane >> 2, t = lane & 3;

    // ---- K, V: asynchronous fp32 landing
    for (int idx = lane; idx < L * 8; idx += 32) {
        const int j = idx >> 3, seg = idx & 7;
        const float* sr = src0 + (int64_t)j * rstep + seg * 4;
        cp_async16(kraw + j * 32 + seg * 4, sr + hid);
        cp_async16(vraw + j * 32 + seg * 4, sr + 2 * hid);
    }
    cp_async_commit();
    for (int idx = lane; idx < (ML - L) * 8; idx += 32) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(kraw + L * 32 + idx * 4) = z;
        *reinterpret_cast<float4*>(vraw + L * 32 + idx * 4) = z;
    }
    // ---- Q fragments of query tile 0 straight from global: rows (g, g+8), columns 16 ks + 2t (+8)
    float2 qn[2][4];
    auto load_q = [&](int mt) {
        const int r0 = 16 * mt + g, r1 = r0 + 8;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int c = 16 * ks + 2 * t;
            const float2 z = make_float2(0.f, 0.f);
            qn[ks][0] = r0 < L ? *reinterpret_cast<const float2*>(src0 + (int64_t)r0 * rstep + c) : z;
            qn[ks][1] = r1 < L ? *reinterpret_cast<const float2*>(src0 + (int64_t)r1 * rstep + c) : z;
            qn[ks][2] = r0 < L ? *reinterpret_cast<const float2*>(src0 + (int64_t)r0 * rstep + c + 8) : z;
            qn[ks][3] = r1 < L ? *reinterpret_cast<const float2*>(src0 + (int64_t)r1 * rstep + c + 8) : z;
        }
    };
    load_q(0);
    cp_async_wait<0>();
    __syncwarp();

    // ---- one pass: K (rotary) and V -> split-bf16 planes, in place (all reads of a tile precede its first write)
    {
        float4 kr[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) kr[i] = *reinterpret_cast<const float4*>(kraw + (lane + 32 * i) * 4);
        __syncwarp();
        uint8_t* khi = reinterpret_cast<uint8_t*>(kraw);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const int idx = lane + 32 * i, j = idx >> 3, c4 = idx & 7;
            float4 v = kr[i];
            if (rot_cos && j < L) {
                const float2 cc = *reinterpret_cast<const float2*>(rot_cos + j * (DH / 2) + 2 * c4);   // pairs 2 c4, 2 c4 + 1
                const float2 sn = *reinterpret_cast<const float2*>(rot_sin + j * (DH / 2) + 2 * c4);
                v = make_float4(v.x * cc.x - v.y * sn.x, v.y * cc.x + v.x * sn.x, v.z * cc.y - v.w * sn.y, v.w * cc.y + v.z * sn.y);
            }
            uint2 hv, lv;
            split_bf16x2(v.x, v.y, hv.x, lv.x);
            split_bf16x2(v.z, v.w, hv.y, lv.y);
            const int off = kv_off(j, c4 >> 1) + ((c4 & 1) << 3);
            *reinterpret_cast<uint2*>(khi + off) = hv;
            *reinterpret_cast<uint2*>(khi + KV_PLANE + off) = lv;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) kr[i] = *reinterpret_cast<const float4*>(vraw + (lane + 32 * i) * 4);
        __syncwarp();
        uint8_t* vhi = reinterpret_cast<uint8_t*>(vraw);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const int idx = lane + 32 * i, j = idx >> 3, c4 = idx & 7;
            uint2 hv, lv;
            split_bf16x2(kr[i].x, kr[i].y, hv.x, lv.x);
            split_bf16x2(kr[i].z, kr[i].w, hv.y, lv.y);
            const int off = kv_off(j, c4 >> 1) + ((c4 & 1) << 3);
            *reinterpret_cast<uint2*>(vhi + off) = hv;
            *reinterpret_cast<uint2*>(vhi + KV_PLANE + off) = lv;
        }
    }
    __syncwarp();

    const uint32_t k_sa = smem_addr_u32(kraw), v_sa = smem_addr_u32(vraw);
    const int lt = lane >> 3, lr = lane & 7;                 // ldmatrix: this lane supplies row lr of tile lt
    const float* pb = pos_bias ? pos_bias + (int64_t)h * L * L : nullptr;
    const bool pair_ok = (L & 1) == 0 && (reinterpret_cast<uintptr_t>(pos_bias) & 7) == 0;

#pragma unroll 1
    for (int mt = 0; mt < 3; ++mt) {
        if (16 * mt >= L) break;
        const int i0 = 16 * mt + g, i1 = i0 + 8;
        // ---- A fragments of this query tile: scale, rotary (reference :325-331), split
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 q2 = qn[ks][e];
                q2.x *= scale; q2.y *= scale;
                if (rot_cos) {
                    const int row = (e & 1) ? i1 : i0;
                    if (row < L) {
                        const int pr = 8 * ks + t + ((e >> 1) << 2);
                        const float c = rot_cos[row * (DH / 2) + pr], sn = rot_sin[row * (DH / 2) + pr];
                        q2 = make_float2(q2.x * c - q2.y * sn, q2.y * c + q2.x * sn);
                    }
                }
                split_bf16x2(q2.x, q2.y, ah[ks][e], al[ks][e]);
            }
        if (16 * (mt + 1) < L) load_q(mt + 1);               // prefetch the next tile's Q behind this tile's math

        // ---- S = Q K^T : B fragments by ldmatrix from the K planes (tiles: 8 keys x 8 dims)
        float acc[5][4];
#pragma unroll
        for (int nt = 0; nt < 5; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int np = 0; np < 2; ++np) {                 // key tiles (2np, 2np+1): regs {b0, b1} of each
                const int key = 8 * (2 * np + (lt >> 1)) + lr, chunk = 2 * ks + (lt & 1);
                uint32_t kh4[4], kl4[4];
                ldmatrix_x4(kh4, k_sa + kv_off(key, chunk));
                ldmatrix_x4(kl4, k_sa + KV_PLANE + kv_off(key, chunk));
                const uint32_t bh0[2] = {kh4[0], kh4[1]}, bl0[2] = {kl4[0], kl4[1]};
                const uint32_t bh1[2] = {kh4[2], kh4[3]}, bl1[2] = {kl4[2], kl4[3]};
                mma3_bf16(acc[2 * np], ah[ks], al[ks], bh0, bl0);
                mma3_bf16(acc[2 * np + 1], ah[ks], al[ks], bh1, bl1);
            }
            {
                const int key = 32 + lr, chunk = 2 * ks + (lt & 1);      // key tile 4: lanes 0-15 supply the addresses
                uint32_t kh2[2], kl2[2];
                ldmatrix_x2(kh2, k_sa + kv_off(key, chunk));
                ldmatrix_x2(kl2, k_sa + KV_PLANE + kv_off(key, chunk));
                mma3_bf16(acc[4], ah[ks], al[ks], kh2, kl2);
            }
        }
        // ---- bias + row softmax in registers (a row lives in the 4 lanes that share g)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = half ? i1 : i0;
            const bool row_ok = i < L;
            if (!row_ok) continue;                 // rows >= L of the last tile: never stored, their P fragments only feed dead O rows
            const unsigned gmask = 0xFu << (lane & ~3);    // the 4 lanes that share this row (row_ok is uniform inside the group)
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                const int j0 = 8 * nt + 2 * t;
                float2 b2 = make_float2(0.f, 0.f);
                if (pb && row_ok) {
                    if (pair_ok) { if (j0 < L) b2 = *reinterpret_cast<const float2*>(pb + i * L + j0); }
                    else { if (j0 < L) b2.x = pb[i * L + j0]; if (j0 + 1 < L) b2.y = pb[i * L + j0 + 1]; }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v = acc[nt][2 * half + e] + (e ? b2.y : b2.x);
                    v = (j0 + e < L) ? v : -INFINITY;
                    acc[nt][2 * half + e] = v;
                    mx = fmaxf(mx, v);
                }
            }
            mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 2));
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 5; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float ex = __expf(acc[nt][2 * half + e] - mx);
                    acc[nt][2 * half + e] = ex;
                    sum += ex;
                }
            sum += __shfl_xor_sync(gmask, sum, 1);
            sum += __shfl_xor_sync(gmask, sum, 2);
            const float inv = 1.f / sum;
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) { acc[nt][2 * half] *= inv; acc[nt][2 * half + 1] *= inv; }
        }
        // ---- O = P V : A fragments from the C fragments of S; B fragments by ldmatrix.trans from the V planes
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            uint32_t ph[4], pl[4];
            split_bf16x2(acc[2 * ks][0], acc[2 * ks][1], ph[0], pl[0]);
            split_bf16x2(acc[2 * ks][2], acc[2 * ks][3], ph[1], pl[1]);
            if (2 * ks + 1 < 5) {
                split_bf16x2(acc[2 * ks + 1][0], acc[2 * ks + 1][1], ph[2], pl[2]);
                split_bf16x2(acc[2 * ks + 1][2], acc[2 * ks + 1][3], ph[3], pl[3]);
            } else {
                ph[2] = pl[2] = ph[3] = pl[3] = 0u;
            }
            if (ks < 2) {
#pragma unroll
                for (int np = 0; np < 2; ++np) {             // dim tiles (2np, 2np+1): tiles {keys 16ks.., keys 16ks+8..} of each
                    const int key = 16 * ks + 8 * (lt & 1) + lr, chunk = 2 * np + (lt >> 1);
                    uint32_t vh4[4], vl4[4];
                    ldmatrix_x4_trans(vh4, v_sa + kv_off(key, chunk));
                    ldmatrix_x4_trans(vl4, v_sa + KV_PLANE + kv_off(key, chunk));
                    const uint32_t bh0[2] = {vh4[0], vh4[1]}, bl0[2] = {vl4[0], vl4[1]};
                    const uint32_t bh1[2] = {vh4[2], vh4[3]}, bl1[2] = {vl4[2], vl4[3]};
                    mma3_bf16(o[2 * np], ph, pl, bh0, bl0);
                    mma3_bf16(o[2 * np + 1], ph, pl, bh1, bl1);
                }
            } else {                                         // keys 32..39 only (40..47 do not exist: zero B halves)
                const int key = 32 + lr, chunk = lt;         // four dim tiles of the same 8 keys
                uint32_t vh4[4], vl4[4];
                ldmatrix_x4_trans(vh4, v_sa + kv_off(key, chunk));
                ldmatrix_x4_trans(vl4, v_sa + KV_PLANE + kv_off(key, chunk));
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const uint32_t bh[2] = {vh4[nt], 0u}, bl[2] = {vl4[nt], 0u};
                    mma3_bf16(o[nt], ph, pl, bh, bl);
                }
            }
        }
        // ---- stage the 16 x 32 output tile, then coalesced row stores
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<float2*>(ost + g * OST_PITCH + 8 * nt + 2 * t) = make_float2(o[nt][0], o[nt][1]);
            *reinterpret_cast<float2*>(ost + (g + 8) * OST_PITCH + 8 * nt + 2 * t) = make_float2(o[nt][2], o[nt][3]);
        }
        __syncwarp();
        for (int idx = lane; idx < 16 * 8; idx += 32) {
            const int r = idx >> 3, seg = idx & 7, i = 16 * mt + r;
            if (i < L) {
                const float4 v = *reinterpret_cast<const float4*>(ost + r * OST_PITCH + seg * 4);
                const int64_t oi = (base + (int64_t)i * row_stride) * hid + h * DH + seg * 4;
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi) = v;
                if (out_sb) store_sb4(out_sb, out_plane, oi, v);
            }
        }
        __syncwarp();
    }
}

inline int v2_set_smem(size_t smem) {
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(attn_softmax_mma16v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_softmax_mma16v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        once.mark();
    }
    return 0;
}

// one block (4 warps) per (frame, head).  Every warp streams 32-row chunks of (k, v) and later q through a
// double-buffered cp.async pipeline (all bytes of the next chunk are in flight while the current one is consumed);
// exp / softmax of a row is computed once and kept in the smem tile; the 32x32 context and its application to q are
// register-tiled: lane (a = lane/8, b = lane%8) owns ctx[d = 8a..8a+7][e = 4b..4b+3].
constexpr int LW = 4;
constexpr int TP = DH + 4;                               // tile row pitch (floats), rows 16-byte aligned
constexpr int TILE = 32 * TP;                            // one 32-row tile

__device__ __forceinline__ void issue_rows(float* dst, const float* src_col, int64_t ld, int n0, int n_pos, int lane) {
    // 32 rows x 8 segments of 16 B
    for (int idx = lane; idx < 32 * 8; idx += 32) {
        const int r = idx >> 3, seg = idx & 7;
        if (n0 + r < n_pos) cp_async16(dst + r * TP + seg * 4, src_col + (int64_t)(n0 + r) * ld + seg * 4);
    }
}

template <bool USE_MMA>
__global__ void __launch_bounds__(32 * LW, 3) attn_linear_kernel(const float* __restrict__ qkv, bf16* __restrict__ out_sb,
                                                                 int64_t out_plane, float* __restrict__ out_f32, int n_pos,
                                                                 int heads) {
    pdl_prologue_done();
    extern __shared__ __align__(16) float s_dynl[];      // [LW][2 stages][2 tiles][32][TP]; aliased by the reduction
    __shared__ float s_red[LW][DH];
    __shared__ float s_max[DH];
    __shared__ float s_rmax[LW][DH];                     // USE_MMA: per-warp running column maxima of k (online softmax)
    float (*s_ctxn)[DH + 4] = reinterpret_cast<float (*)[DH + 4]>(s_dynl + LW * DH * DH);   // aliases tiles (after phase 2)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = lane >> 3, b = lane & 7;
    const int64_t fr = blockIdx.x / heads;
    const int h = blockIdx.x % heads;
    const int hid = heads * DH;
    const int64_t ld = 3 * hid;
    const float* base = qkv + fr * n_pos * ld + h * DH;
    const float scale = 0.17677669529663687f;
    float* wbuf = s_dynl + w * 4 * TILE;                  // stage s: k/q tile at wbuf + s*2*TILE, v tile at + TILE

    // phase 1 (CUDA-core variant only): column max of k over positions (lane = d).  The tensor-core variant keeps a running
    // maximum per warp instead (online softmax: accumulators are rescaled when the maximum moves), which saves this whole
    // extra pass over k (a quarter of the kernel's DRAM reads) and its exposed latency.
    float kmax = -INFINITY;
    if constexpr (!USE_MMA) {
        float mx = -INFINITY;
#pragma unroll 16
        for (int n = w; n < n_pos; n += LW) mx = fmaxf(mx, base[(int64_t)n * ld + hid + lane]);
        s_red[w][lane] = mx;
        __syncthreads();
        if (w == 0) {
            float m = s_red[0][lane];
#pragma unroll
            for (int i = 1; i < LW; ++i) m = fmaxf(m, s_red[i][lane]);
            s_max[lane] = m;
        }
        __syncthreads();
        kmax = s_max[lane];
    }

    // phase 2: ctx[d][e] += exp(k[n][d] - max_d) * v[n][e];  den[d] += exp(..)
    float ctx[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ctx[i][j] = 0.f;
    float cacc[2][4][4];                                  // mma C fragments of the context (USE_MMA)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) cacc[mt][nt][e] = 0.f;
    float den = 0.f;
    const int n_chunks = (n_pos + 31) / 32;
    if (w < n_chunks) {
        issue_rows(wbuf, base + hid, ld, w * 32, n_pos, lane);
        issue_rows(wbuf + TILE, base + 2 * hid, ld, w * 32, n_pos, lane);
    }
    cp_async_commit();
    int st = 0;
    for (int c = w; c < n_chunks; c += LW, st ^= 1) {
        const int cn = c + LW;
        if (cn < n_chunks) {
            issue_rows(wbuf + (st ^ 1) * 2 * TILE, base + hid, ld, cn * 32, n_pos, lane);
            issue_rows(wbuf + (st ^ 1) * 2 * TILE + TILE, base + 2 * hid, ld, cn * 32, n_pos, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        float* kt = wbuf + st * 2 * TILE;
        const float* vt = kt + TILE;
        const int n0 = c * 32;
        const int rows = min(32, n_pos - n0);
        if constexpr (USE_MMA) {
            float cm = -INFINITY;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) cm = fmaxf(cm, (r < rows) ? kt[r * TP + lane] : -INFINITY);
            const float nm = fmaxf(kmax, cm);
            const float f = __expf(kmax - nm);            // first chunk: exp(-inf) = 0 on all-zero accumulators
            kmax = nm;
            den *= f;
            const int g = lane >> 2;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float f0 = __shfl_sync(0xffffffffu, f, g + 16 * mt), f1 = __shfl_sync(0xffffffffu, f, g + 16 * mt + 8);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    cacc[mt][nt][0] *= f0; cacc[mt][nt][1] *= f0; cacc[mt][nt][2] *= f1; cacc[mt][nt][3] *= f1;
                }
            }
        }
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float e = (r < rows) ? __expf(kt[r * TP + lane] - kmax) : 0.f;
            den += e;
            kt[r * TP + lane] = e;
        }
        __syncwarp();
        if constexpr (USE_MMA) {
            // ctx (32 x 32) += ek^T (32 x 32 rows) . v : split-bf16 mma.m16n8k16, A(m = d, k = n) = ek[n][d], B(k = n, n = e) = v[n][e]
            float* vtw = const_cast<float*>(vt);
            for (int r = rows; r < 32; ++r) vtw[r * TP + lane] = 0.f;      // ragged last chunk: no stale rows in the product
            __syncwarp();
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int n_lo = 16 * ks + 2 * t;
                uint32_t bh[4][2], bl[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int e = g + 8 * nt;
                    split_bf16x2(vt[n_lo * TP + e], vt[(n_lo + 1) * TP + e], bh[nt][0], bl[nt][0]);
                    split_bf16x2(vt[(n_lo + 8) * TP + e], vt[(n_lo + 9) * TP + e], bh[nt][1], bl[nt][1]);
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int d0 = g + 16 * mt;
                    uint32_t ah[4], al[4];
                    split_bf16x2(kt[n_lo * TP + d0], kt[(n_lo + 1) * TP + d0], ah[0], al[0]);
                    split_bf16x2(kt[n_lo * TP + d0 + 8], kt[(n_lo + 1) * TP + d0 + 8], ah[1], al[1]);
                    split_bf16x2(kt[(n_lo + 8) * TP + d0], kt[(n_lo + 9) * TP + d0], ah[2], al[2]);
                    split_bf16x2(kt[(n_lo + 8) * TP + d0 + 8], kt[(n_lo + 9) * TP + d0 + 8], ah[3], al[3]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma3_bf16(cacc[mt][nt], ah, al, bh[nt], bl[nt]);
                }
            }
        } else {
#pragma unroll 4
            for (int r = 0; r < rows; ++r) {
                const float4 e0 = *reinterpret_cast<const float4*>(kt + r * TP + 8 * a);
                const float4 e1 = *reinterpret_cast<const float4*>(kt + r * TP + 8 * a + 4);
                const float4 v = *reinterpret_cast<const float4*>(vt + r * TP + 4 * b);
                const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    ctx[i][0] = fmaf(ev[i], v.x, ctx[i][0]); ctx[i][1] = fmaf(ev[i], v.y, ctx[i][1]);
                    ctx[i][2] = fmaf(ev[i], v.z, ctx[i][2]); ctx[i][3] = fmaf(ev[i], v.w, ctx[i][3]);
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncthreads();                                     // all tiles consumed: alias the buffer as [LW][32][32]
    float* s_ctx = s_dynl;
    if constexpr (USE_MMA) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<float2*>(s_ctx + (w * DH + g + 16 * mt) * DH + 8 * nt + 2 * t) = make_float2(cacc[mt][nt][0], cacc[mt][nt][1]);
                *reinterpret_cast<float2*>(s_ctx + (w * DH + g + 16 * mt + 8) * DH + 8 * nt + 2 * t) = make_float2(cacc[mt][nt][2], cacc[mt][nt][3]);
            }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(s_ctx + (w * DH + 8 * a + i) * DH + 4 * b) = make_float4(ctx[i][0], ctx[i][1], ctx[i][2], ctx[i][3]);
    }
    s_red[w][lane] = den;
    s_rmax[w][lane] = kmax;
    __syncthreads();
    for (int i = threadIdx.x; i < DH * DH; i += 32 * LW) {
        const int d = i / DH, e = i % DH;
        float acc = 0.f, dd = 0.f;
        if constexpr (USE_MMA) {                         // warps hold sums relative to their own maxima: bring them to the common one
            float m = s_rmax[0][d];
#pragma unroll
            for (int ww = 1; ww < LW; ++ww) m = fmaxf(m, s_rmax[ww][d]);
#pragma unroll
            for (int ww = 0; ww < LW; ++ww) {
                const float sc = __expf(s_rmax[ww][d] - m);          // a warp without chunks: exp(-inf) = 0
                acc = fmaf(s_ctx[(ww * DH + d) * DH + e], sc, acc);
                dd = fmaf(s_red[ww][d], sc, dd);
            }
        } else {
#pragma unroll
            for (int ww = 0; ww < LW; ++ww) { acc += s_ctx[(ww * DH + d) * DH + e]; dd += s_red[ww][d]; }
        }
        s_ctxn[d][e] = acc / dd;                         // softmax normalisation of k folded into the context
    }
    __syncthreads();
    uint32_t pbh[2][4][2], pbl[2][4][2];                  // B fragments of ctxn (k = d, n = e), kept for the whole phase 3
    if constexpr (USE_MMA) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int d0 = 16 * ks + 2 * t, e = g + 8 * nt;
                split_bf16x2(s_ctxn[d0][e], s_ctxn[d0 + 1][e], pbh[ks][nt][0], pbl[ks][nt][0]);
                split_bf16x2(s_ctxn[d0 + 8][e], s_ctxn[d0 + 9][e], pbh[ks][nt][1], pbl[ks][nt][1]);
            }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(&s_ctxn[8 * a + i][4 * b]);
            ctx[i][0] = t.x; ctx[i][1] = t.y; ctx[i][2] = t.z; ctx[i][3] = t.w;
        }
    }
    __syncthreads();                                     // s_ctxn (aliased) fully consumed before q tiles land on it

    // phase 3: out[n][e] = sum_d ctxn[d][e] * softmax_d(q[n])[d] * scale
    if (w < n_chunks) issue_rows(wbuf, base, ld, w * 32, n_pos, lane);
    cp_async_commit();
    st = 0;
    for (int c = w; c < n_chunks; c += LW, st ^= 1) {
        const int cn = c + LW;
        if (cn < n_chunks) issue_rows(wbuf + (st ^ 1) * 2 * TILE, base, ld, cn * 32, n_pos, lane);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        float* qt = wbuf + st * 2 * TILE;
        const int n0 = c * 32;
        if (n0 + lane < n_pos) {                          // one thread per row: softmax over d, * scale
            float4 x[8];
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = *reinterpret_cast<const float4*>(qt + lane * TP + 4 * i);
                m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i].x = __expf(x[i].x - m); x[i].y = __expf(x[i].y - m); x[i].z = __expf(x[i].z - m); x[i].w = __expf(x[i].w - m);
                sum += (x[i].x + x[i].y) + (x[i].z + x[i].w);
            }
            const float inv = scale / sum;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(qt + lane * TP + 4 * i) = make_float4(x[i].x * inv, x[i].y * inv, x[i].z * inv, x[i].w * inv);
        }
        __syncwarp();
        const int rows = min(32, n_pos - n0);
        if constexpr (USE_MMA) {
            // out (32 rows x 32) = qs (32 x 32) . ctxn : A(m = n, k = d) from the tile, B fragments resident in registers
            const int g = lane >> 2, t = lane & 3;
            float oacc[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) oacc[mt][nt][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int r0 = g + 16 * mt, dcol = 16 * ks + 2 * t;
                    const float2 x0 = *reinterpret_cast<const float2*>(qt + r0 * TP + dcol);
                    const float2 x1 = *reinterpret_cast<const float2*>(qt + (r0 + 8) * TP + dcol);
                    const float2 x2 = *reinterpret_cast<const float2*>(qt + r0 * TP + dcol + 8);
                    const float2 x3 = *reinterpret_cast<const float2*>(qt + (r0 + 8) * TP + dcol + 8);
                    uint32_t ah[4], al[4];
                    split_bf16x2(x0.x, x0.y, ah[0], al[0]);
                    split_bf16x2(x1.x, x1.y, ah[1], al[1]);
                    split_bf16x2(x2.x, x2.y, ah[2], al[2]);
                    split_bf16x2(x3.x, x3.y, ah[3], al[3]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma3_bf16(oacc[mt][nt], ah, al, pbh[ks][nt], pbl[ks][nt]);
                }
            __syncwarp();                                  // all lanes done reading qs: stage the outputs in the same tile
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    *reinterpret_cast<float2*>(qt + (g + 16 * mt) * TP + 8 * nt + 2 * t) = make_float2(oacc[mt][nt][0], oacc[mt][nt][1]);
                    *reinterpret_cast<float2*>(qt + (g + 16 * mt + 8) * TP + 8 * nt + 2 * t) = make_float2(oacc[mt][nt][2], oacc[mt][nt][3]);
                }
            __syncwarp();
            for (int idx = lane; idx < rows * 8; idx += 32) {
                const int r = idx >> 3, seg = idx & 7;
                const float4 o = *reinterpret_cast<const float4*>(qt + r * TP + seg * 4);
                const int64_t oi = (fr * n_pos + n0 + r) * hid + h * DH + seg * 4;
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi) = o;
                if (out_sb) store_sb4(out_sb, out_plane, oi, o);
            }
        } else {
#pragma unroll 2
            for (int r = 0; r < rows; ++r) {
                const float4 q0 = *reinterpret_cast<const float4*>(qt + r * TP + 8 * a);
                const float4 q1 = *reinterpret_cast<const float4*>(qt + r * TP + 8 * a + 4);
                const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    o.x = fmaf(ctx[i][0], qv[i], o.x); o.y = fmaf(ctx[i][1], qv[i], o.y);
                    o.z = fmaf(ctx[i][2], qv[i], o.z); o.w = fmaf(ctx[i][3], qv[i], o.w);
                }
                // reduce over the four d-blocks (lanes that share b)
                o.x += __shfl_xor_sync(0xffffffffu, o.x, 8);  o.y += __shfl_xor_sync(0xffffffffu, o.y, 8);
                o.z += __shfl_xor_sync(0xffffffffu, o.z, 8);  o.w += __shfl_xor_sync(0xffffffffu, o.w, 8);
                o.x += __shfl_xor_sync(0xffffffffu, o.x, 16); o.y += __shfl_xor_sync(0xffffffffu, o.y, 16);
                o.z += __shfl_xor_sync(0xffffffffu, o.z, 16); o.w += __shfl_xor_sync(0xffffffffu, o.w, 16);
                if (a == 0) {
                    const int64_t oi = (fr * n_pos + n0 + r) * hid + h * DH + 4 * b;
                    if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi) = o;
                    if (out_sb) store_sb4(out_sb, out_plane, oi, o);
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

}  // namespace

template <int RIP, int CJ, bool ALIAS_P>
static int launch_attn_softmax(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_seq, int seq_len,
                               int heads, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t row_stride,
                               const float* rot_cos, const float* rot_sin, const float* pos_bias, cudaStream_t st) {
    const int LP = CJ * 8;
    static_assert(!ALIAS_P || (LP * (LP + 4) <= 2 * LP * QP), "P must fit in the Q+K tiles");
    const size_t per_warp = sizeof(float) * (size_t)(3 * LP * QP + (ALIAS_P ? 0 : LP * (LP + 4)));
    int wpb = (int)(100 * 1024 / per_warp);
    if (wpb > 4) wpb = 4;
    if (wpb < 1) wpb = 1;
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(attn_softmax_kernel<RIP, CJ, ALIAS_P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(per_warp * 4 > 200 * 1024 ? 200 * 1024 : per_warp * 4));
        if (e != cudaSuccess) return (int)e;
        once.mark();
    }
    const int64_t units = n_seq * heads;
    const int64_t blocks = (units + wpb - 1) / wpb;
    LFDM_LAUNCH_PDL((attn_softmax_kernel<RIP, CJ, ALIAS_P>), dim3((unsigned)blocks), dim3(32 * wpb), per_warp * wpb, st,
                    qkv, (bf16*)out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride, row_stride,
                    rot_cos, rot_sin, pos_bias, wpb);
    return 0;
}

extern "C" int lfdm_attn_softmax(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_seq,
                                 int seq_len, int heads, int64_t inner, int64_t outer_stride, int64_t inner_stride,
                                 int64_t row_stride, const float* rot_cos, const float* rot_sin, const float* pos_bias,
                                 void* stream) {
    if (!qkv || seq_len <= 0 || seq_len > MAXL || heads <= 0 || inner <= 0) return LFDM_E_BADARG;
    if ((rot_cos == nullptr) != (rot_sin == nullptr)) return LFDM_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
#define LFDM_ATTN_ARGS qkv, out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride, row_stride, rot_cos, rot_sin, pos_bias, st
    if (seq_len <= 16) return launch_attn_softmax<4, 2, true>(LFDM_ATTN_ARGS);
    if (seq_len <= 40) {
        static const bool use_mma = (getenv("LFDM_ATTN_SIMT") == nullptr);      // A/B switch: CUDA-core tiling instead
        if (use_mma && seq_len >= 17) {
            const size_t smem = (size_t)4 * V2_WARP_BYTES;
            int rc = v2_set_smem(smem);
            if (rc) return rc;
            const int64_t units = n_seq * heads;
            if (seq_len == ML)
                LFDM_LAUNCH_PDL(attn_softmax_mma16v2_kernel<true>, dim3((unsigned)((units + 3) / 4)), dim3(128), smem, st,
                                qkv, (bf16*)out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride,
                                row_stride, rot_cos, rot_sin, pos_bias, 0);
            else
                LFDM_LAUNCH_PDL(attn_softmax_mma16v2_kernel<false>, dim3((unsigned)((units + 3) / 4)), dim3(128), smem, st,
                                qkv, (bf16*)out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride,
                                row_stride, rot_cos, rot_sin, pos_bias, 0);
            return 0;
        }
        return launch_attn_softmax<10, 5, true>(LFDM_ATTN_ARGS);
    }
    return launch_attn_softmax<8, 8, false>(LFDM_ATTN_ARGS);        // L <= 64: two row passes of 32
#undef LFDM_ATTN_ARGS
}

extern "C" int lfdm_attn_softmax_pre(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_seq,
                                     int seq_len, int heads, int64_t inner, int64_t outer_stride, int64_t inner_stride,
                                     int64_t row_stride, const float* pos_bias, void* stream) {
    if (!qkv || heads <= 0 || inner <= 0) return LFDM_E_BADARG;
    if (seq_len < 17 || seq_len > ML) return LFDM_E_UNSUPP;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)4 * V2_WARP_BYTES;
    int rc = v2_set_smem(smem);
    if (rc) return rc;
    const int64_t units = n_seq * heads;
    if (seq_len == ML)
        LFDM_LAUNCH_PDL(attn_softmax_mma16v2_kernel<true>, dim3((unsigned)((units + 3) / 4)), dim3(128), smem, st,
                        qkv, (bf16*)out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride,
                        row_stride, (const float*)nullptr, (const float*)nullptr, pos_bias, 1);
    else
        LFDM_LAUNCH_PDL(attn_softmax_mma16v2_kernel<false>, dim3((unsigned)((units + 3) / 4)), dim3(128), smem, st,
                        qkv, (bf16*)out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride,
                        row_stride, (const float*)nullptr, (const float*)nullptr, pos_bias, 1);
    return 0;
}

extern "C" int lfdm_attn_linear(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_frames,
                                int n_pos, int heads, void* stream) {
    if (!qkv || n_pos <= 0 || heads <= 0) return LFDM_E_BADARG;
    const size_t smem = sizeof(float) * LW * 4 * TILE;
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(attn_linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        once.mark();
    }
    static const bool use_mma = (getenv("LFDM_ATTN_SIMT") == nullptr);          // A/B switch: CUDA-core tiling instead
    if (use_mma)
        LFDM_LAUNCH_PDL(attn_linear_kernel<true>, dim3((unsigned)(n_frames * heads)), dim3(32 * LW), smem, (cudaStream_t)stream,
                        qkv, (bf16*)out_sb, out_plane, out_f32, n_pos, heads);
    else
        LFDM_LAUNCH_PDL(attn_linear_kernel<false>, dim3((unsigned)(n_frames * heads)), dim3(32 * LW), smem, (cudaStream_t)stream,
                        qkv, (bf16*)out_sb, out_plane, out_f32, n_pos, heads);
    return 0;
}
