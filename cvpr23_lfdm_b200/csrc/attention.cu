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
#include "common.cuh"

namespace {

constexpr int DH = 32;
constexpr int MAXL = 64;

// one warp per (sequence, head); 4 warps per block
__global__ void __launch_bounds__(128) attn_softmax_kernel(const float* __restrict__ qkv, bf16* __restrict__ out_sb,
                                                           int64_t out_plane, float* __restrict__ out_f32,
                                                           int64_t n_seq, int L, int heads, int64_t inner,
                                                           int64_t outer_stride, int64_t inner_stride,
                                                           int64_t row_stride, const float* __restrict__ rot_cos,
                                                           const float* __restrict__ rot_sin,
                                                           const float* __restrict__ pos_bias) {
    extern __shared__ float s_dyn[];   // [4 warps][3][L][DH+1]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float (*sq)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(s_dyn + (size_t)(w * 3 + 0) * L * (DH + 1));
    float (*sk)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(s_dyn + (size_t)(w * 3 + 1) * L * (DH + 1));
    float (*sv)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(s_dyn + (size_t)(w * 3 + 2) * L * (DH + 1));
    const int64_t unit = (int64_t)blockIdx.x * 4 + w;
    if (unit >= n_seq * heads) return;   // whole warp exits together (no block-level sync below)
    const int64_t s = unit / heads;
    const int h = (int)(unit - s * heads);
    const int hid = heads * DH;
    const int64_t base = (s / inner) * outer_stride + (s % inner) * inner_stride;
    const float scale = 0.17677669529663687f;  // 32^-0.5 (python float dim_head ** -0.5 rounded to fp32)

    // load q (scaled, rotated), k (rotated), v: lane = d
    for (int j = 0; j < L; ++j) {
        const float* r = qkv + (base + (int64_t)j * row_stride) * (3 * hid) + h * DH;
        float q = r[lane] * scale, k = r[hid + lane], v = r[2 * hid + lane];
        if (rot_cos) {
            // interleaved pairs (2p, 2p+1): rot(x)[2p] = -x[2p+1], rot(x)[2p+1] = x[2p]
            float c = rot_cos[j * (DH / 2) + (lane >> 1)], sn = rot_sin[j * (DH / 2) + (lane >> 1)];
            float qo = __shfl_xor_sync(0xffffffffu, q, 1), ko = __shfl_xor_sync(0xffffffffu, k, 1);
            float qr = (lane & 1) ? qo : -qo, kr = (lane & 1) ? ko : -ko;
            q = q * c + qr * sn;
            k = k * c + kr * sn;
        }
        sq[j][lane] = q; sk[j][lane] = k; sv[j][lane] = v;
    }
    __syncwarp();

    const float* pb = pos_bias ? pos_bias + (int64_t)h * L * L : nullptr;
    for (int i = 0; i < L; ++i) {
        // scores for columns lane and lane+32
        float s0 = -INFINITY, s1 = -INFINITY;
        if (lane < L) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) a = fmaf(sq[i][d], sk[lane][d], a);
            if (pb) a += pb[i * L + lane];
            s0 = a;
        }
        if (lane + 32 < L) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) a = fmaf(sq[i][d], sk[lane + 32][d], a);
            if (pb) a += pb[i * L + lane + 32];
            s1 = a;
        }
        float mx = warp_max(fmaxf(s0, s1));
        float p0 = (lane < L) ? expf(s0 - mx) : 0.f;
        float p1 = (lane + 32 < L) ? expf(s1 - mx) : 0.f;
        float den = warp_sum(p0 + p1);
        float inv = 1.f / den;
        p0 *= inv; p1 *= inv;
        float o = 0.f;
        const int l0 = L < 32 ? L : 32;
        for (int j = 0; j < l0; ++j) o = fmaf(__shfl_sync(0xffffffffu, p0, j), sv[j][lane], o);
        for (int j = 32; j < L; ++j) o = fmaf(__shfl_sync(0xffffffffu, p1, j - 32), sv[j][lane], o);
        int64_t orow = base + (int64_t)i * row_stride;
        int64_t oi = orow * hid + h * DH + lane;
        if (out_f32) out_f32[oi] = o;
        if (out_sb) store_sb1(out_sb, out_plane, oi, o);
    }
}

// one block (256 threads, 8 warps) per (frame, head)
__global__ void __launch_bounds__(256, 3) attn_linear_kernel(const float* __restrict__ qkv, bf16* __restrict__ out_sb,
                                                          int64_t out_plane, float* __restrict__ out_f32, int n_pos,
                                                          int heads) {
    __shared__ float s_red[8][DH];
    __shared__ float s_max[DH];
    __shared__ float s_ctx[8][DH][DH + 1];
    __shared__ float s_ctxn[DH][DH + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t fr = blockIdx.x / heads;
    const int h = blockIdx.x % heads;
    const int hid = heads * DH;
    const float* base = qkv + fr * n_pos * (int64_t)(3 * hid) + h * DH;
    const float scale = 0.17677669529663687f;

    // phase 1: column max of k over positions (lane = d)
    float mx = -INFINITY;
    for (int n = w; n < n_pos; n += 8) mx = fmaxf(mx, base[(int64_t)n * 3 * hid + hid + lane]);
    s_red[w][lane] = mx;
    __syncthreads();
    if (w == 0) {
        float m = s_red[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i][lane]);
        s_max[lane] = m;
    }
    __syncthreads();
    const float kmax = s_max[lane];

    // phase 2: ctx[d][e] = sum_n exp(k[n][d]-max_d) v[n][e];  den[d] = sum_n exp(k[n][d]-max_d)
    float ctx[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) ctx[d] = 0.f;
    float den = 0.f;
    for (int n = w; n < n_pos; n += 8) {
        const float* r = base + (int64_t)n * 3 * hid;
        float ek = expf(r[hid + lane] - kmax);
        float v = r[2 * hid + lane];
        den += ek;
#pragma unroll
        for (int d = 0; d < DH; ++d) ctx[d] = fmaf(__shfl_sync(0xffffffffu, ek, d), v, ctx[d]);
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) s_ctx[w][d][lane] = ctx[d];
    s_red[w][lane] = den;
    __syncthreads();
    // reduce across the 8 warps: thread t handles entries t, t+256, ... of the 32x32 ctx
    for (int i = threadIdx.x; i < DH * DH; i += 256) {
        int d = i / DH, e = i % DH;
        float a = 0.f, dd = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) { a += s_ctx[ww][d][e]; dd += s_red[ww][d]; }
        s_ctxn[d][e] = a / dd;   // softmax normalisation of k folded into the context
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < DH; ++d) ctx[d] = s_ctxn[d][lane];

    // phase 3: out[n][e] = sum_d ctxn[d][e] * softmax_d(q[n])[d] * scale
    for (int n = w; n < n_pos; n += 8) {
        float q = base[(int64_t)n * 3 * hid + lane];
        float m = warp_max(q);
        float eq = expf(q - m);
        float s = warp_sum(eq);
        float qs = eq / s * scale;
        float o = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) o = fmaf(ctx[d], __shfl_sync(0xffffffffu, qs, d), o);
        int64_t oi = (fr * n_pos + n) * hid + h * DH + lane;
        if (out_f32) out_f32[oi] = o;
        if (out_sb) store_sb1(out_sb, out_plane, oi, o);
    }
}

}  // namespace

extern "C" int lfdm_attn_softmax(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_seq,
                                 int seq_len, int heads, int64_t inner, int64_t outer_stride, int64_t inner_stride,
                                 int64_t row_stride, const float* rot_cos, const float* rot_sin, const float* pos_bias,
                                 void* stream) {
    if (!qkv || seq_len <= 0 || seq_len > MAXL || heads <= 0 || inner <= 0) return LFDM_E_BADARG;
    if ((rot_cos == nullptr) != (rot_sin == nullptr)) return LFDM_E_BADARG;
    int64_t units = n_seq * heads;
    int64_t blocks = (units + 3) / 4;
    size_t smem = sizeof(float) * 4 * 3 * (size_t)seq_len * (DH + 1);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(float) * 4 * 3 * MAXL * (DH + 1)));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    attn_softmax_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
        qkv, (bf16*)out_sb, out_plane, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride, row_stride,
        rot_cos, rot_sin, pos_bias);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_attn_linear(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_frames,
                                int n_pos, int heads, void* stream) {
    if (!qkv || n_pos <= 0 || heads <= 0) return LFDM_E_BADARG;
    attn_linear_kernel<<<(unsigned)(n_frames * heads), 256, 0, (cudaStream_t)stream>>>(qkv, (bf16*)out_sb, out_plane,
                                                                                     out_f32, n_pos, heads);
    LFDM_CHECK_LAUNCH();
    return 0;
}
