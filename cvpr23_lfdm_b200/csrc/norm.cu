// norm.cu — GroupNorm statistics / apply (+scale-shift, SiLU, residual) and channel LayerNorm over row matrices.
// HBM-bound element-wise kernels: float4 loads/stores, one pass each.
//   GroupNorm: reference Block.forward (DM/modules/video_flow_diffusion.py:203-211) — nn.GroupNorm on a 5-D tensor:
//              statistics span (C/groups, F, H, W) per sample; biased variance; eps inside the sqrt.
//   LayerNorm: reference LayerNorm.forward (:176-179) — over channels, biased variance, gamma only.
#include "common.cuh"

namespace {

// ---------------- GroupNorm statistics: grid (chunks, B*groups); double atomics into stats[b][g][2]
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int64_t rows_per_sample, int c,
                                                       int cpg, int groups, double* __restrict__ stats) {
    const int bg = blockIdx.y;
    const int b = bg / groups, g = bg % groups;
    const int64_t total = rows_per_sample * cpg;
    const float* base = x + (int64_t)b * rows_per_sample * c + g * cpg;
    float s = 0.f, ss = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cpg;
        int cc = (int)(i - r * cpg);
        float v = base[r * c + cc];
        s += v; ss = fmaf(v, v, ss);
    }
    __shared__ double sh[2][8];
    s = warp_sum(s); ss = warp_sum(ss);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = (double)s; sh[1][w] = (double)ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, q = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh[0][i]; q += sh[1][i]; }
        // contributions on a fixed 2^-24 grid: their double sums are exact, so the order the blocks arrive in cannot change the
        // result (same scheme as the fused sums in conv_tc.cu)
        atomicAdd(&stats[bg * 2 + 0], rint(a * 16777216.0) * (1.0 / 16777216.0));
        atomicAdd(&stats[bg * 2 + 1], rint(q * 16777216.0) * (1.0 / 16777216.0));
    }
}

// ---------------- GroupNorm apply: one thread = 4 consecutive channels of one row
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ ss, const float* __restrict__ residual,
                                                       float* __restrict__ out_f32, bf16* __restrict__ out_sb,
                                                       int64_t out_plane, int64_t m, int c, int cpg, int groups,
                                                       int64_t rows_per_sample, float eps, int64_t ss_stride) {
    pdl_prologue_done();
    extern __shared__ float2 s_mr[];   // (mean, rstd) per (b, g)
    const int c4 = c >> 2;
    const int64_t total = m * c4;
    const double inv_n = 1.0 / ((double)rows_per_sample * cpg);
    const int nbg = (int)(m / rows_per_sample) * groups;
    for (int i = threadIdx.x; i < nbg; i += blockDim.x) {
        double mean = stats[i * 2 + 0] * inv_n;
        double var = stats[i * 2 + 1] * inv_n - mean * mean;
        if (var < 0) var = 0;
        s_mr[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    __syncthreads();
    // the 4 channels of a float4 share one group and every parameter vector is 16-byte aligned: vector parameter loads
    const bool quad_params = (cpg & 3) == 0 && ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0 &&
                             (ss == nullptr || ((reinterpret_cast<uintptr_t>(ss) & 15) == 0 && (ss_stride & 3) == 0));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / c4;
        const int ch = (int)(i - row * c4) * 4;
        const int b = (int)(row / rows_per_sample);
        const float4 v = *reinterpret_cast<const float4*>(x + row * c + ch);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (residual) q = *reinterpret_cast<const float4*>(residual + row * c + ch);
        float vv[4] = {v.x, v.y, v.z, v.w};
        float r[4];
        if (quad_params) {
            const float2 mr = s_mr[b * groups + ch / cpg];
            const float4 g4 = *reinterpret_cast<const float4*>(gamma + ch), b4 = *reinterpret_cast<const float4*>(beta + ch);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
            float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
            if (ss) {
                const float4 s4 = *reinterpret_cast<const float4*>(ss + (int64_t)b * ss_stride + ch);
                const float4 h4 = *reinterpret_cast<const float4*>(ss + (int64_t)b * ss_stride + c + ch);
                sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
                sh[0] = h4.x; sh[1] = h4.y; sh[2] = h4.z; sh[3] = h4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float y = (vv[e] - mr.x) * mr.y * gg[e] + bb[e];
                if (ss) y = y * (sc[e] + 1.f) + sh[e];
                r[e] = silu_f(y);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cc = ch + e;
                const float2 mr = s_mr[b * groups + cc / cpg];
                float y = (vv[e] - mr.x) * mr.y * gamma[cc] + beta[cc];
                if (ss) y = y * (ss[(int64_t)b * ss_stride + cc] + 1.f) + ss[(int64_t)b * ss_stride + c + cc];
                r[e] = silu_f(y);
            }
        }
        const float4 o = make_float4(r[0] + q.x, r[1] + q.y, r[2] + q.z, r[3] + q.w);
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * c + ch) = o;
        if (out_sb) store_sb4(out_sb, out_plane, row * c + ch, o);
    }
}

// ---------------- GroupNorm apply, streaming form (c / 4 divides 256, 4 | cpg, aligned parameter vectors, < 2^31 float4): a thread keeps
// its 4 channels for the whole kernel (gamma / beta in registers, no 64-bit division per element) and has U independent rows in
// flight per iteration.  The generic kernel above measured 77 % of the HBM peak at 32x32 with one 16-byte load in flight per thread.
template <int U>
__global__ void __launch_bounds__(256) gn_apply_stream_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ ss, const float* __restrict__ residual,
                                                              float* __restrict__ out_f32, bf16* __restrict__ out_sb,
                                                              int64_t out_plane, int m, int c, int cpg, int groups,
                                                              int rows_per_sample, float eps, int64_t ss_stride) {
    pdl_prologue_done();
    extern __shared__ float2 s_mr[];   // (mean, rstd) per (b, g)
    const int c4 = c >> 2;
    const double inv_n = 1.0 / ((double)rows_per_sample * cpg);
    const int nbg = (m / rows_per_sample) * groups;
    for (int i = threadIdx.x; i < nbg; i += blockDim.x) {
        double mean = stats[i * 2 + 0] * inv_n;
        double var = stats[i * 2 + 1] * inv_n - mean * mean;
        if (var < 0) var = 0;
        s_mr[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    __syncthreads();
    const int rpb = 256 / c4;                                    // rows per block pass
    const int ch = (threadIdx.x % c4) * 4;
    const int g = ch / cpg;
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + ch), b4 = *reinterpret_cast<const float4*>(beta + ch);
    const int row_stride = (int)gridDim.x * rpb;
    for (int row0 = (int)blockIdx.x * rpb + (int)threadIdx.x / c4; row0 < m; row0 += U * row_stride) {
        float4 v[U], q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int row = row0 + u * row_stride;
            if (row < m) {
                v[u] = *reinterpret_cast<const float4*>(x + (int64_t)row * c + ch);
                q[u] = residual ? *reinterpret_cast<const float4*>(residual + (int64_t)row * c + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int row = row0 + u * row_stride;
            if (row >= m) break;
            const int b = row / rows_per_sample;
            const float2 mr = s_mr[b * groups + g];
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ss) {
                const float4 s4 = *reinterpret_cast<const float4*>(ss + (int64_t)b * ss_stride + ch);
                sh = *reinterpret_cast<const float4*>(ss + (int64_t)b * ss_stride + c + ch);
                sc = make_float4(s4.x + 1.f, s4.y + 1.f, s4.z + 1.f, s4.w + 1.f);
            }
            // same expression order as the generic kernel: ((x - mean) * rstd * gamma + beta) * (scale + 1) + shift
            float y0 = (v[u].x - mr.x) * mr.y * g4.x + b4.x, y1 = (v[u].y - mr.x) * mr.y * g4.y + b4.y;
            float y2 = (v[u].z - mr.x) * mr.y * g4.z + b4.z, y3 = (v[u].w - mr.x) * mr.y * g4.w + b4.w;
            if (ss) { y0 = y0 * sc.x + sh.x; y1 = y1 * sc.y + sh.y; y2 = y2 * sc.z + sh.z; y3 = y3 * sc.w + sh.w; }
            const float4 o = make_float4(silu_f(y0) + q[u].x, silu_f(y1) + q[u].y, silu_f(y2) + q[u].z, silu_f(y3) + q[u].w);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + (int64_t)row * c + ch) = o;
            if (out_sb) store_sb4(out_sb, out_plane, (int64_t)row * c + ch, o);
        }
    }
}

// ---------------- LayerNorm over channels: c % 4 == 0, c <= 1024.  A row is handled by G = c4-rounded-up-to-pow2
// (<= 32) lanes or by a full warp with NJ float4 per lane; R rows are in flight per warp iteration so that a warp
// always has >= 8 independent 16-byte loads outstanding (the kernel is pure HBM streaming).
template <int G, int NJ>      // G lanes per row (G <= 32), NJ float4 per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        bf16* __restrict__ out_sb, int64_t out_plane,
                                                        float* __restrict__ out_f32, int64_t m, int c, float eps) {
    pdl_prologue_done();
    constexpr int RPW = 32 / G;                 // rows per warp per pass
    constexpr int PASSES = (NJ >= 4) ? 1 : (4 / NJ);   // independent row passes kept in flight
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, gl = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int c4 = c >> 2;
    float4 g[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int q = gl + G * j;
        g[j] = q < c4 ? *reinterpret_cast<const float4*>(gamma + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t row0 = warp * (RPW * PASSES); row0 < m; row0 += nwarps * (RPW * PASSES)) {
        float4 v[PASSES][NJ];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int64_t row = row0 + p * RPW + sub;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int q = gl + G * j;
                v[p][j] = (row < m && q < c4) ? *reinterpret_cast<const float4*>(x + row * c + q * 4)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int64_t row = row0 + p * RPW + sub;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s += (v[p][j].x + v[p][j].y) + (v[p][j].z + v[p][j].w);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)c;
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (gl + G * j < c4) {
                    const float a = v[p][j].x - mean, b = v[p][j].y - mean, cc = v[p][j].z - mean, d = v[p][j].w - mean;
                    sq += a * a + b * b + cc * cc + d * d;
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = 1.f / sqrtf(sq / (float)c + eps);
            if (row < m) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int q = gl + G * j;
                    if (q < c4) {
                        const float4 o4 = make_float4((v[p][j].x - mean) * rstd * g[j].x, (v[p][j].y - mean) * rstd * g[j].y,
                                                      (v[p][j].z - mean) * rstd * g[j].z, (v[p][j].w - mean) * rstd * g[j].w);
                        if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * c + q * 4) = o4;
                        if (out_sb) store_sb4(out_sb, out_plane, row * c + q * 4, o4);
                    }
                }
            }
        }
    }
}

}  // namespace

extern "C" int lfdm_gn_stats(const float* x, int64_t m, int c, int groups, int rows_per_sample, double* stats,
                             void* stream) {
    if (!x || !stats || groups <= 0 || c % groups || rows_per_sample <= 0 || m % rows_per_sample) return LFDM_E_BADARG;
    int b = (int)(m / rows_per_sample);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * b * groups, st);
    if (e != cudaSuccess) return (int)e;
    int cpg = c / groups;
    int64_t total = (int64_t)rows_per_sample * cpg;
    int chunks = (int)((total + 256 * 16 - 1) / (256 * 16));
    if (chunks < 1) chunks = 1;
    if (chunks > 512) chunks = 512;
    gn_stats_kernel<<<dim3(chunks, b * groups), 256, 0, st>>>(x, rows_per_sample, c, cpg, groups, stats);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_gn_apply(const float* x, const double* stats, const float* gamma, const float* beta,
                             const float* ss, int64_t ss_stride, const float* residual, float* out_f32, void* out_sb, int64_t out_plane,
                             int64_t m, int c, int groups, int rows_per_sample, float eps, void* stream) {
    if (!x || !stats || !gamma || !beta || (c & 3) || c % groups || m % rows_per_sample) return LFDM_E_BADARG;
    int64_t total = m * (c >> 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    size_t smem = sizeof(float2) * (size_t)(m / rows_per_sample) * groups;
    if (smem > 40000) return LFDM_E_UNSUPP;
    {
        static const bool allow = (getenv("LFDM_GN_GENERIC") == nullptr);           // A/B switch
        const int c4 = c >> 2, cpg = c / groups;
        const bool aligned = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(ss)) & 15) == 0 &&
                             (ss_stride & 3) == 0;
        if (allow && c4 <= 256 && (256 % c4) == 0 && (cpg & 3) == 0 && aligned && total < (1ll << 31) && m < (1ll << 31)) {
            const int rpb = 256 / c4;
            int64_t nb = (m + (int64_t)rpb * 4 - 1) / ((int64_t)rpb * 4);       // 4 rows in flight per thread
            if (nb > 148 * 8) nb = 148 * 8;
            if (nb < 1) nb = 1;
            LFDM_LAUNCH_PDL(gn_apply_stream_kernel<4>, dim3((unsigned)nb), dim3(256), smem, (cudaStream_t)stream, x, stats, gamma, beta, ss,
                            residual, out_f32, (bf16*)out_sb, out_plane, (int)m, c, cpg, groups, rows_per_sample, eps, ss_stride);
            return 0;
        }
    }
    LFDM_LAUNCH_PDL(gn_apply_kernel, dim3(blocks), dim3(256), smem, (cudaStream_t)stream, x, stats, gamma, beta, ss, residual,
                    out_f32, (bf16*)out_sb, out_plane, m, c, c / groups, groups, rows_per_sample, eps, ss_stride);
    return 0;
}

extern "C" int lfdm_layernorm(const float* x, const float* gamma, void* out_sb, int64_t out_plane, float* out_f32,
                              int64_t m, int c, float eps, void* stream) {
    if (!x || !gamma || (c & 3) || c > 1024) return LFDM_E_BADARG;
    const int c4 = c >> 2;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t rows_per_block;
#define LFDM_LN_LAUNCH(G, NJ)                                                                                         \
    do {                                                                                                              \
        rows_per_block = 8 * (32 / G) * ((NJ >= 4) ? 1 : (4 / NJ));                                                   \
        int64_t blocks = (m + rows_per_block - 1) / rows_per_block;                                                   \
        if (blocks > 148 * 16) blocks = 148 * 16;                                                                     \
        LFDM_LAUNCH_PDL((layernorm_kernel<G, NJ>), dim3((unsigned)blocks), dim3(256), 0, st, x, gamma, (bf16*)out_sb,   \
                        out_plane, out_f32, m, c, eps);                                                               \
    } while (0)
    if (c4 <= 4) LFDM_LN_LAUNCH(4, 1);
    else if (c4 <= 8) LFDM_LN_LAUNCH(8, 1);
    else if (c4 <= 16) LFDM_LN_LAUNCH(16, 1);
    else if (c4 <= 32) LFDM_LN_LAUNCH(32, 1);
    else if (c4 <= 64) LFDM_LN_LAUNCH(32, 2);
    else if (c4 <= 128) LFDM_LN_LAUNCH(32, 4);
    else LFDM_LN_LAUNCH(32, 8);
#undef LFDM_LN_LAUNCH
    LFDM_CHECK_LAUNCH();
    return 0;
}
