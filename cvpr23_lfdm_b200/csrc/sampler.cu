// sampler.cu — the reverse-diffusion update of GaussianDiffusion (DM/modules/video_flow_diffusion.py):
//   predict_start_from_noise :697-701, dynamic thresholding with torch.quantile :719-732, q_posterior :703-710,
//   p_sample :737-746, ddim_sample body :792-827.
// HBM/L2-bound element-wise passes + an exact per-sample radix select (replaces torch.quantile's full sort).
// All element-wise arithmetic uses round-to-nearest intrinsics in the reference's operation order (no FMA
// contraction), so given identical eps the update is bit-identical to the PyTorch expression.
#include "common.cuh"

namespace {

__device__ __forceinline__ float x0_of(float x, float eps, float c1, float c2) {
    return __fsub_rn(__fmul_rn(c1, x), __fmul_rn(c2, eps));
}

__global__ void __launch_bounds__(256) x0_abs_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                     const float* __restrict__ coef,
                                                     const int32_t* __restrict__ step_idx, float* __restrict__ absx0,
                                                     int64_t total4) {
    pdl_prologue_done();
    const int row = step_idx ? *step_idx : 0;
    const float c1 = coef[row * 8 + 0], c2 = coef[row * 8 + 1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = reinterpret_cast<const float4*>(x)[i], e = reinterpret_cast<const float4*>(eps)[i];
        float4 o;
        o.x = fabsf(x0_of(a.x, e.x, c1, c2)); o.y = fabsf(x0_of(a.y, e.y, c1, c2));
        o.z = fabsf(x0_of(a.z, e.z, c1, c2)); o.w = fabsf(x0_of(a.w, e.w, c1, c2));
        reinterpret_cast<float4*>(absx0)[i] = o;
    }
}

// exact k-th / (k+1)-th order statistic of non-negative floats: 4 passes of 8-bit MSB-first radix select + one
// min pass; one 1024-thread block per sample.  Non-negative IEEE floats order like their uint32 bit patterns.
__global__ void __launch_bounds__(1024) quantile_kernel(const float* __restrict__ absx0, float* __restrict__ s_out,
                                                        int64_t n, int64_t k_lo, float w_hi) {
    pdl_prologue_done();
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sh_prefix, sh_k, sh_ceq;
    __shared__ unsigned int sh_min[32];
    const uint32_t* keys = reinterpret_cast<const uint32_t*>(absx0) + (int64_t)blockIdx.x * n;
    const int tid = threadIdx.x;
    uint32_t prefix = 0, mask = 0;
    uint32_t k = (uint32_t)k_lo;
    uint32_t ceq = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += 1024) {
            uint32_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            int d = 0;
            for (; d < 256; ++d) {
                uint32_t c = hist[d];
                if (acc + c > k) break;
                acc += c;
            }
            if (d > 255) d = 255;
            sh_prefix = prefix | ((uint32_t)d << shift);
            sh_k = k - acc;
            sh_ceq = hist[d];
        }
        __syncthreads();
        prefix = sh_prefix; k = sh_k; ceq = sh_ceq;
        mask |= 0xFFu << shift;
        __syncthreads();
    }
    // prefix = key of rank k_lo; k = its rank among the ceq equal keys
    uint32_t v_lo = prefix, v_hi = prefix;
    if (k + 1 >= ceq) {   // rank k_lo+1 is the smallest key strictly greater than v_lo (if any)
        uint32_t mn = 0xFFFFFFFFu;
        for (int64_t i = tid; i < n; i += 1024) {
            uint32_t key = keys[i];
            if (key > v_lo && key < mn) mn = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((tid & 31) == 0) sh_min[tid >> 5] = mn;
        __syncthreads();
        if (tid < 32) {
            mn = sh_min[tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if (tid == 0) sh_min[0] = mn;
        }
        __syncthreads();
        v_hi = sh_min[0] == 0xFFFFFFFFu ? v_lo : sh_min[0];
    }
    if (tid == 0) {
        float a = __uint_as_float(v_lo), b = __uint_as_float(v_hi);
        // at::lerp: |w| < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        float diff = __fsub_rn(b, a);
        float q = (fabsf(w_hi) < 0.5f) ? __fadd_rn(a, __fmul_rn(w_hi, diff)) : __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.f, w_hi)));
        s_out[blockIdx.x] = fmaxf(q, 1.f);   // s.clamp_(min=1.)
    }
}

__global__ void __launch_bounds__(256) update_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                     const float* __restrict__ noise, const float* __restrict__ s,
                                                     const float* __restrict__ coef,
                                                     const int32_t* __restrict__ step_idx, float* __restrict__ x_out,
                                                     float* __restrict__ x0_out, int64_t n4_per_sample,
                                                     int64_t total4) {
    pdl_prologue_done();
    const int row = step_idx ? *step_idx : 0;
    const float* cf = coef + row * 8;
    const float c1 = cf[0], c2 = cf[1], ca = cf[2], cb = cf[3], sg = cf[4], ce = cf[5];
    const bool ddim = cf[6] != 0.f;
    const bool noclip = cf[7] != 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / n4_per_sample);
        const float sv = s ? s[b] : 1.f;
        float4 xa = reinterpret_cast<const float4*>(x)[i], ea = reinterpret_cast<const float4*>(eps)[i];
        float4 za = make_float4(0.f, 0.f, 0.f, 0.f);
        if (noise) za = reinterpret_cast<const float4*>(noise)[i];
        float xv[4] = {xa.x, xa.y, xa.z, xa.w}, ev[4] = {ea.x, ea.y, ea.z, ea.w}, zv[4] = {za.x, za.y, za.z, za.w};
        float o[4], o0[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x0 = x0_of(xv[e], ev[e], c1, c2);
            if (!noclip) x0 = __fdiv_rn(fminf(fmaxf(x0, -sv), sv), sv);       // clamp(-s, s) / s
            o0[e] = x0;
            float y = ddim ? __fmul_rn(ce, ev[e]) : __fmul_rn(cb, xv[e]);
            float v = __fadd_rn(__fmul_rn(ca, x0), y);
            if (noise) v = __fadd_rn(v, __fmul_rn(sg, zv[e]));
            o[e] = v;
        }
        reinterpret_cast<float4*>(x_out)[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (x0_out) reinterpret_cast<float4*>(x0_out)[i] = make_float4(o0[0], o0[1], o0[2], o0[3]);
    }
}

__global__ void advance_kernel(int32_t* step_idx) {
    pdl_prologue_done();
    *step_idx += 1;
}

}  // namespace

extern "C" int lfdm_sampler_x0(const float* x, const float* eps, const float* coef, const int32_t* step_idx,
                               float* absx0, int64_t n_per_sample, int b, void* stream) {
    if (!x || !eps || !coef || !absx0 || (n_per_sample & 3)) return LFDM_E_BADARG;
    int64_t total4 = n_per_sample / 4 * b;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    LFDM_LAUNCH_PDL(x0_abs_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, eps, coef, step_idx, absx0, total4);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_sampler_quantile(const float* absx0, float* s, int64_t n_per_sample, int b, int64_t k_lo,
                                     float w_hi, void* workspace, void* stream) {
    (void)workspace;
    if (!absx0 || !s || k_lo < 0 || k_lo >= n_per_sample) return LFDM_E_BADARG;
    LFDM_LAUNCH_PDL(quantile_kernel, dim3(b), dim3(1024), 0, (cudaStream_t)stream, absx0, s, n_per_sample, k_lo, w_hi);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_sampler_update(const float* x, const float* eps, const float* noise, const float* s,
                                   const float* coef, int32_t* step_idx, int advance, float* x_out, float* x0_out,
                                   int64_t n_per_sample, int b, void* stream) {
    if (!x || !eps || !coef || !x_out || (n_per_sample & 3)) return LFDM_E_BADARG;
    int64_t total4 = n_per_sample / 4 * b;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    LFDM_LAUNCH_PDL(update_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, eps, noise, s, coef, step_idx, x_out,
                    x0_out, n_per_sample / 4, total4);
    LFDM_CHECK_LAUNCH();
    if (advance && step_idx) {
        LFDM_LAUNCH_PDL(advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_idx);
        LFDM_CHECK_LAUNCH();
    }
    return 0;
}
