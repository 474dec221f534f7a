// conv_simt.cu — general-shape fp32 CUDA-core implicit-GEMM convolution over row matrices (LFDM_ENGINE_SIMT).
//
// Role: (1) the engine for geometries the tcgen05 kernel does not tile (channel counts that are not multiples
// of 64, odd spatial sizes, reflect padding, tiny test models), (2) an independent on-device cross-check of the
// tensor-core path.  Arithmetic: fp32 FMA over exact fp32 weights; SB inputs are read as hi+lo.
// Replaces the cuDNN calls behind nn.Conv3d/ConvTranspose3d/Conv2d/Linear (see include/lfdm_b200.h).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct SimtArgs {
    lfdm_conv_desc d;
    int cin_total;
    int64_t m_out;
    int ktot;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // torch 'reflect' padding (no edge repeat), valid for pad < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// fetch A[m][k] (k = tap*cin_total + c) for output row m
__device__ __forceinline__ float fetch_a(const SimtArgs& a, int nfr, int ho, int wo, int k) {
    const lfdm_conv_desc& d = a.d;
    int tap = k / a.cin_total;
    int c = k - tap * a.cin_total;
    int kh = tap / d.kw, kw = tap - kh * d.kw;
    int hi, wi;
    if (d.mode == LFDM_CONV_DIRECT) {
        hi = ho * d.stride - d.pad + kh;
        wi = wo * d.stride - d.pad + kw;
        if (d.reflect) { hi = reflect_idx(hi, d.h_in); wi = reflect_idx(wi, d.w_in); }
        if (hi < 0 || hi >= d.h_in || wi < 0 || wi >= d.w_in) return 0.f;
    } else if (d.mode == LFDM_CONV_TRANSPOSED) {
        int hn = ho + d.pad - kh, wn = wo + d.pad - kw;
        if (hn < 0 || wn < 0 || (hn % d.stride) || (wn % d.stride)) return 0.f;
        hi = hn / d.stride; wi = wn / d.stride;
        if (hi >= d.h_in || wi >= d.w_in) return 0.f;
    } else {  // UPNEAREST: conv over the x2 nearest-upsampled image
        int hu = ho - d.pad + kh, wu = wo - d.pad + kw;
        int H2 = d.h_in * 2, W2 = d.w_in * 2;
        if (d.reflect) { hu = reflect_idx(hu, H2); wu = reflect_idx(wu, W2); }
        if (hu < 0 || hu >= H2 || wu < 0 || wu >= W2) return 0.f;
        hi = hu >> 1; wi = wu >> 1;
    }
    int src = 0;
    if (c >= d.a_c[0]) { src = 1; c -= d.a_c[0]; }
    int64_t row = ((int64_t)nfr * d.h_in + hi) * d.w_in + wi;
    int64_t idx = row * d.a_c[src] + c;
    if (d.a_f32[src]) return d.a_f32[src][idx];
    const bf16* p = reinterpret_cast<const bf16*>(d.a_sb[src]);
    return join_bf16(p[idx], p[idx + d.a_plane[src]]);
}

__global__ void __launch_bounds__(NT) conv_simt_kernel(SimtArgs a) {
    __shared__ float sA[BK][BM + 4];
    __shared__ float sB[BK][BN + 4];
    const lfdm_conv_desc& d = a.d;
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // A-load assignment: 64 rows x 16 k = 1024 elements, 4 per thread: row = tid / 4, k = (tid % 4) * 4 + e
    const int a_row = tid / 4, a_k0 = (tid % 4) * 4;
    int64_t am = m0 + a_row;
    bool a_valid = am < a.m_out;
    int a_nf = 0, a_ho = 0, a_wo = 0;
    if (a_valid) {
        int64_t hw = (int64_t)d.h_out * d.w_out;
        a_nf = (int)(am / hw);
        int r = (int)(am - (int64_t)a_nf * hw);
        a_ho = r / d.w_out; a_wo = r - a_ho * d.w_out;
    }
    // B-load: 16 k x 64 n = 1024 elements: k = tid / 16, n = (tid % 16) * 4 + e
    const int b_k = tid / 16, b_n0 = (tid % 16) * 4;

    for (int k0 = 0; k0 < a.ktot; k0 += BK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int k = k0 + a_k0 + e;
            float v = 0.f;
            if (a_valid && k < a.ktot) v = fetch_a(a, a_nf, a_ho, a_wo, k);
            sA[a_k0 + e][a_row] = v;
        }
        {
            int k = k0 + b_k;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int n = n0 + b_n0 + e;
                float v = 0.f;
                if (k < a.ktot && n < d.c_out) v = d.w_f32[(int64_t)k * d.c_out + n];
                sB[b_k][b_n0 + e] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sA[kk][tm + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = sB[kk][tn + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue
    const int64_t P = (int64_t)d.h_out * d.w_out;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + tm + i;
        if (m >= a.m_out) continue;
        int64_t rrow = m;
        if (d.res_bcast_f > 0) rrow = (m / ((int64_t)d.res_bcast_f * P)) * P + (m % P);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tn + j;
            if (n >= d.c_out) continue;
            float v = acc[i][j];
            if (d.bias) v += d.bias[n];
            if (d.residual) v += d.residual[rrow * d.c_out + n];
            int64_t o = m * d.c_out + n;
            if (d.out_f32) d.out_f32[o] = apply_act(v, d.f32_act);
            if (d.out_sb) {
                float u = v;
                if (d.sb_scale) u *= d.sb_scale[n];
                if (d.sb_shift) u += d.sb_shift[n];
                u = apply_act(u, d.sb_act);
                store_sb1(reinterpret_cast<bf16*>(d.out_sb), d.out_plane, o, u);
            }
        }
    }
}

}  // namespace

int lfdm_conv_simt(const lfdm_conv_desc* d, cudaStream_t stream) {
    if (!d || !d->w_f32) return LFDM_E_BADARG;
    if (d->gn_stats) return LFDM_E_UNSUPP;  // use lfdm_gn_stats after the conv
    SimtArgs a;
    a.d = *d;
    a.cin_total = d->a_c[0] + d->a_c[1];
    a.m_out = (int64_t)d->nf * d->h_out * d->w_out;
    a.ktot = d->kh * d->kw * a.cin_total;
    if (a.m_out <= 0 || d->c_out <= 0 || a.ktot <= 0) return LFDM_E_BADARG;
    dim3 grid((unsigned)ceil_div64(a.m_out, BM), (unsigned)((d->c_out + BN - 1) / BN));
    conv_simt_kernel<<<grid, NT, 0, stream>>>(a);
    LFDM_CHECK_LAUNCH();
    return 0;
}
