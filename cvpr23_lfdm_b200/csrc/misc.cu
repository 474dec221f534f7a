// misc.cu — embedding MLPs, layout conversion and small element-wise kernels of the LFDM hot path.
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_f(float v, int act) {
    if (act == 1) return silu_f(v);
    if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));   // nn.GELU() (erf form)
    return v;
}

// y[r][n] = act_out(sum_k act_in(x[r][k]) W[n][k] + b[n]); one warp per output column n, W row cached in registers.
// time_mlp (video_flow_diffusion.py:422-428) and the per-block (scale,shift) MLPs (:217-220,230-232).
template <int KPL>   // k values per lane (k <= 32*KPL)
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ y, int rows,
                                                           int k, int n, int act_in, int act_out) {
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (col >= n) return;
    float wr[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        int kk = lane + 32 * j;
        wr[j] = kk < k ? w[(int64_t)col * k + kk] : 0.f;
    }
    const float bias = b ? b[col] : 0.f;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            int kk = lane + 32 * j;
            if (kk < k) acc = fmaf(act_f(x[(int64_t)r * k + kk], act_in), wr[j], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) y[(int64_t)r * n + col] = act_f(acc + bias, act_out);
    }
}

__global__ void sinusoidal_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                  float* __restrict__ out, int rows, int dim) {
    const int half = dim / 2;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * half) return;
    int r = i / half, j = i % half;
    float a = (float)t[r] * freqs[j];
    out[r * dim + j] = sinf(a);
    out[r * dim + half + j] = cosf(a);
}

__global__ void ss_combine_kernel(const float* __restrict__ time_tab, const int32_t* __restrict__ step_idx,
                                  const float* __restrict__ cond_tab, float* __restrict__ ss, int b, int n) {
    pdl_prologue_done();
    const int row = step_idx ? *step_idx : 0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * n) return;
    int j = i % n;
    ss[i] = time_tab[(int64_t)row * n + j] + cond_tab[i];
}

// in[b][c][f][p] -> rows[(b*F+f)*P + p][c_pad]; 32x32 smem transpose tile; grid (P/32, c_pad/32, B*F)
__global__ void __launch_bounds__(256) to_rows_kernel(const float* __restrict__ in, int c, int f, int p, int64_t sb,
                                                      int64_t sc, int64_t sf, int c_pad, bf16* __restrict__ out_sb,
                                                      int64_t out_plane, float* __restrict__ out_f32) {
    pdl_prologue_done();
    __shared__ float tile[32][33];
    const int n = blockIdx.z, bi = n / f, fi = n % f;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int cc = c0 + ty + 8 * j, pp = p0 + tx;
        float v = 0.f;
        if (cc < c && pp < p) v = in[bi * sb + cc * sc + fi * sf + pp];
        tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int pp = p0 + ty + 8 * j, cc = c0 + tx;
        if (pp < p && cc < c_pad) {
            float v = tile[tx][ty + 8 * j];
            int64_t o = ((int64_t)n * p + pp) * c_pad + cc;
            if (out_f32) out_f32[o] = v;
            if (out_sb) store_sb1(out_sb, out_plane, o, v);
        }
    }
}

// rows[(b*F+f)*P+p][ld] -> out[b][c][f][p]
__global__ void __launch_bounds__(256) from_rows_kernel(const float* __restrict__ rows, int ld, int c, int f, int p,
                                                        float* __restrict__ out) {
    pdl_prologue_done();
    __shared__ float tile[32][33];
    const int n = blockIdx.z, bi = n / f, fi = n % f;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int pp = p0 + ty + 8 * j, cc = c0 + tx;
        float v = 0.f;
        if (pp < p && cc < c) v = rows[((int64_t)n * p + pp) * ld + cc];
        tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int cc = c0 + ty + 8 * j, pp = p0 + tx;
        if (cc < c && pp < p) out[(((int64_t)bi * c + cc) * f + fi) * p + pp] = tile[tx][ty + 8 * j];
    }
}

// in[b][c][f][h][w] -> SB rows [M][k_pad], k = (kh*ks + kw)*c + ch, zero padded borders / k >= ks*ks*c.
// One thread produces 4 consecutive k of one row (8-byte stores into each plane); 32-bit index math.
__global__ void __launch_bounds__(256) im2col_small_kernel(const float* __restrict__ in, int c, int f, int h, int w,
                                                           int ks, int pad, int k_pad, bf16* __restrict__ out_sb,
                                                           int64_t out_plane, int64_t total4) {
    const int kreal = ks * ks * c;
    const int kq_per_row = k_pad >> 2;
    const int hw = h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / kq_per_row);
        const int kq = (int)(i - (int64_t)m * kq_per_row);
        const int n = m / hw;
        const int rem = m - n * hw;
        const int y = rem / w, x = rem - y * w;
        const int bi = n / f, fi = n - bi * f;
        const float* base = in + ((int64_t)bi * c * f + fi) * hw;           // + ch * f * hw + yy * w + xx
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = kq * 4 + e;
            float val = 0.f;
            if (k < kreal) {
                const int tap = k / c, ch = k - tap * c;
                const int kh = tap / ks, kw = tap - kh * ks;
                const int yy = y - pad + kh, xx = x - pad + kw;
                if (yy >= 0 && yy < h && xx >= 0 && xx < w) val = base[(int64_t)ch * f * hw + yy * w + xx];
            }
            v[e] = val;
        }
        store_sb4(out_sb, out_plane, (int64_t)m * k_pad + kq * 4, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// Same result, one CTA per output image row: the c x ks x (w + 2 pad) input patch is staged (zero padded) in shared
// memory, every thread owns one 4-wide k group (its four (channel, tap) smem offsets are computed once) and walks the
// pixels of the row; stores are 8-byte, contiguous across the threads of a pixel.  (The gather straight from global
// memory above touches ~100 cache lines per warp load.)
__global__ void __launch_bounds__(256) im2col_row_kernel(const float* __restrict__ in, int c, int f, int h, int w, int ks,
                                                         int pad, int k_pad, bf16* __restrict__ out_sb, int64_t out_plane) {
    pdl_prologue_done();
    extern __shared__ float patch[];                      // [c][ks][w + 2 pad]
    const int wp = w + 2 * pad;
    const int row = blockIdx.x;                           // (n, y)
    const int n = row / h, y = row - n * h;
    const int bi = n / f, fi = n - bi * f;
    const int hw = h * w;
    const float* base = in + ((int64_t)bi * c * f + fi) * hw;
    for (int i = threadIdx.x; i < c * ks * wp; i += blockDim.x) {
        const int xx = i % wp - pad, r = i / wp;
        const int kh = r % ks, ch = r / ks;
        const int yy = y - pad + kh;
        float v = 0.f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = base[(int64_t)ch * f * hw + yy * w + xx];
        patch[i] = v;
    }
    __syncthreads();
    const int kq_per_row = k_pad >> 2, kreal = ks * ks * c;
    const int kq = threadIdx.x % kq_per_row, px0 = threadIdx.x / kq_per_row, pstep = blockDim.x / kq_per_row;
    int off[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = kq * 4 + e;
        off[e] = -1;
        if (k < kreal) {
            const int tap = k / c, ch = k - tap * c;
            const int kh = tap / ks, kw = tap - kh * ks;
            off[e] = (ch * ks + kh) * wp + kw;
        }
    }
    for (int px = px0; px < w; px += pstep) {
        float4 v;
        v.x = off[0] >= 0 ? patch[off[0] + px] : 0.f;
        v.y = off[1] >= 0 ? patch[off[1] + px] : 0.f;
        v.z = off[2] >= 0 ? patch[off[2] + px] : 0.f;
        v.w = off[3] >= 0 ? patch[off[3] + px] : 0.f;
        store_sb4(out_sb, out_plane, ((int64_t)row * w + px) * k_pad + kq * 4, v);
    }
}

__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ in, int h, int w, int c,
                                                       float* __restrict__ out_f32, bf16* __restrict__ out_sb,
                                                       int64_t out_plane, int64_t total) {
    const int c4 = c >> 2, ho = h >> 1, wo = w >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cq = (int)(i % c4);
        int64_t pix = i / c4;
        int x = (int)(pix % wo);
        int y = (int)((pix / wo) % ho);
        int64_t n = pix / ((int64_t)ho * wo);
        const float* b = in + (((int64_t)n * h + 2 * y) * w + 2 * x) * c + cq * 4;
        float4 a0 = *reinterpret_cast<const float4*>(b), a1 = *reinterpret_cast<const float4*>(b + c);
        float4 a2 = *reinterpret_cast<const float4*>(b + (int64_t)w * c), a3 = *reinterpret_cast<const float4*>(b + (int64_t)w * c + c);
        float4 r = make_float4((a0.x + a1.x + a2.x + a3.x) * 0.25f, (a0.y + a1.y + a2.y + a3.y) * 0.25f,
                               (a0.z + a1.z + a2.z + a3.z) * 0.25f, (a0.w + a1.w + a2.w + a3.w) * 0.25f);
        int64_t o = pix * c + cq * 4;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = r;
        if (out_sb) store_sb4(out_sb, out_plane, o, r);
    }
}

// one warp per row: out[b][ch][f][p] = dot(a[m], wa[ch]) + ba[ch]  (ch < na)  |  dot(o[m], wo[ch-na]) + bo
__global__ void __launch_bounds__(256) unet_heads_kernel(const float* __restrict__ a, const float* __restrict__ wa,
                                                         const float* __restrict__ ba, int na,
                                                         const float* __restrict__ o, const float* __restrict__ wo,
                                                         const float* __restrict__ bo, int no, int c, int f, int p,
                                                         int64_t m_total, float* __restrict__ out, int64_t null_off, float cfg) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nch = na + no;
    for (int64_t m = warp; m < m_total; m += nwarps) {
        const int64_t n = m / p;
        const int pp = (int)(m % p);
        const int bi = (int)(n / f), fi = (int)(n % f);
        for (int ch = 0; ch < nch; ++ch) {
            const float* src = ch < na ? a : o;
            const float* wv = ch < na ? wa + (int64_t)ch * c : wo + (int64_t)(ch - na) * c;
            float acc = 0.f;
            for (int k = lane; k < c; k += 32) acc = fmaf(src[m * c + k], wv[k], acc);
            acc = warp_sum(acc);
            if (null_off) {          // classifier-free guidance: rows m (cond) and m + null_off (null) of a 2B batch
                float accn = 0.f;
                for (int k = lane; k < c; k += 32) accn = fmaf(src[(m + null_off) * c + k], wv[k], accn);
                accn = warp_sum(accn);
                const float bb = ch < na ? (ba ? ba[ch] : 0.f) : (bo ? bo[ch - na] : 0.f);
                acc = (accn + bb) + ((acc + bb) - (accn + bb)) * cfg - bb;
            }
            if (lane == 0) {
                float bias = ch < na ? (ba ? ba[ch] : 0.f) : (bo ? bo[ch - na] : 0.f);
                out[(((int64_t)bi * nch + ch) * f + fi) * p + pp] = acc + bias;
            }
        }
    }
}

// c == 64, na + no <= 4: eight lanes per row (two float4 each -> every warp load instruction covers four full rows =
// four 256-byte segments), weights held in registers, three shuffles per output channel.
__global__ void __launch_bounds__(256) unet_heads64_kernel(const float* __restrict__ a, const float* __restrict__ wa,
                                                           const float* __restrict__ ba, int na,
                                                           const float* __restrict__ o, const float* __restrict__ wo,
                                                           const float* __restrict__ bo, int no, int f, int p,
                                                           int64_t m_total, float* __restrict__ out, int64_t null_off, float cfg) {
    pdl_prologue_done();
    const int lane = threadIdx.x & 31, sub = lane & 7, rsel = lane >> 3;
    const int nch = na + no;
    float w[4][8], bias[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        bias[ch] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) w[ch][j] = 0.f;
        if (ch < nch) {
            const float* wv = ch < na ? wa + ch * 64 : wo + (ch - na) * 64;
            const float4 w0 = *reinterpret_cast<const float4*>(wv + sub * 8), w1 = *reinterpret_cast<const float4*>(wv + sub * 8 + 4);
            w[ch][0] = w0.x; w[ch][1] = w0.y; w[ch][2] = w0.z; w[ch][3] = w0.w;
            w[ch][4] = w1.x; w[ch][5] = w1.y; w[ch][6] = w1.z; w[ch][7] = w1.w;
            bias[ch] = ch < na ? (ba ? ba[ch] : 0.f) : (bo ? bo[ch - na] : 0.f);
        }
    }
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m0 = gw * 4; m0 < m_total; m0 += nw * 4) {
        const int64_t m = m0 + rsel;
        if (m >= m_total) continue;        // whole 8-lane groups drop out together (shuffles below stay inside a group)
        const float4 a0 = *reinterpret_cast<const float4*>(a + m * 64 + sub * 8), a1 = *reinterpret_cast<const float4*>(a + m * 64 + sub * 8 + 4);
        const float4 o0 = *reinterpret_cast<const float4*>(o + m * 64 + sub * 8), o1 = *reinterpret_cast<const float4*>(o + m * 64 + sub * 8 + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        float acc[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            acc[ch] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[ch] = fmaf(ch < na ? av[j] : ov[j], w[ch][j], acc[ch]);
        }
        const unsigned gmask = 0xffu << (rsel * 8);
        if (null_off) {              // classifier-free guidance (reference :521-526): eps = null + (cond - null) * scale, with
                                     // logits - null_logits formed first as the reference does; rows m / m + null_off of a 2B batch
            const int64_t mn = m + null_off;
            const float4 c0 = *reinterpret_cast<const float4*>(a + mn * 64 + sub * 8), c1 = *reinterpret_cast<const float4*>(a + mn * 64 + sub * 8 + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(o + mn * 64 + sub * 8), d1 = *reinterpret_cast<const float4*>(o + mn * 64 + sub * 8 + 4);
            const float an[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            const float on[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float accn = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) accn = fmaf(ch < na ? an[j] : on[j], w[ch][j], accn);
                acc[ch] += __shfl_xor_sync(gmask, acc[ch], 1); accn += __shfl_xor_sync(gmask, accn, 1);
                acc[ch] += __shfl_xor_sync(gmask, acc[ch], 2); accn += __shfl_xor_sync(gmask, accn, 2);
                acc[ch] += __shfl_xor_sync(gmask, acc[ch], 4); accn += __shfl_xor_sync(gmask, accn, 4);
                const float lc = acc[ch] + bias[ch], ln = accn + bias[ch];
                acc[ch] = ln + (lc - ln) * cfg;                 // bias already inside both logits
            }
        } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            acc[ch] += __shfl_xor_sync(gmask, acc[ch], 1);
            acc[ch] += __shfl_xor_sync(gmask, acc[ch], 2);
            acc[ch] += __shfl_xor_sync(gmask, acc[ch], 4);
        }
        }
        if (sub < nch) {
            const int64_t n = m / p;
            const int pp = (int)(m - n * p);
            const int bi = (int)(n / f), fi = (int)(n - (int64_t)bi * f);
            const float bz = null_off ? 0.f : 1.f;
            float v = acc[0] + bz * bias[0];
            if (sub == 1) v = acc[1] + bz * bias[1];
            if (sub == 2) v = acc[2] + bz * bias[2];
            if (sub == 3) v = acc[3] + bz * bias[3];
            out[(((int64_t)bi * nch + sub) * f + fi) * p + pp] = v;
        }
    }
}

__global__ void split_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t plane, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        store_sb1(out, plane, i, in[i]);
}

// SB rows [n][h][w][c] -> replicate-padded SB rows [n][h+2][w+2][c] (16-byte chunks of both planes)
__global__ void pad_replicate_kernel(const uint4* __restrict__ in, int64_t in_plane16, uint4* __restrict__ out, int64_t out_plane16,
                                     int n, int h, int w, int c8) {
    const int64_t total = (int64_t)n * (h + 2) * (w + 2) * c8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % c8);
        int64_t p = i / c8;
        const int x = (int)(p % (w + 2)); p /= (w + 2);
        const int y = (int)(p % (h + 2));
        const int64_t img = p / (h + 2);
        const int sx = min(max(x - 1, 0), w - 1), sy = min(max(y - 1, 0), h - 1);
        const int64_t src = ((img * h + sy) * w + sx) * c8 + q;
        out[i] = in[src];
        out[out_plane16 + i] = in[in_plane16 + src];
    }
}

inline int grid_for(int64_t total, int cap = 148 * 32) {
    int64_t b = (total + 255) / 256;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" int lfdm_small_linear(const float* x, const float* w, const float* b, float* y, int rows, int k, int n,
                                 int act_in, int act_out, void* stream) {
    if (!x || !w || !y || rows <= 0 || k <= 0 || n <= 0 || k > 2048) return LFDM_E_BADARG;
    dim3 grid((n + 7) / 8, rows < 64 ? rows : 64);
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 64) small_linear_kernel<2><<<grid, 256, 0, st>>>(x, w, b, y, rows, k, n, act_in, act_out);
    else if (k <= 256) small_linear_kernel<8><<<grid, 256, 0, st>>>(x, w, b, y, rows, k, n, act_in, act_out);
    else if (k <= 1024) small_linear_kernel<32><<<grid, 256, 0, st>>>(x, w, b, y, rows, k, n, act_in, act_out);
    else small_linear_kernel<64><<<grid, 256, 0, st>>>(x, w, b, y, rows, k, n, act_in, act_out);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_sinusoidal(const int64_t* t, const float* freqs, float* out, int rows, int dim, void* stream) {
    if (!t || !freqs || !out || dim < 2 || (dim & 1)) return LFDM_E_BADARG;
    int total = rows * (dim / 2);
    sinusoidal_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, freqs, out, rows, dim);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_ss_combine(const float* time_tab, const int32_t* step_idx, const float* cond_tab, float* ss, int b,
                               int n, void* stream) {
    if (!time_tab || !cond_tab || !ss) return LFDM_E_BADARG;
    int total = b * n;
    LFDM_LAUNCH_PDL(ss_combine_kernel, dim3((total + 255) / 256), dim3(256), 0, (cudaStream_t)stream, time_tab, step_idx, cond_tab, ss, b, n);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_to_rows(const float* in, int b, int c, int f, int p, int64_t sb, int64_t sc, int64_t sf, int c_pad,
                            void* out_sb, int64_t out_plane, float* out_f32, void* stream) {
    if (!in || c_pad < c || (!out_sb && !out_f32)) return LFDM_E_BADARG;
    dim3 grid((p + 31) / 32, (c_pad + 31) / 32, b * f);
    LFDM_LAUNCH_PDL(to_rows_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, in, c, f, p, sb, sc, sf, c_pad, (bf16*)out_sb, out_plane, out_f32);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_from_rows(const float* rows, int ld, int b, int c, int f, int p, float* out, void* stream) {
    if (!rows || !out || ld < c) return LFDM_E_BADARG;
    dim3 grid((p + 31) / 32, (c + 31) / 32, b * f);
    LFDM_LAUNCH_PDL(from_rows_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, rows, ld, c, f, p, out);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_im2col_small(const float* in, int b, int c, int f, int h, int w, int ksize, int pad, int k_pad,
                                 void* out_sb, int64_t out_plane, void* stream) {
    if (!in || !out_sb || k_pad < ksize * ksize * c || (k_pad & 3)) return LFDM_E_BADARG;
    if ((int64_t)b * f * h * w >= (1ll << 31)) return LFDM_E_BADARG;
    int64_t total4 = (int64_t)b * f * h * w * (k_pad >> 2);
    const int kq_per_row = k_pad >> 2;
    const size_t patch_bytes = (size_t)c * ksize * (w + 2 * pad) * sizeof(float);
    if (kq_per_row <= 256 && patch_bytes <= 40 * 1024) {
        const int pstep = 256 / kq_per_row;
        LFDM_LAUNCH_PDL(im2col_row_kernel, dim3((unsigned)(b * f * h)), dim3(pstep * kq_per_row), patch_bytes, (cudaStream_t)stream,
                        in, c, f, h, w, ksize, pad, k_pad, (bf16*)out_sb, out_plane);
    } else
    im2col_small_kernel<<<grid_for(total4), 256, 0, (cudaStream_t)stream>>>(in, c, f, h, w, ksize, pad, k_pad,
                                                                         (bf16*)out_sb, out_plane, total4);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_avgpool2_rows(const float* in, int n, int h, int w, int c, float* out_f32, void* out_sb,
                                  int64_t out_plane, void* stream) {
    if (!in || (c & 3) || (h & 1) || (w & 1)) return LFDM_E_BADARG;
    int64_t total = (int64_t)n * (h / 2) * (w / 2) * (c / 4);
    avgpool2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(in, h, w, c, out_f32, (bf16*)out_sb, out_plane, total);
    LFDM_CHECK_LAUNCH();
    return 0;
}

static int unet_heads_launch(const float* a, const float* wa, const float* ba, int na, const float* o, const float* wo,
                             const float* bo, int no, int c, int b, int f, int p, float* out, int64_t null_off, float cfg,
                             void* stream) {
    if (!a || !o || !wa || !wo || !out) return LFDM_E_BADARG;
    int64_t m = (int64_t)b * f * p;
    int64_t blocks = (m + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (c == 64 && na + no <= 4 && na >= 0 && no >= 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(o) |
                                                         reinterpret_cast<uintptr_t>(wa) | reinterpret_cast<uintptr_t>(wo)) & 15) == 0) {
        int64_t b4 = (m + 31) / 32;            // 8 warps x 4 rows per pass
        if (b4 > 148 * 8) b4 = 148 * 8;
        LFDM_LAUNCH_PDL(unet_heads64_kernel, dim3((unsigned)b4), dim3(256), 0, (cudaStream_t)stream, a, wa, ba, na, o, wo, bo, no, f, p, m, out, null_off, cfg);
    } else
    unet_heads_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, wa, ba, na, o, wo, bo, no, c, f, p, m, out, null_off, cfg);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_unet_heads(const float* a, const float* wa, const float* ba, int na, const float* o,
                               const float* wo, const float* bo, int no, int c, int b, int f, int p, float* out,
                               void* stream) {
    return unet_heads_launch(a, wa, ba, na, o, wo, bo, no, c, b, f, p, out, 0, 1.f, stream);
}

extern "C" int lfdm_unet_heads_cfg(const float* a, const float* wa, const float* ba, int na, const float* o,
                                   const float* wo, const float* bo, int no, int c, int b, int f, int p, float cond_scale,
                                   float* out, void* stream) {
    if (b < 1) return LFDM_E_BADARG;
    return unet_heads_launch(a, wa, ba, na, o, wo, bo, no, c, b, f, p, out, (int64_t)b * f * p, cond_scale, stream);
}

extern "C" int lfdm_pad_replicate_rows(const void* in_sb, int64_t in_plane, void* out_sb, int64_t out_plane, int n, int h, int w,
                                       int c, void* stream) {
    if (!in_sb || !out_sb || (c & 7) || (in_plane & 7) || (out_plane & 7)) return LFDM_E_BADARG;
    const int64_t total = (int64_t)n * (h + 2) * (w + 2) * (c / 8);
    pad_replicate_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)in_sb, in_plane / 8, (uint4*)out_sb,
                                                                            out_plane / 8, n, h, w, c / 8);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_split_bf16(const float* in, void* out_sb, int64_t out_plane, int64_t n, void* stream) {
    if (!in || !out_sb) return LFDM_E_BADARG;
    split_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, (bf16*)out_sb, out_plane, n);
    LFDM_CHECK_LAUNCH();
    return 0;
}
