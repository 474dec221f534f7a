// motion.cu — dense-motion / region kernels of the full LFAE branch (kernels K14/K15):
//   lfdm_antialias_down : AntiAliasInterpolation2d (LFAE/modules/util.py:217-264)
//   lfdm_region_moments : spatial softmax(T) -> mean / covariance -> closed-form 2x2 symmetric SVD on the device
//                         (RegionPredictor.forward region_predictor.py:84-117; replaces torch.svd(covar.cpu()) :21)
//   lfdm_motion_prep    : heatmap differences + sparse motions + 11x deformed source in ONE per-pixel kernel
//                         (PixelwiseFlowPredictor pixelwise_flow_predictor.py:48-102; util.py:22-48)
//   lfdm_motion_finish  : softmax over regions + weighted aggregation of the sparse motions + occlusion sigmoid
//                         (pixelwise_flow_predictor.py:124-135)
//   lfdm_rows_mean      : global average of a row matrix per image (BGMotionPredictor bg_motion_predictor.py:47)
#include "common.cuh"

namespace {

__device__ __forceinline__ float coord(int i, int n) { return 2.f * ((float)i / (float)(n - 1)) - 1.f; }   // util.py:59-60

__global__ void antialias_kernel(const float* __restrict__ in, const float* __restrict__ kern, float* __restrict__ out,
                                 int nc, int h, int w, int ks, int ka, int s) {
    const int ho = h / s + ((h % s) ? 1 : 0), wo = w / s + ((w % s) ? 1 : 0);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nc * ho * wo) return;
    int x = (int)(i % wo), y = (int)((i / wo) % ho);
    int64_t c = i / ((int64_t)ho * wo);
    const float* p = in + c * h * w;
    float acc = 0.f;
    for (int a = 0; a < ks; ++a) {
        int yy = y * s - ka + a;
        if (yy < 0 || yy >= h) continue;
        for (int b = 0; b < ks; ++b) {
            int xx = x * s - ka + b;
            if (xx < 0 || xx >= w) continue;
            acc = fmaf(p[yy * w + xx], kern[a * ks + b], acc);
        }
    }
    out[i] = acc;
}

// one block per (n, k): logits rows [(n*hw + p)][K]
__global__ void __launch_bounds__(256) region_moments_kernel(const float* __restrict__ logits, int K, int h, int w,
                                                             float inv_temp, float* __restrict__ heatmap,
                                                             float* __restrict__ shift, float* __restrict__ covar,
                                                             float* __restrict__ affine, float* __restrict__ u_out,
                                                             float* __restrict__ d_out) {
    __shared__ float red[8];
    __shared__ float bc[6];
    const int n = blockIdx.x / K, k = blockIdx.x % K;
    const int hw = h * w, tid = threadIdx.x;
    const float* base = logits + (int64_t)n * hw * K + k;
    auto block_reduce = [&](float v, bool is_max) -> float {
        v = is_max ? warp_max(v) : warp_sum(v);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        float r = red[0];
        for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
        return r;
    };
    float mx = -INFINITY;
    for (int p = tid; p < hw; p += 256) mx = fmaxf(mx, base[(int64_t)p * K] * inv_temp);
    mx = block_reduce(mx, true);
    float se = 0.f;
    for (int p = tid; p < hw; p += 256) se += expf(base[(int64_t)p * K] * inv_temp - mx);
    se = block_reduce(se, false);
    float mxs = 0.f, mys = 0.f;
    float* hm = heatmap + ((int64_t)n * K + k) * hw;
    for (int p = tid; p < hw; p += 256) {
        float r = expf(base[(int64_t)p * K] * inv_temp - mx) / se;
        hm[p] = r;
        mxs += r * coord(p % w, w);
        mys += r * coord(p / w, h);
    }
    mxs = block_reduce(mxs, false);
    mys = block_reduce(mys, false);
    float cxx = 0.f, cxy = 0.f, cyy = 0.f;
    for (int p = tid; p < hw; p += 256) {
        float r = hm[p];
        float dx = coord(p % w, w) - mxs, dy = coord(p / w, h) - mys;
        cxx += dx * dx * r; cxy += dx * dy * r; cyy += dy * dy * r;
    }
    cxx = block_reduce(cxx, false);
    cxy = block_reduce(cxy, false);
    cyy = block_reduce(cyy, false);
    if (tid == 0) {
        const int64_t o = (int64_t)n * K + k;
        shift[o * 2] = mxs; shift[o * 2 + 1] = mys;
        covar[o * 4] = cxx; covar[o * 4 + 1] = cxy; covar[o * 4 + 2] = cxy; covar[o * 4 + 3] = cyy;
        // symmetric PSD 2x2: eigen-decomposition == SVD.  l1 >= l2 >= 0
        float hd = 0.5f * (cxx - cyy), mid = 0.5f * (cxx + cyy);
        float rad = sqrtf(hd * hd + cxy * cxy);
        float l1 = mid + rad, l2 = fmaxf(mid - rad, 0.f);
        float vx, vy;   // eigenvector of l1
        if (fabsf(cxy) > 1e-30f || hd != 0.f) {
            if (hd >= 0.f) { vx = hd + rad; vy = cxy; } else { vx = cxy; vy = rad - hd; }
            float nrm = sqrtf(vx * vx + vy * vy);
            if (nrm > 0.f) { vx /= nrm; vy /= nrm; } else { vx = 1.f; vy = 0.f; }
        } else { vx = 1.f; vy = 0.f; }
        // sign convention (the reference's LAPACK choice is data dependent, region_predictor.py:21): U is a
        // reflection (det = -1) with a non-positive first entry, the most frequent LAPACK outcome
        if (vx > 0.f) { vx = -vx; vy = -vy; }
        float u00 = vx, u10 = vy, u01 = vy, u11 = -vx;
        float s1 = sqrtf(l1), s2 = sqrtf(l2);
        u_out[o * 4] = u00; u_out[o * 4 + 1] = u01; u_out[o * 4 + 2] = u10; u_out[o * 4 + 3] = u11;
        d_out[o * 4] = s1; d_out[o * 4 + 1] = 0.f; d_out[o * 4 + 2] = 0.f; d_out[o * 4 + 3] = s2;
        affine[o * 4] = u00 * s1; affine[o * 4 + 1] = u01 * s2; affine[o * 4 + 2] = u10 * s1; affine[o * 4 + 3] = u11 * s2;
    }
}

struct RegionP { float icd[4], ics[4], aff[4], sd[2], ss[2]; };

__device__ __forceinline__ void inv2(const float* m, float* o) {
    float det = m[0] * m[3] - m[1] * m[2];
    float r = 1.f / det;
    o[0] = m[3] * r; o[1] = -m[1] * r; o[2] = -m[2] * r; o[3] = m[0] * r;
}

// grid (ceil(hw/128), N); down: (N,3,h,w) planar
__global__ void __launch_bounds__(128) motion_prep_kernel(const float* __restrict__ down, const float* __restrict__ d_shift,
                                                          const float* __restrict__ d_covar, const float* __restrict__ d_aff,
                                                          const float* __restrict__ s_shift, const float* __restrict__ s_covar,
                                                          const float* __restrict__ s_aff, const float* __restrict__ bg,
                                                          int K, int h, int w, int revert_axis_swap, int nch,
                                                          float* __restrict__ hg_in, float* __restrict__ sparse) {
    __shared__ RegionP rp[32];
    __shared__ float sbg[9];
    const int n = blockIdx.y, tid = threadIdx.x, hw = h * w;
    if (tid < K) {
        const int64_t o = (int64_t)n * K + tid;
        RegionP r;
        inv2(d_covar + o * 4, r.icd);
        inv2(s_covar + o * 4, r.ics);
        if (d_aff) {
            float id[4];
            inv2(d_aff + o * 4, id);
            const float* a = s_aff + o * 4;
            r.aff[0] = a[0] * id[0] + a[1] * id[2]; r.aff[1] = a[0] * id[1] + a[1] * id[3];
            r.aff[2] = a[2] * id[0] + a[3] * id[2]; r.aff[3] = a[2] * id[1] + a[3] * id[3];
            if (revert_axis_swap) {
                float sg = r.aff[0] > 0.f ? 1.f : (r.aff[0] < 0.f ? -1.f : 0.f);
                for (int i = 0; i < 4; ++i) r.aff[i] *= sg;
            }
        } else { r.aff[0] = 1.f; r.aff[1] = 0.f; r.aff[2] = 0.f; r.aff[3] = 1.f; }
        r.sd[0] = d_shift[o * 2]; r.sd[1] = d_shift[o * 2 + 1];
        r.ss[0] = s_shift[o * 2]; r.ss[1] = s_shift[o * 2 + 1];
        rp[tid] = r;
    }
    if (tid < 9) sbg[tid] = bg ? bg[n * 9 + tid] : ((tid % 4 == 0) ? 1.f : 0.f);
    __syncthreads();
    const int p = blockIdx.x * 128 + tid;
    if (p >= hw) return;
    const int x = p % w, y = p / w;
    const float gx = coord(x, w), gy = coord(y, h);
    const int cin = (K + 1) * (nch + 1);
    float* row = hg_in + ((int64_t)n * hw + p) * cin;
    const float* img = down + (int64_t)n * nch * hw;
    for (int k = 0; k <= K; ++k) {
        float heat = 0.f, mx, my;
        if (k == 0) {
            float X = sbg[0] * gx + sbg[1] * gy + sbg[2];
            float Y = sbg[3] * gx + sbg[4] * gy + sbg[5];
            float Z = sbg[6] * gx + sbg[7] * gy + sbg[8];
            mx = X / Z; my = Y / Z;
        } else {
            const RegionP& r = rp[k - 1];
            float dx = gx - r.sd[0], dy = gy - r.sd[1];
            float qd = dx * (r.icd[0] * dx + r.icd[1] * dy) + dy * (r.icd[2] * dx + r.icd[3] * dy);
            float sx = gx - r.ss[0], sy = gy - r.ss[1];
            float qs = sx * (r.ics[0] * sx + r.ics[1] * sy) + sy * (r.ics[2] * sx + r.ics[3] * sy);
            heat = expf(-0.5f * qd) - expf(-0.5f * qs);
            mx = r.aff[0] * dx + r.aff[1] * dy + r.ss[0];
            my = r.aff[2] * dx + r.aff[3] * dy + r.ss[1];
        }
        float* sp = sparse + ((((int64_t)n * (K + 1) + k) * h + y) * w + x) * 2;
        sp[0] = mx; sp[1] = my;
        row[k * (nch + 1)] = heat;
        // F.grid_sample(bilinear, zeros, align_corners=False) of the down-sampled source
        float ix = ((mx + 1.f) * (float)w - 1.f) / 2.f, iy = ((my + 1.f) * (float)h - 1.f) / 2.f;
        float fx = floorf(ix), fy = floorf(iy);
        int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        float wnw = (fx + 1.f - ix) * (fy + 1.f - iy), wne = (ix - fx) * (fy + 1.f - iy);
        float wsw = (fx + 1.f - ix) * (iy - fy), wse = (ix - fx) * (iy - fy);
        bool vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h;
        for (int c = 0; c < nch; ++c) {
            const float* ip = img + (int64_t)c * hw;
            float a = 0.f;
            if (vy0 && vx0) a += ip[y0 * w + x0] * wnw;
            if (vy0 && vx1) a += ip[y0 * w + x1] * wne;
            if (vy1 && vx0) a += ip[y1 * w + x0] * wsw;
            if (vy1 && vx1) a += ip[y1 * w + x1] * wse;
            row[k * (nch + 1) + 1 + c] = a;
        }
    }
}

// logits rows [N*hw][K+1 (+1 occlusion)] -> flow (N,h,w,2), occ (N,1,h,w)
__global__ void motion_finish_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ sparse, int K,
                                     int hw, int64_t total, int has_occ, float* __restrict__ flow, float* __restrict__ occ) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t n = i / hw;
    const int p = (int)(i % hw);
    const float* l = logits + i * ld;
    float mx = -INFINITY;
    for (int k = 0; k <= K; ++k) mx = fmaxf(mx, l[k]);
    float se = 0.f;
    for (int k = 0; k <= K; ++k) se += expf(l[k] - mx);
    float fx = 0.f, fy = 0.f;
    for (int k = 0; k <= K; ++k) {
        float m = expf(l[k] - mx) / se;
        const float* sp = sparse + (((int64_t)n * (K + 1) + k) * hw + p) * 2;
        fx += sp[0] * m; fy += sp[1] * m;
    }
    flow[i * 2] = fx; flow[i * 2 + 1] = fy;
    if (has_occ) occ[i] = 1.f / (1.f + expf(-l[K + 1]));
}

__global__ void __launch_bounds__(256) rows_mean_kernel(const float* __restrict__ rows, int p, int c, float* __restrict__ out) {
    const int n = blockIdx.x;
    for (int ch = threadIdx.x; ch < c; ch += 256) {
        float a = 0.f;
        for (int i = 0; i < p; ++i) a += rows[((int64_t)n * p + i) * c + ch];
        out[(int64_t)n * c + ch] = a / (float)p;
    }
}

}  // namespace

extern "C" int lfdm_antialias_down(const float* in, const float* kern, float* out, int n, int c, int h, int w, int ks,
                                   int ka, int s, void* stream) {
    if (!in || !kern || !out || s <= 0) return LFDM_E_BADARG;
    int ho = (h + s - 1) / s, wo = (w + s - 1) / s;
    int64_t total = (int64_t)n * c * ho * wo;
    antialias_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, kern, out, n * c, h, w, ks, ka, s);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_region_moments(const float* logits, int n, int k, int h, int w, float temperature, float* heatmap,
                                   float* shift, float* covar, float* affine, float* u, float* d, void* stream) {
    if (!logits || !heatmap || !shift || !covar || !affine || !u || !d || k <= 0) return LFDM_E_BADARG;
    region_moments_kernel<<<n * k, 256, 0, (cudaStream_t)stream>>>(logits, k, h, w, 1.f / temperature, heatmap, shift,
                                                                 covar, affine, u, d);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_motion_prep(const float* down, const float* d_shift, const float* d_covar, const float* d_affine,
                                const float* s_shift, const float* s_covar, const float* s_affine, const float* bg, int n,
                                int k, int nch, int h, int w, int revert_axis_swap, float* hg_in, float* sparse,
                                void* stream) {
    if (!down || !d_shift || !d_covar || !s_shift || !s_covar || !hg_in || !sparse || k > 32) return LFDM_E_BADARG;
    if ((d_affine == nullptr) != (s_affine == nullptr)) return LFDM_E_BADARG;
    dim3 grid((h * w + 127) / 128, n);
    motion_prep_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(down, d_shift, d_covar, d_affine, s_shift, s_covar, s_affine,
                                                              bg, k, h, w, revert_axis_swap, nch, hg_in, sparse);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_motion_finish(const float* logits, int ld, const float* sparse, int n, int k, int hw, int has_occ,
                                  float* flow, float* occ, void* stream) {
    if (!logits || !sparse || !flow || (has_occ && !occ)) return LFDM_E_BADARG;
    int64_t total = (int64_t)n * hw;
    motion_finish_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(logits, ld, sparse, k, hw, total,
                                                                                         has_occ, flow, occ);
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_rows_mean(const float* rows, int n, int p, int c, float* out, void* stream) {
    if (!rows || !out) return LFDM_E_BADARG;
    rows_mean_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(rows, p, c, out);
    LFDM_CHECK_LAUNCH();
    return 0;
}
