// render.cu — output stage of the demo scripts on the GPU: uint8 quantisation of the generated videos and
// the 5-panel frame [source | generated | warped source | sampling-grid figure | confidence] of demo/demo_mug.py:126-145
// (sample_img :68-74, misc.conf2fig misc.py:76-80, misc.grid2fig misc.py:44-63).  At B200 sampling speed the reference's
// per-frame matplotlib figure + PIL paste dominates the demo; here one sample's 40 frames are composed by three small kernels
// and leave the device as one uint8 tensor (the host only encodes the GIF, asynchronously).
//   * photographic panels: clamp(x + mean/255, 0, 1) * 255 truncated to uint8, exactly sample_img;
//   * confidence panel: nearest up-sampling of conf * 255 truncated, exactly conf2fig;
//   * grid panel: the identity grid (light grey) and the warped sampling grid (matplotlib "C0" blue) as anti-aliased polylines on
//     white, axes scaled to the data limits with matplotlib's 5 % margins, y pointing up as in the reference figure.  This is a
//     rasterisation, not matplotlib: it is visually equivalent, not pixel-identical (documented in DESIGN.md).
#include "common.cuh"

namespace {

// limits[f] = {xmin, xmax, ymin, ymax} of identity grid U warped grid, with 5 % margins (matplotlib autoscale default)
__global__ void grid_limits_kernel(const float* __restrict__ grid, int f_total, int hw, float* __restrict__ limits) {
    const int f = blockIdx.x;
    const float* gx = grid + (int64_t)f * hw;                       // grid layout (2, F, h, w)
    const float* gy = grid + ((int64_t)f_total + f) * hw;
    float x0 = -1.f, x1 = 1.f, y0 = -1.f, y1 = 1.f;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        const float x = gx[i], y = gy[i];
        x0 = fminf(x0, x); x1 = fmaxf(x1, x); y0 = fminf(y0, y); y1 = fmaxf(y1, y);
    }
    __shared__ float s[4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if (lane == 0) { s[0][w] = x0; s[1][w] = x1; s[2][w] = y0; s[3][w] = y1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int i = 1; i < nw; ++i) {
            x0 = fminf(x0, s[0][i]); x1 = fmaxf(x1, s[1][i]); y0 = fminf(y0, s[2][i]); y1 = fmaxf(y1, s[3][i]);
        }
        const float mx = 0.05f * (x1 - x0), my = 0.05f * (y1 - y0);
        limits[f * 4 + 0] = x0 - mx; limits[f * 4 + 1] = x1 + mx; limits[f * 4 + 2] = y0 - my; limits[f * 4 + 3] = y1 + my;
    }
}

// one thread per polyline segment (identity grid: layer 0, warped grid: layer 1): anti-aliased coverage, atomicMax per pixel
__global__ void grid_segments_kernel(const float* __restrict__ grid, const float* __restrict__ limits, int f_total, int h, int w,
                                     int H, int W, float half_width, unsigned int* __restrict__ cover) {
    const int per_dir = h * (w - 1) + (h - 1) * w;     // horizontal + vertical segments of one grid
    const int64_t total = (int64_t)f_total * 2 * per_dir;
    const int hw = h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / (2 * per_dir));
        int rem = (int)(i - (int64_t)f * 2 * per_dir);
        const int layer = rem / per_dir;
        rem -= layer * per_dir;
        int ia, ib;
        if (rem < h * (w - 1)) { const int y = rem / (w - 1), x = rem % (w - 1); ia = y * w + x; ib = ia + 1; }
        else { rem -= h * (w - 1); const int y = rem / w, x = rem % w; ia = y * w + x; ib = ia + w; }
        float ax, ay, bx, by;
        if (layer == 0) {                               // identity grid: linspace(-1, 1) in both directions (misc.py:47-50)
            ax = -1.f + 2.f * (ia % w) / (float)(w - 1); ay = -1.f + 2.f * (ia / w) / (float)(h - 1);
            bx = -1.f + 2.f * (ib % w) / (float)(w - 1); by = -1.f + 2.f * (ib / w) / (float)(h - 1);
        } else {
            ax = grid[(int64_t)f * hw + ia]; ay = grid[((int64_t)f_total + f) * hw + ia];
            bx = grid[(int64_t)f * hw + ib]; by = grid[((int64_t)f_total + f) * hw + ib];
        }
        const float* lim = limits + f * 4;
        const float sx = (float)W / (lim[1] - lim[0]), sy = (float)H / (lim[3] - lim[2]);
        // data -> pixel centres; matplotlib's y axis points up
        const float pax = (ax - lim[0]) * sx - 0.5f, pay = (lim[3] - ay) * sy - 0.5f;
        const float pbx = (bx - lim[0]) * sx - 0.5f, pby = (lim[3] - by) * sy - 0.5f;
        const float pad = half_width + 1.f;
        const int x_lo = max(0, (int)floorf(fminf(pax, pbx) - pad)), x_hi = min(W - 1, (int)ceilf(fmaxf(pax, pbx) + pad));
        const int y_lo = max(0, (int)floorf(fminf(pay, pby) - pad)), y_hi = min(H - 1, (int)ceilf(fmaxf(pay, pby) + pad));
        const float dx = pbx - pax, dy = pby - pay;
        const float inv_len2 = 1.f / fmaxf(dx * dx + dy * dy, 1e-12f);
        unsigned int* cv = cover + ((int64_t)(f * 2 + layer) * H) * W;
        for (int y = y_lo; y <= y_hi; ++y)
            for (int x = x_lo; x <= x_hi; ++x) {
                const float t = fminf(fmaxf(((x - pax) * dx + (y - pay) * dy) * inv_len2, 0.f), 1.f);
                const float ex = pax + t * dx - x, ey = pay + t * dy - y;
                const float c = fminf(fmaxf(half_width + 0.5f - sqrtf(ex * ex + ey * ey), 0.f), 1.f);
                if (c > 0.f) atomicMax(cv + (int64_t)y * W + x, (unsigned int)(c * 255.f + 0.5f));
            }
    }
}

__device__ __forceinline__ uint8_t quant(float v, float m) {      // sample_img (demo_mug.py:68-74)
    float t = v + m;
    t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    return (uint8_t)(t * 255.f);
}

__global__ void compose_panels_kernel(const float* __restrict__ src, const float* __restrict__ out_vid,
                                      const float* __restrict__ warped_vid, const float* __restrict__ conf,
                                      const unsigned int* __restrict__ cover, float m0, float m1, float m2, int f_total, int H,
                                      int W, int h, int w, uint8_t* __restrict__ out) {
    const int64_t total = (int64_t)f_total * H * 5 * W;
    const int64_t plane = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x5 = (int)(i % (5 * W));
        const int y = (int)((i / (5 * W)) % H);
        const int f = (int)(i / ((int64_t)5 * W * H));
        const int panel = x5 / W, x = x5 - panel * W;
        uint8_t r, g, b;
        if (panel == 0) {
            const int64_t o = (int64_t)y * W + x;
            r = quant(src[o], m0); g = quant(src[plane + o], m1); b = quant(src[2 * plane + o], m2);
        } else if (panel <= 2) {
            const float* v = panel == 1 ? out_vid : warped_vid;     // (3, F, H, W)
            const int64_t o = ((int64_t)f * H + y) * W + x;
            r = quant(v[o], m0); g = quant(v[(int64_t)f_total * plane + o], m1); b = quant(v[2 * (int64_t)f_total * plane + o], m2);
        } else if (panel == 3) {
            const float c0 = cover[((int64_t)(f * 2) * H + y) * W + x] * (1.f / 255.f);        // identity grid, light grey
            const float c1 = cover[((int64_t)(f * 2 + 1) * H + y) * W + x] * (1.f / 255.f);    // warped grid, C0
            float cr = 255.f, cg = 255.f, cb = 255.f;
            cr += c0 * (211.f - cr); cg += c0 * (211.f - cg); cb += c0 * (211.f - cb);
            cr += c1 * (31.f - cr); cg += c1 * (119.f - cg); cb += c1 * (180.f - cb);
            r = (uint8_t)(cr + 0.5f); g = (uint8_t)(cg + 0.5f); b = (uint8_t)(cb + 0.5f);
        } else {
            const int sy = min(h - 1, (int)floorf(y * ((float)h / (float)H))), sx = min(w - 1, (int)floorf(x * ((float)w / (float)W)));
            r = g = b = (uint8_t)(conf[((int64_t)f * h + sy) * w + sx] * 255.f);                 // misc.conf2fig
        }
        out[i * 3] = r; out[i * 3 + 1] = g; out[i * 3 + 2] = b;
    }
}

}  // namespace

extern "C" int lfdm_render_panels(const float* src, const float* out_vid, const float* warped_vid, const float* grid,
                                  const float* conf, const float* mean3, int f, int H, int W, int h, int w, float line_width,
                                  void* workspace, uint8_t* out, void* stream) {
    if (!src || !out_vid || !warped_vid || !grid || !conf || !workspace || !out || f < 1 || h < 2 || w < 2) return LFDM_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    float* limits = reinterpret_cast<float*>(workspace);                       // [f][4]
    unsigned int* cover = reinterpret_cast<unsigned int*>(limits + (((int64_t)f * 4 + 63) / 64) * 64);   // [f][2][H][W]
    cudaError_t e = cudaMemsetAsync(cover, 0, (size_t)f * 2 * H * W * sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    grid_limits_kernel<<<f, 256, 0, st>>>(grid, f, h * w, limits);
    const int64_t nseg = (int64_t)f * 2 * (h * (w - 1) + (h - 1) * w);
    grid_segments_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(grid, limits, f, h, w, H, W, 0.5f * line_width, cover);
    const float m0 = mean3 ? mean3[0] : 0.f, m1 = mean3 ? mean3[1] : 0.f, m2 = mean3 ? mean3[2] : 0.f;   // host pointer
    const int64_t total = (int64_t)f * H * 5 * W;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    compose_panels_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, out_vid, warped_vid, conf, cover, m0, m1, m2, f, H, W, h, w, out);
    LFDM_CHECK_LAUNCH();
    return 0;
}
