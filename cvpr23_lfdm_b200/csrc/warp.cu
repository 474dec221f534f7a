// warp.cu — K12: latent-flow warp + occlusion blend of the LFAE decoder (HBM-bandwidth-bound).
// One kernel replaces the reference's 7-kernel chain per call site
//   F.interpolate(flow, bilinear) -> F.grid_sample(bilinear, zeros, align_corners=False) -> F.interpolate(occ) ->
//   skip*occ + prev*(1-occ)                       (LFAE/modules/generator.py:60-88, call sites :147,149,154,157,162)
// The 32x32 flow / occlusion latents are up-sampled on the fly (never materialised); the source feature map is
// channels-last so each bilinear tap is a contiguous float4 run; every output element is written exactly once.
#include "common.cuh"

namespace {

struct Taps {
    int x0, x1, y0, y1;          // clamped-valid indices are checked via the v* flags
    float wnw, wne, wsw, wse;
    bool vx0, vx1, vy0, vy1;
};

// torch upsample_bilinear2d, align_corners=False: src = max(0, scale*(dst+0.5)-0.5)
__device__ __forceinline__ void up_coord(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

__device__ __forceinline__ float bil4(float v00, float v01, float v10, float v11, float lh0, float lh1, float lw0,
                                      float lw1) {
    return lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11);
}

// flow/occ sampling at output pixel (y,x) of an (hs,ws) map from (hf,wf) latents; returns grid coords + occlusion
__device__ __forceinline__ void latent_at(const float* __restrict__ flow, const float* __restrict__ occ, int64_t n,
                                          int y, int x, int hs, int ws, int hf, int wf, float& gx, float& gy,
                                          float& oc) {
    int y0, y1, x0, x1;
    float lh0, lh1, lw0, lw1;
    up_coord(y, (float)hf / (float)hs, hf, y0, y1, lh0, lh1);
    up_coord(x, (float)wf / (float)ws, wf, x0, x1, lw0, lw1);
    const float* f = flow + n * hf * wf * 2;
    const float2 f00 = *reinterpret_cast<const float2*>(f + (y0 * wf + x0) * 2);
    const float2 f01 = *reinterpret_cast<const float2*>(f + (y0 * wf + x1) * 2);
    const float2 f10 = *reinterpret_cast<const float2*>(f + (y1 * wf + x0) * 2);
    const float2 f11 = *reinterpret_cast<const float2*>(f + (y1 * wf + x1) * 2);
    gx = bil4(f00.x, f01.x, f10.x, f11.x, lh0, lh1, lw0, lw1);
    gy = bil4(f00.y, f01.y, f10.y, f11.y, lh0, lh1, lw0, lw1);
    oc = 1.f;
    if (occ) {
        const float* o = occ + n * hf * wf;
        oc = bil4(o[y0 * wf + x0], o[y0 * wf + x1], o[y1 * wf + x0], o[y1 * wf + x1], lh0, lh1, lw0, lw1);
    }
}

// torch grid_sampler_2d bilinear / zeros / align_corners=False
__device__ __forceinline__ Taps make_taps(float gx, float gy, int hs, int ws) {
    Taps t;
    float ix = ((gx + 1.f) * (float)ws - 1.f) / 2.f;
    float iy = ((gy + 1.f) * (float)hs - 1.f) / 2.f;
    float fx = floorf(ix), fy = floorf(iy);
    t.x0 = (int)fx; t.y0 = (int)fy; t.x1 = t.x0 + 1; t.y1 = t.y0 + 1;
    float ex = fx + 1.f, ey = fy + 1.f;
    t.wnw = (ex - ix) * (ey - iy);
    t.wne = (ix - fx) * (ey - iy);
    t.wsw = (ex - ix) * (iy - fy);
    t.wse = (ix - fx) * (iy - fy);
    t.vx0 = t.x0 >= 0 && t.x0 < ws; t.vx1 = t.x1 >= 0 && t.x1 < ws;
    t.vy0 = t.y0 >= 0 && t.y0 < hs; t.vy1 = t.y1 >= 0 && t.y1 < hs;
    return t;
}

__device__ __forceinline__ float4 fma4(float4 acc, float4 v, float w) {
    acc.x += v.x * w; acc.y += v.y * w; acc.z += v.z * w; acc.w += v.w * w;
    return acc;
}

// One warp owns 32 consecutive output pixels.  Phase 1: lane p computes the per-pixel set-up ONCE (up-sampled flow and
// occlusion, 4 tap offsets + weights) — per pixel, not per float4 as a naive mapping would.  Phase 2: the warp walks the
// pixels, broadcasting a pixel's set-up with shuffles while the lanes sweep its channels with 16-byte accesses, so
// every load/store instruction covers contiguous 128..512 B runs of the channels-last rows.
__global__ void __launch_bounds__(256) warp_rows_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                        const float* __restrict__ occ, const float* __restrict__ prev,
                                                        float* __restrict__ out_f32, bf16* __restrict__ out_sb,
                                                        int64_t out_plane, const float* __restrict__ sb_scale,
                                                        const float* __restrict__ sb_shift, int sb_act, int64_t n_img,
                                                        int frames_per_src, int hs, int ws, int c, int hf, int wf) {
    const int lane = threadIdx.x & 31;
    const int qpp = c >> 2;                                  // float4 per pixel
    int lpp = 1;                                             // lanes per pixel: largest power of two <= min(32, qpp)
    while (lpp * 2 <= qpp && lpp < 32) lpp *= 2;
    const int ppi = 32 / lpp;                                // pixels per iteration
    const int64_t total_pix = n_img * hs * ws;
    const int64_t n_groups = (total_pix + 31) >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int hw = hs * ws;
    for (int64_t g = warp0; g < n_groups; g += nwarps) {
        // ---- phase 1: per-pixel set-up (lane = pixel)
        const int64_t pix = (g << 5) + lane;
        int off[4] = {-1, -1, -1, -1};
        float wgt[4] = {0.f, 0.f, 0.f, 0.f};
        float oc = 1.f;
        if (pix < total_pix) {
            const int n = (int)(pix / hw);
            const int rem = (int)(pix - (int64_t)n * hw);
            const int y = rem / ws, x = rem - y * ws;
            float gx, gy;
            latent_at(flow, occ, n, y, x, hs, ws, hf, wf, gx, gy, oc);
            const Taps t = make_taps(gx, gy, hs, ws);
            const int sbase = (n / frames_per_src) * hw;     // source image row offset (in pixels)
            if (t.vy0 && t.vx0) off[0] = sbase + t.y0 * ws + t.x0;
            if (t.vy0 && t.vx1) off[1] = sbase + t.y0 * ws + t.x1;
            if (t.vy1 && t.vx0) off[2] = sbase + t.y1 * ws + t.x0;
            if (t.vy1 && t.vx1) off[3] = sbase + t.y1 * ws + t.x1;
            wgt[0] = t.wnw; wgt[1] = t.wne; wgt[2] = t.wsw; wgt[3] = t.wse;
        }
        // ---- phase 2: sweep channels
        const int sub = lane / lpp, ql = lane - sub * lpp;
        for (int it = 0; it < 32; it += ppi) {
            const int pl = it + sub;                          // pixel (lane index) this lane works on
            const int o0 = __shfl_sync(0xffffffffu, off[0], pl), o1 = __shfl_sync(0xffffffffu, off[1], pl);
            const int o2 = __shfl_sync(0xffffffffu, off[2], pl), o3 = __shfl_sync(0xffffffffu, off[3], pl);
            const float w0 = __shfl_sync(0xffffffffu, wgt[0], pl), w1 = __shfl_sync(0xffffffffu, wgt[1], pl);
            const float w2 = __shfl_sync(0xffffffffu, wgt[2], pl), w3 = __shfl_sync(0xffffffffu, wgt[3], pl);
            const float ocp = __shfl_sync(0xffffffffu, oc, pl);
            const int64_t p = (g << 5) + pl;
            if (p >= total_pix) continue;                     // uniform per sub-group; shuffles above are done by all
            for (int q = ql; q < qpp; q += lpp) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o0 >= 0) acc = fma4(acc, *reinterpret_cast<const float4*>(src + (int64_t)o0 * c + q * 4), w0);
                if (o1 >= 0) acc = fma4(acc, *reinterpret_cast<const float4*>(src + (int64_t)o1 * c + q * 4), w1);
                if (o2 >= 0) acc = fma4(acc, *reinterpret_cast<const float4*>(src + (int64_t)o2 * c + q * 4), w2);
                if (o3 >= 0) acc = fma4(acc, *reinterpret_cast<const float4*>(src + (int64_t)o3 * c + q * 4), w3);
                const int64_t o = p * c + q * 4;
                float4 r;
                if (occ) {
                    r = make_float4(acc.x * ocp, acc.y * ocp, acc.z * ocp, acc.w * ocp);
                    if (prev) {
                        const float4 pv = *reinterpret_cast<const float4*>(prev + o);
                        const float om = 1.f - ocp;
                        r.x += pv.x * om; r.y += pv.y * om; r.z += pv.z * om; r.w += pv.w * om;
                    }
                } else {
                    r = acc;
                }
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = r;
                if (out_sb) {
                    float u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cc = q * 4 + e;
                        float v = u[e];
                        if (sb_scale) v *= sb_scale[cc];
                        if (sb_shift) v += sb_shift[cc];
                        u[e] = apply_act(v, sb_act);
                    }
                    store_sb4(out_sb, out_plane, o, make_float4(u[0], u[1], u[2], u[3]));
                }
            }
        }
    }
}

// Batched variant for the channel counts of the decoder (c = 64 / 128 / 256): LPP lanes sweep one pixel with QPL float4 each
// and U pixel slots are processed together -- all 4*U*QPL tap loads and the U*QPL `prev` loads of a lane are issued before the
// first use, so a warp keeps >= 20 independent 16-byte loads in flight (the plain kernel above issues one pixel's loads, waits,
// stores, and is latency-bound at ~45 % of the HBM peak).
__device__ __forceinline__ float4 ldg_nc4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int LPP, int QPL, int U, int MINB>
__global__ void __launch_bounds__(128, MINB) warp_rows_batched_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                                const float* __restrict__ occ, const float* __restrict__ prev,
                                                                float* __restrict__ out_f32, bf16* __restrict__ out_sb,
                                                                int64_t out_plane, const float* __restrict__ sb_scale,
                                                                const float* __restrict__ sb_shift, int sb_act, int64_t n_img,
                                                                int frames_per_src, int hs, int ws, int hf, int wf, int taps_l1, int prefetch) {
    constexpr int C = LPP * QPL * 4;
    constexpr int PPI = 32 / LPP;                            // pixels per slot row
    const int lane = threadIdx.x & 31;
    const int64_t total_pix = n_img * hs * ws;
    const int64_t n_groups = (total_pix + 31) >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int hw = hs * ws;
    const int sub = lane / LPP, ql = lane - sub * LPP;
    float4 sc4[QPL], sh4[QPL];
#pragma unroll
    for (int qq = 0; qq < QPL; ++qq) {
        sc4[qq] = sb_scale ? *reinterpret_cast<const float4*>(sb_scale + (ql + qq * LPP) * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
        sh4[qq] = sb_shift ? *reinterpret_cast<const float4*>(sb_shift + (ql + qq * LPP) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t g = warp0; g < n_groups; g += nwarps) {
        // the `prev` rows of this warp's NEXT group are the only loads that come from HBM (taps hit L1/L2): pull them into L2 now
        if (prefetch && prev && occ && g + nwarps < n_groups) {
            const char* nx = reinterpret_cast<const char*>(prev + ((g + nwarps) << 5) * C);
            const int64_t lim = (total_pix - ((g + nwarps) << 5)) * C * 4;
#pragma unroll
            for (int l = 0; l < (32 * C * 4) / (128 * 32); ++l) {
                const int64_t b = ((int64_t)l * 32 + lane) * 128;
                if (b < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + b));
            }
        }
        // ---- phase 1: per-pixel set-up (lane = pixel)
        const int64_t pix = (g << 5) + lane;
        int off[4] = {-1, -1, -1, -1};
        float wgt[4] = {0.f, 0.f, 0.f, 0.f};
        float oc = 1.f;
        if (pix < total_pix) {
            const int n = (int)(pix / hw);
            const int rem = (int)(pix - (int64_t)n * hw);
            const int y = rem / ws, x = rem - y * ws;
            float gx, gy;
            latent_at(flow, occ, n, y, x, hs, ws, hf, wf, gx, gy, oc);
            const Taps t = make_taps(gx, gy, hs, ws);
            const int sbase = (n / frames_per_src) * hw;
            if (t.vy0 && t.vx0) off[0] = sbase + t.y0 * ws + t.x0;
            if (t.vy0 && t.vx1) off[1] = sbase + t.y0 * ws + t.x1;
            if (t.vy1 && t.vx0) off[2] = sbase + t.y1 * ws + t.x0;
            if (t.vy1 && t.vx1) off[3] = sbase + t.y1 * ws + t.x1;
            wgt[0] = t.wnw; wgt[1] = t.wne; wgt[2] = t.wsw; wgt[3] = t.wse;
        }
        // ---- phase 2: U pixel slots at a time
#pragma unroll 1
        for (int it = 0; it < 32; it += PPI * U) {
            int o[U][4];
            float w[U][4], ocp[U];
            bool live[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pl = it + u * PPI + sub;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[u][k] = __shfl_sync(0xffffffffu, off[k], pl);
                    w[u][k] = __shfl_sync(0xffffffffu, wgt[k], pl);
                }
                ocp[u] = __shfl_sync(0xffffffffu, oc, pl);
                live[u] = ((g << 5) + pl) < total_pix;
            }
            float4 tap[U][QPL][4], pv[U][QPL];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int qq = 0; qq < QPL; ++qq) {
                    const int q4 = (ql + qq * LPP) * 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        // a flow field is smooth: neighbouring pixels share taps, so the source rows are worth keeping in L1
                        // (the streamed `prev` rows are not)
                        tap[u][qq][k] = (live[u] && o[u][k] >= 0)
                                            ? (taps_l1 ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)o[u][k] * C + q4))
                                                       : ldg_nc4(src + (int64_t)o[u][k] * C + q4))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    pv[u][qq] = (live[u] && prev && occ) ? ldg_nc4(prev + ((g << 5) + it + u * PPI + sub) * C + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!live[u]) continue;
                const int64_t p = (g << 5) + it + u * PPI + sub;
#pragma unroll
                for (int qq = 0; qq < QPL; ++qq) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc = fma4(acc, tap[u][qq][k], w[u][k]);     // invalid taps hold zeros (same sums as the plain kernel)
                    float4 r;
                    if (occ) {
                        r = make_float4(acc.x * ocp[u], acc.y * ocp[u], acc.z * ocp[u], acc.w * ocp[u]);
                        if (prev) {
                            const float om = 1.f - ocp[u];
                            r.x += pv[u][qq].x * om; r.y += pv[u][qq].y * om; r.z += pv[u][qq].z * om; r.w += pv[u][qq].w * om;
                        }
                    } else {
                        r = acc;
                    }
                    const int64_t oidx = p * C + (ql + qq * LPP) * 4;
                    if (out_f32) *reinterpret_cast<float4*>(out_f32 + oidx) = r;
                    if (out_sb) {
                        float4 t4 = r;
                        if (sb_scale) { t4.x *= sc4[qq].x; t4.y *= sc4[qq].y; t4.z *= sc4[qq].z; t4.w *= sc4[qq].w; }
                        if (sb_shift) { t4.x += sh4[qq].x; t4.y += sh4[qq].y; t4.z += sh4[qq].z; t4.w += sh4[qq].w; }
                        store_sb4(out_sb, out_plane, oidx,
                                  make_float4(apply_act(t4.x, sb_act), apply_act(t4.y, sb_act), apply_act(t4.z, sb_act), apply_act(t4.w, sb_act)));
                    }
                }
            }
        }
    }
}

// 3-channel planar image: one thread per output pixel
__global__ void __launch_bounds__(256) warp_image_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                         const float* __restrict__ occ, const float* __restrict__ prev,
                                                         int prev_ld, float* __restrict__ out, int b, int f, int h,
                                                         int w, int hf, int wf) {
    const int64_t hw = (int64_t)h * w;
    const int64_t total = (int64_t)b * f * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const int64_t n = i / hw;
        const int bi = (int)(n / f), fi = (int)(n % f);
        float gx, gy, oc;
        latent_at(flow, occ, n, y, x, h, w, hf, wf, gx, gy, oc);
        const Taps t = make_taps(gx, gy, h, w);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* sp = src + ((int64_t)bi * 3 + ch) * hw;
            float acc = 0.f;
            if (t.vy0 && t.vx0) acc += sp[(int64_t)t.y0 * w + t.x0] * t.wnw;
            if (t.vy0 && t.vx1) acc += sp[(int64_t)t.y0 * w + t.x1] * t.wne;
            if (t.vy1 && t.vx0) acc += sp[(int64_t)t.y1 * w + t.x0] * t.wsw;
            if (t.vy1 && t.vx1) acc += sp[(int64_t)t.y1 * w + t.x1] * t.wse;
            float r = acc;
            if (occ) {
                r = acc * oc;
                if (prev) r += prev[i * prev_ld + ch] * (1.f - oc);
            }
            out[(((int64_t)bi * 3 + ch) * f + fi) * hw + (int64_t)y * w + x] = r;
        }
    }
}

}  // namespace

extern "C" int lfdm_warp_blend_rows(const float* src, const float* flow, const float* occ, const float* prev,
                                    float* out_f32, void* out_sb, int64_t out_plane, const float* sb_scale,
                                    const float* sb_shift, int sb_act, int n, int frames_per_src, int hs, int ws, int c,
                                    int hf, int wf, void* stream) {
    if (!src || !flow || (c & 3) || frames_per_src <= 0 || (!out_f32 && !out_sb)) return LFDM_E_BADARG;
    int64_t groups = ((int64_t)n * hs * ws + 31) / 32;
    int64_t blocks = (groups + 7) / 8;                  // 8 warps per block, one 32-pixel group per warp iteration
    if (blocks > 148 * 16) blocks = 148 * 16;
    if ((int64_t)n * hs * ws >= (1ll << 31) / 1 || (int64_t)(n / frames_per_src + 1) * hs * ws >= (1ll << 31)) return LFDM_E_BADARG;
    static const int taps_l1 = (getenv("LFDM_WARP_TAPS_NO_L1") == nullptr) ? 1 : 0;      // A/B switch: taps bypass L1
    static const bool plain_env = (getenv("LFDM_WARP_PLAIN") != nullptr);       // A/B switch: one pixel slot at a time
    const bool plain = plain_env || (prev && (const void*)prev == (const void*)out_f32);   // in-place blend: no read-only path for prev
    const bool al16 = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(prev) | reinterpret_cast<uintptr_t>(sb_scale) |
                        reinterpret_cast<uintptr_t>(sb_shift)) & 15) == 0;
    // 128-thread blocks: at ~150 registers per thread three of them fit an SM (12 warps; a 256-thread block would sit alone with 8).
    // The kernel is bound by instruction issue and dependent-issue waits (ncu: 90 warp instructions per pixel, 3 warps per
    // scheduler), not by the memory system - see profiles/r02_ncu_warp_rows.md.
    static const int prefetch = (getenv("LFDM_WARP_NO_PREFETCH") == nullptr) ? 1 : 0;
    // slots per pass / blocks per SM: measured per channel count (profiles/r02_ncu_warp_rows.md); LFDM_WARP_VARIANT=0|1|2 forces
    // 4 slots x 3 blocks | 4 slots x 4 blocks (128 registers) | 2 slots x 5 blocks for every channel count
    static const int variant = getenv("LFDM_WARP_VARIANT") ? atoi(getenv("LFDM_WARP_VARIANT")) : -1;
    int64_t bblocks = (groups + 3) / 4;
    if (bblocks > 148 * 24) bblocks = 148 * 24;
#define LFDM_WARP_BATCHED(LPP, QPL, U, MINB)                                                                                  \
    warp_rows_batched_kernel<LPP, QPL, U, MINB><<<(unsigned)bblocks, 128, 0, (cudaStream_t)stream>>>(                           \
        src, flow, occ, prev, out_f32, (bf16*)out_sb, out_plane, sb_scale, sb_shift, sb_act, n, frames_per_src, hs, ws, hf, wf, taps_l1, \
        prefetch)
    const bool ok = !plain && al16;
    if (ok && c == 64 && variant == 1) LFDM_WARP_BATCHED(16, 1, 4, 4);
    else if (ok && c == 128 && variant == 1) LFDM_WARP_BATCHED(32, 1, 4, 4);
    else if (ok && c == 256 && variant == 1) LFDM_WARP_BATCHED(32, 2, 2, 4);
    else if (ok && c == 64 && variant == 2) LFDM_WARP_BATCHED(16, 1, 2, 5);
    else if (ok && c == 128 && variant == 2) LFDM_WARP_BATCHED(32, 1, 2, 5);
    else if (ok && c == 256 && variant == 2) LFDM_WARP_BATCHED(32, 2, 1, 5);
    else if (ok && c == 64 && variant == 0) LFDM_WARP_BATCHED(16, 1, 4, 3);
    else if (ok && c == 128 && variant == 0) LFDM_WARP_BATCHED(32, 1, 4, 3);
    else if (ok && c == 256 && variant == 0) LFDM_WARP_BATCHED(32, 2, 2, 3);
    else if (ok && c == 64) LFDM_WARP_BATCHED(16, 1, 2, 5);
    else if (ok && c == 128) LFDM_WARP_BATCHED(32, 1, 4, 4);
    else if (ok && c == 256) LFDM_WARP_BATCHED(32, 2, 1, 5);
    else
    warp_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, flow, occ, prev, out_f32, (bf16*)out_sb,
                                                                        out_plane, sb_scale, sb_shift, sb_act, n,
                                                                        frames_per_src, hs, ws, c, hf, wf);
#undef LFDM_WARP_BATCHED
    LFDM_CHECK_LAUNCH();
    return 0;
}

extern "C" int lfdm_warp_blend_image(const float* src, const float* flow, const float* occ, const float* prev,
                                     int prev_ld, float* out, int b, int f, int h, int w, int hf, int wf,
                                     void* stream) {
    if (!src || !flow || !out) return LFDM_E_BADARG;
    int64_t total = (int64_t)b * f * h * w;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    warp_image_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, flow, occ, prev, prev_ld, out, b, f, h, w,
                                                                         hf, wf);
    LFDM_CHECK_LAUNCH();
    return 0;
}
