/*
 * lfdm_b200.h — C-ABI of liblfdm_b200.so: the sm_100a kernels of the LFDM sampling + LFAE decode hot path.
 *
 * The reference (nihaomiao/CVPR23_LFDM) has no FFI layer: its "plugin boundary" for this path is the Python
 * nn.Module API.  The host-side mirror of that API lives in cvpr23_lfdm_b200/{dm,lfae}/ and calls
 * ONLY the entry points declared here (through ctypes; INTEGRATION.md shows the binding).  Each entry point cites the
 * reference call site(s) whose arithmetic it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless it says "host".
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns 0 on success, a
 *     cudaError_t value (>0) on a CUDA failure, or a negative LFDM_E_* code on an argument error.  No host sync,
 *     no allocation: all entry points are CUDA-graph capturable.
 *   - activations are "row matrices": X[M][C], row m = ((n*H + h)*W + w), n = b*F + f (frames are batch items,
 *     every 3-D conv of the reference has a (1,k,k) kernel).  Two storage formats:
 *       F32 : float  [M][C]
 *       SB  : split-bf16, two planes of __nv_bfloat16 [M][C]: hi = bf16(x), lo = bf16(x - hi)   (x ~ hi + lo,
 *             |err| <= 2^-17 |x|).  SB is what the tcgen05 GEMM consumes; hi*hi + hi*lo + lo*hi in fp32 TMEM
 *             accumulators reproduces fp32 products to ~2^-16 relative.  `plane` arguments give the element offset
 *             between the hi and the lo plane.
 */
#ifndef LFDM_B200_H
#define LFDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFDM_E_BADARG   (-1)
#define LFDM_E_UNSUPP   (-2)   /* shape not supported by the requested engine (caller may use the SIMT engine) */
#define LFDM_E_NODRIVER (-3)   /* cuTensorMapEncodeTiled not resolvable */

#define LFDM_ENGINE_SIMT 0     /* fp32 CUDA-core implicit GEMM, any shape                                   */
#define LFDM_ENGINE_TC   1     /* tcgen05 / TMEM / TMA implicit GEMM, split-bf16 x3, power-of-two geometry   */

#define LFDM_ACT_NONE    0
#define LFDM_ACT_RELU    1
#define LFDM_ACT_SIGMOID 2

#define LFDM_CONV_DIRECT     0 /* out(ho,wo) += in(ho*s - pad + kh, wo*s - pad + kw)                          */
#define LFDM_CONV_TRANSPOSED 1 /* ConvTranspose stride 2: in((ho + pad - kh)/2, ...) when divisible           */
#define LFDM_CONV_UPNEAREST  2 /* nearest x2 upsample then stride-1 conv: in((ho - pad + kh) >> 1, ...)       */

/* One 2-D convolution / GEMM over row matrices, with fused epilogue.
 * Replaces: nn.Conv3d (1,k,k) / nn.ConvTranspose3d / nn.Conv2d / nn.Linear call sites of
 *   DM/modules/video_flow_diffusion.py:156-167,199,224,246-247,300-301,410-411,495,508 and
 *   LFAE/modules/util.py:78-81,102-103,122-123,142-143, LFAE/modules/generator.py:51 (+ folded BatchNorm).
 * v[m][n] = sum_k A[m][k] W[n][k] + bias[n] + residual[m'][n];   out_f32 = f32_act(v);
 * out_sb = split(sb_act(sb_scale[n]*v + sb_shift[n]));   gn_stats[b][g] += (sum v, sum v^2)                */
typedef struct lfdm_conv_desc {
    /* A operand: up to two sources concatenated along channels (virtual torch.cat, video_flow_diffusion.py:580,587) */
    const void*  a_sb[2];      /* SB hi plane (or NULL)                                                     */
    const float* a_f32[2];     /* F32 (or NULL); exactly one of a_sb[i]/a_f32[i] is set for a used source    */
    int64_t      a_plane[2];   /* SB plane offset in elements                                                */
    int32_t      a_c[2];       /* channels of each source; a_c[1] = 0 when unused                            */
    int32_t      nf, h_in, w_in;
    int32_t      h_out, w_out;
    int32_t      kh, kw, pad, stride;
    int32_t      mode;         /* LFDM_CONV_*                                                                 */
    int32_t      reflect;      /* 1: reflect padding (padding_mode='reflect', video_flow_diffusion.py:162).  TC engine:
                                  LFDM_CONV_UPNEAREST only, and a_sb must then be the REPLICATE-PADDED input
                                  [(h_in+2) x (w_in+2)] from lfdm_pad_replicate_rows (reflect on the x2 map == clamp here) */
    /* weights */
    const float* w_f32;        /* SIMT: [kh*kw][Cin_total][Cout]                                              */
    const void*  w_sb;         /* TC  : packed by lfdm_pack, hi plane                         */
    int64_t      w_plane;
    const float* bias;         /* [Cout] or NULL                                                              */
    int32_t      c_out;
    /* epilogue */
    const float* residual;     /* F32 [M'][Cout] or NULL                                                      */
    int32_t      res_bcast_f;  /* >0: residual row = (m / (F*P))*P + m % P with F = res_bcast_f, P = h_out*w_out
                                  (frame-invariant term, hoisted init_conv(fea) of video_flow_diffusion.py:713-714) */
    float*       out_f32;      /* or NULL                                                                     */
    int32_t      f32_act;
    void*        out_sb;       /* or NULL                                                                     */
    int64_t      out_plane;
    int32_t      sb_act;
    const float* sb_scale;     /* [Cout] or NULL (1)                                                          */
    const float* sb_shift;     /* [Cout] or NULL (0)                                                          */
    double*      gn_stats;     /* [B][groups][2] accumulators (pre-zeroed) or NULL (TC engine only)           */
    int32_t      gn_cpg;       /* channels per group                                                          */
    int32_t      rows_per_sample;
    /* fused q/k rotary embedding of a temporal-attention qkv projection (TC engine, plain F32 output only): columns
     * [0, rot_cols) are rotated pairwise (2i, 2i+1) by the angle of frame f = (row / rot_rows_per_frame) % rot_frames,
     * tables [rot_frames][16] (dim_head 32); columns [0, rot_scale_cols) are multiplied by rot_scale first
     * (q = q * scale; q, k = rotary(q), rotary(k): Attention.forward video_flow_diffusion.py:325-331).  NULL = off. */
    const float* rot_cos;
    const float* rot_sin;
    int32_t      rot_frames, rot_rows_per_frame, rot_cols, rot_scale_cols;
    float        rot_scale;
    /* stream-K work space of the TC engine (optional; NULL = whole tiles only).  Layers with few long tiles (4x4 / 8x8 levels)
     * are cut into one equal K-block range per CTA; a tile cut in two is finished by the CTA that owns its end and the other
     * CTA hands over its fp32 partial accumulator: sk_workspace >= #SMs * 128 * 128 * 4 bytes, sk_flags = sk_slots (>= #SMs)
     * int32 zeros (the kernel leaves them at zero).  One work space can serve every conv of a stream.             */
    void*        sk_workspace;
    int64_t      sk_workspace_bytes;
    int32_t*     sk_flags;
    int32_t      sk_slots;
} lfdm_conv_desc;

int lfdm_conv(const lfdm_conv_desc* d, int engine, void* stream);

/* --- normalisation ------------------------------------------------------------------------------------------- */
/* GroupNorm statistics of F32 x[M][C]: stats[b][g] = (sum, sumsq) in double.  video_flow_diffusion.py:200,205 */
int lfdm_gn_stats(const float* x, int64_t m, int c, int groups, int rows_per_sample, double* stats, void* stream);
/* y = silu(gn(x)*(scale+1)+shift) (+ residual): Block.forward + ResnetBlock residual, video_flow_diffusion.py:203-211,237.
 * ss: row b at ss + b*ss_stride holds (scale[C] | shift[C]), or NULL.                                           */
int lfdm_gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, const float* ss,
                  int64_t ss_stride, const float* residual, float* out_f32, void* out_sb, int64_t out_plane, int64_t m, int c,
                  int groups, int rows_per_sample, float eps, void* stream);
/* channel LayerNorm (biased var, gamma only): video_flow_diffusion.py:176-179 -> SB (and/or F32)              */
int lfdm_layernorm(const float* x, const float* gamma, void* out_sb, int64_t out_plane, float* out_f32,
                   int64_t m, int c, float eps, void* stream);

/* --- attention cores ----------------------------------------------------------------------------------------- */
/* softmax(q k^T * . + bias) v over sequences gathered from qkv[M][3*heads*32].
 * sequence s: rows base(s) + j*row_stride, j < seq_len, base(s) = (s / inner)*outer_stride + (s % inner)*inner_stride.
 * rot_cos/rot_sin: [seq_len][16] or NULL; pos_bias: [heads][seq_len][seq_len] or NULL.
 * Attention.forward video_flow_diffusion.py:303-363 (+ EinopsToAndFrom :270-283).  seq_len <= 64.              */
int lfdm_attn_softmax(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_seq, int seq_len,
                      int heads, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t row_stride,
                      const float* rot_cos, const float* rot_sin, const float* pos_bias, void* stream);
/* same core for q|k that already carry scale and rotary (written by lfdm_conv with rot_cos set): no scaling, no rotary
 * here.  Only the tensor-core kernel of 17 <= seq_len <= 40 implements it; LFDM_E_UNSUPP otherwise.            */
int lfdm_attn_softmax_pre(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_seq, int seq_len,
                          int heads, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t row_stride,
                          const float* pos_bias, void* stream);
/* SpatialLinearAttention core video_flow_diffusion.py:253-263: per (frame, head) over n = hw positions.        */
int lfdm_attn_linear(const float* qkv, void* out_sb, int64_t out_plane, float* out_f32, int64_t n_frames, int n_pos,
                     int heads, void* stream);

/* Whole temporal-attention block (Residual(PreNorm(EinopsToAndFrom(Attention))), video_flow_diffusion.py:132-138,
 * 170-190,270-283,286-363) as ONE tcgen05 kernel: out = x + to_out(attention(to_qkv(LayerNorm(x)))) (+ out_bias).
 * x / out rows [(b*frames + f)*pixels + p][c]; sequences run over f.  c == 64 and frames == 40 only (LFDM_E_UNSUPP
 * otherwise: the caller then composes lfdm_layernorm / lfdm_conv / lfdm_attn_softmax).  wq_packed / wo_packed: per-head
 * split-bf16 weight slices in the kernel's shared-memory image (heads x 24576 B: [hi|lo] planes of the 96 x 64 rows
 * (q_h; k_h; v_h) of to_qkv; heads x 8192 B: 64 rows [w_hi(32) | w_lo(32)] of to_out[:, 32h:32h+32]), 128-byte rows
 * with 16-byte chunk index XOR (row & 7) -- built once by the host (engine/ops.py: pack_fused_attention).
 * debug: optional diagnostics [M][3*hid + heads*40 + hid] = rotated q | rotated k | v | softmax rows | head outputs. */
int lfdm_attn_temporal_fused(const float* x, const float* gamma, const void* wq_packed, const void* wo_packed,
                             const float* out_bias, const float* rot_cos, const float* rot_sin, const float* pos_bias,
                             float* out_f32, void* out_sb, int64_t out_plane, int n_b, int frames, int pixels, int c,
                             int heads, float eps, float* debug, void* stream);

/* Whole spatial linear-attention block (Residual(PreNorm(SpatialLinearAttention)), video_flow_diffusion.py:132-138,
 * 170-190,240-265) on tcgen05: out = x + to_out(linear_attention(to_qkv(LayerNorm(x)))) + out_bias, q|k|v never in HBM.
 * x / out rows [frame * pos + p][c], frames independent.  c == 64, heads == 8 (x 32), pos % 128 == 0 only (LFDM_E_UNSUPP
 * otherwise: the caller then composes lfdm_layernorm / lfdm_conv / lfdm_attn_linear).  wk / wv / wq_packed: split-bf16
 * operand images built by the host packer (engine/ops.py: pack_fused_linear_attention); wout: [64][256] fp32.
 * Work space owned by the caller (no allocation inside): partials [frames][LFDM_LINATTN_MAXP][256][34] fp32,
 * g_images [frames][65536] bytes (per frame: the context folded with to_out, B operand of the apply pass) and
 * xn_images [frames * pos / 128][32768] bytes (LayerNorm(x) as split-bf16 operand images, written by pass 1, read by pass 3).
 * Three launches on `stream` (context partials / merge / apply).                                                          */
#define LFDM_LINATTN_MAXP 4
int lfdm_attn_linear_fused(const float* x, const float* gamma, const void* wk_packed, const void* wv_packed, const void* wq_packed,
                           const float* wout, const float* out_bias, float* partials, void* g_images, void* xn_images,
                           float* out_f32, void* out_sb, int64_t out_plane, int frames, int pos, int c, int heads, float eps,
                           void* stream);

/* --- embeddings ---------------------------------------------------------------------------------------------- */
/* y[r][n] = act_out( sum_k act_in(x[r][k]) W[n][k] + b[n] ), small-M GEMV-class (time_mlp :422-428, block mlp :217-220).
 * act_in/out: 0 none, 1 silu, 2 gelu(erf).                                                                    */
int lfdm_small_linear(const float* x, const float* w, const float* b, float* y, int rows, int k, int n,
                      int act_in, int act_out, void* stream);
/* SinusoidalPosEmb :146-153: out[r] = (sin(t*freqs), cos(t*freqs)); freqs[dim/2] = exp(arange * -log(1e4)/(half-1)) */
int lfdm_sinusoidal(const int64_t* t, const float* freqs, float* out, int rows, int dim, void* stream);
/* ss[b][:] = a[a_row(b)][:] + c[b][:]   (time-table row + per-sample cond term)                                 */
int lfdm_ss_combine(const float* time_tab, const int32_t* step_idx, const float* cond_tab, float* ss, int b, int n,
                    void* stream);

/* --- sampler (GaussianDiffusion.p_mean_variance / p_sample / ddim_sample, video_flow_diffusion.py:697-746,792-827) */
/* x0 = c1*x - c2*eps ; absx0 = |x0|.  coef: device table [n_steps][8], row = *step_idx (or 0 if NULL):
 *   {c1, c2, a, b, sigma, c_eps, ddim, noclip}:  ddim=0: x' = a*x0c + b*x + sigma*z ; ddim=1: x' = a*x0c + c_eps*eps + sigma*z;
 *   noclip=1 skips the clamp (clip_denoised=False) */
int lfdm_sampler_x0(const float* x, const float* eps, const float* coef, const int32_t* step_idx, float* absx0,
                    int64_t n_per_sample, int b, void* stream);
/* exact torch.quantile(|x0|, q) (linear interpolation) per sample via radix select; s[b] = max(1, quantile).     */
int lfdm_sampler_quantile(const float* absx0, float* s, int64_t n_per_sample, int b, int64_t k_lo, float w_hi,
                          void* workspace, void* stream);
/* x_out = a*clamp(x0,-s,s)/s + b*x + c_eps*eps + sigma*z  (s == NULL: static clamp to [-1,1]); advances *step_idx   */
int lfdm_sampler_update(const float* x, const float* eps, const float* noise, const float* s, const float* coef,
                        int32_t* step_idx, int advance, float* x_out, float* x0_out, int64_t n_per_sample, int b,
                        void* stream);

/* --- LFAE warp / blend (Generator.deform_input + apply_optical, generator.py:60-88) --------------------------- */
/* out[n][y][x][c] = bilinear_sample(src[n / frames_per_src], up(flow[n]))*up(occ[n]) + prev[n]*(1-up(occ[n]))
 * src rows F32 [Ns*Hs*Ws][C]; flow [N][hf][wf][2]; occ [N][hf][wf] (NULL: no occlusion); prev F32 rows or NULL.  */
int lfdm_warp_blend_rows(const float* src, const float* flow, const float* occ, const float* prev, float* out_f32,
                         void* out_sb, int64_t out_plane, const float* sb_scale, const float* sb_shift, int sb_act,
                         int n, int frames_per_src, int hs, int ws, int c, int hf, int wf, void* stream);
/* planar variant for the 3-channel image: src NCHW (Ns,3,H,W); prev rows F32 [N*H*W][prev_ld] (sigmoid output) or NULL;
 * out NCDHW (B,3,F,H,W) written at [b][c][f].  generator.py:147,162.                                           */
int lfdm_warp_blend_image(const float* src, const float* flow, const float* occ, const float* prev, int prev_ld,
                          float* out, int b, int f, int h, int w, int hf, int wf, void* stream);

/* --- layout / small ops -------------------------------------------------------------------------------------- */
/* in[b][c][f][p] (strides sb, sc, sf, 1) -> rows [(b*F+f)*P + p][c_pad] (zero padded), SB and/or F32.           */
int lfdm_to_rows(const float* in, int b, int c, int f, int p, int64_t sb, int64_t sc, int64_t sf, int c_pad,
                 void* out_sb, int64_t out_plane, float* out_f32, void* stream);
/* rows F32 [(b*F+f)*P+p][ld] (first c channels) -> out[b][c][f][p] contiguous.                                   */
int lfdm_from_rows(const float* rows, int ld, int b, int c, int f, int p, float* out, void* stream);
/* 7x7 (k x k) im2col of a few-channel planar tensor in[b][c][f][h][w] -> SB rows [M][k_pad], k = tap*c + ch      */
int lfdm_im2col_small(const float* in, int b, int c, int f, int h, int w, int ksize, int pad, int k_pad,
                      void* out_sb, int64_t out_plane, void* stream);
/* SB rows [n][h][w][c] -> SB rows [n][h+2][w+2][c] with replicated borders (operand of the reflect-padded up-sampling conv) */
int lfdm_pad_replicate_rows(const void* in_sb, int64_t in_plane, void* out_sb, int64_t out_plane, int n, int h, int w, int c,
                            void* stream);
/* 2x2 average pool on rows (DownBlock2d, util.py:124,131)                                                        */
int lfdm_avgpool2_rows(const float* in, int n, int h, int w, int c, float* out_f32, void* out_sb, int64_t out_plane,
                       void* stream);
/* final 1x1 heads (video_flow_diffusion.py:495,508,588): out[b][0:2|2][f][p] from two F32 row matrices            */
int lfdm_unet_heads(const float* a, const float* wa, const float* ba, int na, const float* o, const float* wo,
                    const float* bo, int no, int c, int b, int f, int p, float* out, void* stream);

/* classifier-free guidance form (Unet3D.forward_with_cond_scale, video_flow_diffusion.py:521-526): a / o hold a 2B batch,
 * rows [0, B*f*p) evaluated with the condition and rows [B*f*p, 2*B*f*p) with the null condition;
 * out[b] = null_logits + (logits - null_logits) * cond_scale, b < B.                                              */
int lfdm_unet_heads_cfg(const float* a, const float* wa, const float* ba, int na, const float* o, const float* wo,
                        const float* bo, int no, int c, int b, int f, int p, float cond_scale, float* out, void* stream);

/* --- full-LFAE branch: dense motion / regions --------------------------------------- */
/* AntiAliasInterpolation2d (LFAE/modules/util.py:256-264): depthwise ks x ks Gaussian (zero pad ka) + ::s subsample, NCHW */
int lfdm_antialias_down(const float* in, const float* kern, float* out, int n, int c, int h, int w, int ks, int ka,
                        int s, void* stream);
/* RegionPredictor head (region_predictor.py:84-117): logits rows [(n*hw+p)][K] -> heatmap (N,K,h,w), shift (N,K,2),
 * covar/affine/u/d (N,K,2,2); the 2x2 SVD is closed form on the device (replaces torch.svd(covar.cpu()), :21).  */
int lfdm_region_moments(const float* logits, int n, int k, int h, int w, float temperature, float* heatmap,
                        float* shift, float* covar, float* affine, float* u, float* d, void* stream);
/* PixelwiseFlowPredictor front half (pixelwise_flow_predictor.py:48-102): hourglass input rows [N*hw][(K+1)*(nch+1)]
 * (heatmap difference | deformed source per region) and sparse motions (N,K+1,h,w,2).                            */
int lfdm_motion_prep(const float* down, const float* d_shift, const float* d_covar, const float* d_affine,
                     const float* s_shift, const float* s_covar, const float* s_affine, const float* bg, int n, int k,
                     int nch, int h, int w, int revert_axis_swap, float* hg_in, float* sparse, void* stream);
/* back half (:124-135): softmax over the K+1 mask logits, flow = sum_k mask_k * sparse_k, occlusion = sigmoid      */
int lfdm_motion_finish(const float* logits, int ld, const float* sparse, int n, int k, int hw, int has_occ,
                       float* flow, float* occ, void* stream);
/* per-image mean over positions of a row matrix [N*p][c] -> [N][c] (bg_motion_predictor.py:47)                    */
int lfdm_rows_mean(const float* rows, int n, int p, int c, float* out, void* stream);

/* --- output stage (demo/demo_mug.py:126-145) --------------------------------------------------------------- */
/* uint8 5-panel frames of ONE sample: out[f][H][5W][3] = [source | generated | warped source | sampling-grid figure | confidence].
 * src (3,H,W); out_vid / warped_vid (3,F,H,W); grid (2,F,h,w) absolute sampling grid in [-1,1]; conf (1,F,h,w) (device pointers);
 * mean3: HOST pointer to the 3 per-channel means already divided by 255 (demo MEAN / 255), or NULL.
 * Photographic panels: clamp(x + mean, 0, 1) * 255 truncated (sample_img :68-74); confidence: nearest up-sampling of conf * 255
 * truncated (misc.conf2fig); grid panel: identity grid + warped grid polylines, anti-aliased (replaces misc.grid2fig's matplotlib
 * figure: visually equivalent, not pixel-identical).  workspace: >= f*4*4 (rounded up to 256 B) + f*2*H*W*4 bytes.            */
int lfdm_render_panels(const float* src, const float* out_vid, const float* warped_vid, const float* grid, const float* conf,
                       const float* mean3, int f, int H, int W, int h, int w, float line_width, void* workspace, uint8_t* out,
                       void* stream);

/* --- weight packing (host-side helper, device pointers) ------------------------------------------------------ */
/* fp32 [rows][cols] -> SB planes                                                                                 */
int lfdm_split_bf16(const float* in, void* out_sb, int64_t out_plane, int64_t n, void* stream);

/* library info: returns sm arch the kernels were compiled for (100) and fills `has_tc` with 1                    */
int lfdm_version(int* arch, int* has_tc);

#ifdef __cplusplus
}
#endif
#endif /* LFDM_B200_H */
