"""Thin host-side wrappers over the C-ABI kernels (tensor plumbing only: allocation, pointers, weight packing)."""
import os
import torch
from .. import _lib as L
from .._lib import SB, ptr, stream, check, lib

# engine policy: "tc" (tcgen05 where the geometry allows, SIMT otherwise), "simt" (force CUDA-core engine)
DEFAULT_ENGINE = os.environ.get("LFDM_ENGINE", "tc")
FUSED_GN_STATS = os.environ.get("LFDM_FUSED_GN_STATS", "1") == "1"
# when set to a list, every conv launch appends (name, engine, algorithmic_flops, start_event, end_event):
# bench.py uses it for the live CUDA-event roofline measurement (never enabled inside timed regions)
PROFILE = None


def f32(m, c, device):
    return torch.empty((m, c), dtype=torch.float32, device=device)


def split_planes(w):
    """fp32 tensor -> (2, ...) bf16 hi/lo planes (same rounding as the device split)."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo], 0).contiguous()


STREAM_K = os.environ.get("LFDM_CONV_NO_STREAMK") is None
_SK_WS = {}


def streamk_workspace(dev):
    """one stream-K work space per device (lfdm_conv_desc.sk_*): every conv of the stream shares it (kernels are ordered)"""
    key = (dev.type, dev.index)
    ws = _SK_WS.get(key)
    if ws is None:
        slots = 256
        ws = (torch.empty(slots * 128 * 128, dtype=torch.float32, device=dev), torch.zeros(slots, dtype=torch.int32, device=dev), slots)
        _SK_WS[key] = ws
    return ws


class ConvLayer:
    """One convolution / linear layer with weights packed for both engines.

    weight: (Cout, Cin, kh, kw) [DIRECT / UPNEAREST] or (Cin, Cout, kh, kw) [TRANSPOSED] fp32 (already BN-folded).
    src_channels: channel split of the virtual-concat sources (sum == Cin)."""

    def __init__(self, weight, bias, *, mode=L.CONV_DIRECT, stride=1, pad=0, reflect=False, src_channels=None,
                 engine=None, name=""):
        dev = weight.device
        w = weight.detach().float()
        self.name = name
        self.mode, self.stride, self.pad, self.reflect = mode, stride, pad, int(bool(reflect))
        if mode == L.CONV_TRANSPOSED:
            cin, cout, kh, kw = w.shape
            w_oc = w.permute(1, 0, 2, 3).contiguous()        # (Cout, Cin, kh, kw)
        else:
            cout, cin, kh, kw = w.shape
            w_oc = w
        self.cin, self.cout, self.kh, self.kw = cin, cout, kh, kw
        self.src_channels = list(src_channels) if src_channels else [cin]
        assert sum(self.src_channels) == cin
        self.bias = bias.detach().float().contiguous() if bias is not None else None
        self.engine = engine or DEFAULT_ENGINE
        # SIMT packing: [tap][Cin][Cout]
        self.w_f32 = w_oc.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout).contiguous()
        # TC packing (see conv_tc.cu / DESIGN.md): [plane][tap][Cout_pad][Cin]
        self.w_sb = None
        if self.engine == "tc" and self._tc_static_ok():
            if cout % 128 == 0:
                bn = 128
            elif cout % 64 == 0:
                bn = 64
            elif cout % 32 == 0:
                bn = 32
            else:
                bn = 16
            cpad = (cout + bn - 1) // bn * bn
            if mode == L.CONV_DIRECT:
                taps = w_oc.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
            elif mode == L.CONV_TRANSPOSED:
                lists = {0: [1, 3], 1: [0, 2]}
                taps = torch.stack([w_oc[:, :, lists[p][a], lists[q][b]]
                                    for p in (0, 1) for q in (0, 1) for a in (0, 1) for b in (0, 1)], 0)
            else:  # UPNEAREST: pre-summed 2x2 sub-kernels per output parity
                groups = {0: [[0], [1, 2]], 1: [[0, 1], [2]]}
                taps = torch.stack([
                    sum(w_oc[:, :, i, j] for i in groups[p][a] for j in groups[q][b])
                    for p in (0, 1) for q in (0, 1) for a in (0, 1) for b in (0, 1)], 0)
            if cpad != cout:
                taps = torch.cat([taps, torch.zeros(taps.shape[0], cpad - cout, cin, device=dev)], 1)
            self.w_sb = split_planes(taps.contiguous())
            self.w_plane = self.w_sb[0].numel()

    def _tc_static_ok(self):
        if any(c % 64 for c in self.src_channels) or len(self.src_channels) > 2:
            return False
        if self.reflect and not (self.mode == L.CONV_UPNEAREST and len(self.src_channels) == 1):
            return False
        if not (self.cout % 32 == 0 or self.cout <= 16):
            return False
        if self.mode == L.CONV_DIRECT and self.stride == 1:
            return self.kh * self.kw <= 52 and self.pad == self.kh // 2 and self.kh == self.kw
        if self.mode == L.CONV_DIRECT and self.stride == 2:
            return self.kh == 4 and self.kw == 4 and self.pad == 1
        if self.mode == L.CONV_TRANSPOSED:
            return self.kh == 4 and self.kw == 4 and self.pad == 1 and self.stride == 2
        if self.mode == L.CONV_UPNEAREST:
            return self.kh == 3 and self.kw == 3 and self.pad == 1
        return False

    def out_hw(self, h, w):
        if self.mode == L.CONV_DIRECT:
            return (h + 2 * self.pad - self.kh) // self.stride + 1, (w + 2 * self.pad - self.kw) // self.stride + 1
        if self.mode == L.CONV_TRANSPOSED:
            return (h - 1) * self.stride - 2 * self.pad + self.kh, (w - 1) * self.stride - 2 * self.pad + self.kw
        return 2 * h + 2 * self.pad - self.kh + 1, 2 * w + 2 * self.pad - self.kw + 1

    def __call__(self, srcs, nf, h, w, *, out_f32=None, out_sb=None, residual=None, res_bcast_f=0, f32_act=0,
                 sb_act=0, sb_scale=None, sb_shift=None, gn_stats=None, gn_groups=8, rows_per_sample=0, stats_zeroed=False,
                 rot=None):
        """srcs: list of SB or F32 row matrices (one per virtual-concat source).
        rot = (cos, sin, frames, rows_per_frame, rot_cols, scale_cols, scale): fuse q-scale + q/k rotary into the epilogue
        (tcgen05 engine, plain F32 output); `self.rot_applied` tells the caller whether the engine took it."""
        ho, wo = self.out_hw(h, w)
        a_sb = [None, None]
        a_f32 = [None, None]
        a_plane = [0, 0]
        a_c = [0, 0]
        all_sb = True
        for i, s in enumerate(srcs):
            a_c[i] = self.src_channels[i]
            if isinstance(s, SB):
                assert s.c == a_c[i] and s.m == nf * h * w, (self.name, s.m, s.c, nf, h, w, a_c[i])
                a_sb[i] = s.t
                a_plane[i] = s.plane
            else:
                assert s.shape == (nf * h * w, a_c[i]), (self.name, tuple(s.shape), nf, h, w, a_c[i])
                a_f32[i] = s
                all_sb = False
        fused_stats = gn_stats is not None and FUSED_GN_STATS
        if self.reflect and self.w_sb is not None and all_sb:
            # tcgen05 engine: reflect padding of the x2 nearest up-sampled map == clamping on the low-res map -> replicate-pad once
            src = srcs[0]
            padded = SB(nf * (h + 2) * (w + 2), src.c, src.t.device)
            check(lib().lfdm_pad_replicate_rows(ptr(src.t), src.plane, ptr(padded.t), padded.plane, nf, h, w, src.c, stream()),
                  "lfdm_pad_replicate_rows")
            a_sb[0], a_plane[0] = padded.t, padded.plane
        kw = dict(a_sb=a_sb, a_f32=a_f32, a_plane=a_plane, a_c=a_c, nf=nf, h_in=h, w_in=w, h_out=ho, w_out=wo,
                  kh=self.kh, kw=self.kw, pad=self.pad, stride=self.stride, mode=self.mode, reflect=self.reflect,
                  w_f32=self.w_f32, w_sb=self.w_sb, w_plane=self.w_plane if self.w_sb is not None else 0,
                  bias=self.bias, c_out=self.cout, residual=residual, res_bcast_f=res_bcast_f,
                  out_f32=out_f32, f32_act=f32_act, out_sb=out_sb.t if out_sb is not None else None,
                  out_plane=out_sb.plane if out_sb is not None else 0, sb_act=sb_act, sb_scale=sb_scale,
                  sb_shift=sb_shift, gn_stats=None, gn_cpg=0, rows_per_sample=rows_per_sample)
        if STREAM_K and self.w_sb is not None and all_sb:
            ws, flags, slots = streamk_workspace(self.w_sb.device)
            kw.update(sk_workspace=ws, sk_workspace_bytes=ws.numel() * 4, sk_flags=flags, sk_slots=slots)
        self.rot_applied = False
        if rot is not None:
            kw.update(rot_cos=rot[0], rot_sin=rot[1], rot_frames=int(rot[2]), rot_rows_per_frame=int(rot[3]),
                      rot_cols=int(rot[4]), rot_scale_cols=int(rot[5]), rot_scale=float(rot[6]))
        rc = L.E_UNSUPP
        self.last_engine = "simt"
        ev0 = None
        if PROFILE is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if self.w_sb is not None and all_sb:
            if fused_stats:
                cpg = self.cout // gn_groups
                if cpg % 8 == 0 and rows_per_sample % 128 == 0 and self.mode == L.CONV_DIRECT:
                    if not stats_zeroed:
                        gn_stats.zero_()
                    kw["gn_stats"] = gn_stats
                    kw["gn_cpg"] = cpg
                else:
                    fused_stats = False
            rc = L.conv(kw, L.ENGINE_TC)
            if rc == L.E_UNSUPP and rot is not None:       # this geometry cannot fuse the rotary: plain projection instead
                for k in ("rot_cos", "rot_sin"):
                    kw[k] = None
                rc = L.conv(kw, L.ENGINE_TC)
            elif rc == 0 and rot is not None:
                self.rot_applied = True
            self.last_engine = "tc"
            if rc == L.E_UNSUPP:
                self.last_engine = "simt"
                kw["gn_stats"] = None
                fused_stats = False
        else:
            fused_stats = False
        if rc == L.E_UNSUPP:
            kw["gn_stats"] = None
            kw["rot_cos"] = kw["rot_sin"] = None
            rc = L.conv(kw, L.ENGINE_SIMT)
        check(rc, f"lfdm_conv[{self.name}]")
        if ev0 is not None:
            ev1.record()
            eff_taps = self.kh * self.kw if self.mode == L.CONV_DIRECT else 4   # post-hoist: 2x2 sub-kernels per output
            flops = 2.0 * nf * ho * wo * self.cout * self.cin * eff_taps
            # algorithmic bytes: every operand once (activations 4 B/element in either format, packed weights 4 B/element)
            m_in, m_out = nf * h * w, nf * ho * wo
            abytes = 4.0 * (m_in * self.cin + self.cout * self.cin * self.kh * self.kw
                            + m_out * self.cout * ((out_f32 is not None) + (out_sb is not None))
                            + (0 if residual is None else residual.numel()))
            PROFILE.append((self.name, self.last_engine, flops, ev0, ev1, abytes))
        if gn_stats is not None and not fused_stats:
            assert out_f32 is not None
            check(lib().lfdm_gn_stats(ptr(out_f32), out_f32.shape[0], self.cout, gn_groups, rows_per_sample,
                                      ptr(gn_stats), stream()), "lfdm_gn_stats")
        return ho, wo


def gn_apply(x, stats, gamma, beta, ss, residual, out_f32, out_sb, groups, rows_per_sample, eps=1e-5):
    """ss: None or a (B, 2C) view (row stride arbitrary, unit inner stride) of the (scale | shift) table."""
    m, c = x.shape
    ss_ptr, ss_stride = None, 0
    if ss is not None:
        assert ss.stride(1) == 1 and ss.shape[1] == 2 * c
        ss_ptr, ss_stride = L.C.c_void_p(ss.data_ptr()), ss.stride(0)
    check(lib().lfdm_gn_apply(ptr(x), ptr(stats), ptr(gamma), ptr(beta), ss_ptr, ss_stride, ptr(residual), ptr(out_f32),
                              ptr(out_sb.t) if out_sb is not None else None,
                              out_sb.plane if out_sb is not None else 0, m, c, groups, rows_per_sample, eps, stream()),
          "lfdm_gn_apply")


def layernorm(x, gamma, out_sb=None, out_f32=None, eps=1e-5):
    m, c = x.shape
    check(lib().lfdm_layernorm(ptr(x), ptr(gamma), ptr(out_sb.t) if out_sb is not None else None,
                               out_sb.plane if out_sb is not None else 0, ptr(out_f32), m, c, eps, stream()),
          "lfdm_layernorm")


def attn_softmax(qkv, out_sb, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride, row_stride,
                 rot_cos=None, rot_sin=None, pos_bias=None):
    check(lib().lfdm_attn_softmax(ptr(qkv), ptr(out_sb.t) if out_sb is not None else None,
                                  out_sb.plane if out_sb is not None else 0, ptr(out_f32), n_seq, seq_len, heads, inner,
                                  outer_stride, inner_stride, row_stride, ptr(rot_cos), ptr(rot_sin), ptr(pos_bias),
                                  stream()), "lfdm_attn_softmax")


def attn_softmax_pre(qkv, out_sb, out_f32, n_seq, seq_len, heads, inner, outer_stride, inner_stride, row_stride, pos_bias=None):
    """q | k already scaled + rotated by the qkv projection's epilogue (ConvLayer(..., rot=...))"""
    check(lib().lfdm_attn_softmax_pre(ptr(qkv), ptr(out_sb.t) if out_sb is not None else None,
                                      out_sb.plane if out_sb is not None else 0, ptr(out_f32), n_seq, seq_len, heads, inner,
                                      outer_stride, inner_stride, row_stride, ptr(pos_bias), stream()), "lfdm_attn_softmax_pre")


def attn_linear(qkv, out_sb, out_f32, n_frames, n_pos, heads):
    check(lib().lfdm_attn_linear(ptr(qkv), ptr(out_sb.t) if out_sb is not None else None,
                                 out_sb.plane if out_sb is not None else 0, ptr(out_f32), n_frames, n_pos, heads,
                                 stream()), "lfdm_attn_linear")


def _sw128_image(rows_bf16):
    """(R, 64) bf16 rows of 128 bytes -> the kernel's shared-memory image: 16-byte chunk c of row r stored at chunk
    c ^ (r & 7) (the 128-byte swizzle of the UMMA / TMA operand layout; 8-row groups are contiguous 1024-byte atoms)."""
    r = rows_bf16.shape[0]
    assert rows_bf16.shape[1] == 64 and r % 8 == 0
    ch = rows_bf16.reshape(r, 8, 8)
    idx = torch.arange(8, device=rows_bf16.device)[None, :] ^ (torch.arange(r, device=rows_bf16.device)[:, None] & 7)   # (r, 8)
    return torch.gather(ch, 1, idx[:, :, None].expand(r, 8, 8)).reshape(r, 64).contiguous()     # out[r, c'] = in[r, c' ^ (r&7)]


def pack_fused_attention(qkv_w, out_w, heads):
    """to_qkv weight (3*hid, 64) and to_out weight (64, hid) -> per-head split-bf16 operand images of
    lfdm_attn_temporal_fused (include/lfdm_b200.h): (heads, 2, 96, 64) and (heads, 64, 64) bf16."""
    hid = heads * 32
    qkv_w = qkv_w.detach().float().reshape(3 * hid, -1)
    out_w = out_w.detach().float().reshape(-1, hid)
    assert qkv_w.shape[1] == 64 and out_w.shape[0] == 64
    wq, wo = [], []
    for h in range(heads):
        # q rows carry the softmax scale 32^-0.5 (reference :325 scales q after the projection; the kernel does not)
        wh = torch.cat([qkv_w[h * 32:(h + 1) * 32] * (32 ** -0.5), qkv_w[hid + h * 32:hid + (h + 1) * 32],
                        qkv_w[2 * hid + h * 32:2 * hid + (h + 1) * 32]], 0)                   # (96, 64): q_h; k_h; v_h
        pl = split_planes(wh)                                                                  # (2, 96, 64)
        wq.append(torch.stack([_sw128_image(pl[0]), _sw128_image(pl[1])], 0))
        oh = split_planes(out_w[:, h * 32:(h + 1) * 32])                                       # (2, 64, 32)
        wo.append(_sw128_image(torch.cat([oh[0], oh[1]], 1)))                                  # rows [w_hi(32) | w_lo(32)]
    return torch.stack(wq, 0).contiguous(), torch.stack(wo, 0).contiguous()


def attn_temporal_fused(x, gamma, wq, wo, out_bias, cos, sin, pos_bias, out_f32, out_sb, n_b, frames, pixels, heads, eps,
                        debug=None):
    """returns the C-ABI code (0, or L.E_UNSUPP when the geometry is not covered by the fused kernel)"""
    rc = lib().lfdm_attn_temporal_fused(ptr(x), ptr(gamma), ptr(wq), ptr(wo), ptr(out_bias), ptr(cos), ptr(sin), ptr(pos_bias),
                                        ptr(out_f32), ptr(out_sb.t) if out_sb is not None else None,
                                        out_sb.plane if out_sb is not None else 0, n_b, frames, pixels, x.shape[1], heads,
                                        eps, ptr(debug), stream())
    if rc != L.E_UNSUPP:
        check(rc, "lfdm_attn_temporal_fused")
    return rc


LINATTN_MAXP = 4      # LFDM_LINATTN_MAXP of include/lfdm_b200.h


def pack_fused_linear_attention(qkv_w, out_w):
    """to_qkv weight (768, 64) and to_out weight (64, 256) of SpatialLinearAttention (8 heads x 32) -> operand images of
    lfdm_attn_linear_fused: wk / wv (2 halves, 2 planes, 128 rows, 64) rows (head, d|e); wq (4 chunks, 2 planes, 64 rows, 64);
    wout fp32 (64, 256)."""
    hid = 256
    qkv_w = qkv_w.detach().float().reshape(3 * hid, -1)
    assert qkv_w.shape[1] == 64 and out_w.numel() == 64 * hid
    # q and k only ever enter exp(): their projections carry log2(e) so that the kernels use ex2 without a multiply
    l2e = 1.4426950408889634
    wq, wk, wv = qkv_w[:hid] * l2e, qkv_w[hid:2 * hid] * l2e, qkv_w[2 * hid:]

    def img(w, rows):            # (256, 64) -> (256 / rows, 2, rows, 64) swizzled split-bf16 images
        out = []
        for i in range(0, hid, rows):
            pl = split_planes(w[i:i + rows])
            out.append(torch.stack([_sw128_image(pl[0]), _sw128_image(pl[1])], 0))
        return torch.stack(out, 0).contiguous()
    return img(wk, 128), img(wv, 128), img(wq, 64), out_w.detach().float().reshape(64, hid).contiguous()


def attn_linear_fused(x, gamma, packed, out_bias, out_f32, out_sb, frames, pos, eps, work=None):
    """returns (rc, work): the C-ABI code (0, or L.E_UNSUPP when the geometry is not covered) and the work-space tensors
    (partials, g_images, xn_images) so that callers can keep them across calls"""
    wk, wv, wq, wout = packed
    if work is None or work[0].shape[0] != frames:
        work = (torch.empty((frames, LINATTN_MAXP, 256, 34), device=x.device),
                torch.empty((frames, 65536), dtype=torch.uint8, device=x.device),
                torch.empty((frames * pos // 128, 32768), dtype=torch.uint8, device=x.device))
    rc = lib().lfdm_attn_linear_fused(ptr(x), ptr(gamma), ptr(wk), ptr(wv), ptr(wq), ptr(wout), ptr(out_bias), ptr(work[0]), ptr(work[1]),
                                      ptr(work[2]), ptr(out_f32), ptr(out_sb.t) if out_sb is not None else None,
                                      out_sb.plane if out_sb is not None else 0, frames, pos, x.shape[1], 8, eps, stream())
    if rc != L.E_UNSUPP:
        check(rc, "lfdm_attn_linear_fused")
    return rc, work


def small_linear(x, w, b, y, act_in=0, act_out=0):
    rows, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k and y.shape == (rows, n)
    check(lib().lfdm_small_linear(ptr(x), ptr(w), ptr(b), ptr(y), rows, k, n, act_in, act_out, stream()),
          "lfdm_small_linear")


def to_rows(x5, c_pad=None, out_sb=None, out_f32=None):
    """x5: (B, C, F, H, W) (any strides with contiguous HW) -> rows [(b*F+f)*P + p][c_pad]."""
    b, c, f, h, w = x5.shape
    p = h * w
    assert x5.stride(4) == 1 and x5.stride(3) == w
    check(lib().lfdm_to_rows(L.C.c_void_p(x5.data_ptr()), b, c, f, p, x5.stride(0), x5.stride(1), x5.stride(2),
                             c_pad or c, ptr(out_sb.t) if out_sb is not None else None,
                             out_sb.plane if out_sb is not None else 0, ptr(out_f32), stream()), "lfdm_to_rows")


def from_rows(rows, b, c, f, p, out):
    check(lib().lfdm_from_rows(ptr(rows), rows.shape[1], b, c, f, p, ptr(out), stream()), "lfdm_from_rows")
