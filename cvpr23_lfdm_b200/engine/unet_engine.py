"""UnetEngine — runs reference `Unet3D.forward` (DM/modules/video_flow_diffusion.py:528-588) on the sm_100a kernels.

Host code here is tensor plumbing only: it packs weights once, allocates row-matrix buffers and issues kernel
launches through the C-ABI in the reference's data-flow order.  Legal algebraic hoists:
  * torch.cat([x, fea]) is never materialised: init_conv(x ‖ fea) = conv_x(x) + conv_fea(fea) and fea is constant
    over frames and steps, so conv_fea runs once per sample (`prepare_fea`) and enters as a frame-broadcast residual;
  * skip/concat tensors are "virtual": consumers walk two operand descriptors;
  * W·SiLU([t_emb ; cond]) = W_t·SiLU(t_emb) + W_c·SiLU(cond): per-step (scale, shift) come from a time table row
    plus a per-sample cond term (`build_tables` / `ss_from_tables`).
"""
import math
import os
import torch
from .. import _lib as L
from .._lib import SB, ptr, stream, check, lib
from . import ops
from .ops import ConvLayer, f32

# LFDM_FUSED_ROTARY=1: q*scale + q/k rotary of the temporal attention applied by the epilogue of its qkv projection
# (lfdm_conv rot_* + lfdm_attn_softmax_pre) instead of inside the attention kernel.  Off by default: measured on B200 it
# moves ~0.1 ms per 32x32 call from the attention kernel (16 warps/SM) into the 8 epilogue warps of an HBM-bound GEMM
# and loses overall (conv 6.45 -> 6.83 ms per evaluation).
FUSE_ROTARY = os.environ.get("LFDM_FUSED_ROTARY") is not None
# LFDM_FUSED_ATTN=0: compose the temporal-attention block from layernorm / qkv conv / attention core / out conv instead
# of the single tcgen05 kernel lfdm_attn_temporal_fused (C = 64, 40 frames) -- A/B switch and cross-check.
FUSED_ATTN = os.environ.get("LFDM_FUSED_ATTN", "1") == "1"
# LFDM_FUSED_LINATTN=0: same switch for the spatial linear-attention block (lfdm_attn_linear_fused: C = 64, 8 heads,
# positions per frame a multiple of 128).
FUSED_LINATTN = os.environ.get("LFDM_FUSED_LINATTN", "1") == "1"


def _rel_pos_bucket(n, num_buckets=32, max_distance=32):
    """integer bucket table of RelativePositionBias (reference :85-102); index math done on the host once."""
    q = torch.arange(n)
    rel = q[None, :] - q[:, None]
    neg = -rel
    nb = num_buckets // 2
    ret = (neg < 0).long() * nb
    a = neg.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(a.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    return ret + torch.where(a < max_exact, a, large)


class _Resnet:
    def __init__(self, mod, src_channels, name):
        w = lambda conv: conv.weight.detach().squeeze(2)
        self.c_out = mod.block1.proj.out_channels
        self.conv1 = ConvLayer(w(mod.block1.proj), mod.block1.proj.bias, pad=1, src_channels=src_channels, name=name + ".block1")
        self.conv2 = ConvLayer(w(mod.block2.proj), mod.block2.proj.bias, pad=1, name=name + ".block2")
        self.g1, self.b1 = mod.block1.norm.weight.detach().float().contiguous(), mod.block1.norm.bias.detach().float().contiguous()
        self.g2, self.b2 = mod.block2.norm.weight.detach().float().contiguous(), mod.block2.norm.bias.detach().float().contiguous()
        self.groups = mod.block1.norm.num_groups
        self.eps = mod.block1.norm.eps
        self.res = None
        if isinstance(mod.res_conv, torch.nn.Conv3d):
            self.res = ConvLayer(w(mod.res_conv), mod.res_conv.bias, pad=0, src_channels=src_channels, name=name + ".res")
        self.has_mlp = mod.mlp is not None
        self.ss_off = None


class _Attn:
    def __init__(self, ln, qkv_w, out_w, out_b, heads, name, linear=False):
        self.gamma = ln.gamma.detach().float().reshape(-1).contiguous()
        self.eps = ln.eps
        c = self.gamma.numel()
        self.qkv = ConvLayer(qkv_w.detach().reshape(qkv_w.shape[0], c, 1, 1), None, name=name + ".qkv")
        self.out = ConvLayer(out_w.detach().reshape(out_w.shape[0], -1, 1, 1), out_b, name=name + ".out")
        self.heads = heads
        self.hid = qkv_w.shape[0] // 3
        self.out_bias = out_b.detach().float().contiguous() if out_b is not None else None
        self.fused = None          # (wq, wo) operand images of the fused temporal block (C == 64)
        self.fused_lin = None      # operand images of the fused linear-attention block (C == 64, 8 heads)
        if c == 64 and qkv_w.shape[0] == 3 * heads * 32:
            if linear:
                if heads == 8:
                    self.fused_lin = ops.pack_fused_linear_attention(qkv_w, out_w)
            else:
                self.fused = ops.pack_fused_attention(qkv_w, out_w, heads)


class UnetEngine:
    def __init__(self, unet):
        self.dev = unet.init_conv.weight.device
        self.unet_channels = unet.channels
        self.dim = unet.dim
        self.heads = unet.attn_heads
        self.out_grid_dim, self.out_conf_dim = unet.out_grid_dim, unet.out_conf_dim
        dev = self.dev
        # ---- init conv (two packings: generic 259->64 over padded channels-last rows; hoisted x-part / fea-part)
        wi = unet.init_conv.weight.detach().squeeze(2).float()         # (Co, Cin, k, k)
        self.init_dim = wi.shape[0]
        self.k_init = wi.shape[-1]
        cin = wi.shape[1]
        self.cin_pad = (cin + 63) // 64 * 64
        wpad = torch.zeros(self.init_dim, self.cin_pad, self.k_init, self.k_init, device=dev)
        wpad[:, :cin] = wi
        self.init_full = ConvLayer(wpad, unet.init_conv.bias, pad=self.k_init // 2, name="init_conv")
        self._init_w = wi
        self._init_b = unet.init_conv.bias.detach().float().contiguous()
        self._hoist = None
        # ---- temporal attention helpers
        self.rel_emb = unet.time_rel_pos_bias.relative_attention_bias.weight.detach().float()
        self.rel_buckets, self.rel_maxdist = unet.time_rel_pos_bias.num_buckets, unet.time_rel_pos_bias.max_distance
        self.rot_freqs = unet.init_temporal_attn.fn.fn.fn.rotary_emb.freqs.detach().float()
        self._frame_tabs = {}

        def t_attn(res_mod, name):      # Residual(PreNorm(EinopsToAndFrom(Attention)))
            pre = res_mod.fn
            att = pre.fn.fn
            return _Attn(pre.norm, att.to_qkv.weight, att.to_out.weight, None, att.heads, name)

        def l_attn(res_mod, name):      # Residual(PreNorm(SpatialLinearAttention))
            pre = res_mod.fn
            att = pre.fn
            return _Attn(pre.norm, att.to_qkv.weight, att.to_out.weight, att.to_out.bias, att.heads, name, linear=True)

        self.init_tattn = t_attn(unet.init_temporal_attn, "init_temporal_attn")
        # ---- time / cond MLPs
        self.tm_w1 = unet.time_mlp[1].weight.detach().float().contiguous()
        self.tm_b1 = unet.time_mlp[1].bias.detach().float().contiguous()
        self.tm_w2 = unet.time_mlp[3].weight.detach().float().contiguous()
        self.tm_b2 = unet.time_mlp[3].bias.detach().float().contiguous()
        half = unet.dim // 2
        # frequencies exactly as the reference computes them (float32 torch ops on the host, :148-150)
        self.sin_freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).float().to(dev)
        self.time_dim = self.tm_w2.shape[0]
        self.cond_dim = unet.cond_in_dim if unet.has_cond else 0

        self.stages_down, self.stages_up = [], []
        cond_blocks = []

        def resnet(mod, src, name):
            r = _Resnet(mod, src, name)
            if r.has_mlp:
                cond_blocks.append((r, mod.mlp[1]))
            return r

        self.has_lin = not isinstance(unet.downs[0][2], torch.nn.Identity)
        for i, (b1, b2, sa, ta, down) in enumerate(unet.downs):
            st = dict(b1=resnet(b1, None, f"downs.{i}.0"), b2=resnet(b2, None, f"downs.{i}.1"),
                      sa=l_attn(sa, f"downs.{i}.2") if self.has_lin else None, ta=t_attn(ta, f"downs.{i}.3"), down=None)
            if isinstance(down, torch.nn.Conv3d):
                st["down"] = ConvLayer(down.weight.detach().squeeze(2), down.bias, stride=2, pad=1, name=f"downs.{i}.4")
            self.stages_down.append(st)
        self.mid1 = resnet(unet.mid_block1, None, "mid_block1")
        pre = unet.mid_spatial_attn.fn
        att = pre.fn.fn
        self.mid_sattn = _Attn(pre.norm, att.to_qkv.weight, att.to_out.weight, None, att.heads, "mid_spatial_attn")
        self.mid_tattn = t_attn(unet.mid_temporal_attn, "mid_temporal_attn")
        self.mid2 = resnet(unet.mid_block2, None, "mid_block2")
        for i, (b1, b2, sa, ta, up) in enumerate(unet.ups):
            cin_tot = b1.block1.proj.in_channels
            st = dict(b1=resnet(b1, [cin_tot // 2, cin_tot // 2], f"ups.{i}.0"), b2=resnet(b2, None, f"ups.{i}.1"),
                      sa=l_attn(sa, f"ups.{i}.2") if self.has_lin else None, ta=t_attn(ta, f"ups.{i}.3"), up=None)
            if isinstance(up, torch.nn.ConvTranspose3d):
                st["up"] = ConvLayer(up.weight.detach().squeeze(2), up.bias, mode=L.CONV_TRANSPOSED, stride=2, pad=1,
                                     name=f"ups.{i}.4")
            elif isinstance(up, torch.nn.Sequential):
                conv = up[1]
                st["up"] = ConvLayer(conv.weight.detach().squeeze(2), conv.bias, mode=L.CONV_UPNEAREST, pad=1,
                                     reflect=(conv.padding_mode == "reflect"), name=f"ups.{i}.4")
            self.stages_up.append(st)
        d = unet.dim
        self.head_a = resnet(unet.final_conv[0], [d, d], "final_conv.0")
        self.head_o = resnet(unet.occlusion_map[0], [d, d], "occlusion_map.0")
        self.wa = unet.final_conv[1].weight.detach().float().reshape(unet.out_grid_dim, d).contiguous()
        self.ba = unet.final_conv[1].bias.detach().float().contiguous()
        self.wo = unet.occlusion_map[1].weight.detach().float().reshape(unet.out_conf_dim, d).contiguous()
        self.bo = unet.occlusion_map[1].bias.detach().float().contiguous()

        # ---- concatenated (scale, shift) MLP of all conditioned blocks: one [sum 2C, time_dim + cond_dim] matrix
        off = 0
        ws, bs = [], []
        for r, lin in cond_blocks:
            r.ss_off = off
            off += lin.weight.shape[0]
            ws.append(lin.weight.detach().float())
            bs.append(lin.bias.detach().float())
        self.ss_total = off
        self.ss_w = torch.cat(ws, 0).contiguous()
        self.ss_b = torch.cat(bs, 0).contiguous()
        self.ss_w_time = self.ss_w[:, :self.time_dim].contiguous()
        self.ss_w_cond = self.ss_w[:, self.time_dim:].contiguous() if self.cond_dim else None
        self.taps = None   # optional dict filled with F32 copies of intermediate activations (debug / tests)
        self._time_tab_cache = {}   # tuple(times) -> (len(times), ss_total) table: depends only on the weights of this engine

    # ------------------------------------------------------------------------------------------------
    # embeddings
    # ------------------------------------------------------------------------------------------------
    def time_embed(self, time):
        """time_mlp (reference :422-428): (R,) int64 -> (R, time_dim)"""
        r = time.shape[0]
        dev = self.dev
        s = torch.empty((r, self.dim), device=dev)
        check(lib().lfdm_sinusoidal(ptr(time.contiguous()), ptr(self.sin_freqs), ptr(s), r, self.dim, stream()), "lfdm_sinusoidal")
        h = torch.empty((r, self.time_dim), device=dev)
        ops.small_linear(s, self.tm_w1, self.tm_b1, h, 0, 2)
        t = torch.empty((r, self.time_dim), device=dev)
        ops.small_linear(h, self.tm_w2, self.tm_b2, t, 0, 0)
        return t

    def scale_shift(self, time, cond):
        """all (scale, shift) vectors of one evaluation: (B, ss_total).  reference :552-562 + :230-232"""
        t = self.time_embed(time)
        if self.cond_dim:
            t = torch.cat((t, cond.float()), dim=-1).contiguous()
        ss = torch.empty((t.shape[0], self.ss_total), device=self.dev)
        ops.small_linear(t, self.ss_w, self.ss_b, ss, 1, 0)
        return ss

    def build_tables(self, times, cond):
        """hoisted form: time table (len(times), ss_total) and cond table (B, ss_total) with
        ss(step, b) = time_tab[step] + cond_tab[b]."""
        key = tuple(int(v) for v in times.tolist())
        time_tab = self._time_tab_cache.get(key)
        if time_tab is None:
            t = self.time_embed(times)
            time_tab = torch.empty((t.shape[0], self.ss_total), device=self.dev)
            ops.small_linear(t, self.ss_w_time, None, time_tab, 1, 0)
            if len(self._time_tab_cache) >= 4:
                self._time_tab_cache.clear()
            self._time_tab_cache[key] = time_tab
        if self.cond_dim:
            cond_tab = torch.empty((cond.shape[0], self.ss_total), device=self.dev)
            ops.small_linear(cond.float().contiguous(), self.ss_w_cond, self.ss_b, cond_tab, 1, 0)
        else:
            cond_tab = self.ss_b[None].contiguous()
        return time_tab, cond_tab

    def ss_from_tables(self, time_tab, cond_tab, step_idx, b):
        ss = torch.empty((b, self.ss_total), device=self.dev)
        ct = cond_tab if cond_tab.shape[0] == b else cond_tab.expand(b, -1).contiguous()
        check(lib().lfdm_ss_combine(ptr(time_tab), ptr(step_idx), ptr(ct), ptr(ss), b, self.ss_total, stream()), "lfdm_ss_combine")
        return ss

    def _frame_tables(self, f):
        if f not in self._frame_tabs:
            bucket = _rel_pos_bucket(f, self.rel_buckets, self.rel_maxdist).to(self.dev)
            bias = self.rel_emb[bucket].permute(2, 0, 1).contiguous()          # (heads, f, f)
            ang = torch.outer(torch.arange(f, device=self.dev).float(), self.rot_freqs)   # (f, 16)
            self._frame_tabs[f] = (bias, ang.cos().contiguous(), ang.sin().contiguous())
        return self._frame_tabs[f]

    # ------------------------------------------------------------------------------------------------
    # blocks
    # ------------------------------------------------------------------------------------------------
    def _resnet(self, r, srcs_sb, x_f32, nf, h, w, rps, ss, b, want_sb, want_f32=True):
        dev = self.dev
        m = nf * h * w
        c = r.c_out
        stats, stats2 = self._next_stats(b, r.groups), self._next_stats(b, r.groups)
        h1 = f32(m, c, dev)
        r.conv1(srcs_sb, nf, h, w, out_f32=h1, gn_stats=stats, gn_groups=r.groups, rows_per_sample=rps, stats_zeroed=True)
        a1 = SB(m, c, dev)
        ss_view = ss[:, r.ss_off:r.ss_off + 2 * c] if r.has_mlp else None
        ops.gn_apply(h1, stats, r.g1, r.b1, ss_view, None, None, a1, r.groups, rps, r.eps)
        h2 = h1  # reuse buffer: conv1 output is dead after gn_apply
        r.conv2([a1], nf, h, w, out_f32=h2, gn_stats=stats2, gn_groups=r.groups, rows_per_sample=rps, stats_zeroed=True)
        if r.res is not None:
            res = f32(m, c, dev)
            r.res(srcs_sb, nf, h, w, out_f32=res)
        else:
            res = x_f32
        out = f32(m, c, dev) if want_f32 else None
        out_sb = SB(m, c, dev) if want_sb else None
        ops.gn_apply(h2, stats2, r.g2, r.b2, None, res, out, out_sb, r.groups, rps, r.eps)
        return out, out_sb

    def _attn_common(self, at, x_f32, nf, h, w, core, want_sb, rot=None):
        dev = self.dev
        m, c = x_f32.shape
        n = SB(m, c, dev)
        ops.layernorm(x_f32, at.gamma, out_sb=n, eps=at.eps)
        qkv = f32(m, 3 * at.hid, dev)
        at.qkv([n], nf, h, w, out_f32=qkv, rot=rot)
        o = SB(m, at.hid, dev)
        if rot is not None:
            core(qkv, o, at.qkv.rot_applied)
        else:
            core(qkv, o)
        out = f32(m, c, dev)
        out_sb = SB(m, c, dev) if want_sb else None
        at.out([o], nf, h, w, out_f32=out, out_sb=out_sb, residual=x_f32)
        return out, out_sb

    def _temporal(self, at, x_f32, b, f, h, w, want_sb):
        p = h * w
        bias, cos, sin = self._frame_tables(f)
        if FUSED_ATTN and at.fused is not None:
            m, c = x_f32.shape
            out = f32(m, c, self.dev)
            out_sb = SB(m, c, self.dev) if want_sb else None
            rc = ops.attn_temporal_fused(x_f32, at.gamma, at.fused[0], at.fused[1], at.out_bias, cos, sin, bias, out, out_sb,
                                         b, f, p, at.heads, at.eps)
            if rc == 0:
                return out, out_sb
        def core(qkv, o, pre=False):
            if pre:        # q*scale and the q / k rotary were applied by the qkv projection's epilogue
                ops.attn_softmax_pre(qkv, o, None, b * p, f, at.heads, p, f * p, 1, p, bias)
            else:
                ops.attn_softmax(qkv, o, None, b * p, f, at.heads, p, f * p, 1, p, cos, sin, bias)
        rot = None
        if FUSE_ROTARY and 17 <= f <= 40:
            rot = (cos, sin, f, p, 2 * at.hid, at.hid, 32 ** -0.5)
        return self._attn_common(at, x_f32, b * f, h, w, core, want_sb, rot=rot)

    def _mid_spatial(self, at, x_f32, b, f, h, w):
        p = h * w
        core = lambda qkv, o: ops.attn_softmax(qkv, o, None, b * f, p, at.heads, 1, p, 0, 1, None, None, None)
        return self._attn_common(at, x_f32, b * f, h, w, core, False)

    def _linear(self, at, x_f32, b, f, h, w):
        p = h * w
        if FUSED_LINATTN and at.fused_lin is not None and p % 128 == 0:
            m, c = x_f32.shape
            out = f32(m, c, self.dev)
            rc, _ = ops.attn_linear_fused(x_f32, at.gamma, at.fused_lin, at.out_bias, out, None, b * f, p, at.eps)
            if rc == 0:
                return out, None
        core = lambda qkv, o: ops.attn_linear(qkv, o, None, b * f, p, at.heads)
        return self._attn_common(at, x_f32, b * f, h, w, core, False)

    def _tap(self, name, x_f32, b, f, h, w):
        if self.taps is not None:
            c = x_f32.shape[1]
            self.taps[name] = x_f32.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3).clone()

    # ------------------------------------------------------------------------------------------------
    # init conv variants
    # ------------------------------------------------------------------------------------------------
    def _hoisted_layers(self):
        if self._hoist is None:
            k, c0 = self.k_init, self.unet_channels
            n_x = 3                                   # diffused channels (flow x, flow y, occlusion)
            wx = self._init_w[:, :n_x]                # (Co, 3, k, k)
            kreal = k * k * n_x
            kpad = (kreal + 63) // 64 * 64
            wflat = torch.zeros(self.init_dim, kpad, device=self.dev)
            wflat[:, :kreal] = wx.permute(0, 2, 3, 1).reshape(self.init_dim, kreal)   # k index = tap*3 + ch
            lx = ConvLayer(wflat.reshape(self.init_dim, kpad, 1, 1), self._init_b, name="init_conv.x")
            lf = ConvLayer(self._init_w[:, n_x:].contiguous(), None, pad=k // 2, name="init_conv.fea")
            self._hoist = (lx, lf, kpad, n_x)
        return self._hoist

    def prepare_fea(self, fea):
        """conv_fea(fea) of the split init_conv, once per sample: fea (B, C_fea, H, W) -> F32 rows [B*H*W][init_dim]"""
        lx, lf, kpad, n_x = self._hoisted_layers()
        b, cf, h, w = fea.shape
        assert cf == self.unet_channels - n_x
        rows = SB(b * h * w, cf, self.dev)
        ops.to_rows(fea.float().contiguous().reshape(b, cf, 1, h, w), out_sb=rows)
        out = f32(b * h * w, self.init_dim, self.dev)
        lf([rows], b, h, w, out_f32=out)
        return out

    # ------------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------------
    def forward(self, x, time, cond, cfg_scale=None):
        """generic Unet3D.forward: x (B, channels, F, H, W) NCDHW fp32 -> (B, 3, F, H, W)
        (cfg_scale: the batch is [cond half ; null half], see forward_hoisted)"""
        b, c, f, h, w = x.shape
        assert c == self.unet_channels
        ss = self.scale_shift(time, cond)
        rows = SB(b * f * h * w, self.cin_pad, self.dev)
        ops.to_rows(x.float(), c_pad=self.cin_pad, out_sb=rows)
        m = b * f * h * w
        x0, x0_sb = f32(m, self.init_dim, self.dev), SB(m, self.init_dim, self.dev)
        self.init_full([rows], b * f, h, w, out_f32=x0, out_sb=x0_sb)
        return self._body(x0, x0_sb, ss, b, f, h, w, cfg_scale)

    def forward_hoisted(self, x3, fea_conv, ss, cfg_scale=None):
        """x3 (B, 3, F, H, W); fea_conv from prepare_fea; ss (B, ss_total) -> (B, 3, F, H, W).
        cfg_scale: classifier-free guidance as ONE 2B batch (reference :521-526 runs two forwards): x3 / fea_conv / ss hold
        [conditional half ; null-condition half]; the result is null + (cond - null) * cfg_scale of shape (B/2, 3, F, H, W)."""
        lx, lf, kpad, n_x = self._hoisted_layers()
        b, c, f, h, w = x3.shape
        m = b * f * h * w
        cols = SB(m, kpad, self.dev)
        check(lib().lfdm_im2col_small(ptr(x3), b, c, f, h, w, self.k_init, self.k_init // 2, kpad, ptr(cols.t),
                                      cols.plane, stream()), "lfdm_im2col_small")
        x0, x0_sb = f32(m, self.init_dim, self.dev), SB(m, self.init_dim, self.dev)
        lx([cols], b * f, h, w, out_f32=x0, out_sb=x0_sb, residual=fea_conv, res_bcast_f=f)
        return self._body(x0, x0_sb, ss, b, f, h, w, cfg_scale)

    def _next_stats(self, b, groups):
        """GroupNorm accumulators of one evaluation come from one pool zeroed by a single fill (not one per conv)"""
        n = b * groups * 2
        if self._stats_pool is None or self._stats_off + n > self._stats_pool.numel():
            self._stats_pool = torch.zeros(max(64 * n, 4096), dtype=torch.float64, device=self.dev)
            self._stats_off = 0
        t = self._stats_pool[self._stats_off:self._stats_off + n].view(b, groups, 2)
        self._stats_off += n
        return t

    def _body(self, r_f32, r_sb, ss, b, f, h, w, cfg_scale=None):
        nf = b * f
        self._stats_pool, self._stats_off = None, 0
        self._tap("init_conv", r_f32, b, f, h, w)
        x, x_sb = self._temporal(self.init_tattn, r_f32, b, f, h, w, want_sb=True)
        self._tap("init", x, b, f, h, w)
        skips = []
        for i, st in enumerate(self.stages_down):
            rps = f * h * w
            x, x_sb = self._resnet(st["b1"], [x_sb], x, nf, h, w, rps, ss, b, want_sb=True)
            x, _ = self._resnet(st["b2"], [x_sb], x, nf, h, w, rps, ss, b, want_sb=False)
            if st["sa"] is not None:
                x, _ = self._linear(st["sa"], x, b, f, h, w)
            x, x_sb = self._temporal(st["ta"], x, b, f, h, w, want_sb=True)
            skips.append(x_sb)
            if st["down"] is not None:
                ho, wo = st["down"].out_hw(h, w)
                y, y_sb = f32(nf * ho * wo, x.shape[1], self.dev), SB(nf * ho * wo, x.shape[1], self.dev)
                st["down"]([x_sb], nf, h, w, out_f32=y, out_sb=y_sb)
                x, x_sb, h, w = y, y_sb, ho, wo
            self._tap(f"down{i}", x, b, f, h, w)
        rps = f * h * w
        x, _ = self._resnet(self.mid1, [x_sb], x, nf, h, w, rps, ss, b, want_sb=False)
        x, _ = self._mid_spatial(self.mid_sattn, x, b, f, h, w)
        x, x_sb = self._temporal(self.mid_tattn, x, b, f, h, w, want_sb=True)
        x, x_sb = self._resnet(self.mid2, [x_sb], x, nf, h, w, rps, ss, b, want_sb=True)
        self._tap("mid", x, b, f, h, w)
        for i, st in enumerate(self.stages_up):
            rps = f * h * w
            skip_sb = skips.pop()
            x, x_sb = self._resnet(st["b1"], [x_sb, skip_sb], None, nf, h, w, rps, ss, b, want_sb=True)
            x, _ = self._resnet(st["b2"], [x_sb], x, nf, h, w, rps, ss, b, want_sb=False)
            if st["sa"] is not None:
                x, _ = self._linear(st["sa"], x, b, f, h, w)
            x, x_sb = self._temporal(st["ta"], x, b, f, h, w, want_sb=True)
            if st["up"] is not None:
                ho, wo = st["up"].out_hw(h, w)
                need_f32 = self.taps is not None
                y = f32(nf * ho * wo, x.shape[1], self.dev) if need_f32 else None
                y_sb = SB(nf * ho * wo, x.shape[1], self.dev)
                st["up"]([x_sb], nf, h, w, out_f32=y, out_sb=y_sb)
                x, x_sb, h, w = y, y_sb, ho, wo
            if x is not None:
                self._tap(f"up{i}", x, b, f, h, w)
        rps = f * h * w
        a, _ = self._resnet(self.head_a, [x_sb, r_sb], None, nf, h, w, rps, ss, b, want_sb=False)
        o, _ = self._resnet(self.head_o, [x_sb, r_sb], None, nf, h, w, rps, ss, b, want_sb=False)
        if cfg_scale is not None:
            assert b % 2 == 0
            out = torch.empty((b // 2, self.out_grid_dim + self.out_conf_dim, f, h, w), device=self.dev)
            check(lib().lfdm_unet_heads_cfg(ptr(a), ptr(self.wa), ptr(self.ba), self.out_grid_dim, ptr(o), ptr(self.wo),
                                            ptr(self.bo), self.out_conf_dim, a.shape[1], b // 2, f, h * w, float(cfg_scale),
                                            ptr(out), stream()), "lfdm_unet_heads_cfg")
            return out
        out = torch.empty((b, self.out_grid_dim + self.out_conf_dim, f, h, w), device=self.dev)
        check(lib().lfdm_unet_heads(ptr(a), ptr(self.wa), ptr(self.ba), self.out_grid_dim, ptr(o), ptr(self.wo),
                                    ptr(self.bo), self.out_conf_dim, a.shape[1], b, f, h * w, ptr(out), stream()),
              "lfdm_unet_heads")
        return out
