"""GPU (-m gpu): per-kernel parity of the sm_100a kernels (called through the C-ABI) against plain PyTorch fp32 on
the CPU.  Tolerances: the north-star's rtol=1e-3 / atol=1e-4 (fp32); bit-exact where the arithmetic is order-
defined (sampler update, quantile, layout kernels)."""
import math
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


def dev():
    return torch.device("cuda:0")


def rows_of(x4):
    """(N, C, H, W) -> rows [(n*H+h)*W+w][C]"""
    n, c, h, w = x4.shape
    return x4.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def nchw_of(rows, n, h, w):
    return rows.reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


def make_sb(rows):
    from cvpr23_lfdm_b200._lib import SB
    from cvpr23_lfdm_b200.engine.ops import split_planes
    s = SB(rows.shape[0], rows.shape[1], dev())
    s.t.copy_(split_planes(rows.to(dev())))
    return s


def close(a, b, what, rtol=RTOL, atol=ATOL):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    lim = atol + rtol * b.abs()
    bad = (err > lim).sum().item()
    assert bad == 0, f"{what}: {bad}/{a.numel()} elements out of tolerance, max abs err {err.max().item():.3e}, ref rms {b.pow(2).mean().sqrt().item():.3e}"


# ----------------------------------------------------------------------------------------------------------------
# convolution engines
# ----------------------------------------------------------------------------------------------------------------
def ref_conv(x, w, b, mode, stride, pad, reflect):
    from cvpr23_lfdm_b200 import _lib as L
    if mode == L.CONV_TRANSPOSED:
        return F.conv_transpose2d(x, w, b, stride=stride, padding=pad)
    if mode == L.CONV_UPNEAREST:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    if reflect:
        x = F.pad(x, (pad, pad, pad, pad), mode="reflect")
        return F.conv2d(x, w, b, stride=stride)
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def run_conv(engine, n, cins, cout, h, w, k, pad, mode=0, stride=1, reflect=False, bias=True, residual=False,
             res_bcast=0, f32_act=0, sb_aff=False, sb_act=0, gn_groups=0, a_f32=False, seed=0, fps=1, f32_only=False):
    from cvpr23_lfdm_b200 import _lib as L
    from cvpr23_lfdm_b200.engine.ops import ConvLayer, f32
    from cvpr23_lfdm_b200._lib import SB
    g = torch.Generator().manual_seed(seed)
    cin = sum(cins)
    xs = [torch.randn(n, c, h, w, generator=g) for c in cins]
    x = torch.cat(xs, 1)
    if mode == L.CONV_TRANSPOSED:
        wt = torch.randn(cin, cout, k, k, generator=g) / math.sqrt(cin * k * k)
    else:
        wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g) if bias else None
    y = ref_conv(x, wt, b, mode, stride, pad, reflect)
    _, _, ho, wo = y.shape
    layer = ConvLayer(wt.to(dev()), b.to(dev()) if bias else None, mode=mode, stride=stride, pad=pad, reflect=reflect,
                      src_channels=cins, engine="tc" if engine == L.ENGINE_TC else "simt")
    if engine == L.ENGINE_TC:
        assert layer.w_sb is not None, "geometry must be TC-eligible for this test"
    srcs = [rows_of(t).to(dev()) if (a_f32 and i == 0) else make_sb(rows_of(t)) for i, t in enumerate(xs)]
    m_out = n * ho * wo
    res_rows = None
    yv = rows_of(y)
    if residual:
        if res_bcast:
            frames = res_bcast
            r4 = torch.randn(n // frames, cout, ho, wo, generator=g)
            res_rows = rows_of(r4).to(dev())
            yv = rows_of(y + r4.repeat_interleave(frames, 0))
        else:
            r4 = torch.randn(n, cout, ho, wo, generator=g)
            res_rows = rows_of(r4).to(dev())
            yv = rows_of(y + r4)
    out_f32 = f32(m_out, cout, dev())
    out_sb = None if f32_only else SB(m_out, cout, dev())
    sc = sh = None
    if sb_aff:
        sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    stats = torch.zeros((n // fps, gn_groups, 2), dtype=torch.float64, device=dev()) if gn_groups else None
    layer(srcs, n, h, w, out_f32=out_f32, out_sb=out_sb, residual=res_rows, res_bcast_f=res_bcast, f32_act=f32_act,
          sb_act=sb_act, sb_scale=sc.to(dev()) if sb_aff else None, sb_shift=sh.to(dev()) if sb_aff else None,
          gn_stats=stats, gn_groups=gn_groups or 8, rows_per_sample=fps * ho * wo)
    torch.cuda.synchronize()
    if engine == L.ENGINE_TC:
        assert layer.last_engine == "tc", "tcgen05 engine rejected this geometry"
    act = {0: lambda v: v, 1: torch.relu, 2: torch.sigmoid}
    close(out_f32, act[f32_act](yv), "out_f32")
    u = yv
    if sb_aff:
        u = u * sc[None] + sh[None]
    if out_sb is not None:
        close(out_sb.float(), act[sb_act](u), "out_sb", rtol=RTOL, atol=ATOL)
    if gn_groups:
        cpg = cout // gn_groups
        v = yv.reshape(n // fps, fps * ho * wo, gn_groups, cpg).double()
        s_ref = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], -1)
        close(stats, s_ref, "gn_stats", rtol=1e-4, atol=1e-2)
        from cvpr23_lfdm_b200.engine import ops as _ops
        if engine == L.ENGINE_TC and _ops.FUSED_GN_STATS:
            # per-CTA sums are added in a fixed order and the cross-CTA double atomics are exact (fixed 2^-24 grid): bit-identical runs
            first, out_first = stats.clone(), out_f32.clone()
            for _ in range(4):
                stats.zero_()
                layer(srcs, n, h, w, out_f32=out_f32, out_sb=out_sb, residual=res_rows, res_bcast_f=res_bcast, f32_act=f32_act,
                      sb_act=sb_act, sb_scale=sc.to(dev()) if sb_aff else None, sb_shift=sh.to(dev()) if sb_aff else None,
                      gn_stats=stats, gn_groups=gn_groups or 8, rows_per_sample=fps * ho * wo)
                torch.cuda.synchronize()
                assert torch.equal(stats, first), "GroupNorm sums differ between two runs of the same launch"
                assert torch.equal(out_f32, out_first)


SIMT_CASES = [
    dict(n=2, cins=[12], cout=20, h=9, w=7, k=3, pad=1),
    dict(n=2, cins=[8, 12], cout=16, h=8, w=8, k=3, pad=1, residual=True, f32_act=1, sb_aff=True, sb_act=1),
    dict(n=1, cins=[11], cout=16, h=8, w=8, k=7, pad=3, a_f32=True),
    dict(n=2, cins=[16], cout=16, h=8, w=8, k=4, pad=1, stride=2),
    dict(n=2, cins=[16], cout=24, h=4, w=4, k=4, pad=1, stride=2, mode=1),
    dict(n=2, cins=[16], cout=8, h=5, w=6, k=3, pad=1, mode=2),
    dict(n=2, cins=[16], cout=8, h=5, w=6, k=3, pad=1, mode=2, reflect=True),
    dict(n=1, cins=[8], cout=8, h=6, w=6, k=3, pad=1, reflect=True),
    dict(n=4, cins=[16], cout=3, h=8, w=8, k=1, pad=0, residual=True, res_bcast=2, f32_act=2),
    dict(n=2, cins=[70], cout=66, h=10, w=10, k=7, pad=0),
]


@pytest.mark.parametrize("case", SIMT_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k not in ("n",)))
def test_conv_simt(case):
    from cvpr23_lfdm_b200 import _lib as L
    run_conv(L.ENGINE_SIMT, **case)


TC_CASES = [
    dict(n=2, cins=[64], cout=64, h=32, w=32, k=3, pad=1),                                  # 32x32 tile 4 rows
    dict(n=2, cins=[64], cout=64, h=32, w=32, k=1, pad=0, bias=False),                      # plain GEMM
    dict(n=2, cins=[128], cout=256, h=16, w=16, k=3, pad=1, residual=True),                 # BN=128, 16x16
    dict(n=4, cins=[64, 64], cout=64, h=8, w=8, k=3, pad=1, gn_groups=8, fps=2),                   # concat, bnf=2, fused GN
    dict(n=8, cins=[256, 256], cout=128, h=4, w=4, k=3, pad=1, gn_groups=8, fps=8),                # 4x4, bnf=8
    dict(n=8, cins=[128], cout=128, h=4, w=4, k=1, pad=0),
    dict(n=1, cins=[256], cout=64, h=32, w=32, k=7, pad=3, bias=False),                     # 49 taps
    dict(n=2, cins=[64], cout=64, h=32, w=32, k=4, pad=1, stride=2),                        # parity views
    dict(n=2, cins=[128], cout=128, h=8, w=8, k=4, pad=1, stride=2, mode=1),                # ConvTranspose phases
    dict(n=2, cins=[128], cout=64, h=16, w=16, k=3, pad=1, mode=2, f32_act=1),              # up-nearest phases
    dict(n=1, cins=[64], cout=3, h=128, w=128, k=7, pad=3, f32_act=2),                      # ragged N, W=128
    dict(n=2, cins=[192], cout=64, h=32, w=32, k=1, pad=0, residual=True, res_bcast=2),     # hoisted init conv
    dict(n=2, cins=[64], cout=768, h=16, w=16, k=1, pad=0, bias=False),                     # qkv projection
    dict(n=2, cins=[256], cout=256, h=32, w=32, k=3, pad=1, residual=True, sb_aff=True, sb_act=1),   # LFAE ResBlock conv2
    dict(n=1, cins=[64], cout=128, h=128, w=128, k=3, pad=1, f32_act=1),                    # LFAE down0
    dict(n=1, cins=[64], cout=64, h=64, w=64, k=3, pad=1),                                  # 64x64 (bw=64, bh=2)
    dict(n=2, cins=[64], cout=768, h=16, w=16, k=1, pad=0, bias=False, f32_only=True),      # TMA-store epilogue (qkv)
    dict(n=4, cins=[64], cout=64, h=8, w=8, k=3, pad=1, gn_groups=8, fps=2, f32_only=True), # TMA-store + GroupNorm sums
    dict(n=2, cins=[128, 128], cout=128, h=16, w=16, k=3, pad=1, gn_groups=8, fps=1, f32_only=True),
    dict(n=40, cins=[64], cout=64, h=32, w=32, k=3, pad=1, gn_groups=8, fps=20, f32_only=True),   # halo mode, 2-3 tiles per CTA
    dict(n=3, cins=[64], cout=64, h=16, w=8, k=3, pad=1),                                         # halo mode, bw=8 bh=16
    dict(n=2, cins=[64], cout=128, h=64, w=32, k=3, pad=1, residual=True),                        # halo mode, 16 row tiles
    dict(n=2, cins=[64], cout=64, h=16, w=16, k=3, pad=1, mode=2, reflect=True),                  # NATOPS up-conv: reflect == clamp on the low-res map
    dict(n=8, cins=[256], cout=256, h=4, w=4, k=3, pad=1, mode=2, reflect=True, f32_act=1),       # ... 4x4 -> 8x8, BN=128 (8 frames per 128-row tile)
    # stream-K (more tiles than SMs, last wave mostly empty): tiles cut between CTAs, partial accumulators handed over
    dict(n=320, cins=[512], cout=512, h=4, w=4, k=3, pad=1, gn_groups=8, fps=40, f32_only=True),  # 160 tiles x 72 K-blocks (the 4x4 level)
    dict(n=320, cins=[256], cout=256, h=8, w=8, k=3, pad=1, residual=True),                       # 320 tiles x 36 K-blocks (the 8x8 level)
    dict(n=640, cins=[64], cout=64, h=8, w=8, k=3, pad=1, gn_groups=8, fps=20),                   # 320 tiles x 9 K-blocks, BN = 64
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k not in ("n",)))
def test_conv_tc(case):
    from cvpr23_lfdm_b200 import _lib as L
    run_conv(L.ENGINE_TC, **case)


# ----------------------------------------------------------------------------------------------------------------
# norms
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,groups,with_ss,with_res", [(16, 8, True, False), (64, 8, True, True), (128, 8, False, True)])
def test_groupnorm(c, groups, with_ss, with_res):
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import SB, lib, ptr, stream, check
    g = torch.Generator().manual_seed(1)
    b, f, h, w = 2, 3, 8, 8
    x = torch.randn(b, c, f, h, w, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    ss_full = torch.randn(b, 2 * c + 10, generator=g)
    ss = ss_full[:, 5:5 + 2 * c]
    res = torch.randn(b, c, f, h, w, generator=g)
    y = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    if with_ss:
        y = y * (ss[:, :c, None, None, None] + 1) + ss[:, c:, None, None, None]
    y = F.silu(y)
    if with_res:
        y = y + res
    to_rows5 = lambda t: t.permute(0, 2, 3, 4, 1).reshape(-1, c).contiguous()
    xr = to_rows5(x).to(dev())
    stats = torch.empty((b, groups, 2), dtype=torch.float64, device=dev())
    check(lib().lfdm_gn_stats(ptr(xr), xr.shape[0], c, groups, f * h * w, ptr(stats), stream()), "gn_stats")
    out, out_sb = torch.empty_like(xr), SB(xr.shape[0], c, dev())
    ssd = ss_full.to(dev())[:, 5:5 + 2 * c] if with_ss else None
    ops.gn_apply(xr, stats, gamma.to(dev()), beta.to(dev()), ssd, to_rows5(res).to(dev()) if with_res else None, out, out_sb,
                 groups, f * h * w)
    close(out, to_rows5(y), "gn_apply f32")
    close(out_sb.float(), to_rows5(y), "gn_apply sb")


@pytest.mark.parametrize("c", [16, 64, 512])
def test_layernorm(c):
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import SB
    g = torch.Generator().manual_seed(2)
    x = torch.randn(300, c, generator=g) * 3 + 1
    gamma = torch.randn(c, generator=g)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    y = (x - x.mean(1, keepdim=True)) / (var + 1e-5).sqrt() * gamma
    out, out_sb = torch.empty(300, c, device=dev()), SB(300, c, dev())
    ops.layernorm(x.to(dev()), gamma.to(dev()), out_sb=out_sb, out_f32=out)
    close(out, y, "layernorm f32")
    close(out_sb.float(), y, "layernorm sb")


# ----------------------------------------------------------------------------------------------------------------
# attention cores
# ----------------------------------------------------------------------------------------------------------------
def test_attn_temporal_and_spatial():
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import SB
    g = torch.Generator().manual_seed(3)
    b, f, p, heads = 2, 40, 16, 4
    hid = heads * 32
    qkv = torch.randn(b, f, p, 3 * hid, generator=g)        # rows ordered (b, f, p)
    bias = torch.randn(heads, f, f, generator=g)
    freqs = O.rotary_freqs(32)
    # temporal: sequences over f for each (b, p)
    x = qkv.permute(0, 2, 1, 3)                               # (b, p, f, 3hid)
    q, k, v = [t.reshape(b, p, f, heads, 32).transpose(-2, -3) for t in x.chunk(3, -1)]
    q = O.rotary_apply(q * 32 ** -0.5, freqs)
    k = O.rotary_apply(k, freqs)
    sim = torch.einsum("...hid,...hjd->...hij", q, k) + bias
    att = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    o = torch.einsum("...hij,...hjd->...hid", att, v).transpose(-2, -3).reshape(b, p, f, hid)
    ref_t = o.permute(0, 2, 1, 3).reshape(b * f * p, hid)
    ang = torch.outer(torch.arange(f).float(), freqs)
    qkv_d = qkv.reshape(b * f * p, 3 * hid).to(dev())
    out, out_sb = torch.empty(b * f * p, hid, device=dev()), SB(b * f * p, hid, dev())
    ops.attn_softmax(qkv_d, out_sb, out, b * p, f, heads, p, f * p, 1, p, ang.cos().contiguous().to(dev()),
                     ang.sin().contiguous().to(dev()), bias.to(dev()))
    close(out, ref_t, "temporal attention f32")
    close(out_sb.float(), ref_t, "temporal attention sb")
    # spatial (mid block): sequences over p for each (b, f); no rotary / bias
    q, k, v = [t.reshape(b, f, p, heads, 32).transpose(-2, -3) for t in qkv.chunk(3, -1)]
    sim = torch.einsum("...hid,...hjd->...hij", q * 32 ** -0.5, k)
    att = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    ref_s = torch.einsum("...hij,...hjd->...hid", att, v).transpose(-2, -3).reshape(b * f * p, hid)
    ops.attn_softmax(qkv_d, None, out, b * f, p, heads, 1, p, 0, 1)
    close(out, ref_s, "spatial attention")


def test_qkv_projection_with_fused_rotary_and_prescaled_attention():
    """lfdm_conv(rot_*) + lfdm_attn_softmax_pre == projection, then q*scale, rotary(q), rotary(k), attention (reference order)"""
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200.engine.ops import ConvLayer, f32
    g = torch.Generator().manual_seed(77)
    b, f, hh, ww, heads, c = 2, 24, 8, 4, 2, 64
    p, hid = hh * ww, heads * 32
    x = torch.randn(b * f, c, hh, ww, generator=g)
    wq = torch.randn(3 * hid, c, 1, 1, generator=g) / 8
    bias = torch.randn(heads, f, f, generator=g)
    freqs = O.rotary_freqs(32)
    ang = torch.outer(torch.arange(f).float(), freqs)
    cos, sin = ang.cos().contiguous().to(dev()), ang.sin().contiguous().to(dev())
    layer = ConvLayer(wq.to(dev()), None, pad=0)
    qkv = f32(b * f * p, 3 * hid, dev())
    layer([make_sb(rows_of(x))], b * f, hh, ww, out_f32=qkv, rot=(cos, sin, f, p, 2 * hid, hid, 32 ** -0.5))
    assert layer.rot_applied and layer.last_engine == "tc"
    y = rows_of(F.conv2d(x, wq)).reshape(b, f, p, 3 * hid)                  # rows (b, f, p)
    q, k, v = [t.reshape(b, f, p, heads, 32).permute(0, 2, 3, 1, 4) for t in y.chunk(3, -1)]   # (b, p, h, f, d)
    q = O.rotary_apply(q * 32 ** -0.5, freqs)
    k = O.rotary_apply(k, freqs)
    back = lambda t: t.permute(0, 3, 1, 2, 4).reshape(b * f * p, hid)
    close(qkv, torch.cat([back(q), back(k), back(v)], 1), "qkv with fused rotary")
    sim = torch.einsum("...id,...jd->...ij", q, k) + bias[None, None]
    att = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    ref = back(torch.einsum("...ij,...jd->...id", att, v))
    out = torch.empty(b * f * p, hid, device=dev())
    ops.attn_softmax_pre(qkv, None, out, b * p, f, heads, p, f * p, 1, p, bias.to(dev()))
    close(out, ref, "attention on pre-rotated q|k")


def _temporal_block_reference(x_rows, gamma, wqkv, wout, out_bias, pos_bias, b, f, p, heads, eps=1e-5):
    """float64 restatement of Residual(PreNorm(EinopsToAndFrom(Attention))) (reference :132-138,170-190,270-283,286-363)
    on rows [(b*f + fi)*p + pi][c]; returns out rows and the intermediates the fused kernel can dump."""
    from oracle import lfdm_oracle as O
    hid = heads * 32
    x = x_rows.double()
    var = x.var(dim=1, unbiased=False, keepdim=True)
    xn = (x - x.mean(1, keepdim=True)) / (var + eps).sqrt() * gamma.double()
    qkv = (xn @ wqkv.double().t()).reshape(b, f, p, 3 * hid).permute(0, 2, 1, 3)       # (b, p, f, 3 hid)
    q, k, v = [t.reshape(b, p, f, heads, 32).transpose(-2, -3) for t in qkv.chunk(3, -1)]   # (b, p, h, f, d)
    freqs = O.rotary_freqs(32).double()
    q = O.rotary_apply(q * 32 ** -0.5, freqs)
    k = O.rotary_apply(k, freqs)
    sim = torch.einsum("...hid,...hjd->...hij", q, k)
    if pos_bias is not None:
        sim = sim + pos_bias.double()
    att = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    o = torch.einsum("...hij,...hjd->...hid", att, v)                                   # (b, p, h, f, d)
    rows = lambda t: t.permute(0, 3, 1, 2, 4).reshape(b * f * p, -1)                     # -> [(b, f, p)][h*d]
    o_rows = rows(o)
    out = o_rows @ wout.double().t() + x
    if out_bias is not None:
        out = out + out_bias.double()
    att_rows = att.permute(0, 3, 1, 2, 4).reshape(b * f * p, heads * f)                  # row (b, fi, p): [h][j]
    return out.float(), dict(q=rows(q).float(), k=rows(k).float(), v=rows(v).float(), att=att_rows.float(), o=o_rows.float())


@pytest.mark.parametrize("b,p,heads,with_bias", [(2, 10, 8, False), (1, 7, 3, True), (3, 400, 8, False)])
def test_temporal_attention_block_fused(b, p, heads, with_bias):
    """lfdm_attn_temporal_fused (one tcgen05 kernel: LN -> qkv -> rotary -> softmax(QK^T + bias) V -> to_out -> + x) against
    a float64 restatement, including the kernel's diagnostic dump of every intermediate; partial last tile, tiles that
    straddle samples, several tiles per CTA (the third case has 400 tiles > 148 SMs)."""
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import SB
    g = torch.Generator().manual_seed(500 + p)
    f, c, hid = 40, 64, heads * 32
    m = b * f * p
    x = torch.randn(m, c, generator=g) * 1.5 + 0.3
    gamma = torch.randn(c, generator=g)
    wqkv = torch.randn(3 * hid, c, generator=g) / 8
    wout = torch.randn(c, hid, generator=g) / 16
    ob = torch.randn(c, generator=g) if with_bias else None
    pos = torch.randn(heads, f, f, generator=g)
    ref, mid = _temporal_block_reference(x, gamma, wqkv, wout, ob, pos, b, f, p, heads)
    ang = torch.outer(torch.arange(f).float(), O.rotary_freqs(32))
    wq_img, wo_img = ops.pack_fused_attention(wqkv.to(dev()), wout.to(dev()), heads)
    out, out_sb = torch.empty(m, c, device=dev()), SB(m, c, dev())
    dbg = torch.zeros(m, 3 * hid + heads * f + hid, device=dev())
    rc = ops.attn_temporal_fused(x.to(dev()), gamma.to(dev()), wq_img, wo_img, ob.to(dev()) if with_bias else None,
                                 ang.cos().contiguous().to(dev()), ang.sin().contiguous().to(dev()), pos.to(dev()),
                                 out, out_sb, b, f, p, heads, 1e-5, debug=dbg)
    assert rc == 0
    torch.cuda.synchronize()
    d = dbg.cpu()
    close(d[:, :hid], mid["q"], "fused block: rotated q")
    close(d[:, hid:2 * hid], mid["k"], "fused block: rotated k")
    close(d[:, 2 * hid:3 * hid], mid["v"], "fused block: v")
    close(d[:, 3 * hid:3 * hid + heads * f], mid["att"], "fused block: softmax rows", atol=2e-5)
    close(d[:, 3 * hid + heads * f:], mid["o"], "fused block: head outputs")
    close(out, ref, "fused block: out f32")
    close(out_sb.float(), ref, "fused block: out sb")
    # without the diagnostic buffer (the production call) the result is the same
    out2 = torch.empty(m, c, device=dev())
    assert ops.attn_temporal_fused(x.to(dev()), gamma.to(dev()), wq_img, wo_img, ob.to(dev()) if with_bias else None,
                                   ang.cos().contiguous().to(dev()), ang.sin().contiguous().to(dev()), pos.to(dev()),
                                   out2, None, b, f, p, heads, 1e-5) == 0
    assert torch.equal(out2, out)


def _linear_block_reference(x_rows, gamma, wqkv, wout, out_bias, frames, pos, heads=8, eps=1e-5):
    """float64 restatement of Residual(PreNorm(SpatialLinearAttention)) (reference :132-138,170-190,240-265) on rows
    [frame * pos + p][c]"""
    hid = heads * 32
    x = x_rows.double()
    var = x.var(dim=1, unbiased=False, keepdim=True)
    xn = (x - x.mean(1, keepdim=True)) / (var + eps).sqrt() * gamma.double()
    qkv = (xn @ wqkv.double().t()).reshape(frames, pos, 3, heads, 32)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]              # (frames, pos, heads, 32)
    q = q.softmax(dim=-1) * 32 ** -0.5                              # over d
    k = k.softmax(dim=1)                                             # over the positions of the frame
    ctx = torch.einsum("fnhd,fnhe->fhde", k, v)
    o = torch.einsum("fhde,fnhd->fnhe", ctx, q).reshape(frames * pos, hid)
    out = o @ wout.double().t() + out_bias.double() + x
    return out.float()


@pytest.mark.parametrize("frames,pos,wscale", [(3, 128, 1.0), (5, 256, 1.0), (2, 1024, 1.0), (40, 1024, 1.0), (7, 512, 6.0)])
def test_linear_attention_block_fused(frames, pos, wscale):
    """lfdm_attn_linear_fused (context partials -> merge -> apply, all GEMM stages on tcgen05) against a float64 restatement:
    one / several tiles per frame, frames cut into segments across CTAs (40 x 1024: 640 tiles of 64 rows over 148 CTAs), and
    large logits (wscale 6: the running maximum of a row grows by more than 2^8 between tiles -> accumulator re-basing)."""
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import SB
    g = torch.Generator().manual_seed(900 + frames + pos)
    c, hid = 64, 256
    m = frames * pos
    x = torch.randn(m, c, generator=g) * 1.5 + 0.3
    gamma = torch.randn(c, generator=g)
    wqkv = torch.randn(3 * hid, c, generator=g) / 8 * wscale
    wout = torch.randn(c, hid, generator=g) / 16
    ob = torch.randn(c, generator=g)
    ref = _linear_block_reference(x, gamma, wqkv, wout, ob, frames, pos)
    packed = ops.pack_fused_linear_attention(wqkv.to(dev()), wout.to(dev()))
    out, out_sb = torch.empty(m, c, device=dev()), SB(m, c, dev())
    rc, work = ops.attn_linear_fused(x.to(dev()), gamma.to(dev()), packed, ob.to(dev()), out, out_sb, frames, pos, 1e-5)
    assert rc == 0
    torch.cuda.synchronize()
    close(out, ref, "fused linear block: out f32")
    close(out_sb.float(), ref, "fused linear block: out sb")
    out2 = torch.empty(m, c, device=dev())
    rc, _ = ops.attn_linear_fused(x.to(dev()), gamma.to(dev()), packed, ob.to(dev()), out2, None, frames, pos, 1e-5, work=work)
    assert rc == 0 and torch.equal(out2, out)


@pytest.mark.parametrize("f", [17, 24, 33, 39])
def test_attn_temporal_ragged_lengths(f):
    """rotary + bias on frame counts that leave partial 16-row query tiles / odd bias rows (ldmatrix kernel, 17 <= L <= 40)"""
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine import ops
    g = torch.Generator().manual_seed(60 + f)
    b, p, heads = 2, 6, 3
    hid = heads * 32
    qkv = torch.randn(b, f, p, 3 * hid, generator=g)
    bias = torch.randn(heads, f, f, generator=g)
    freqs = O.rotary_freqs(32)
    x = qkv.permute(0, 2, 1, 3)
    q, k, v = [t.reshape(b, p, f, heads, 32).transpose(-2, -3) for t in x.chunk(3, -1)]
    q = O.rotary_apply(q * 32 ** -0.5, freqs)
    k = O.rotary_apply(k, freqs)
    sim = torch.einsum("...hid,...hjd->...hij", q, k) + bias
    att = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    o = torch.einsum("...hij,...hjd->...hid", att, v).transpose(-2, -3).reshape(b, p, f, hid)
    ref = o.permute(0, 2, 1, 3).reshape(b * f * p, hid)
    ang = torch.outer(torch.arange(f).float(), freqs)
    out = torch.empty(b * f * p, hid, device=dev())
    ops.attn_softmax(qkv.reshape(b * f * p, 3 * hid).to(dev()), None, out, b * p, f, heads, p, f * p, 1, p,
                     ang.cos().contiguous().to(dev()), ang.sin().contiguous().to(dev()), bias.to(dev()))
    close(out, ref, f"temporal attention f={f}")


@pytest.mark.parametrize("L", [5, 17, 24, 31, 33, 40, 49, 64])
def test_attn_softmax_lengths(L):
    """every register-tile configuration of the softmax attention core (L <= 16, <= 40, <= 64 with two row passes)"""
    from cvpr23_lfdm_b200.engine import ops
    g = torch.Generator().manual_seed(30 + L)
    n_seq, heads = 5, 3
    hid = heads * 32
    qkv = torch.randn(n_seq, L, 3 * hid, generator=g)
    q, k, v = [t.reshape(n_seq, L, heads, 32).transpose(1, 2) for t in qkv.chunk(3, -1)]
    sim = torch.einsum("shid,shjd->shij", q * 32 ** -0.5, k)
    att = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
    ref = torch.einsum("shij,shjd->shid", att, v).transpose(1, 2).reshape(n_seq * L, hid)
    out = torch.empty(n_seq * L, hid, device=dev())
    qd = qkv.reshape(n_seq * L, 3 * hid).to(dev())
    ops.attn_softmax(qd, None, out, n_seq, L, heads, 1, L, 0, 1)
    close(out, ref, f"softmax attention L={L}")


@pytest.mark.parametrize("n_pos", [16, 50, 64, 1024])
def test_attn_linear(n_pos):
    from cvpr23_lfdm_b200.engine import ops
    g = torch.Generator().manual_seed(4)
    nf, heads = 3, 2
    hid = heads * 32
    qkv = torch.randn(nf, n_pos, 3 * hid, generator=g) * 1.5
    # a rising trend on k: the running column maximum of the single-pass (online softmax) kernel moves in every chunk
    qkv[:, :, hid:2 * hid] += torch.linspace(-6.0, 6.0, n_pos)[None, :, None]
    q, k, v = [t.reshape(nf, n_pos, heads, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, -1)]    # (nf, h, d, n)
    q = q.softmax(dim=-2) * 32 ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q)                                                 # (nf, h, e, n)
    ref = o.permute(0, 3, 1, 2).reshape(nf * n_pos, hid)
    out = torch.empty(nf * n_pos, hid, device=dev())
    ops.attn_linear(qkv.reshape(nf * n_pos, 3 * hid).to(dev()), None, out, nf, n_pos, heads)
    close(out, ref, "linear attention")


# ----------------------------------------------------------------------------------------------------------------
# embeddings
# ----------------------------------------------------------------------------------------------------------------
def test_small_linear_and_sinusoidal():
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import lib, ptr, stream, check
    g = torch.Generator().manual_seed(5)
    for rows, k, n, ai, ao in [(2, 64, 256, 0, 2), (8, 1024, 300, 1, 0), (50, 256, 64, 1, 0), (3, 24, 40, 0, 0)]:
        x, w, b = torch.randn(rows, k, generator=g), torch.randn(n, k, generator=g) / math.sqrt(k), torch.randn(n, generator=g)
        act = {0: lambda v: v, 1: F.silu, 2: F.gelu}
        ref = act[ao](F.linear(act[ai](x), w, b))
        y = torch.empty(rows, n, device=dev())
        ops.small_linear(x.to(dev()), w.to(dev()), b.to(dev()), y, ai, ao)
        close(y, ref, f"small_linear {rows}x{k}x{n}")
    t = torch.tensor([0, 1, 17, 500, 999])
    dim = 64
    half = dim // 2
    fr = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).float()
    out = torch.empty(5, dim, device=dev())
    td, frd = t.to(dev()), fr.to(dev())     # keep the device copies alive across the async launch
    check(lib().lfdm_sinusoidal(ptr(td), ptr(frd), ptr(out), 5, dim, stream()), "sinusoidal")
    close(out, O.sinusoidal_emb(t, dim), "sinusoidal", rtol=1e-4, atol=1e-5)


# ----------------------------------------------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_per", [3 * 5 * 8 * 8, 3 * 40 * 32 * 32])
def test_sampler_update_bit_exact(n_per):
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200._lib import lib, ptr, stream, check
    g = torch.Generator().manual_seed(6)
    b = 3
    buf = O.diffusion_buffers(1000)
    for t in (999, 400, 0):
        x, eps, z = [torch.randn(b, n_per, generator=g) for _ in range(3)]
        ref = O.p_sample_step(buf, x, t, eps, z)
        sigma = (0.5 * buf["posterior_log_variance_clipped"][t]).exp() * (0.0 if t == 0 else 1.0)
        coef = torch.stack([buf["sqrt_recip_alphas_cumprod"][t], buf["sqrt_recipm1_alphas_cumprod"][t],
                            buf["posterior_mean_coef1"][t], buf["posterior_mean_coef2"][t], sigma, torch.tensor(0.),
                            torch.tensor(0.), torch.tensor(0.)])[None].contiguous().to(dev())
        xd, ed, zd = x.to(dev()), eps.to(dev()), z.to(dev())
        ab = torch.empty_like(xd)
        check(lib().lfdm_sampler_x0(ptr(xd), ptr(ed), ptr(coef), None, ptr(ab), n_per, b, stream()), "x0")
        s = torch.empty(b, device=dev())
        r = torch.tensor(0.9, dtype=torch.float32) * (n_per - 1)
        k_lo, w_hi = int(torch.floor(r).item()), float((r - torch.floor(r)).item())
        check(lib().lfdm_sampler_quantile(ptr(ab), ptr(s), n_per, b, k_lo, w_hi, None, stream()), "quantile")
        x0 = buf["sqrt_recip_alphas_cumprod"][t] * x - buf["sqrt_recipm1_alphas_cumprod"][t] * eps
        s_ref = torch.quantile(x0.abs(), 0.9, dim=-1).clamp(min=1.0)
        assert torch.equal(s.cpu(), s_ref), (s.cpu(), s_ref)
        out = torch.empty_like(xd)
        check(lib().lfdm_sampler_update(ptr(xd), ptr(ed), ptr(zd), ptr(s), ptr(coef), None, 0, ptr(out), None, n_per, b,
                                        stream()), "update")
        assert torch.equal(out.cpu(), ref), f"t={t}: max diff {(out.cpu() - ref).abs().max().item()}"


def test_sampler_quantile_ties_and_ddim():
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200._lib import lib, ptr, stream, check
    g = torch.Generator().manual_seed(7)
    n_per, b = 4000, 2
    # heavy ties: quantised values
    a = (torch.rand(b, n_per, generator=g) * 20).round() / 4
    s = torch.empty(b, device=dev())
    for q in (0.9, 0.5, 0.999):
        r = torch.tensor(q, dtype=torch.float32) * (n_per - 1)
        k_lo, w_hi = int(torch.floor(r).item()), float((r - torch.floor(r)).item())
        a_d = a.to(dev())
        check(lib().lfdm_sampler_quantile(ptr(a_d), ptr(s), n_per, b, k_lo, w_hi, None, stream()), "quantile")
        assert torch.equal(s.cpu(), torch.quantile(a, q, dim=-1).clamp(min=1.0))
    buf = O.diffusion_buffers(1000)
    x, eps, z = [torch.randn(b, n_per, generator=g) for _ in range(3)]
    for (time, time_next) in O.ddim_times(1000, 4):
        ref = O.ddim_step(buf, x, time, time_next, eps, z if time_next > 0 else None)
        alpha, alpha_next = buf["alphas_cumprod_prev"][time], buf["alphas_cumprod_prev"][time_next]
        sigma = 1.0 * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = ((1 - alpha_next) - sigma ** 2).sqrt()
        coef = torch.stack([buf["sqrt_recip_alphas_cumprod"][time], buf["sqrt_recipm1_alphas_cumprod"][time],
                            alpha_next.sqrt(), torch.tensor(0.), sigma, c, torch.tensor(1.), torch.tensor(0.)])[None].contiguous().to(dev())
        xd, ed, zd = x.to(dev()), eps.to(dev()), z.to(dev())
        ab = torch.empty_like(xd)
        check(lib().lfdm_sampler_x0(ptr(xd), ptr(ed), ptr(coef), None, ptr(ab), n_per, b, stream()), "x0")
        r = torch.tensor(0.9, dtype=torch.float32) * (n_per - 1)
        check(lib().lfdm_sampler_quantile(ptr(ab), ptr(s), n_per, b, int(torch.floor(r).item()),
                                          float((r - torch.floor(r)).item()), None, stream()), "quantile")
        out = torch.empty_like(xd)
        check(lib().lfdm_sampler_update(ptr(xd), ptr(ed), ptr(zd) if time_next > 0 else None, ptr(s), ptr(coef), None, 0,
                                        ptr(out), None, n_per, b, stream()), "update")
        assert torch.equal(out.cpu(), ref), f"ddim {time}->{time_next}: {(out.cpu() - ref).abs().max().item()}"


# ----------------------------------------------------------------------------------------------------------------
# warp / blend
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hs,c,with_prev", [(8, 16, False), (16, 8, True), (32, 64, True)])
def test_warp_rows(hs, c, with_prev):
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine.lfae_engine import _warp_rows
    from cvpr23_lfdm_b200._lib import SB
    g = torch.Generator().manual_seed(8)
    ns, fps, hf = 2, 3, 8
    n = ns * fps
    src = torch.randn(ns, c, hs, hs, generator=g)
    flow = torch.rand(n, hf, hf, 2, generator=g) * 2.6 - 1.3
    occ = torch.rand(n, 1, hf, hf, generator=g)
    prev = torch.randn(n, c, hs, hs, generator=g) if with_prev else None
    ref = O.apply_optical(prev, src.repeat_interleave(fps, 0), flow, occ)
    out, out_sb = torch.empty(n * hs * hs, c, device=dev()), SB(n * hs * hs, c, dev())
    _warp_rows(rows_of(src).to(dev()), flow.to(dev()), occ.reshape(n, hf, hf).to(dev()),
               rows_of(prev).to(dev()) if with_prev else None, out, out_sb, n, fps, hs, hs, c, hf, hf)
    close(out, rows_of(ref), "warp rows f32", rtol=1e-4, atol=1e-5)
    close(out_sb.float(), rows_of(ref), "warp rows sb")


def test_warp_image():
    from oracle import lfdm_oracle as O
    from cvpr23_lfdm_b200.engine.lfae_engine import _warp_image
    g = torch.Generator().manual_seed(9)
    b, f, hw, hf = 2, 3, 32, 8
    img = torch.rand(b, 3, hw, hw, generator=g)
    flow = torch.rand(b * f, hf, hf, 2, generator=g) * 2.4 - 1.2
    occ = torch.rand(b * f, 1, hf, hf, generator=g)
    prev = torch.rand(b * f, 3, hw, hw, generator=g)
    rep = img.repeat_interleave(f, 0)
    ref_def = O.deform_input(rep, flow).reshape(b, f, 3, hw, hw).permute(0, 2, 1, 3, 4)
    ref_bl = O.apply_optical(prev, rep, flow, occ).reshape(b, f, 3, hw, hw).permute(0, 2, 1, 3, 4)
    out = torch.empty(b, 3, f, hw, hw, device=dev())
    _warp_image(img.to(dev()), flow.to(dev()), None, None, 0, out, b, f, hw, hw, hf, hf)
    close(out, ref_def, "deformed image", rtol=1e-4, atol=1e-5)
    _warp_image(img.to(dev()), flow.to(dev()), occ.reshape(b * f, hf, hf).to(dev()), rows_of(prev).to(dev()), 3, out, b, f,
                hw, hw, hf, hf)
    close(out, ref_bl, "blended image", rtol=1e-4, atol=1e-5)


# ----------------------------------------------------------------------------------------------------------------
# layout / small ops
# ----------------------------------------------------------------------------------------------------------------
def test_layout_kernels():
    from cvpr23_lfdm_b200.engine import ops
    from cvpr23_lfdm_b200._lib import SB, lib, ptr, stream, check
    g = torch.Generator().manual_seed(10)
    b, c, f, h, w = 2, 11, 3, 6, 10
    x = torch.randn(b, c, f, h, w, generator=g)
    rows_ref = x.permute(0, 2, 3, 4, 1).reshape(-1, c)
    out = torch.empty(b * f * h * w, 16, device=dev())
    sb = SB(b * f * h * w, 16, dev())
    ops.to_rows(x.to(dev()), c_pad=16, out_sb=sb, out_f32=out)
    assert torch.equal(out.cpu()[:, :c], rows_ref) and (out.cpu()[:, c:] == 0).all()
    close(sb.float()[:, :c], rows_ref, "to_rows sb", rtol=1e-5, atol=1e-6)
    back = torch.empty(b, c, f, h * w, device=dev())
    ops.from_rows(out, b, c, f, h * w, back)
    assert torch.equal(back.cpu().reshape(x.shape), x)
    # im2col (7x7, pad 3) vs unfold
    x3 = torch.randn(2, 3, 2, 8, 8, generator=g)
    cols = SB(2 * 2 * 64, 192, dev())
    x3d = x3.to(dev())
    check(lib().lfdm_im2col_small(ptr(x3d), 2, 3, 2, 8, 8, 7, 3, 192, ptr(cols.t), cols.plane, stream()), "im2col")
    xf = x3.permute(0, 2, 1, 3, 4).reshape(4, 3, 8, 8)
    un = F.unfold(xf, 7, padding=3).reshape(4, 3, 49, 64).permute(0, 3, 2, 1).reshape(4 * 64, 147)   # k = tap*3 + ch
    got = cols.float().cpu()
    close(got[:, :147], un, "im2col", rtol=1e-5, atol=1e-6)
    assert (got[:, 147:] == 0).all()
    # avgpool
    y = torch.randn(3, 8, 6, 4, generator=g)          # (n, c, h, w)
    pooled = torch.empty(3 * 3 * 2, 8, device=dev())
    yd = rows_of(y).to(dev())
    check(lib().lfdm_avgpool2_rows(ptr(yd), 3, 6, 4, 8, ptr(pooled), None, 0, stream()), "avgpool")
    close(pooled, rows_of(F.avg_pool2d(y, 2)), "avgpool", rtol=1e-5, atol=1e-6)
    # heads
    a, o = torch.randn(2 * 3 * 16, 16, generator=g), torch.randn(2 * 3 * 16, 16, generator=g)
    wa, ba, wo, bo = torch.randn(2, 16, generator=g), torch.randn(2, generator=g), torch.randn(1, 16, generator=g), torch.randn(1, generator=g)
    outh = torch.empty(2, 3, 3, 4, 4, device=dev())
    ad, wad, bad, od, wod, bod = [v.to(dev()) for v in (a, wa, ba, o, wo, bo)]
    check(lib().lfdm_unet_heads(ptr(ad), ptr(wad), ptr(bad), 2, ptr(od), ptr(wod), ptr(bod), 1, 16, 2, 3, 16, ptr(outh),
                                stream()), "heads")
    ref = torch.cat([F.linear(a, wa, ba), F.linear(o, wo, bo)], 1).reshape(2, 3, 16, 3).permute(0, 3, 1, 2).reshape(2, 3, 3, 4, 4)
    close(outh, ref, "heads", rtol=1e-4, atol=1e-5)
    # heads, c = 64 fast path (eight lanes per row), ragged row count (45 rows: last group of four is partial)
    a, o = torch.randn(3 * 15, 64, generator=g), torch.randn(3 * 15, 64, generator=g)
    wa, ba, wo, bo = torch.randn(2, 64, generator=g), torch.randn(2, generator=g), torch.randn(1, 64, generator=g), torch.randn(1, generator=g)
    outh = torch.empty(1, 3, 3, 15, device=dev())
    ad, wad, bad, od, wod, bod = [v.to(dev()) for v in (a, wa, ba, o, wo, bo)]
    check(lib().lfdm_unet_heads(ptr(ad), ptr(wad), ptr(bad), 2, ptr(od), ptr(wod), ptr(bod), 1, 64, 1, 3, 15, ptr(outh),
                                stream()), "heads64")
    ref = torch.cat([F.linear(a, wa, ba), F.linear(o, wo, bo)], 1).reshape(1, 3, 15, 3).permute(0, 3, 1, 2)
    close(outh, ref, "heads64", rtol=1e-4, atol=1e-5)
