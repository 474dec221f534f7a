#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the BASELINE config-2 geometry (MUG-128, batch 8/GPU): CUDA-event timing of each hot
kernel in isolation with an L2 flush between iterations, printed as JSON lines with algorithmic FLOPs / bytes and the
fraction of the measured peaks (MEASURED_PEAKS.json, burst figures: kernels timed alone).  Also the command ncu wraps:

  ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 6 -o gpurun_out/prof_conv \\
      python tools/profile_kernels.py --only conv3x3_c64_32 --iters 2
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cvpr23_lfdm_b200 import _lib as L  # noqa: E402
from cvpr23_lfdm_b200._lib import SB, lib, ptr, stream, check  # noqa: E402
from cvpr23_lfdm_b200.engine import ops  # noqa: E402
from cvpr23_lfdm_b200.engine.ops import ConvLayer  # noqa: E402
from cvpr23_lfdm_b200.engine.lfae_engine import _warp_rows, _warp_image  # noqa: E402

B, F = 8, 40
dev = torch.device("cuda:0")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def rand_sb(m, c):
    s = SB(m, c, dev)
    s.t.copy_(ops.split_planes(torch.randn(m, c, device=dev)))
    return s


def conv_case(cin, cout, hw, k, gn=False, residual=False, srcs=1):
    nf = B * F
    m = nf * hw * hw
    layer = ConvLayer(torch.randn(cout, cin * srcs, k, k, device=dev) * 0.05, torch.randn(cout, device=dev), pad=k // 2,
                      src_channels=[cin] * srcs)
    a = [rand_sb(m, cin) for _ in range(srcs)]
    out = torch.empty(m, cout, device=dev)
    res = torch.randn(m, cout, device=dev) if residual else None
    stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev) if gn else None

    def run():
        if stats is not None:
            stats.zero_()
        layer(a, nf, hw, hw, out_f32=out, residual=res, gn_stats=stats, gn_groups=8, rows_per_sample=F * hw * hw, stats_zeroed=True)
    flops = 2.0 * m * cout * cin * srcs * k * k
    byt = m * cin * srcs * 4 + m * cout * 4 * (2 if residual else 1)
    return run, flops, byt, lambda: layer.last_engine


def cases():
    c = {}
    c["conv3x3_c64_32"] = lambda: conv_case(64, 64, 32, 3, gn=True)
    c["conv3x3_c64_32_nogn"] = lambda: conv_case(64, 64, 32, 3, gn=False)
    c["conv1x1_c64_32_nogn"] = lambda: conv_case(64, 64, 32, 1, gn=False)
    c["conv3x3_c256_32_nogn"] = lambda: conv_case(256, 256, 32, 3, gn=False)
    c["conv3x3_c64x2_32"] = lambda: conv_case(64, 64, 32, 3, gn=True, srcs=2)
    c["conv3x3_c128_16"] = lambda: conv_case(128, 128, 16, 3, gn=True)
    c["conv3x3_c256_8"] = lambda: conv_case(256, 256, 8, 3, gn=True)
    c["conv3x3_c512_4"] = lambda: conv_case(512, 512, 4, 3, gn=True)
    c["qkv_c64_32"] = lambda: conv_case(64, 768, 32, 1)
    c["qkv_c128_16"] = lambda: conv_case(128, 768, 16, 1)
    c["out_c256_32"] = lambda: conv_case(256, 64, 32, 1, residual=True)

    def attn_t():
        p, heads = 1024, 8
        m = B * F * p
        qkv = torch.randn(m, 768, device=dev)
        o = SB(m, 256, dev)
        ang = torch.outer(torch.arange(F, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32)))
        cs, sn, bias = ang.cos().contiguous(), ang.sin().contiguous(), torch.randn(heads, F, F, device=dev)
        run = lambda: ops.attn_softmax(qkv, o, None, B * p, F, heads, p, F * p, 1, p, cs, sn, bias)
        return run, 4.0 * B * p * heads * F * F * 32, m * 768 * 4 + m * 256 * 4, lambda: "mma.sync split-bf16"
    c["attn_temporal_32"] = attn_t

    def attn_tc():
        # same problem, sequences contiguous in memory ((b, pixel) major, frame minor): what the gather pattern costs
        p, heads = 1024, 8
        m = B * F * p
        qkv = torch.randn(m, 768, device=dev)
        o = SB(m, 256, dev)
        ang = torch.outer(torch.arange(F, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32)))
        cs, sn, bias = ang.cos().contiguous(), ang.sin().contiguous(), torch.randn(heads, F, F, device=dev)
        run = lambda: ops.attn_softmax(qkv, o, None, B * p, F, heads, 1, F, 0, 1, cs, sn, bias)
        return run, 4.0 * B * p * heads * F * F * 32, m * 768 * 4 + m * 256 * 4, lambda: "mma.sync split-bf16, contiguous sequences"
    c["attn_temporal_32_contig"] = attn_tc

    def tblock(fused, hw=32):
        def mk():
            p, heads, ch = hw * hw, 8, 64
            m = B * F * p
            x = torch.randn(m, ch, device=dev)
            gamma = torch.randn(ch, device=dev)
            wqkv, wout = torch.randn(768, ch, 1, 1, device=dev) / 8, torch.randn(ch, 256, 1, 1, device=dev) / 16
            ang = torch.outer(torch.arange(F, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32)))
            cs, sn, bias = ang.cos().contiguous(), ang.sin().contiguous(), torch.randn(heads, F, F, device=dev)
            out, out_sb = torch.empty(m, ch, device=dev), SB(m, ch, dev)
            # algorithmic work of the block: LN + to_qkv + attention cores + to_out; bytes: x read, out written twice (F32 + SB)
            flops = 2.0 * m * ch * 768 + 4.0 * B * p * heads * F * F * 32 + 2.0 * m * 256 * ch
            byt = m * ch * 4 * 3
            if fused:
                wq_img, wo_img = ops.pack_fused_attention(wqkv, wout, heads)
                run = lambda: ops.attn_temporal_fused(x, gamma, wq_img, wo_img, None, cs, sn, bias, out, out_sb, B, F, p, heads, 1e-5)
                return run, flops, byt, lambda: "tcgen05 fused block (split-bf16 x3)"
            lq, lo = ConvLayer(wqkv, None), ConvLayer(wout, None)
            n_sb, qkv, o = SB(m, ch, dev), torch.empty(m, 768, device=dev), SB(m, 256, dev)

            def run():
                ops.layernorm(x, gamma, out_sb=n_sb)
                lq([n_sb], B * F, hw, hw, out_f32=qkv)
                ops.attn_softmax(qkv, o, None, B * p, F, heads, p, F * p, 1, p, cs, sn, bias)
                lo([o], B * F, hw, hw, out_f32=out, out_sb=out_sb, residual=x)
            return run, flops, byt, lambda: "4 kernels: layernorm + tcgen05 qkv + mma.sync core + tcgen05 out"
        return mk
    c["tblock_fused_32"] = tblock(True)
    c["tblock_unfused_32"] = tblock(False)
    c["tblock_fused_16"] = tblock(True, 16)
    c["tblock_unfused_16"] = tblock(False, 16)

    def lblock(fused, hw=32):
        def mk():
            p, ch = hw * hw, 64
            m = B * F * p
            x = torch.randn(m, ch, device=dev)
            gamma = torch.randn(ch, device=dev)
            wqkv, wout = torch.randn(768, ch, 1, 1, device=dev) / 8, torch.randn(ch, 256, 1, 1, device=dev) / 16
            ob = torch.randn(ch, device=dev)
            out = torch.empty(m, ch, device=dev)
            # algorithmic work: LN + to_qkv + (context + apply) + to_out; bytes: x read, out written
            flops = 2.0 * m * ch * 768 + 4.0 * B * F * 8 * 32 * 32 * p + 2.0 * m * 256 * ch
            byt = m * ch * 4 * 2
            if fused:
                packed = ops.pack_fused_linear_attention(wqkv, wout)
                work = [None]

                def run():
                    _, work[0] = ops.attn_linear_fused(x, gamma, packed, ob, out, None, B * F, p, 1e-5, work=work[0])
                return run, flops, byt, lambda: "tcgen05 fused linear block (3 launches, split-bf16 x3)"
            lq, lo = ConvLayer(wqkv, None), ConvLayer(wout, ob)
            n_sb, qkv, o = SB(m, ch, dev), torch.empty(m, 768, device=dev), SB(m, 256, dev)

            def run():
                ops.layernorm(x, gamma, out_sb=n_sb)
                lq([n_sb], B * F, hw, hw, out_f32=qkv)
                ops.attn_linear(qkv, o, None, B * F, p, 8)
                lo([o], B * F, hw, hw, out_f32=out, residual=x)
            return run, flops, byt, lambda: "4 kernels: layernorm + tcgen05 qkv + mma.sync core + tcgen05 out"
        return mk
    c["lblock_fused_32"] = lblock(True)
    c["lblock_unfused_32"] = lblock(False)

    def attn_l():
        p, heads = 1024, 8
        m = B * F * p
        qkv = torch.randn(m, 768, device=dev)
        o = SB(m, 256, dev)
        run = lambda: ops.attn_linear(qkv, o, None, B * F, p, heads)
        return run, 4.0 * B * F * heads * 32 * 32 * p, m * 768 * 4 * (4 / 3) + m * 256 * 4, lambda: "mma.sync split-bf16"
    c["attn_linear_32"] = attn_l

    def gn():
        m, ch = B * F * 1024, 64
        x, res = torch.randn(m, ch, device=dev), torch.randn(m, ch, device=dev)
        stats = torch.empty(B, 8, 2, dtype=torch.float64, device=dev)
        check(lib().lfdm_gn_stats(ptr(x), m, ch, 8, F * 1024, ptr(stats), stream()), "gn_stats")
        g, bt, ss = torch.ones(ch, device=dev), torch.zeros(ch, device=dev), torch.randn(B, 2 * ch, device=dev)
        o, osb = torch.empty_like(x), SB(m, ch, dev)
        run = lambda: ops.gn_apply(x, stats, g, bt, ss, res, o, osb, 8, F * 1024)
        return run, 0.0, m * ch * 4 * 4, lambda: "hbm"
    c["gn_apply_c64_32"] = gn

    def ln():
        m, ch = B * F * 1024, 64
        x, g, osb = torch.randn(m, ch, device=dev), torch.ones(ch, device=dev), SB(m, ch, dev)
        run = lambda: ops.layernorm(x, g, out_sb=osb)
        return run, 0.0, m * ch * 4 * 2, lambda: "hbm"
    c["layernorm_c64_32"] = ln

    def warp(hs, ch, smooth=False):
        def mk():
            n = B * F
            src = torch.randn(B * hs * hs, ch, device=dev)
            if smooth:
                # what the sampler produces: the identity grid plus a displacement of a few latent pixels
                ax = torch.linspace(-1, 1, 32, device=dev)
                ident = torch.stack(torch.meshgrid(ax, ax, indexing="xy"), -1)
                flow = (ident[None] + 0.15 * torch.randn(n, 1, 1, 2, device=dev) + 0.02 * torch.randn(n, 32, 32, 2, device=dev)).contiguous()
            else:
                flow = torch.rand(n, 32, 32, 2, device=dev) * 2 - 1            # worst case: every tap a random row of the source
            occ = torch.rand(n, 32, 32, device=dev)
            prev = torch.randn(n * hs * hs, ch, device=dev)
            out = SB(n * hs * hs, ch, dev)
            run = lambda: _warp_rows(src, flow, occ, prev, None, out, n, F, hs, hs, ch, 32, 32)
            byt = src.numel() * 4 + flow.numel() * 4 + occ.numel() * 4 + prev.numel() * 4 + n * hs * hs * ch * 4
            return run, 0.0, byt, lambda: "hbm"
        return mk
    c["warp_rows_128_c64"] = warp(128, 64)
    c["warp_rows_64_c128"] = warp(64, 128)
    c["warp_rows_32_c256"] = warp(32, 256)
    c["warp_rows_128_c64_smooth"] = warp(128, 64, True)
    c["warp_rows_64_c128_smooth"] = warp(64, 128, True)
    c["warp_rows_32_c256_smooth"] = warp(32, 256, True)

    def wimg():
        img = torch.rand(B, 3, 128, 128, device=dev)
        flow = torch.rand(B * F, 32, 32, 2, device=dev) * 2 - 1
        occ = torch.rand(B * F, 32, 32, device=dev)
        prev = torch.rand(B * F * 128 * 128, 3, device=dev)
        out = torch.empty(B, 3, F, 128, 128, device=dev)
        run = lambda: _warp_image(img, flow, occ, prev, 3, out, B, F, 128, 128, 32, 32)
        return run, 0.0, img.numel() * 4 + flow.numel() * 4 + occ.numel() * 4 + prev.numel() * 4 + out.numel() * 4, lambda: "hbm"
    c["warp_image_128"] = wimg

    def sampler():
        n = 3 * F * 32 * 32
        x, e, z = [torch.randn(B, n, device=dev) for _ in range(3)]
        coef = torch.tensor([[1.5, 1.1, 0.3, 0.7, 0.1, 0, 0, 0]], device=dev)
        ab, s, out = torch.empty_like(x), torch.empty(B, device=dev), torch.empty_like(x)

        def run():
            check(lib().lfdm_sampler_x0(ptr(x), ptr(e), ptr(coef), None, ptr(ab), n, B, stream()), "x0")
            check(lib().lfdm_sampler_quantile(ptr(ab), ptr(s), n, B, 110591, 0.09375, None, stream()), "q")
            check(lib().lfdm_sampler_update(ptr(x), ptr(e), ptr(z), ptr(s), ptr(coef), None, 0, ptr(out), None, n, B, stream()), "u")
        return run, 0.0, B * n * 4 * 4, lambda: "hbm"
    c["sampler_step"] = sampler
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--noflush", action="store_true", help="keep the L2 warm between iterations (diagnostic only)")
    args = ap.parse_args()
    hbm, tf, src = peaks()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for name, mk in cases().items():
        if args.only and args.only not in name:
            continue
        run, flops, byt, eng = mk()
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            if not args.noflush:
                flush.fill_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(1_000_000)      # park the GPU ~0.5 ms: the host enqueues the launch ahead, no launch gap inside the interval
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        rec = {"kernel": name, "engine": eng(), "ms": round(ms, 4), "algorithmic_gbytes": round(byt / 1e9, 4),
               "gb_per_s": round(byt / ms / 1e6, 1), "hbm_frac": round(byt / ms / 1e6 / hbm, 3)}
        if flops:
            rec.update({"algorithmic_gflop": round(flops / 1e9, 2), "tflops": round(flops / ms / 1e9, 2),
                        "tensor_frac": round(flops / ms / 1e9 / tf, 4)})
        rec["peaks"] = src
        print(json.dumps(rec), flush=True)
        del run
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
