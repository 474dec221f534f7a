"""Engines of the full-LFAE ("real video") branch: PixelwiseFlowPredictor (inside Generator.forward), RegionPredictor and
BGMotionPredictor (kernels K14/K15).

All three are small 32x32-and-below networks (0.5-2.5 GFLOP per frame); their hourglass convolutions run on the
general-shape CUDA-core conv engine (channel counts 3/6/35/44/108 and 2x2/1x1 maps do not tile the tcgen05 path),
BatchNorm is folded, the per-pixel motion algebra and the region soft-argmax + 2x2 SVD are single kernels
(csrc/motion.cu) — in particular the reference's device->host->device trip for torch.svd (region_predictor.py:21)
is gone."""
import torch
from .. import _lib as L
from .._lib import ptr, stream, check, lib
from . import ops
from .ops import ConvLayer, f32
from .lfae_engine import _fold_post_bn


class _Hourglass:
    """Encoder/Decoder of LFAE/modules/util.py:153-214 on F32 row matrices."""

    def __init__(self, hg, name):
        self.downs = [ConvLayer(*_fold_post_bn(b.conv, b.norm), pad=1, engine="simt", name=f"{name}.down{i}")
                      for i, b in enumerate(hg.encoder.down_blocks)]
        self.ups = []
        for i, b in enumerate(hg.decoder.up_blocks):
            w, bias = _fold_post_bn(b.conv, b.norm)
            self.ups.append((w, bias, f"{name}.up{i}"))
        self._up_layers = {}
        self.out_filters = hg.out_filters

    def _up(self, i, split):
        key = (i, tuple(split))
        if key not in self._up_layers:
            w, bias, name = self.ups[i]
            self._up_layers[key] = ConvLayer(w, bias, mode=L.CONV_UPNEAREST, pad=1, src_channels=split, engine="simt", name=name)
        return self._up_layers[key]

    def __call__(self, x_rows, c_in, n, h, w):
        """returns the list of virtual-concat sources [(rows, channels)] of the decoder output at (h, w)"""
        dev = x_rows.device
        outs = [(x_rows, c_in, h, w)]
        cur, ch, cw = x_rows, h, w
        for layer in self.downs:
            y = f32(n * ch * cw, layer.cout, dev)
            layer([cur], n, ch, cw, out_f32=y, f32_act=L.ACT_RELU)
            ch, cw = ch // 2, cw // 2
            p = f32(n * ch * cw, layer.cout, dev)
            check(lib().lfdm_avgpool2_rows(ptr(y), n, ch * 2, cw * 2, layer.cout, ptr(p), None, 0, stream()), "lfdm_avgpool2_rows")
            outs.append((p, layer.cout, ch, cw))
            cur = p
        srcs = [outs.pop()]               # deepest feature
        for i in range(len(self.ups)):
            split = [s[1] for s in srcs]
            layer = self._up(i, split)
            _, _, sh, sw = srcs[0]
            y = f32(n * 4 * sh * sw, layer.cout, dev)
            layer([s[0] for s in srcs], n, sh, sw, out_f32=y, f32_act=L.ACT_RELU)
            skip = outs.pop()
            srcs = [(y, layer.cout, 2 * sh, 2 * sw), skip]      # torch.cat([out, skip], dim=1)
        return srcs


def _down_image(mod, x):
    """AntiAliasInterpolation2d: (N,C,H,W) -> (N,C,H*s,W*s)"""
    n, c, h, w = x.shape
    s = mod.int_inv_scale
    ks = mod.weight.shape[-1]
    out = torch.empty((n, c, (h + s - 1) // s, (w + s - 1) // s), device=x.device)
    kern = mod.weight[0, 0].detach().float().contiguous()
    check(lib().lfdm_antialias_down(ptr(x), ptr(kern), ptr(out), n, c, h, w, ks, mod.ka, s, stream()), "lfdm_antialias_down")
    return out


class RegionEngine:
    def __init__(self, mod):
        self.mod = mod
        self.hg = _Hourglass(mod.predictor, "region.hg")
        self._heads = {}
        self.k = mod.regions.out_channels

    def forward(self, x):
        mod = self.mod
        if mod.jacobian is not None or not mod.pca_based:
            raise NotImplementedError("only the pca_based RegionPredictor of the released configs is implemented")
        x = x.float().contiguous()
        if mod.scale_factor != 1:
            x = _down_image(mod.down, x)
        n, c, h, w = x.shape
        dev = x.device
        rows = f32(n * h * w, c, dev)
        ops.to_rows(x.reshape(n, c, 1, h, w), out_f32=rows)
        srcs = self.hg(rows, c, n, h, w)
        split = tuple(s[1] for s in srcs)
        if split not in self._heads:
            self._heads[split] = ConvLayer(mod.regions.weight.detach().float(), mod.regions.bias, pad=mod.regions.padding[0],
                                           src_channels=list(split), engine="simt", name="region.regions")
        head = self._heads[split]
        ho, wo = head.out_hw(h, w)
        logits = f32(n * ho * wo, self.k, dev)
        head([s[0] for s in srcs], n, h, w, out_f32=logits)
        heat = torch.empty((n, self.k, ho, wo), device=dev)
        shift = torch.empty((n, self.k, 2), device=dev)
        covar, affine = torch.empty((n, self.k, 2, 2), device=dev), torch.empty((n, self.k, 2, 2), device=dev)
        u, d = torch.empty((n * self.k, 2, 2), device=dev), torch.empty((n * self.k, 2, 2), device=dev)
        check(lib().lfdm_region_moments(ptr(logits), n, self.k, ho, wo, float(mod.temperature), ptr(heat), ptr(shift),
                                        ptr(covar), ptr(affine), ptr(u), ptr(d), stream()), "lfdm_region_moments")
        return {"shift": shift, "covar": covar, "heatmap": heat, "affine": affine, "u": u, "d": d}


class BGEngine:
    def __init__(self, mod):
        self.mod = mod
        if mod.bg_type != "zero":
            self.downs = [ConvLayer(*_fold_post_bn(b.conv, b.norm), pad=1, engine="simt", name=f"bg.down{i}")
                          for i, b in enumerate(mod.encoder.down_blocks)]
            self.fc_w = mod.fc.weight.detach().float().contiguous()
            self.fc_b = mod.fc.bias.detach().float().contiguous()

    def forward(self, src, drv):
        mod = self.mod
        bs = src.shape[0]
        dev = src.device
        out = torch.eye(3, device=dev).unsqueeze(0).repeat(bs, 1, 1)
        if mod.bg_type == "zero":
            return out
        x = torch.cat([src.float(), drv.float()], dim=1).contiguous()
        n, c, h, w = x.shape
        cur = f32(n * h * w, c, dev)
        ops.to_rows(x.reshape(n, c, 1, h, w), out_f32=cur)
        for layer in self.downs:
            y = f32(n * h * w, layer.cout, dev)
            layer([cur], n, h, w, out_f32=y, f32_act=L.ACT_RELU)
            h, w = h // 2, w // 2
            p = f32(n * h * w, layer.cout, dev)
            check(lib().lfdm_avgpool2_rows(ptr(y), n, h * 2, w * 2, layer.cout, ptr(p), None, 0, stream()), "lfdm_avgpool2_rows")
            cur = p
        cch = cur.shape[1]
        mean = torch.empty((n, cch), device=dev)
        check(lib().lfdm_rows_mean(ptr(cur), n, h * w, cch, ptr(mean), stream()), "lfdm_rows_mean")
        pred = torch.empty((n, self.fc_w.shape[0]), device=dev)
        ops.small_linear(mean, self.fc_w, self.fc_b, pred)
        if mod.bg_type == "shift":
            out[:, :2, 2] = pred
        elif mod.bg_type == "affine":
            out[:, :2, :] = pred.view(bs, 2, 3)
        else:
            out[:, :2, :] = pred[:, :6].view(bs, 2, 3)
            out[:, 2, :2] = pred[:, 6:].view(bs, 2)
        return out


class PixelwiseFlowEngine:
    def __init__(self, mod):
        self.mod = mod
        if not (mod.use_covar_heatmap and mod.use_deformed_source):
            raise NotImplementedError("only use_covar_heatmap=True / use_deformed_source=True (released configs)")
        self.hg = _Hourglass(mod.hourglass, "pfp.hg")
        self._heads = {}
        self.k = mod.num_regions

    def forward(self, source_image, drv, src, bg_params=None):
        mod = self.mod
        x = source_image.float().contiguous()
        if mod.scale_factor != 1:
            x = _down_image(mod.down, x)
        n, nch, h, w = x.shape
        dev = x.device
        k = self.k
        cin = (k + 1) * (nch + 1)
        hg_in = f32(n * h * w, cin, dev)
        sparse = torch.empty((n, k + 1, h, w, 2), device=dev)
        c = lambda t: t.float().contiguous()
        has_aff = "affine" in drv
        check(lib().lfdm_motion_prep(ptr(x), ptr(c(drv["shift"])), ptr(c(drv["covar"])), ptr(c(drv["affine"])) if has_aff else None,
                                     ptr(c(src["shift"])), ptr(c(src["covar"])), ptr(c(src["affine"])) if has_aff else None,
                                     ptr(c(bg_params)) if bg_params is not None else None, n, k, nch, h, w,
                                     int(bool(mod.revert_axis_swap)), ptr(hg_in), ptr(sparse), stream()), "lfdm_motion_prep")
        srcs = self.hg(hg_in, cin, n, h, w)
        split = tuple(s[1] for s in srcs)
        has_occ = mod.occlusion is not None
        if split not in self._heads:
            wgt = mod.mask.weight.detach().float()
            bias = mod.mask.bias.detach().float()
            if has_occ:                       # mask (K+1) and occlusion (1) heads share their input: one conv, K+2 outputs
                wgt = torch.cat([wgt, mod.occlusion.weight.detach().float()], 0)
                bias = torch.cat([bias, mod.occlusion.bias.detach().float()], 0)
            self._heads[split] = ConvLayer(wgt, bias, pad=3, src_channels=list(split), engine="simt", name="pfp.heads")
        head = self._heads[split]
        logits = f32(n * h * w, head.cout, dev)
        head([s[0] for s in srcs], n, h, w, out_f32=logits)
        flow = torch.empty((n, h, w, 2), device=dev)
        occ = torch.empty((n, 1, h, w), device=dev) if has_occ else None
        check(lib().lfdm_motion_finish(ptr(logits), head.cout, ptr(sparse), n, k, h * w, int(has_occ), ptr(flow), ptr(occ),
                                       stream()), "lfdm_motion_finish")
        out = {"optical_flow": flow}
        if has_occ:
            out["occlusion_map"] = occ
        return out
