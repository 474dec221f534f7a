"""SamplerEngine — the reverse-diffusion loops of the reference GaussianDiffusion
(DM/modules/video_flow_diffusion.py:712-830) on the sm_100a kernels.

One sampling step = [ (scale,shift) row lookup -> UNet (hoisted init conv) -> x0/|x0| -> exact per-sample quantile ->
fused posterior update ] ; all shapes are static, so the step is captured once into a CUDA graph and replayed
(`LFDM_CUDA_GRAPH=0` disables capture).  The step index lives on the device (`step_idx`), every t-dependent scalar
is read from a per-step coefficient table, so the captured graph is identical for all steps.  Noise is drawn with
torch.randn on the run device in the reference's call order (or from `GaussianDiffusion.noise_fn`)."""
import os
import torch
from .. import _lib as L
from .._lib import ptr, stream, check, lib

USE_GRAPH = os.environ.get("LFDM_CUDA_GRAPH", "1") == "1"


class SamplerEngine:
    def __init__(self, gd):
        self.gd = gd
        self.device = gd.betas.device
        if self.device.type != "cuda":
            raise RuntimeError("cvpr23_lfdm_b200.GaussianDiffusion samples only on a CUDA (sm_100a) device; no CPU fallback")
        self._coef_cache = {}
        self._loops = {}      # (engine, shape, guided, steps) -> persistent buffers + captured step graph

    # ---- per-step coefficient tables (float32 torch arithmetic on the registered buffers, as the reference does; built
    # ---- vectorised over all steps and cached: they only depend on the schedule buffers) --------------------------------
    def _buf_key(self):
        g = self.gd
        return (g.betas._version, g.betas.data_ptr(), g.sampling_timesteps, float(g.ddim_sampling_eta))

    def _ddpm_rows(self, ts, clip=True):
        g = self.gd
        key = ("ddpm", tuple(ts), bool(clip)) + self._buf_key()
        hit = self._coef_cache.get(key)
        if hit is not None:
            return hit
        t = torch.tensor(list(ts), device=self.device, dtype=torch.long)
        sigma = (0.5 * g.posterior_log_variance_clipped[t]).exp() * (t != 0).to(torch.float32)
        z = torch.zeros_like(sigma)
        rows = torch.stack([g.sqrt_recip_alphas_cumprod[t], g.sqrt_recipm1_alphas_cumprod[t], g.posterior_mean_coef1[t],
                            g.posterior_mean_coef2[t], sigma, z, z, torch.full_like(sigma, 0.0 if clip else 1.0)], 1).contiguous()
        return self._remember(key, rows)

    def _ddim_pairs(self):
        g = self.gd
        times = torch.linspace(0., g.num_timesteps, steps=g.sampling_timesteps + 2)[:-1]
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def _ddim_rows(self, pairs, clip=True):
        g = self.gd
        key = ("ddim", tuple(pairs), bool(clip)) + self._buf_key()
        hit = self._coef_cache.get(key)
        if hit is not None:
            return hit
        eta = g.ddim_sampling_eta
        t0 = torch.tensor([p[0] for p in pairs], device=self.device, dtype=torch.long)
        t1 = torch.tensor([p[1] for p in pairs], device=self.device, dtype=torch.long)
        alpha, alpha_next = g.alphas_cumprod_prev[t0], g.alphas_cumprod_prev[t1]
        sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = ((1 - alpha_next) - sigma ** 2).sqrt()
        rows = torch.stack([g.sqrt_recip_alphas_cumprod[t0], g.sqrt_recipm1_alphas_cumprod[t0], alpha_next.sqrt(),
                            torch.zeros_like(c), sigma, c, torch.ones_like(c), torch.full_like(c, 0.0 if clip else 1.0)],
                           1).contiguous()
        return self._remember(key, rows)

    def _remember(self, key, rows):
        if len(self._coef_cache) >= 8:
            self._coef_cache.clear()
        self._coef_cache[key] = rows
        return rows

    def _rank(self, n):
        r = torch.tensor(self.gd.dynamic_thres_percentile, dtype=torch.float32) * (n - 1)   # at::quantile rank (fp32)
        lo = torch.floor(r)
        return int(lo.item()), float((r - lo).item())

    # ---- kernels ----------------------------------------------------------------------------------------------
    def _update(self, x, eps, noise, coef, step_idx, advance, x_out, clip_denoised=True, x0_out=None):
        b = x.shape[0]
        n = x[0].numel()
        s = None
        if clip_denoised and self.gd.use_dynamic_thres:
            absx0 = torch.empty_like(x)
            check(lib().lfdm_sampler_x0(ptr(x), ptr(eps), ptr(coef), ptr(step_idx), ptr(absx0), n, b, stream()), "lfdm_sampler_x0")
            s = torch.empty((b,), device=x.device)
            k_lo, w_hi = self._rank(n)
            check(lib().lfdm_sampler_quantile(ptr(absx0), ptr(s), n, b, k_lo, w_hi, None, stream()), "lfdm_sampler_quantile")
        check(lib().lfdm_sampler_update(ptr(x), ptr(eps), ptr(noise), ptr(s), ptr(coef), ptr(step_idx), int(advance),
                                        ptr(x_out), ptr(x0_out), n, b, stream()), "lfdm_sampler_update")
        return s

    def _eps_generic(self, x, t, fea, cond, cond_scale):
        """reference data flow for a foreign denoise_fn: cat([x, fea.repeat]) -> forward_with_cond_scale"""
        fea5 = fea.unsqueeze(2).repeat(1, 1, x.size(2), 1, 1)
        return self.gd.denoise_fn.forward_with_cond_scale(torch.cat([x, fea5], dim=1), t, cond=cond, cond_scale=cond_scale)

    @staticmethod
    def _t_groups(t):
        """{timestep: [sample indices]} - the reference indexes its schedule buffers with a per-sample t
        (video_flow_diffusion.py:714-735 `extract(a, t, x_shape)`); its own loops pass a uniform t, which is the one-group case."""
        groups = {}
        for i, a in enumerate(t.tolist()):
            groups.setdefault(int(a), []).append(i)
        return groups

    def _update_by_t(self, x, eps, noise, t, out, clip_denoised):
        """one `_update` per distinct timestep (dynamic thresholding is per sample, so sub-batches are exact)"""
        groups = self._t_groups(t)
        if len(groups) == 1:
            self._update(x, eps, noise, self._ddpm_rows(list(groups), clip_denoised), None, False, out, clip_denoised)
            return
        for ti, idx in groups.items():
            ix = torch.tensor(idx, device=x.device)
            oi = torch.empty((len(idx),) + tuple(x.shape[1:]), device=x.device)
            self._update(x[ix].contiguous(), eps[ix].contiguous(), None if noise is None else noise[ix].contiguous(),
                         self._ddpm_rows([ti], clip_denoised), None, False, oi, clip_denoised)
            out[ix] = oi

    # ---- reference API -----------------------------------------------------------------------------------------
    def p_mean_variance(self, x, t, fea, clip_denoised, cond=None, cond_scale=1.):
        g = self.gd
        x = x.contiguous().float()
        eps = self._eps_generic(x, t, fea, cond, cond_scale).contiguous()
        mean = torch.empty_like(x)
        self._update_by_t(x, eps, None, t, mean, clip_denoised)
        shp = (x.shape[0],) + (1,) * (x.ndim - 1)
        return mean, g.posterior_variance[t].reshape(shp), g.posterior_log_variance_clipped[t].reshape(shp)

    def p_sample(self, x, t, fea, cond=None, cond_scale=1., clip_denoised=True):
        x = x.contiguous().float()
        eps = self._eps_generic(x, t, fea, cond, cond_scale).contiguous()
        noise = self.gd._randn(x.shape, x.device)
        out = torch.empty_like(x)
        self._update_by_t(x, eps, noise, t, out, clip_denoised)
        return out

    # ---- the hot loops -------------------------------------------------------------------------------------------
    def _run_loop(self, fea, shape, cond, cond_scale, coef, times, draw_noise, clip_denoised=True):
        """times: python list of UNet timesteps per step; draw_noise[i]: whether step i consumes a noise draw.

        Everything a step reads lives in persistent device buffers (image, noise, hoisted init-conv term, (scale, shift)
        tables, coefficient table, step counter) kept per (engine, shape, guided, schedule): a repeated `sample()` call
        refills them and replays the step graph captured by the first call instead of rebuilding tables and re-capturing."""
        g = self.gd
        dev = self.device
        unet = g.denoise_fn
        b = shape[0]
        if not hasattr(unet, "engine"):       # foreign denoiser: plain reference data flow, our sampler kernels
            img = g._randn(shape, dev).contiguous()
            for i, t in enumerate(times):
                tt = torch.full((b,), t, device=dev, dtype=torch.long)
                eps = self._eps_generic(img, tt, fea, cond, cond_scale).contiguous()
                noise = g._randn(shape, dev) if draw_noise[i] else None
                self._update(img, eps, noise, coef[i:i + 1].contiguous(), None, False, img, clip_denoised)
            return img
        eng = unet.engine()
        n = len(times)
        guided = bool(unet.has_cond) and cond_scale not in (0, 1)      # reference :515-526: 0 -> null only, 1 -> cond only
        null_only = bool(unet.has_cond) and cond_scale == 0
        key = (id(eng), tuple(shape), guided, float(cond_scale) if guided else 1.0, tuple(times), tuple(draw_noise),
               bool(clip_denoised), coef.data_ptr())
        st = self._loops.get(key)
        if st is None or st["eng"] is not eng or st["coef"] is not coef:      # `coef` is pinned by the entry: no address reuse
            if len(self._loops) >= 2:           # each entry pins one step's activation pool: keep at most two
                self._loops.clear()
            st = {"eng": eng, "coef": coef, "graph": None, "img": torch.empty(shape, device=dev), "noise": torch.empty(shape, device=dev),
                  "step_idx": torch.zeros((1,), dtype=torch.int32, device=dev), "fea_conv": None, "cond_tab": None,
                  "x2": torch.empty((2 * b,) + tuple(shape[1:]), device=dev) if guided else None}
            self._loops[key] = st
        img, noise_buf, step_idx = st["img"], st["noise"], st["step_idx"]
        img.copy_(g._randn(shape, dev))
        step_idx.zero_()
        # ---- per-call inputs into the persistent buffers
        fea_conv = eng.prepare_fea(fea)
        if cond is not None:
            cond = cond.to(dev).float()
        null_emb = unet.null_cond_emb.detach().to(dev).float().expand(b, -1).contiguous() if unet.has_cond else None
        tvec = torch.tensor(times, device=dev, dtype=torch.long)
        time_tab, cond_tab = eng.build_tables(tvec, null_emb if null_only else cond)
        if guided:                               # one 2B batch: [conditional half ; null-condition half]
            _, null_tab = eng.build_tables(tvec, null_emb)
            cond_tab = torch.cat([cond_tab.expand(b, -1), null_tab.expand(b, -1)], 0)
            fea_conv = torch.cat([fea_conv, fea_conv], 0)
        for name, val in (("fea_conv", fea_conv), ("cond_tab", cond_tab.contiguous())):
            if st[name] is None or st[name].shape != val.shape:
                st[name] = val.clone()
                st["graph"] = None
            else:
                st[name].copy_(val)
        if st.get("time_tab") is not time_tab:  # cached inside the engine: same tensor for the same schedule
            st["time_tab"] = time_tab
            st["graph"] = None
        fea_conv, cond_tab = st["fea_conv"], st["cond_tab"]
        bb = 2 * b if guided else b

        def step(noise):
            ss = eng.ss_from_tables(time_tab, cond_tab, step_idx, bb)
            if guided:
                x2 = st["x2"]
                x2[:b].copy_(img)
                x2[b:].copy_(img)
                eps = eng.forward_hoisted(x2, fea_conv, ss, cfg_scale=float(cond_scale))
            else:
                eps = eng.forward_hoisted(img, fea_conv, ss)
            self._update(img, eps, noise, coef, step_idx, True, img, clip_denoised)

        graphable = USE_GRAPH and g.noise_fn is None
        i = 0
        self.last_stats = {"steps": n, "calls_per_step": st.get("calls_per_step", 0), "other_calls": 0,
                           "graph_reused": st["graph"] is not None}
        if graphable and n >= 4:
            if st["graph"] is None:
                # leading steps run eagerly (also warms every kernel / allocator pool), then the step is captured once
                while i < 2:
                    c0 = L.launch_count
                    step(g._randn(shape, dev) if draw_noise[i] else None)
                    st["calls_per_step"] = self.last_stats["calls_per_step"] = L.launch_count - c0
                    i += 1
            while i < n:
                if draw_noise[i]:
                    if st["graph"] is None:
                        torch.cuda.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph):
                            noise_buf.normal_()
                            step(noise_buf)
                        st["graph"] = graph         # capture does not execute: replay for this step
                    st["graph"].replay()
                else:
                    step(None)
                i += 1
            return img.clone()
        while i < n:
            c0 = L.launch_count
            step(g._randn(shape, dev) if draw_noise[i] else None)
            self.last_stats["calls_per_step"] = L.launch_count - c0
            i += 1
        return img.clone()

    def p_sample_loop(self, fea, shape, cond=None, cond_scale=1.):
        g = self.gd
        times = list(reversed(range(g.num_timesteps)))
        coef = self._ddpm_rows(times)
        return self._run_loop(fea, shape, cond, cond_scale, coef, times, [True] * len(times))

    def ddim_sample(self, fea, shape, cond=None, cond_scale=1., clip_denoised=True):
        pairs = self._ddim_pairs()
        coef = self._ddim_rows(pairs, clip_denoised)
        times = [p[0] for p in pairs]
        draw = [p[1] > 0 for p in pairs]
        return self._run_loop(fea, shape, cond, cond_scale, coef, times, draw, clip_denoised)
