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
        self._graphs = {}

    # ---- per-step coefficient tables (float32 torch arithmetic on the registered buffers, as the reference does) ----
    def _ddpm_rows(self, ts, clip=True):
        g = self.gd
        rows = []
        for t in ts:
            sigma = (0.5 * g.posterior_log_variance_clipped[t]).exp() * (0.0 if t == 0 else 1.0)
            rows.append(torch.stack([g.sqrt_recip_alphas_cumprod[t], g.sqrt_recipm1_alphas_cumprod[t],
                                     g.posterior_mean_coef1[t], g.posterior_mean_coef2[t], sigma,
                                     torch.zeros_like(sigma), torch.zeros_like(sigma),
                                     torch.full_like(sigma, 0.0 if clip else 1.0)]))
        return torch.stack(rows).contiguous()

    def _ddim_pairs(self):
        g = self.gd
        times = torch.linspace(0., g.num_timesteps, steps=g.sampling_timesteps + 2)[:-1]
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def _ddim_rows(self, pairs, clip=True):
        g = self.gd
        eta = g.ddim_sampling_eta
        rows = []
        for time, time_next in pairs:
            alpha = g.alphas_cumprod_prev[time]
            alpha_next = g.alphas_cumprod_prev[time_next]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = ((1 - alpha_next) - sigma ** 2).sqrt()
            rows.append(torch.stack([g.sqrt_recip_alphas_cumprod[time], g.sqrt_recipm1_alphas_cumprod[time],
                                     alpha_next.sqrt(), torch.zeros_like(c), sigma, c, torch.ones_like(c),
                                     torch.full_like(c, 0.0 if clip else 1.0)]))
        return torch.stack(rows).contiguous()

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

    def _uniform_t(self, t):
        v = t.tolist()
        if any(a != v[0] for a in v):
            raise NotImplementedError("per-sample different timesteps inside one p_sample call: the reference's loops "
                                      "always pass a uniform t (video_flow_diffusion.py:756,795)")
        return int(v[0])

    # ---- reference API -----------------------------------------------------------------------------------------
    def p_mean_variance(self, x, t, fea, clip_denoised, cond=None, cond_scale=1.):
        g = self.gd
        ti = self._uniform_t(t)
        x = x.contiguous().float()
        eps = self._eps_generic(x, t, fea, cond, cond_scale).contiguous()
        coef = self._ddpm_rows([ti], clip_denoised)
        mean = torch.empty_like(x)
        self._update(x, eps, None, coef, None, False, mean, clip_denoised)
        shp = (x.shape[0],) + (1,) * (x.ndim - 1)
        return mean, g.posterior_variance[t].reshape(shp), g.posterior_log_variance_clipped[t].reshape(shp)

    def p_sample(self, x, t, fea, cond=None, cond_scale=1., clip_denoised=True):
        ti = self._uniform_t(t)
        x = x.contiguous().float()
        eps = self._eps_generic(x, t, fea, cond, cond_scale).contiguous()
        noise = self.gd._randn(x.shape, x.device)
        coef = self._ddpm_rows([ti], clip_denoised)
        out = torch.empty_like(x)
        self._update(x, eps, noise, coef, None, False, out, clip_denoised)
        return out

    # ---- the hot loops -------------------------------------------------------------------------------------------
    def _run_loop(self, fea, shape, cond, cond_scale, coef, times, draw_noise, clip_denoised=True):
        """times: python list of UNet timesteps per step; draw_noise[i]: whether step i consumes a noise draw."""
        g = self.gd
        dev = self.device
        unet = g.denoise_fn
        b = shape[0]
        img = g._randn(shape, dev).contiguous()
        if not hasattr(unet, "engine"):       # foreign denoiser: plain reference data flow, our sampler kernels
            for i, t in enumerate(times):
                tt = torch.full((b,), t, device=dev, dtype=torch.long)
                eps = self._eps_generic(img, tt, fea, cond, cond_scale).contiguous()
                noise = g._randn(shape, dev) if draw_noise[i] else None
                self._update(img, eps, noise, coef[i:i + 1].contiguous(), None, False, img, clip_denoised)
            return img
        eng = unet.engine()
        fea_conv = eng.prepare_fea(fea)
        tvec = torch.tensor(times, device=dev, dtype=torch.long)
        if cond is not None:
            cond = cond.to(dev).float()
        guided = not (cond_scale == 1 or not unet.has_cond)
        null_only = cond_scale == 0 and unet.has_cond
        null_emb = unet.null_cond_emb.detach().to(dev).float().expand(b, -1).contiguous() if unet.has_cond else None
        tabs = eng.build_tables(tvec, null_emb if null_only else cond)
        tabs_null = eng.build_tables(tvec, null_emb) if guided else None
        step_idx = torch.zeros((1,), dtype=torch.int32, device=dev)

        def step(noise):
            ss = eng.ss_from_tables(tabs[0], tabs[1], step_idx, b)
            eps = eng.forward_hoisted(img, fea_conv, ss)
            if guided:
                ssn = eng.ss_from_tables(tabs_null[0], tabs_null[1], step_idx, b)
                eps_null = eng.forward_hoisted(img, fea_conv, ssn)
                eps = eps_null + (eps - eps_null) * cond_scale      # reference :526
            self._update(img, eps, noise, coef, step_idx, True, img, clip_denoised)

        graphable = USE_GRAPH and g.noise_fn is None
        n = len(times)
        i = 0
        calls0 = L.launch_count
        self.last_stats = {"steps": n, "calls_per_step": 0, "other_calls": 0}
        if graphable and n >= 4:
            # leading steps run eagerly (also warms every kernel / allocator pool), then one graph per noise pattern
            noise_buf = torch.empty(shape, device=dev)
            while i < 2:
                c0 = L.launch_count
                step(g._randn(shape, dev) if draw_noise[i] else None)
                self.last_stats["calls_per_step"] = L.launch_count - c0
                i += 1
            graph = None
            while i < n:
                if draw_noise[i]:
                    if graph is None:
                        torch.cuda.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph):
                            noise_buf.normal_()
                            step(noise_buf)
                        # capture does not execute: replay for this step
                    graph.replay()
                else:
                    step(None)
                i += 1
            return img
        while i < n:
            c0 = L.launch_count
            step(g._randn(shape, dev) if draw_noise[i] else None)
            self.last_stats["calls_per_step"] = L.launch_count - c0
            i += 1
        return img

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
