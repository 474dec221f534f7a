"""TEST INFRASTRUCTURE: generates tests/golden/*.pt by running the UNMODIFIED reference (through oracle/ref_shim.py)
on this CPU container.  Run:  python oracle/make_golden.py     (needs /root/reference; deterministic: fixed seeds)

Fixtures
  tiny_unet.pt      tiny Unet3D (state_dict + inputs + outputs for cond_scale 1 and 2)
  tiny_sampler.pt   GaussianDiffusion over the tiny UNet: teacher-forced p_sample steps + a 4-step DDIM chain
  tiny_lfae.pt      tiny Generator/RegionPredictor/BGMotionPredictor with randomised BN statistics
  full_fingerprint.pt  strided slices of the FULL-SIZE mug128 model (seed 1234, random init) outputs:
                       one UNet eval, one decoded frame, 3-step DDIM sample_one_video (recipe of SURVEY.md §8c)
"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TINY_UNET = dict(dim=16, cond_dim=24, dim_mults=(1, 2), channels=3 + 8, attn_heads=2, attn_dim_head=32)
TINY_GEN = dict(num_channels=3, num_regions=4, block_expansion=16, max_features=64, num_down_blocks=2,
                num_bottleneck_blocks=2, skips=True, revert_axis_swap=True,
                pixelwise_flow_predictor_params=dict(block_expansion=16, max_features=64, num_blocks=2,
                                                     scale_factor=0.25, use_deformed_source=True,
                                                     use_covar_heatmap=True, estimate_occlusion_map=True))
TINY_RP = dict(num_regions=4, num_channels=3, estimate_affine=True, temperature=0.1, block_expansion=16,
               max_features=64, scale_factor=0.25, num_blocks=2, pca_based=True, fast_svd=False)
TINY_BG = dict(num_channels=3, block_expansion=16, max_features=64, num_blocks=2, bg_type='affine')


def randomize_bn(mod, gen):
    for m in mod.modules():
        if hasattr(m, "running_mean") and m.running_mean is not None:
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) * 0.8 + 0.6)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=gen) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)


def main():
    ns = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---------------- tiny UNet
    torch.manual_seed(11)
    unet = ns.Unet3D(**TINY_UNET).eval()
    # give the zero-init-free layers some spread so every path matters
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 11, 5, 8, 8, generator=g)
    t = torch.tensor([731, 12])
    cond = torch.randn(2, 24, generator=g)
    y1 = unet.forward_with_cond_scale(x, t, cond=cond, cond_scale=1.0)
    y2 = unet.forward_with_cond_scale(x, t, cond=cond, cond_scale=2.0)
    torch.save(dict(cfg=TINY_UNET, sd=unet.state_dict(), x=x, t=t, cond=cond, y_scale1=y1, y_scale2=y2),
               os.path.join(OUT, "tiny_unet.pt"))

    # ---------------- tiny sampler
    gd = ns.GaussianDiffusion(unet, image_size=8, num_frames=5, sampling_timesteps=1000, timesteps=1000, loss_type='l2',
                              use_dynamic_thres=True, null_cond_prob=0.1).eval()
    fea = torch.randn(2, 8, 8, 8, generator=g)
    steps = []
    for tt in (999, 500, 1, 0):
        xin = torch.randn(2, 3, 5, 8, 8, generator=g)
        torch.manual_seed(100 + tt)
        out = gd.p_sample(xin, torch.full((2,), tt, dtype=torch.long), fea, cond=cond, cond_scale=1.0)
        mean, var, logvar = gd.p_mean_variance(xin, torch.full((2,), tt, dtype=torch.long), fea, True, cond=cond, cond_scale=1.0)
        steps.append(dict(t=tt, x=xin, seed=100 + tt, out=out, mean=mean))
    gd_ddim = ns.GaussianDiffusion(unet, image_size=8, num_frames=5, sampling_timesteps=4, timesteps=1000, loss_type='l2',
                                   use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0).eval()
    torch.manual_seed(77)
    ddim = gd_ddim.sample(fea, cond=cond, cond_scale=1.0)
    gd_short = ns.GaussianDiffusion(unet, image_size=8, num_frames=5, sampling_timesteps=6, timesteps=6, loss_type='l2',
                                    use_dynamic_thres=True, null_cond_prob=0.1).eval()
    torch.manual_seed(78)
    ddpm6 = gd_short.sample(fea, cond=cond, cond_scale=1.0)
    torch.save(dict(fea=fea, cond=cond, steps=steps, ddim4=ddim, ddim_seed=77, ddpm6=ddpm6, ddpm6_seed=78),
               os.path.join(OUT, "tiny_sampler.pt"))

    # ---------------- tiny LFAE
    torch.manual_seed(21)
    gen = ns.Generator(**TINY_GEN).eval()
    rp = ns.RegionPredictor(**TINY_RP).eval()
    bg = ns.BGMotionPredictor(**TINY_BG).eval()
    bg.fc.weight.data.normal_(0, 0.01)      # the reference zero-inits this layer; make the encoder matter
    gb = torch.Generator().manual_seed(9)
    for m in (gen, rp, bg):
        randomize_bn(m, gb)
    img = torch.rand(2, 3, 32, 32, generator=gb)
    drv = torch.rand(2, 3, 32, 32, generator=gb)
    flow = torch.rand(2, 8, 8, 2, generator=gb) * 2.4 - 1.2      # some samples fall outside [-1, 1]
    occ = torch.rand(2, 1, 8, 8, generator=gb)
    fwf = gen.forward_with_flow(img, flow, occ)
    fea_t = gen.compute_fea(img)
    src_rp, drv_rp = rp(img), rp(drv)
    bgp = bg(img, drv)
    full = gen(img, source_region_params=src_rp, driving_region_params=drv_rp, bg_params=bgp)
    keep = lambda d: {k: v for k, v in d.items() if k in ("shift", "covar", "affine", "heatmap")}
    torch.save(dict(gen_cfg=TINY_GEN, rp_cfg=TINY_RP, bg_cfg=TINY_BG, gen_sd=gen.state_dict(), rp_sd=rp.state_dict(),
                    bg_sd=bg.state_dict(), img=img, drv=drv, flow=flow, occ=occ, fwf=fwf, fea=fea_t, src_rp=keep(src_rp),
                    drv_rp=keep(drv_rp), bg=bgp, full=full), os.path.join(OUT, "tiny_lfae.pt"))

    # ---------------- full-size fingerprints (weights are NOT stored: rebuilt from the seed on both sides)
    torch.manual_seed(1234)
    m = ns.FlowDiffusion(is_train=False, sampling_timesteps=3, img_size=32, num_frames=40,
                         config_pth=os.path.join(ROOT, "config", "mug128.yaml") if os.path.exists(os.path.join(ROOT, "config", "mug128.yaml")) else '/root/reference/config/mug128.yaml',
                         pretrained_pth="").eval()
    gf = torch.Generator().manual_seed(4321)
    x = torch.randn(1, 3, 40, 32, 32, generator=gf)
    fea = torch.randn(1, 256, 32, 32, generator=gf).abs()
    cond = torch.randn(1, 768, generator=gf)
    t = torch.tensor([640])
    y = m.unet.forward_with_cond_scale(torch.cat([x, fea.unsqueeze(2).repeat(1, 1, 40, 1, 1)], 1), t, cond=cond, cond_scale=1.0)
    img = torch.rand(1, 3, 128, 128, generator=gf)
    flow = torch.rand(1, 32, 32, 2, generator=gf) * 2.2 - 1.1
    occ = torch.rand(1, 1, 32, 32, generator=gf)
    dec = m.generator.forward_with_flow(img, flow, occ)
    feat = m.generator.compute_fea(img)
    torch.manual_seed(1234)
    img2 = torch.rand(1, 3, 128, 128)
    cond2 = torch.randn(1, 768)
    m.set_sample_input(img2, cond2)
    torch.manual_seed(99)
    m.sample_one_video(1.0)
    fp = dict(
        unet_in=dict(seed=4321, t=640), unet_out_slice=y[:, :, ::4, ::4, ::4].clone(), unet_out_mean=y.mean(), unet_out_std=y.std(),
        dec_pred_slice=dec["prediction"][:, :, ::8, ::8].clone(), dec_def_slice=dec["deformed"][:, :, ::8, ::8].clone(),
        fea_slice=feat[:, ::16, ::4, ::4].clone(),
        sample=dict(noise_seed=99, out_mean=m.sample_out_vid.mean(), warped_mean=m.sample_warped_vid.mean(),
                    grid_mean=m.sample_vid_grid.mean(), conf_mean=m.sample_vid_conf.mean(),
                    grid_slice=m.sample_vid_grid[:, :, ::8, ::4, ::4].clone(), conf_slice=m.sample_vid_conf[:, :, ::8, ::4, ::4].clone(),
                    out_slice=m.sample_out_vid[:, :, ::8, ::16, ::16].clone()))
    torch.save(fp, os.path.join(OUT, "full_fingerprint.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
