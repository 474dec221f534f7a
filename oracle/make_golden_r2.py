"""TEST INFRASTRUCTURE: round-2 golden vectors from the UNMODIFIED reference (through oracle/ref_shim.py), CPU only.
Run:  python oracle/make_golden_r2.py      (needs /root/reference; deterministic: fixed seeds)

Fixtures (tests/golden/)
  r2_guided.pt        tiny UNet, classifier-free guidance inside the sampler loops: teacher-forced guided p_sample
                      steps (cond_scale 2.0), 4-step DDIM and 6-step DDPM chains with cond_scale 2.0, a cond_scale 0 chain
  r2_forward_full.pt  FULL-SIZE mug128 FlowDiffusion.forward() (real-video branch, video_flow_diffusion_model.py:116-143)
                      on 3 driving frames: RegionPredictor / BGMotionPredictor / Generator.forward at 128x128, strided slices
  r2_mhad_region.pt   FULL-SIZE RegionPredictor with the mhad128 / natops128 option `pad: 0` (region_predictor.py:33-35)
  r2_config1.pt       BASELINE config 1 run IN FULL on the reference: MUG-128, B=1, 40 frames, 50 DDIM steps, seed 1234
                      (written only with --config1: ~2-3 min of CPU)
Weights of the full-size models are NOT stored: both sides rebuild them from the seed (+ `perturb_lfae`).
"""
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402
from oracle.make_golden import TINY_UNET  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def perturb_lfae(model, seed=4242):
    """Deterministic perturbation applied identically on the reference and on the B200 side: random-init LFAE modules are
    degenerate for parity purposes (BatchNorm statistics at 0/1, `bg_predictor.fc` zero-initialised => identity bg motion)."""
    g = torch.Generator().manual_seed(seed)
    for mod in (model.generator, model.region_predictor, model.bg_predictor):
        for m in mod.modules():
            if hasattr(m, "running_mean") and m.running_mean is not None:
                dev = m.running_mean.device
                m.running_mean.copy_((torch.randn(m.running_mean.shape, generator=g) * 0.1).to(dev))
                m.running_var.copy_((torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75).to(dev))
    fc = model.bg_predictor.fc
    fc.weight.data.copy_((torch.randn(fc.weight.shape, generator=g) * 0.01).to(fc.weight.device))


def guided(ns):
    torch.manual_seed(11)
    unet = ns.Unet3D(**TINY_UNET).eval()
    g = torch.Generator().manual_seed(6)
    cond = torch.randn(2, 24, generator=g)
    fea = torch.randn(2, 8, 8, 8, generator=g)
    gd = ns.GaussianDiffusion(unet, image_size=8, num_frames=5, sampling_timesteps=1000, timesteps=1000, loss_type='l2',
                              use_dynamic_thres=True, null_cond_prob=0.1).eval()
    steps = []
    for tt in (999, 500, 1, 0):
        xin = torch.randn(2, 3, 5, 8, 8, generator=g)
        torch.manual_seed(200 + tt)
        out = gd.p_sample(xin, torch.full((2,), tt, dtype=torch.long), fea, cond=cond, cond_scale=2.0)
        steps.append(dict(t=tt, x=xin, seed=200 + tt, out=out))

    def chain(sampling, timesteps, seed, cs):
        d = ns.GaussianDiffusion(unet, image_size=8, num_frames=5, sampling_timesteps=sampling, timesteps=timesteps,
                                 loss_type='l2', use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0).eval()
        torch.manual_seed(seed)
        return d.sample(fea, cond=cond, cond_scale=cs)
    torch.save(dict(sd=unet.state_dict(), cfg=TINY_UNET, fea=fea, cond=cond, steps=steps,
                    ddim4_cs2=chain(4, 1000, 81, 2.0), ddim4_seed=81, ddpm6_cs2=chain(6, 6, 82, 2.0), ddpm6_seed=82,
                    ddpm6_cs0=chain(6, 6, 83, 0.0), ddpm6_cs0_seed=83), os.path.join(OUT, "r2_guided.pt"))


def forward_full(ns):
    torch.manual_seed(1234)
    m = ns.FlowDiffusion(is_train=False, sampling_timesteps=3, img_size=32, num_frames=40,
                         config_pth=os.path.join(ROOT, "config", "mug128.yaml"), pretrained_pth="").eval()
    perturb_lfae(m)
    g = torch.Generator().manual_seed(777)
    ref_img = torch.rand(2, 3, 128, 128, generator=g)
    # driving frames = smooth deformations of the reference image + noise, so the regions actually move
    vid = torch.stack([torch.roll(ref_img, shifts=(3 * i, -2 * i), dims=(2, 3)) * (1 - 0.05 * i) +
                       0.05 * torch.rand(2, 3, 128, 128, generator=g) for i in range(3)], 2)
    m.set_train_input(ref_img, vid, ["a", "b"])
    m.forward()
    src = m.region_predictor(ref_img)
    drv = m.region_predictor(vid[:, :, 2])
    bg = m.bg_predictor(ref_img, vid[:, :, 2])
    sl = lambda t: t[..., ::8, ::8].clone()
    torch.save(dict(seed=777, grid=m.real_vid_grid.clone(), conf=m.real_vid_conf.clone(), out_slice=sl(m.real_out_vid),
                    warped_slice=sl(m.real_warped_vid), fea_slice=m.ref_img_fea[:, ::16, ::4, ::4].clone(),
                    out_mean=m.real_out_vid.mean(), src_shift=src["shift"], src_covar=src["covar"], src_affine=src["affine"],
                    src_heat_slice=src["heatmap"][..., ::4, ::4].clone(), drv_shift=drv["shift"], drv_covar=drv["covar"],
                    drv_affine=drv["affine"], bg=bg), os.path.join(OUT, "r2_forward_full.pt"))


def mhad_region(ns):
    import yaml
    mp = yaml.safe_load(open(os.path.join(ROOT, "config", "mhad128.yaml")))["model_params"]
    torch.manual_seed(4321)
    rp = ns.RegionPredictor(num_regions=mp["num_regions"], num_channels=mp["num_channels"],
                            estimate_affine=mp["estimate_affine"], **mp["region_predictor_params"]).eval()
    g = torch.Generator().manual_seed(55)
    for mm in rp.modules():
        if hasattr(mm, "running_mean") and mm.running_mean is not None:
            mm.running_mean.copy_(torch.randn(mm.running_mean.shape, generator=g) * 0.1)
            mm.running_var.copy_(torch.rand(mm.running_var.shape, generator=g) * 0.5 + 0.75)
    img = torch.rand(2, 3, 128, 128, generator=g)
    r = rp(img)
    torch.save(dict(img_seed=55, shift=r["shift"], covar=r["covar"], affine=r["affine"], heatmap=r["heatmap"]),
               os.path.join(OUT, "r2_mhad_region.pt"))


def config1(ns):
    """BASELINE config 1 in full: demo_mug.py path, 50 DDIM steps on CPU, random-init weights, seed 1234."""
    torch.manual_seed(1234)
    m = ns.FlowDiffusion(is_train=False, sampling_timesteps=50, img_size=32, num_frames=40,
                         config_pth=os.path.join(ROOT, "config", "mug128.yaml"), pretrained_pth="").eval()
    torch.manual_seed(1234)
    img = torch.rand(1, 3, 128, 128)
    cond = torch.randn(1, 768)
    m.set_sample_input(img, cond)
    torch.manual_seed(99)
    t0 = time.time()
    m.sample_one_video(1.0)
    dt = time.time() - t0
    torch.save(dict(noise_seed=99, seconds=dt, threads=torch.get_num_threads(), frames_per_s=40.0 / dt,
                    grid=m.sample_vid_grid.clone(), conf=m.sample_vid_conf.clone(),
                    out_slice=m.sample_out_vid[:, :, :, ::4, ::4].clone(), out_mean=m.sample_out_vid.mean(),
                    warped_mean=m.sample_warped_vid.mean()), os.path.join(OUT, "r2_config1.pt"))
    print(f"config 1 on the reference: {dt:.1f} s, {40.0 / dt:.3f} frames/s, {torch.get_num_threads()} threads")


def main():
    ns = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    guided(ns)
    forward_full(ns)
    mhad_region(ns)
    if "--config1" in sys.argv:
        config1(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
