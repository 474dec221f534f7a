"""GPU (-m gpu): round-2 parity cases that close the holes VERDICT r1 lists: the benchmarked batch (B = 8) at full size,
classifier-free guidance inside the sampler loops, FlowDiffusion.forward (real-video branch) and the region / background
predictors at 128x128 against goldens of the UNMODIFIED reference, the mhad/natops `pad: 0` option, the NATOPS UNet options at
dim = 64 (tcgen05 engine exercised), repeated sample() calls (cached step graph), engine-cache invalidation.
Every comparison also records its measured error (fixture `parity_log` -> profiles/r02_parity_errors.md)."""
import os
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-3, 1e-4


def close(a, b, what, rtol=RTOL, atol=ATOL, log=None):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if log is not None:
        log(what, a, b, tol=[rtol, atol])
    err = (a - b).abs()
    bad = (err > atol + rtol * b.abs()).sum().item()
    assert bad == 0, f"{what}: {bad}/{a.numel()} out of tolerance; max abs err {err.max().item():.3e}; ref rms {b.pow(2).mean().sqrt().item():.3e}"


@pytest.fixture(scope="module")
def full_model():
    import cvpr23_lfdm_b200 as P
    torch.manual_seed(1234)
    m = P.FlowDiffusion(is_train=False, sampling_timesteps=3, img_size=32, num_frames=40,
                        config_pth=os.path.join(ROOT, "config", "mug128.yaml"), pretrained_pth="")
    return m.cuda().eval()


# ------------------------------------------------------------------------------------------------ the benchmarked batch
def test_full_unet_and_decode_batch8_vs_oracle(full_model, parity_log):
    """B = 8 (the bench batch): per-sample GroupNorm flushes, per-sample (scale, shift) rows, frame-broadcast init-conv
    residual, the fused temporal block across sample boundaries; then all 40 frames of all 8 samples through decode_video."""
    from oracle import lfdm_oracle as O
    g = torch.Generator().manual_seed(808)
    b = 8
    x = torch.randn(b, 3, 40, 32, 32, generator=g)
    fea = torch.randn(b, 256, 32, 32, generator=g).abs()
    cond = torch.randn(b, 768, generator=g)
    t = torch.full((b,), 417)
    sd = {k: v.detach().cpu() for k, v in full_model.unet.state_dict().items()}
    ref = torch.cat([O.unet3d_forward(sd, torch.cat([x[i:i + 1], fea[i:i + 1].unsqueeze(2).repeat(1, 1, 40, 1, 1)], 1),
                                      t[i:i + 1], cond[i:i + 1]) for i in range(b)], 0)
    eng = full_model.unet.engine()
    ss = eng.scale_shift(t.cuda(), cond.cuda())
    got = eng.forward_hoisted(x.cuda().contiguous(), eng.prepare_fea(fea.cuda()), ss)
    close(got, ref, "full UNet eval B=8 (hoisted, bench batch)", log=parity_log)
    # decode: 8 samples x 40 frames in one batch vs the oracle frame by frame
    img = torch.rand(b, 3, 128, 128, generator=g)
    grid = torch.rand(b, 2, 40, 32, 32, generator=g) * 2.2 - 1.1
    conf = torch.rand(b, 1, 40, 32, 32, generator=g)
    gsd = {k: v.detach().cpu() for k, v in full_model.generator.state_dict().items()}
    pred, deformed = full_model.generator.decode_video(img.cuda(), grid.cuda(), conf.cuda())
    for f in (0, 17, 39):
        r = O.generator_forward_with_flow(gsd, img, grid[:, :, f].permute(0, 2, 3, 1), conf[:, :, f])
        close(pred[:, :, f], r["prediction"], f"decode_video B=8 frame {f} prediction", log=parity_log)
        close(deformed[:, :, f], r["deformed"], f"decode_video B=8 frame {f} deformed", log=parity_log)


def test_unet_eval_is_bitwise_reproducible(full_model):
    """No float atomics on the default path (GroupNorm sums, stream-K partials and attention partials are all added in a fixed
    order): the same eval gives the same bits, so a sampled video is a function of its seed only."""
    g = torch.Generator().manual_seed(99)
    b = 8
    x = torch.randn(b, 3, 40, 32, 32, generator=g).cuda()
    fea = torch.randn(b, 256, 32, 32, generator=g).abs().cuda()
    cond = torch.randn(b, 768, generator=g).cuda()
    t = torch.full((b,), 300).cuda()
    eng = full_model.unet.engine()
    ss = eng.scale_shift(t, cond)
    pf = eng.prepare_fea(fea)
    first = eng.forward_hoisted(x, pf, ss).clone()
    for _ in range(3):
        again = eng.forward_hoisted(x, pf, ss)
        assert torch.equal(again, first), f"eval differs between runs by {(again - first).abs().max().item():.3e}"


# ------------------------------------------------------------------------------------------------ guidance in the loops
def _tiny_guided(golden):
    import cvpr23_lfdm_b200 as P
    g = golden("r2_guided.pt")
    u = P.Unet3D(**g["cfg"])
    u.load_state_dict(g["sd"])
    return g, u.cuda().eval()


def test_guided_sampler_matches_reference_golden(golden, parity_log):
    """cond_scale = 2 inside p_sample / p_sample_loop / ddim_sample (reference :521-526 via :714): one 2B batch here"""
    import cvpr23_lfdm_b200 as P
    g, u = _tiny_guided(golden)
    fea, cond = g["fea"].cuda(), g["cond"].cuda()
    gd = P.GaussianDiffusion(u, image_size=8, num_frames=5, sampling_timesteps=1000, timesteps=1000, loss_type='l2',
                             use_dynamic_thres=True, null_cond_prob=0.1).cuda().eval()
    for st in g["steps"]:
        t = torch.full((2,), st["t"], dtype=torch.long, device="cuda")
        torch.manual_seed(st["seed"])
        ref_noise = torch.randn_like(st["x"])
        gd.noise_fn = lambda shape, device, n=ref_noise: n
        out = gd.p_sample(st["x"].cuda(), t, fea, cond=cond, cond_scale=2.0)
        close(out, st["out"], f"guided p_sample t={st['t']} (teacher forced)", rtol=2e-3, atol=2e-3, log=parity_log)

    # per-sample timesteps (the reference's `extract(a, t, x_shape)` indexes the schedule per sample, :714-735): sample 0 of one
    # golden step and sample 1 of another in ONE call must reproduce the two golden outputs
    if len(g["steps"]) >= 2 and g["steps"][0]["t"] != g["steps"][1]["t"]:
        s0, s1 = g["steps"][0], g["steps"][1]
        noises = []
        for st in (s0, s1):
            torch.manual_seed(st["seed"])
            noises.append(torch.randn_like(st["x"]))
        mixed_noise = torch.stack([noises[0][0], noises[1][1]])
        gd.noise_fn = lambda shape, device, n=mixed_noise: n
        xm = torch.stack([s0["x"][0], s1["x"][1]]).cuda()
        tm = torch.tensor([s0["t"], s1["t"]], dtype=torch.long, device="cuda")
        out = gd.p_sample(xm, tm, fea, cond=cond, cond_scale=2.0)
        close(out, torch.stack([s0["out"][0], s1["out"][1]]), "guided p_sample with per-sample timesteps", rtol=2e-3, atol=2e-3,
              log=parity_log)

    def chain(sampling, timesteps, seed, cs):
        d = P.GaussianDiffusion(u, image_size=8, num_frames=5, sampling_timesteps=sampling, timesteps=timesteps,
                                loss_type='l2', use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0).cuda().eval()
        torch.manual_seed(seed)
        draws = [torch.randn(2, 3, 5, 8, 8) for _ in range(sampling + 2)]
        it = iter(draws)
        d.noise_fn = lambda shape, device: next(it)
        return d.sample(fea, cond=cond, cond_scale=cs)
    close(chain(4, 1000, g["ddim4_seed"], 2.0), g["ddim4_cs2"], "guided ddim 4-step chain (cond_scale 2)", rtol=5e-3, atol=5e-3, log=parity_log)
    close(chain(6, 6, g["ddpm6_seed"], 2.0), g["ddpm6_cs2"], "guided ddpm 6-step chain (cond_scale 2)", rtol=5e-3, atol=5e-3, log=parity_log)
    close(chain(6, 6, g["ddpm6_cs0_seed"], 0.0), g["ddpm6_cs0"], "null-condition ddpm 6-step chain (cond_scale 0)", rtol=5e-3, atol=5e-3, log=parity_log)


def test_guided_graph_loop_is_one_2b_batch_and_matches_two_pass(full_model, parity_log):
    """the captured guided loop (one 2B batch, lerp in the head kernel) == two evaluations + lerp; cond_scale 0 runs ONE evaluation"""
    import cvpr23_lfdm_b200 as P
    u = full_model.unet
    g = torch.Generator().manual_seed(91)
    x = torch.randn(2, 259, 40, 32, 32, generator=g).cuda()
    t = torch.tensor([700, 700]).cuda()
    c = torch.randn(2, 768, generator=g).cuda()
    one = u.forward_with_cond_scale(x, t, cond=c, cond_scale=2.0)
    lc = u.forward(x, t, cond=c, null_cond_prob=0.)
    ln = u.forward(x, t, cond=c, null_cond_prob=1.)
    close(one, ln + (lc - ln) * 2.0, "CFG: one 2B batch vs two passes", rtol=1e-3, atol=1e-4, log=parity_log)
    gd = P.GaussianDiffusion(u, image_size=32, num_frames=40, sampling_timesteps=5, timesteps=1000, loss_type='l2',
                             use_dynamic_thres=True).cuda().eval()
    fea = torch.rand(1, 256, 32, 32, device="cuda")
    from cvpr23_lfdm_b200 import _lib as L
    torch.manual_seed(3)
    gd.sample(fea, cond=c[:1], cond_scale=1.0)
    n1 = gd._engine().last_stats["calls_per_step"]
    torch.manual_seed(3)
    gd.sample(fea, cond=c[:1], cond_scale=0.0)
    n0 = gd._engine().last_stats["calls_per_step"]
    torch.manual_seed(3)
    gd.sample(fea, cond=c[:1], cond_scale=3.0)
    n3 = gd._engine().last_stats["calls_per_step"]
    assert n0 == n1, (n0, n1)                 # cond_scale 0: a single (null) evaluation per step, reference :515-516
    assert n3 <= n1 + 3, (n3, n1)             # guided: still ONE UNet pass per step (2B batch), not two


# ------------------------------------------------------------------------------------------------ real-video branch
def test_flowdiffusion_forward_matches_reference_golden(golden, full_model, parity_log):
    """FlowDiffusion.forward (video_flow_diffusion_model.py:116-143): RegionPredictor / BGMotionPredictor / Generator.forward
    at 128x128 on 3 driving frames, against the unmodified reference (full size, seed-1234 weights + perturb_lfae)."""
    import cvpr23_lfdm_b200 as P
    from oracle.make_golden_r2 import perturb_lfae
    fp = golden("r2_forward_full.pt")
    torch.manual_seed(1234)
    m = P.FlowDiffusion(is_train=False, sampling_timesteps=3, img_size=32, num_frames=40,
                        config_pth=os.path.join(ROOT, "config", "mug128.yaml"), pretrained_pth="").cuda().eval()
    perturb_lfae(m)
    g = torch.Generator().manual_seed(fp["seed"])
    ref_img = torch.rand(2, 3, 128, 128, generator=g)
    vid = torch.stack([torch.roll(ref_img, shifts=(3 * i, -2 * i), dims=(2, 3)) * (1 - 0.05 * i) +
                       0.05 * torch.rand(2, 3, 128, 128, generator=g) for i in range(3)], 2)
    src = m.region_predictor(ref_img.cuda())
    drv = m.region_predictor(vid[:, :, 2].cuda())
    for k, r in (("src", src), ("drv", drv)):
        close(r["shift"], fp[f"{k}_shift"], f"RegionPredictor@128 {k} shift", rtol=1e-3, atol=2e-4, log=parity_log)
        close(r["covar"], fp[f"{k}_covar"], f"RegionPredictor@128 {k} covar", rtol=1e-3, atol=2e-4, log=parity_log)
        a, ref = r["affine"].cpu(), fp[f"{k}_affine"]
        sign = torch.sign((a * ref).sum(dim=-2, keepdim=True))
        close(a * sign, ref, f"RegionPredictor@128 {k} affine (up to column sign)", rtol=5e-3, atol=1e-3, log=parity_log)
    close(src["heatmap"][..., ::4, ::4], fp["src_heat_slice"], "RegionPredictor@128 heatmap", rtol=1e-3, atol=2e-4, log=parity_log)
    close(m.bg_predictor(ref_img.cuda(), vid[:, :, 2].cuda()), fp["bg"], "BGMotionPredictor@128", log=parity_log)
    m.set_train_input(ref_img, vid, ["a", "b"])
    m.forward()
    close(m.real_vid_grid, fp["grid"], "FlowDiffusion.forward real_vid_grid", rtol=2e-3, atol=5e-4, log=parity_log)
    close(m.real_vid_conf, fp["conf"], "FlowDiffusion.forward real_vid_conf", rtol=2e-3, atol=5e-4, log=parity_log)
    close(m.real_out_vid[..., ::8, ::8], fp["out_slice"], "FlowDiffusion.forward real_out_vid", rtol=2e-3, atol=5e-4, log=parity_log)
    close(m.real_warped_vid[..., ::8, ::8], fp["warped_slice"], "FlowDiffusion.forward real_warped_vid", rtol=2e-3, atol=5e-4, log=parity_log)
    close(m.ref_img_fea[:, ::16, ::4, ::4], fp["fea_slice"], "FlowDiffusion.forward ref_img_fea", log=parity_log)
    # the batched branch (all driving frames in one pass) against the reference's per-frame control flow
    batched = [t.clone() for t in (m.real_vid_grid, m.real_vid_conf, m.real_out_vid, m.real_warped_vid, m.ref_img_fea)]
    m.forward_per_frame()
    for name, a, r in zip(("grid", "conf", "out", "warped", "fea"), batched,
                          (m.real_vid_grid, m.real_vid_conf, m.real_out_vid, m.real_warped_vid, m.ref_img_fea)):
        assert a.shape == r.shape, (name, a.shape, r.shape)
        close(a, r, f"FlowDiffusion.forward batched vs per-frame: {name}", rtol=1e-4, atol=1e-5, log=parity_log)


def test_region_predictor_pad0_matches_reference_golden(golden, parity_log):
    """mhad128 / natops128 `pad: 0` (region_predictor.py:33-35: 7x7 heads without padding -> 26x26 heatmaps)"""
    import yaml
    import cvpr23_lfdm_b200 as P
    fp = golden("r2_mhad_region.pt")
    mp = yaml.safe_load(open(os.path.join(ROOT, "config", "mhad128.yaml")))["model_params"]
    torch.manual_seed(4321)
    rp = P.RegionPredictor(num_regions=mp["num_regions"], num_channels=mp["num_channels"],
                           estimate_affine=mp["estimate_affine"], **mp["region_predictor_params"])
    g = torch.Generator().manual_seed(fp["img_seed"])
    for mm in rp.modules():
        if hasattr(mm, "running_mean") and mm.running_mean is not None:
            mm.running_mean.copy_(torch.randn(mm.running_mean.shape, generator=g) * 0.1)
            mm.running_var.copy_(torch.rand(mm.running_var.shape, generator=g) * 0.5 + 0.75)
    img = torch.rand(2, 3, 128, 128, generator=g)
    r = rp.cuda().eval()(img.cuda())
    assert r["heatmap"].shape == fp["heatmap"].shape
    close(r["heatmap"], fp["heatmap"], "RegionPredictor pad=0 heatmap", rtol=1e-3, atol=2e-4, log=parity_log)
    close(r["shift"], fp["shift"], "RegionPredictor pad=0 shift", rtol=1e-3, atol=2e-4, log=parity_log)
    close(r["covar"], fp["covar"], "RegionPredictor pad=0 covar", rtol=1e-3, atol=2e-4, log=parity_log)
    a, ref = r["affine"].cpu(), fp["affine"]
    close(a * torch.sign((a * ref).sum(dim=-2, keepdim=True)), ref, "RegionPredictor pad=0 affine (up to column sign)",
          rtol=5e-3, atol=1e-3, log=parity_log)


# ------------------------------------------------------------------------------------------------ NATOPS options on the TC engine
def test_natops_options_dim64_vs_oracle(parity_log):
    """demo_natops.py:23-32 options (nearest x2 + reflect-padded 3x3 up-conv, learned null condition, CFG) at dim = 64:
    every 64-multiple layer runs on the tcgen05 engine, 40 frames -> the fused temporal block"""
    import cvpr23_lfdm_b200 as P
    from oracle import lfdm_oracle as O
    torch.manual_seed(32)
    u = P.Unet3D(dim=64, cond_dim=24, dim_mults=(1, 2), channels=3 + 64, attn_heads=8, use_deconv=False,
                 padding_mode="reflect", learn_null_cond=True)
    sd = {k: v.detach().clone() for k, v in u.state_dict().items()}
    u = u.cuda().eval()
    x, t, c = torch.randn(2, 67, 40, 16, 16), torch.tensor([900, 3]), torch.randn(2, 24)
    for cs in (1.0, 2.5):
        ref = O.unet3d_forward_with_cond_scale(sd, x, t, c, cs, heads=8, padding_mode="reflect")
        got = u.forward_with_cond_scale(x.cuda(), t.cuda(), cond=c.cuda(), cond_scale=cs)
        close(got, ref, f"natops options dim=64 cond_scale={cs}", rtol=2e-3, atol=2e-4, log=parity_log)
    eng = u.engine()
    assert eng.stages_up[0]["up"] is not None and eng.stages_up[0]["up"].reflect == 1


# ------------------------------------------------------------------------------------------------ host-side state
def test_repeated_sample_reuses_step_graph(full_model, parity_log):
    import cvpr23_lfdm_b200 as P
    gd = P.GaussianDiffusion(full_model.unet, image_size=32, num_frames=40, sampling_timesteps=6, timesteps=1000,
                             loss_type='l2', use_dynamic_thres=True).cuda().eval()
    fea = torch.rand(1, 256, 32, 32, device="cuda")
    cond = torch.randn(1, 768, device="cuda")
    outs = []
    for i in range(3):
        torch.manual_seed(5)
        outs.append(gd.sample(fea if i < 2 else fea * 0.5, cond=cond, cond_scale=1.0).clone())
        assert gd._engine().last_stats["graph_reused"] == (i > 0)
    close(outs[1], outs[0], "second sample() call (cached tables + replayed graph) vs first", rtol=1e-3, atol=1e-4, log=parity_log)
    assert (outs[2] - outs[0]).abs().max().item() > 1e-3          # new inputs really flow into the persistent buffers


def test_lfae_engines_follow_parent_load_state_dict():
    """engines cached inside RegionPredictor / BGMotionPredictor / Generator must notice a load_state_dict issued
    through a PARENT module (nn.Module.load_state_dict never calls the child override) and BN running-stat changes"""
    import cvpr23_lfdm_b200 as P
    torch.manual_seed(8)
    cfg = dict(num_regions=4, num_channels=3, estimate_affine=True, temperature=0.1, block_expansion=16, max_features=64,
               scale_factor=0.25, num_blocks=2, pca_based=True, fast_svd=False)
    holder = torch.nn.ModuleDict({"rp": P.RegionPredictor(**cfg)}).cuda().eval()
    img = torch.rand(1, 3, 32, 32, device="cuda")
    a = holder["rp"](img)["shift"].clone()
    other = torch.nn.ModuleDict({"rp": P.RegionPredictor(**cfg)})
    holder.load_state_dict(other.state_dict())                   # through the parent
    b = holder["rp"](img)["shift"].clone()
    ref = other.cuda().eval()["rp"](img)["shift"]
    assert torch.allclose(b, ref, atol=1e-6) and not torch.allclose(a, b, atol=1e-6)
    for m in holder.modules():                                   # BatchNorm running statistics are buffers
        if hasattr(m, "running_var") and m.running_var is not None:
            m.running_var.mul_(1.7)
    c = holder["rp"](img)["shift"]
    assert not torch.allclose(c, b, atol=1e-6)


# ------------------------------------------------------------------------------------------------ BASELINE config 1, in full
def test_config1_full_50step_run_vs_reference(golden, parity_log):
    """BASELINE.json configs[0] / BASELINE.md 3.3: MUG-128, B = 1, 40 frames, 50 DDIM steps, seed-1234 weights and inputs, CPU
    noise tape seed 99 -- the unmodified reference ran this in full on the CPU (oracle/make_golden_r2.py --config1); the same
    seeds go through the B200 path.  A 50-step free-running chain amplifies fp32-level differences (x0 = 6.4e4 * eps at
    t = 980, dynamic-threshold clamp), so the bound here is on the video, with the measured error recorded."""
    import cvpr23_lfdm_b200 as P
    fp = golden("r2_config1.pt")
    torch.manual_seed(1234)
    m = P.FlowDiffusion(is_train=False, sampling_timesteps=50, img_size=32, num_frames=40,
                        config_pth=os.path.join(ROOT, "config", "mug128.yaml"), pretrained_pth="").cuda().eval()
    torch.manual_seed(1234)
    img = torch.rand(1, 3, 128, 128)
    cond = torch.randn(1, 768)
    m.set_sample_input(img, cond.cuda())
    gen = torch.Generator().manual_seed(fp["noise_seed"])
    m.diffusion.noise_fn = lambda shape, device: torch.randn(shape, generator=gen)
    m.sample_one_video(1.0)
    e_grid = parity_log("config 1 (50 DDIM steps, full run): sample_vid_grid", m.sample_vid_grid, fp["grid"])
    e_conf = parity_log("config 1 (50 DDIM steps, full run): sample_vid_conf", m.sample_vid_conf, fp["conf"])
    e_out = parity_log("config 1 (50 DDIM steps, full run): sample_out_vid", m.sample_out_vid[:, :, :, ::4, ::4], fp["out_slice"])
    assert abs(m.sample_out_vid.mean().item() - fp["out_mean"].item()) < 2e-3
    assert e_out[0] < 5e-2 and e_grid[0] < 5e-2 and e_conf[0] < 5e-2, (e_grid, e_conf, e_out)


# ------------------------------------------------------------------------------------------------ drop-in + output stage
def test_demo_mug_call_sequence_with_string_labels_and_panels(tmp_path, parity_log):
    """demo/demo_mug.py:86-145 through the `DM.` alias package: FlowDiffusion(...) -> set_sample_input(img, [label]) ->
    sample_one_video(cond_scale) with STRING labels (label table instead of torch.hub BERT), then the output stage:
    uint8 quantisation as sample_img (:68-74), confidence panel as misc.conf2fig, 5-panel composition, GIF on a worker thread."""
    import numpy as np
    from DM.modules.video_flow_diffusion_model import FlowDiffusion          # the alias the reference's demo imports
    from cvpr23_lfdm_b200.dm import text as T
    from cvpr23_lfdm_b200.output import AsyncGifWriter
    import misc
    g = torch.Generator().manual_seed(17)
    table = {s: torch.randn(768, generator=g) for s in T.LABELS["mug"]}
    T.clear_text_embeddings()
    T.register_text_embeddings(table)
    try:
        torch.manual_seed(1234)
        model = FlowDiffusion(is_train=True, sampling_timesteps=4, config_pth=os.path.join(ROOT, "config", "mug128.yaml"),
                              pretrained_pth="")
        model.cuda()
        model.eval()
        ref_imgs = torch.rand(1, 3, 128, 128, generator=g).cuda()
        writer = AsyncGifWriter()
        for n, label in enumerate(["anger", "surprise"]):
            torch.manual_seed(7)
            model.set_sample_input(sample_img=ref_imgs, sample_text=[label])
            model.sample_one_video(cond_scale=1.0)
            out_str = model.sample_out_vid.clone()
            torch.manual_seed(7)
            model.set_sample_input(sample_img=ref_imgs, sample_text=table[label][None].cuda())
            model.sample_one_video(cond_scale=1.0)
            # two free-running 4-step chains: equal up to the run-to-run order of the GroupNorm partial-sum atomics
            close(out_str, model.sample_out_vid, f"string label '{label}' == its table embedding passed as a tensor",
                  rtol=5e-3, atol=5e-3, log=parity_log)
            frames = model.render_sample_panels(0, mean=(0.0, 0.0, 0.0))
            assert frames.shape == (40, 128, 640, 3) and frames.dtype == torch.uint8
            writer.submit(frames, str(tmp_path / f"{n:04d}_{label}.gif"))
            fr = frames.cpu().numpy()
            for fi in (0, 23, 39):
                def sample_img(batch):       # demo_mug.py:68-74 (MEAN = 0)
                    a = batch[0].permute(1, 2, 0).cpu().numpy().copy()
                    a[a < 0] = 0
                    a[a > 1] = 1
                    a *= 255
                    return np.array(a, np.uint8)
                assert np.array_equal(fr[fi, :, 0:128], sample_img(ref_imgs))
                assert np.array_equal(fr[fi, :, 128:256], sample_img(model.sample_out_vid[:, :, fi]))
                assert np.array_equal(fr[fi, :, 256:384], sample_img(model.sample_warped_vid[:, :, fi]))
                conf = misc.conf2fig(model.sample_vid_conf[0, :, fi])
                assert np.array_equal(fr[fi, :, 512:640, 0], conf) and np.array_equal(fr[fi, :, 512:640, 2], conf)
                gp = fr[fi, :, 384:512].astype(np.int32)
                # grid figure: white background, blue (C0) warped-grid strokes and grey identity strokes present
                assert (gp.sum(-1) == 765).mean() > 0.05
                assert ((gp[..., 2] - gp[..., 0]) > 60).mean() > 0.02
        writer.close()
        from PIL import Image
        im = Image.open(str(tmp_path / "0001_surprise.gif"))
        assert im.n_frames == 40 and im.size == (640, 128)
    finally:
        T.clear_text_embeddings()


def test_encoder_cache_is_not_fooled_by_address_reuse():
    """the encoder output kept between compute_fea and decode_video is keyed on the tensor object, not its address: a new
    image that the caching allocator places at a freed image's address must be re-encoded (demo loops do exactly this)"""
    import cvpr23_lfdm_b200 as P
    torch.manual_seed(3)
    gen = P.Generator(num_channels=3, num_regions=4, block_expansion=64, max_features=256, num_down_blocks=2,
                      num_bottleneck_blocks=1, skips=True).cuda().eval()
    a = torch.rand(1, 3, 64, 64, device="cuda")
    fa = gen.compute_fea(a).clone()
    addr = a.data_ptr()
    del a
    b = torch.rand(1, 3, 64, 64, device="cuda")
    fb = gen.compute_fea(b).clone()
    fresh = gen.compute_fea(b.clone())
    assert torch.equal(fb, fresh)
    if b.data_ptr() == addr:
        assert not torch.equal(fb, fa)
    b.mul_(0.5)                                    # in-place edit of the cached tensor -> version bump -> re-encode
    assert torch.equal(gen.compute_fea(b), gen.compute_fea(b.clone()))
