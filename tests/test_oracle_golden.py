"""CPU: the oracle restatement vs the committed golden vectors (generated from the UNMODIFIED reference by
oracle/make_golden.py).  Bit-exact expected: same torch build generated them; tolerance 1e-6 guards other CPUs."""
import torch
import pytest
from oracle import lfdm_oracle as O

TOL = dict(rtol=1e-5, atol=1e-6)


def test_tiny_unet(golden):
    g = golden("tiny_unet.pt")
    kw = dict(heads=g["cfg"]["attn_heads"])
    y1 = O.unet3d_forward_with_cond_scale(g["sd"], g["x"], g["t"], g["cond"], 1.0, **kw)
    y2 = O.unet3d_forward_with_cond_scale(g["sd"], g["x"], g["t"], g["cond"], 2.0, **kw)
    torch.testing.assert_close(y1, g["y_scale1"], **TOL)
    torch.testing.assert_close(y2, g["y_scale2"], **TOL)


def test_tiny_sampler_steps(golden):
    g, u = golden("tiny_sampler.pt"), golden("tiny_unet.pt")
    buf = O.diffusion_buffers(1000)
    kw = dict(heads=u["cfg"]["attn_heads"])
    fea5 = g["fea"].unsqueeze(2).repeat(1, 1, 5, 1, 1)
    for st in g["steps"]:
        t = torch.full((2,), st["t"], dtype=torch.long)
        eps = O.unet3d_forward_with_cond_scale(u["sd"], torch.cat([st["x"], fea5], 1), t, g["cond"], 1.0, **kw)
        torch.manual_seed(st["seed"])
        noise = torch.randn_like(st["x"])
        out = O.p_sample_step(buf, st["x"], st["t"], eps, noise)
        torch.testing.assert_close(out, st["out"], **TOL)
        mean = O.p_sample_step(buf, st["x"], st["t"], eps, torch.zeros_like(noise))
        torch.testing.assert_close(mean, st["mean"], **TOL)


def test_tiny_ddim_and_ddpm_chain(golden):
    g, u = golden("tiny_sampler.pt"), golden("tiny_unet.pt")
    kw = dict(heads=u["cfg"]["attn_heads"])
    shape = (2, 3, 5, 8, 8)
    torch.manual_seed(g["ddim_seed"])
    out = O.sample_loop(u["sd"], g["fea"], g["cond"], shape, lambda i: torch.randn(shape), 4, 1000, unet_kw=kw)
    torch.testing.assert_close(out, g["ddim4"], rtol=1e-4, atol=1e-5)
    torch.manual_seed(g["ddpm6_seed"])
    out = O.sample_loop(u["sd"], g["fea"], g["cond"], shape, lambda i: torch.randn(shape), 6, 6, unet_kw=kw)
    torch.testing.assert_close(out, g["ddpm6"], rtol=1e-4, atol=1e-5)


def test_tiny_lfae(golden):
    g = golden("tiny_lfae.pt")
    r = O.generator_forward_with_flow(g["gen_sd"], g["img"], g["flow"], g["occ"])
    torch.testing.assert_close(r["prediction"], g["fwf"]["prediction"], **TOL)
    torch.testing.assert_close(r["deformed"], g["fwf"]["deformed"], **TOL)
    torch.testing.assert_close(O.generator_compute_fea(g["gen_sd"], g["img"]), g["fea"], **TOL)
    src = O.region_predictor(g["rp_sd"], g["img"])
    drv = O.region_predictor(g["rp_sd"], g["drv"])
    for k in ("shift", "covar", "heatmap"):
        torch.testing.assert_close(src[k], g["src_rp"][k], **TOL)
    torch.testing.assert_close(O.bg_motion_predictor(g["bg_sd"], g["img"], g["drv"]), g["bg"], **TOL)
    full = O.generator_forward(g["gen_sd"], g["img"], g["drv_rp"], g["src_rp"], g["bg"])
    for k in ("prediction", "deformed", "optical_flow", "occlusion_map", "bottle_neck_feat"):
        torch.testing.assert_close(full[k], g["full"][k], rtol=1e-4, atol=1e-5)
