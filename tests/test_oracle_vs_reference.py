"""CPU, only where /root/reference exists (this container): pins the oracle and the product's parameter tree to the
UNMODIFIED reference executed through oracle/ref_shim.py."""
import pytest
import torch
from oracle.ref_shim import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def ns():
    from oracle.ref_shim import import_reference
    return import_reference()


def test_oracle_matches_reference_tiny(ns):
    from oracle import lfdm_oracle as O
    from oracle.make_golden import TINY_UNET
    torch.manual_seed(3)
    unet = ns.Unet3D(**TINY_UNET).eval()
    x, t, cond = torch.randn(2, 11, 5, 8, 8), torch.tensor([999, 0]), torch.randn(2, 24)
    with torch.no_grad():
        for cs in (0.0, 1.0, 3.0):
            ref = unet.forward_with_cond_scale(x, t, cond=cond, cond_scale=cs)
            mine = O.unet3d_forward_with_cond_scale(unet.state_dict(), x, t, cond, cs, heads=2)
            assert torch.equal(ref, mine)


def test_oracle_matches_reference_natops_style(ns):
    """nearest-upsample + reflect-padded 3x3 conv instead of ConvTranspose, learned null condition (NATOPS options,
    video_flow_diffusion.py:156-163,395-399): oracle == unmodified reference, bit for bit"""
    from oracle import lfdm_oracle as O
    from oracle.make_golden import TINY_UNET
    torch.manual_seed(5)
    cfg = dict(TINY_UNET, use_deconv=False, padding_mode="reflect", learn_null_cond=True)
    unet = ns.Unet3D(**cfg).eval()
    x, t, cond = torch.randn(2, 11, 5, 8, 8), torch.tensor([500, 7]), torch.randn(2, 24)
    with torch.no_grad():
        for cs in (1.0, 2.0):
            ref = unet.forward_with_cond_scale(x, t, cond=cond, cond_scale=cs)
            mine = O.unet3d_forward_with_cond_scale(unet.state_dict(), x, t, cond, cs, heads=2, padding_mode="reflect")
            assert torch.equal(ref, mine)


def test_state_dict_and_init_identity(ns):
    """same seed -> bit-identical parameters & key order as the reference (drop-in checkpoints, seeded goldens)"""
    import cvpr23_lfdm_b200 as P
    cfg = "/root/reference/config/mug128.yaml"
    for kw in (dict(), dict(learn_null_cond=True, use_deconv=False, padding_mode="reflect")):
        torch.manual_seed(1234)
        r = ns.FlowDiffusion(is_train=False, sampling_timesteps=3, config_pth=cfg, pretrained_pth="", **kw)
        torch.manual_seed(1234)
        m = P.FlowDiffusion(is_train=False, sampling_timesteps=3, config_pth=cfg, pretrained_pth="", **kw)
        a, b = r.state_dict(), m.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_repo_config_matches_reference_model_params():
    import yaml
    for name in ("mug128", "mhad128", "natops128"):
        ref = yaml.safe_load(open(f"/root/reference/config/{name}.yaml"))["model_params"]
        ours = yaml.safe_load(open(f"config/{name}.yaml"))["model_params"]
        ref.pop("avd_network_params", None)
        assert ref == ours
