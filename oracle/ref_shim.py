"""TEST INFRASTRUCTURE ONLY (oracle/): import the UNMODIFIED reference from /root/reference on a CPU box.

The reference (nihaomiao/CVPR23_LFDM) hard-imports a few packages that are absent from this
image (einops_exts, rotary_embedding_torch, matplotlib, skimage, imageio, flow_vis) and calls
`.cuda()` in constructors (DM/modules/video_flow_diffusion_model.py:41,50;
DM/modules/video_flow_diffusion.py:440,560).  This module installs *shims* in `sys.modules`
(no reference file is touched or copied) so that the reference's own arithmetic can be executed
here to (a) pin the oracle restatement in `oracle/lfdm_oracle.py` and (b) generate the golden
fixtures under `tests/golden/` (see `oracle/make_golden.py`).

Third-party restatements (sources not vendored in /root/reference, pinned in requirements.txt):
  * einops_exts==0.0.3  rearrange_many(tensors, pattern, **kw) -> map(rearrange)
  * rotary_embedding_torch==0.1.5  RotaryEmbedding(dim): freqs = 1/10000^(arange(0,dim,2)/dim);
    rotate_queries_or_keys(t, seq_dim=-2): f = outer(arange(n), freqs) repeated pairwise
    ('n f -> n (f r)', r=2); t*cos(f) + rotate_half(t)*sin(f); rotate_half on interleaved pairs
    (x1,x2)->(-x2,x1).  No reference test pins this: "parity unpinned" at that boundary.

Nothing under cvpr23_lfdm_b200/ may import this file.
"""
import sys
import types
import os

REFERENCE_ROOT = os.environ.get("LFDM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "DM", "modules"))


def _install_stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims(patch_cuda=None):
    import torch
    from einops import rearrange, repeat

    if "einops_exts" not in sys.modules:
        def rearrange_many(tensors, pattern, **kw):
            return tuple(rearrange(t, pattern, **kw) for t in tensors)
        _install_stub("einops_exts", rearrange_many=rearrange_many)

    if "rotary_embedding_torch" not in sys.modules:
        class RotaryEmbedding(torch.nn.Module):
            def __init__(self, dim, theta=10000):
                super().__init__()
                freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
                self.register_buffer("freqs", freqs)

            @staticmethod
            def _rotate_half(x):
                x = rearrange(x, "... (d r) -> ... d r", r=2)
                x1, x2 = x.unbind(dim=-1)
                x = torch.stack((-x2, x1), dim=-1)
                return rearrange(x, "... d r -> ... (d r)")

            def rotate_queries_or_keys(self, t, seq_dim=-2):
                n = t.shape[seq_dim]
                pos = torch.arange(n, device=t.device).type_as(self.freqs)
                f = torch.einsum("..., f -> ... f", pos, self.freqs)
                f = repeat(f, "... n -> ... (n r)", r=2)
                rot = f.shape[-1]
                tl, tm, tr = t[..., :0], t[..., :rot], t[..., rot:]
                tm = tm * f.cos() + self._rotate_half(tm) * f.sin()
                return torch.cat((tl, tm, tr), dim=-1)
        _install_stub("rotary_embedding_torch", RotaryEmbedding=RotaryEmbedding)

    # import-time-only baggage of LFAE/modules/util.py:17-18, misc.py:9-12, demo scripts
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _install_stub("matplotlib")
        plt = _install_stub("matplotlib.pyplot")
        col = _install_stub("matplotlib.collections", LineCollection=object)
        mpl.pyplot = plt
        mpl.collections = col
        mpl.use = lambda *a, **k: None
    try:
        import skimage  # noqa: F401
    except Exception:
        sk = _install_stub("skimage")
        sk.draw = _install_stub("skimage.draw", disk=lambda *a, **k: None)
    for name in ("imageio", "flow_vis"):
        try:
            __import__(name)
        except Exception:
            _install_stub(name)

    if patch_cuda is None:
        patch_cuda = not torch.cuda.is_available()
    if patch_cuda and not getattr(torch.nn.Module, "_lfdm_cuda_patched", False):
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module._lfdm_cuda_patched = True



_NS = None
_REF_TOP = ("DM", "LFAE", "misc", "sync_batchnorm")


def import_reference():
    """Returns a namespace with the reference's hot-path classes, imported from REFERENCE_ROOT.

    This repo ships alias packages with the reference's own top-level names (`DM`, `LFAE`, `misc`: the drop-in import
    path of INTEGRATION.md), so the reference is imported with those names temporarily unbound and its modules are
    then detached from `sys.modules` again — both implementations can live in one process without shadowing."""
    global _NS
    if _NS is not None:
        return _NS
    install_shims()
    stash = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_TOP}
    for k in stash:
        del sys.modules[k]
    # the reference's DM/ and LFAE/ are namespace packages (no __init__.py): a regular package of the same name
    # anywhere on sys.path would win, so every path entry that holds one is hidden during the import
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [p_ for p_ in saved_path if p_ != REFERENCE_ROOT and not any(
        os.path.exists(os.path.join(p_ or ".", t, "__init__.py")) or os.path.exists(os.path.join(p_ or ".", t + ".py"))
        for t in _REF_TOP)]
    ns = types.SimpleNamespace()
    try:
        import DM.modules.video_flow_diffusion as vfd
        import DM.modules.video_flow_diffusion_model as vfdm
        import LFAE.modules.generator as gen
        import LFAE.modules.region_predictor as rp
        import LFAE.modules.bg_motion_predictor as bg
        import LFAE.modules.util as util
        for m in (vfd, vfdm, gen, rp, bg, util):
            assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), m.__file__
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_TOP]:
            del sys.modules[k]
        sys.modules.update(stash)
    ns.vfd, ns.vfdm, ns.gen, ns.rp, ns.bg, ns.util = vfd, vfdm, gen, rp, bg, util
    ns.Unet3D, ns.GaussianDiffusion = vfd.Unet3D, vfd.GaussianDiffusion
    ns.FlowDiffusion = vfdm.FlowDiffusion
    ns.Generator, ns.RegionPredictor, ns.BGMotionPredictor = gen.Generator, rp.RegionPredictor, bg.BGMotionPredictor
    _NS = ns
    return ns
