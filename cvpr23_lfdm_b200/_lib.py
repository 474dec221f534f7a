"""ctypes binding of liblfdm_b200.so (C-ABI declared in include/lfdm_b200.h).

Fails loudly: there is no PyTorch / CPU fallback behind these calls.  Every wrapper launches on
`torch.cuda.current_stream()` so calls are ordered with torch work and are CUDA-graph capturable."""
import ctypes as C
import os
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblfdm_b200.so")

ENGINE_SIMT, ENGINE_TC = 0, 1
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
CONV_DIRECT, CONV_TRANSPOSED, CONV_UPNEAREST = 0, 1, 2
E_UNSUPP = -2


class LfdmError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("a_sb", C.c_void_p * 2), ("a_f32", C.c_void_p * 2), ("a_plane", C.c_int64 * 2), ("a_c", C.c_int32 * 2),
        ("nf", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32), ("h_out", C.c_int32), ("w_out", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("pad", C.c_int32), ("stride", C.c_int32), ("mode", C.c_int32),
        ("reflect", C.c_int32),
        ("w_f32", C.c_void_p), ("w_sb", C.c_void_p), ("w_plane", C.c_int64), ("bias", C.c_void_p), ("c_out", C.c_int32),
        ("residual", C.c_void_p), ("res_bcast_f", C.c_int32), ("out_f32", C.c_void_p), ("f32_act", C.c_int32),
        ("out_sb", C.c_void_p), ("out_plane", C.c_int64), ("sb_act", C.c_int32), ("sb_scale", C.c_void_p),
        ("sb_shift", C.c_void_p), ("gn_stats", C.c_void_p), ("gn_cpg", C.c_int32), ("rows_per_sample", C.c_int32),
        ("rot_cos", C.c_void_p), ("rot_sin", C.c_void_p), ("rot_frames", C.c_int32), ("rot_rows_per_frame", C.c_int32),
        ("rot_cols", C.c_int32), ("rot_scale_cols", C.c_int32), ("rot_scale", C.c_float),
        ("sk_workspace", C.c_void_p), ("sk_workspace_bytes", C.c_int64), ("sk_flags", C.c_void_p), ("sk_slots", C.c_int32),
    ]


_lib = None


def lib():
    """Loads the CUDA library or raises (never falls back)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LfdmError(f"{LIB_PATH} is missing: run `python -m cvpr23_lfdm_b200.build` "
                            "(or __graft_entry__.build()). The B200 path has no CPU/PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l):
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        "lfdm_conv": [C.POINTER(ConvDesc), i32, vp],
        "lfdm_gn_stats": [vp, i64, i32, i32, i32, vp, vp],
        "lfdm_gn_apply": [vp, vp, vp, vp, vp, i64, vp, vp, vp, i64, i64, i32, i32, i32, f32, vp],
        "lfdm_layernorm": [vp, vp, vp, i64, vp, i64, i32, f32, vp],
        "lfdm_attn_softmax": [vp, vp, i64, vp, i64, i32, i32, i64, i64, i64, i64, vp, vp, vp, vp],
        "lfdm_attn_softmax_pre": [vp, vp, i64, vp, i64, i32, i32, i64, i64, i64, i64, vp, vp],
        "lfdm_attn_linear": [vp, vp, i64, vp, i64, i32, i32, vp],
        "lfdm_attn_temporal_fused": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, f32, vp, vp],
        "lfdm_attn_linear_fused": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, vp],
        "lfdm_small_linear": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "lfdm_sinusoidal": [vp, vp, vp, i32, i32, vp],
        "lfdm_ss_combine": [vp, vp, vp, vp, i32, i32, vp],
        "lfdm_sampler_x0": [vp, vp, vp, vp, vp, i64, i32, vp],
        "lfdm_sampler_quantile": [vp, vp, i64, i32, i64, f32, vp, vp],
        "lfdm_sampler_update": [vp, vp, vp, vp, vp, vp, i32, vp, vp, i64, i32, vp],
        "lfdm_warp_blend_rows": [vp, vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
        "lfdm_warp_blend_image": [vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp],
        "lfdm_to_rows": [vp, i32, i32, i32, i32, i64, i64, i64, i32, vp, i64, vp, vp],
        "lfdm_from_rows": [vp, i32, i32, i32, i32, i32, vp, vp],
        "lfdm_im2col_small": [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp],
        "lfdm_pad_replicate_rows": [vp, i64, vp, i64, i32, i32, i32, i32, vp],
        "lfdm_avgpool2_rows": [vp, i32, i32, i32, i32, vp, vp, i64, vp],
        "lfdm_unet_heads": [vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp],
        "lfdm_unet_heads_cfg": [vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp],
        "lfdm_render_panels": [vp, vp, vp, vp, vp, C.POINTER(C.c_float), i32, i32, i32, i32, i32, f32, vp, vp, vp],
        "lfdm_split_bf16": [vp, vp, i64, i64, vp],
        "lfdm_antialias_down": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "lfdm_region_moments": [vp, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp],
        "lfdm_motion_prep": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
        "lfdm_motion_finish": [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp],
        "lfdm_rows_mean": [vp, i32, i32, i32, vp, vp],
        "lfdm_version": [C.POINTER(C.c_int), C.POINTER(C.c_int)],
    }
    for name, args in sig.items():
        fn = getattr(l, name)          # AttributeError here == a declared symbol is missing from the .so
        fn.argtypes = args
        fn.restype = C.c_int


EXPORTED = ["lfdm_conv", "lfdm_gn_stats", "lfdm_gn_apply", "lfdm_layernorm", "lfdm_attn_softmax", "lfdm_attn_softmax_pre",
            "lfdm_attn_linear", "lfdm_attn_temporal_fused", "lfdm_attn_linear_fused",
            "lfdm_small_linear", "lfdm_sinusoidal", "lfdm_ss_combine", "lfdm_sampler_x0", "lfdm_sampler_quantile",
            "lfdm_sampler_update", "lfdm_warp_blend_rows", "lfdm_warp_blend_image", "lfdm_to_rows", "lfdm_from_rows",
            "lfdm_im2col_small", "lfdm_pad_replicate_rows", "lfdm_avgpool2_rows", "lfdm_unet_heads", "lfdm_unet_heads_cfg", "lfdm_split_bf16", "lfdm_version",
            "lfdm_antialias_down", "lfdm_region_moments", "lfdm_motion_prep", "lfdm_motion_finish", "lfdm_rows_mean", "lfdm_render_panels"]

launch_count = 0   # number of C-ABI calls issued (bench.py reports kernels launched per step from this)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "kernel operands must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def check(rc, what):
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = {-1: "bad argument", -2: "unsupported shape for this engine", -3: "cuTensorMapEncodeTiled unavailable"}.get(
            rc, f"CUDA/driver error {rc}")
        raise LfdmError(f"{what} failed: {msg}")


class SB:
    """split-bf16 row matrix: tensor (2, M, C) bf16 (hi plane, lo plane)."""
    __slots__ = ("t", "m", "c")

    def __init__(self, m, c, device):
        self.t = torch.empty((2, m, c), dtype=torch.bfloat16, device=device)
        self.m, self.c = m, c

    @property
    def plane(self):
        return self.m * self.c

    def float(self):
        return self.t[0].float() + self.t[1].float()


def conv(desc_kwargs, engine):
    """desc_kwargs: dict of ConvDesc fields (tensors are converted to pointers)."""
    d = ConvDesc()
    for k, v in desc_kwargs.items():
        if k in ("a_sb", "a_f32"):
            arr = getattr(d, k)
            for i, t in enumerate(v):
                arr[i] = t.data_ptr() if t is not None else None
        elif k in ("a_plane", "a_c"):
            arr = getattr(d, k)
            for i, x in enumerate(v):
                arr[i] = int(x)
        elif isinstance(v, torch.Tensor):
            setattr(d, k, v.data_ptr())
        elif v is None:
            setattr(d, k, None)
        else:
            setattr(d, k, v)
    rc = lib().lfdm_conv(C.byref(d), engine, stream())
    return rc
