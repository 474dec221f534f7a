"""cvpr23_lfdm_b200 — B200-native (sm_100a) implementation of the LFDM sampling + LFAE decode hot path.

Public surface mirrors the reference (nihaomiao/CVPR23_LFDM) for this path; see DESIGN.md."""
from .dm.video_flow_diffusion import Unet3D, GaussianDiffusion
from .dm.video_flow_diffusion_model import FlowDiffusion
from .lfae.generator import Generator
from .lfae.region_predictor import RegionPredictor
from .lfae.bg_motion_predictor import BGMotionPredictor

__all__ = ["Unet3D", "GaussianDiffusion", "FlowDiffusion", "Generator", "RegionPredictor", "BGMotionPredictor"]
