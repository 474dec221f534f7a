"""Host-side mirror of reference LFAE/modules/generator.py `Generator`.
Same sub-module tree / state_dict keys / init order; compute_fea, forward_with_flow and forward
run on the sm_100a kernels through `engine.lfae_engine.GeneratorEngine`."""
import torch
from torch import nn
from .util import ResBlock2d, SameBlock2d, UpBlock2d, DownBlock2d, module_state_key
from .pixelwise_flow_predictor import PixelwiseFlowPredictor


class Generator(nn.Module):
    def __init__(self, num_channels, num_regions, block_expansion, max_features, num_down_blocks,
                 num_bottleneck_blocks, pixelwise_flow_predictor_params=None, skips=False, revert_axis_swap=True):
        super().__init__()
        self.pixelwise_flow_predictor = PixelwiseFlowPredictor(
            num_regions=num_regions, num_channels=num_channels, revert_axis_swap=revert_axis_swap,
            **pixelwise_flow_predictor_params) if pixelwise_flow_predictor_params is not None else None
        self.first = SameBlock2d(num_channels, block_expansion, kernel_size=(7, 7), padding=(3, 3))
        feat = lambda i: min(max_features, block_expansion * (2 ** i))
        self.down_blocks = nn.ModuleList([DownBlock2d(feat(i), feat(i + 1), kernel_size=(3, 3), padding=(1, 1))
                                          for i in range(num_down_blocks)])
        self.up_blocks = nn.ModuleList([UpBlock2d(feat(num_down_blocks - i), feat(num_down_blocks - i - 1),
                                                  kernel_size=(3, 3), padding=(1, 1)) for i in range(num_down_blocks)])
        self.bottleneck = torch.nn.Sequential()
        for i in range(num_bottleneck_blocks):
            self.bottleneck.add_module('r' + str(i), ResBlock2d(feat(num_down_blocks), kernel_size=(3, 3), padding=(1, 1)))
        self.final = nn.Conv2d(block_expansion, num_channels, kernel_size=(7, 7), padding=(3, 3))
        self.num_channels, self.skips = num_channels, skips
        self._eng = None

    def engine(self):
        from ..engine.lfae_engine import GeneratorEngine
        dev = self.final.weight.device
        if dev.type != "cuda":
            raise RuntimeError("cvpr23_lfdm_b200.Generator runs only on a CUDA (sm_100a) device; no CPU fallback")
        if self.training:
            raise NotImplementedError(
                "Generator is in train mode: batch-statistics BatchNorm is a training feature (out of scope); the "
                "reference samples with the LFAE in eval mode (video_flow_diffusion_model.py:44,53,60, demo_mug.py:105). "
                "Call .eval() first.")
        key = module_state_key(self)
        if self._eng is None or self._eng_key != key:
            self._eng = GeneratorEngine(self)
            self._eng_key = key
        return self._eng

    @torch.no_grad()
    def compute_fea(self, source_image):
        """reference generator.py:130-134 -> (B, 256, H/4, W/4)"""
        return self.engine().compute_fea(source_image)

    @torch.no_grad()
    def forward_with_flow(self, source_image, optical_flow, occlusion_map):
        """reference generator.py:136-166 -> {'deformed','prediction'}"""
        return self.engine().forward_with_flow(source_image, optical_flow, occlusion_map)

    @torch.no_grad()
    def decode_video(self, source_image, grid, conf):
        """Extension: all F frames in one batched pass with the frame-invariant encoder run once
        (the reference re-runs it per frame, video_flow_diffusion_model.py:206-214).
        grid (B,2,F,h,w), conf (B,1,F,h,w) -> (prediction, deformed) each (B,3,F,H,W)."""
        return self.engine().decode_video(source_image, grid, conf)

    @torch.no_grad()
    def forward_video(self, source_image, driving_region_params, source_region_params, bg_params, num_frames):
        """Extension: `forward` over `num_frames` driving frames per source image in one batch (driving params / bg_params with
        B*F rows, sample-major) -> prediction / deformed (B,3,F,H,W), optical_flow (B,F,h,w,2), occlusion_map (B,F,1,h,w)."""
        return self.engine().forward_video(source_image, driving_region_params, source_region_params, bg_params, num_frames)

    @torch.no_grad()
    def forward(self, source_image, driving_region_params, source_region_params, bg_params=None):
        """reference generator.py:90-128"""
        return self.engine().forward(source_image, driving_region_params, source_region_params, bg_params)
