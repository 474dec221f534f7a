"""Parameter container for the dense-motion network (reference LFAE/modules/pixelwise_flow_predictor.py:17-46)."""
from torch import nn
from .util import Hourglass, AntiAliasInterpolation2d, _Container


class PixelwiseFlowPredictor(_Container):
    def __init__(self, block_expansion, num_blocks, max_features, num_regions, num_channels,
                 estimate_occlusion_map=False, scale_factor=1, region_var=0.01, use_covar_heatmap=False,
                 use_deformed_source=True, revert_axis_swap=False):
        super().__init__()
        self.hourglass = Hourglass(block_expansion=block_expansion,
                                   in_features=(num_regions + 1) * (num_channels * use_deformed_source + 1),
                                   max_features=max_features, num_blocks=num_blocks)
        self.mask = nn.Conv2d(self.hourglass.out_filters, num_regions + 1, kernel_size=(7, 7), padding=(3, 3))
        self.occlusion = nn.Conv2d(self.hourglass.out_filters, 1, kernel_size=(7, 7), padding=(3, 3)) \
            if estimate_occlusion_map else None
        self.num_regions, self.scale_factor, self.region_var = num_regions, scale_factor, region_var
        self.use_covar_heatmap, self.use_deformed_source = use_covar_heatmap, use_deformed_source
        self.revert_axis_swap = revert_axis_swap
        if self.scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, self.scale_factor)
