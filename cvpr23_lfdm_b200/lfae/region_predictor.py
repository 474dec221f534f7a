"""Host-side mirror of reference LFAE/modules/region_predictor.py `RegionPredictor`."""
import torch
from torch import nn
from .util import Hourglass, AntiAliasInterpolation2d, module_state_key


class RegionPredictor(nn.Module):
    def __init__(self, block_expansion, num_regions, num_channels, max_features, num_blocks, temperature,
                 estimate_affine=False, scale_factor=1, pca_based=False, fast_svd=False, pad=3):
        super().__init__()
        self.predictor = Hourglass(block_expansion, in_features=num_channels, max_features=max_features,
                                   num_blocks=num_blocks)
        self.regions = nn.Conv2d(self.predictor.out_filters, num_regions, kernel_size=(7, 7), padding=pad)
        if estimate_affine and not pca_based:
            self.jacobian = nn.Conv2d(self.predictor.out_filters, 4, kernel_size=(7, 7), padding=pad)
            self.jacobian.weight.data.zero_()
            self.jacobian.bias.data.copy_(torch.tensor([1, 0, 0, 1], dtype=torch.float))
        else:
            self.jacobian = None
        self.temperature, self.scale_factor, self.pca_based, self.fast_svd, self.pad = \
            temperature, scale_factor, pca_based, fast_svd, pad
        if self.scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, self.scale_factor)
        self._eng = None

    @torch.no_grad()
    def forward(self, x):
        """reference region_predictor.py:77-117 -> dict(shift, covar, heatmap, affine, u, d)"""
        from ..engine.lfae_engine import RegionPredictorEngine
        if x.device.type != "cuda":
            raise RuntimeError("RegionPredictor runs only on CUDA (sm_100a); no CPU fallback")
        key = module_state_key(self)
        if self._eng is None or self._eng_key != key:      # rebuilt after load_state_dict / .to() / in-place updates
            self._eng = RegionPredictorEngine(self)
            self._eng_key = key
        return self._eng.forward(x)
