"""Host-side mirror of reference LFAE/modules/bg_motion_predictor.py `BGMotionPredictor`."""
import torch
from torch import nn
from .util import Encoder, module_state_key


class BGMotionPredictor(nn.Module):
    def __init__(self, block_expansion, num_channels, max_features, num_blocks, bg_type='zero'):
        super().__init__()
        assert bg_type in ['zero', 'shift', 'affine', 'perspective']
        self.bg_type = bg_type
        if bg_type != 'zero':
            self.encoder = Encoder(block_expansion, in_features=num_channels * 2, max_features=max_features,
                                   num_blocks=num_blocks)
            feat = min(max_features, block_expansion * (2 ** num_blocks))
            n_out, init = {'perspective': (8, [1, 0, 0, 0, 1, 0, 0, 0]), 'affine': (6, [1, 0, 0, 0, 1, 0]),
                           'shift': (2, [0, 0])}[bg_type]
            self.fc = nn.Linear(feat, n_out)
            self.fc.weight.data.zero_()
            self.fc.bias.data.copy_(torch.tensor(init, dtype=torch.float))
        self._eng = None

    @torch.no_grad()
    def forward(self, source_image, driving_image):
        """reference bg_motion_predictor.py:42-57 -> (B,3,3)"""
        from ..engine.lfae_engine import BGPredictorEngine
        if source_image.device.type != "cuda":
            raise RuntimeError("BGMotionPredictor runs only on CUDA (sm_100a); no CPU fallback")
        key = module_state_key(self)
        if self._eng is None or self._eng_key != key:      # rebuilt after load_state_dict / .to() / in-place updates
            self._eng = BGPredictorEngine(self)
            self._eng_key = key
        return self._eng.forward(source_image, driving_image)
