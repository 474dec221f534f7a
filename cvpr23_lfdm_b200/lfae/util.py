"""Parameter containers for the LFAE building blocks (reference LFAE/modules/util.py:70-264).
BatchNorm is always evaluated in inference mode (the reference's SynchronizedBatchNorm2d reduces to
F.batch_norm in eval, sync_batchnorm/batchnorm.py:50-53); the engine folds it into the adjacent conv."""
import torch
from torch import nn

BatchNorm2d = nn.BatchNorm2d  # same state_dict keys as SynchronizedBatchNorm2d


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError(f"{type(self).__name__}: parameter container; math runs in the CUDA engine")


class ResBlock2d(_Container):
    def __init__(self, in_features, kernel_size, padding):
        super().__init__()
        self.conv1 = nn.Conv2d(in_features, in_features, kernel_size=kernel_size, padding=padding)
        self.conv2 = nn.Conv2d(in_features, in_features, kernel_size=kernel_size, padding=padding)
        self.norm1 = BatchNorm2d(in_features, affine=True)
        self.norm2 = BatchNorm2d(in_features, affine=True)


class _ConvNorm(_Container):
    def __init__(self, in_features, out_features, kernel_size=3, padding=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(in_features, out_features, kernel_size=kernel_size, padding=padding, groups=groups)
        self.norm = BatchNorm2d(out_features, affine=True)


class UpBlock2d(_ConvNorm):
    pass


class SameBlock2d(_Container):
    def __init__(self, in_features, out_features, groups=1, kernel_size=3, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(in_features, out_features, kernel_size=kernel_size, padding=padding, groups=groups)
        self.norm = BatchNorm2d(out_features, affine=True)


class DownBlock2d(_ConvNorm):
    def __init__(self, in_features, out_features, kernel_size=3, padding=1, groups=1):
        super().__init__(in_features, out_features, kernel_size, padding, groups)
        self.pool = nn.AvgPool2d(kernel_size=(2, 2))


class Encoder(_Container):
    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        self.down_blocks = nn.ModuleList([
            DownBlock2d(in_features if i == 0 else min(max_features, block_expansion * (2 ** i)),
                        min(max_features, block_expansion * (2 ** (i + 1))), kernel_size=3, padding=1)
            for i in range(num_blocks)])


class Decoder(_Container):
    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        ups = []
        for i in range(num_blocks)[::-1]:
            cin = (1 if i == num_blocks - 1 else 2) * min(max_features, block_expansion * (2 ** (i + 1)))
            ups.append(UpBlock2d(cin, min(max_features, block_expansion * (2 ** i)), kernel_size=3, padding=1))
        self.up_blocks = nn.ModuleList(ups)
        self.out_filters = block_expansion + in_features


class Hourglass(_Container):
    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        self.encoder = Encoder(block_expansion, in_features, num_blocks, max_features)
        self.decoder = Decoder(block_expansion, in_features, num_blocks, max_features)
        self.out_filters = self.decoder.out_filters


class AntiAliasInterpolation2d(_Container):
    """Gaussian band-limit + integer subsample (reference util.py:217-264); buffer `weight`."""

    def __init__(self, channels, scale):
        super().__init__()
        sigma = (1 / scale - 1) / 2
        ks = 2 * round(sigma * 4) + 1
        self.ka = ks // 2
        self.kb = self.ka - 1 if ks % 2 == 0 else self.ka
        g = torch.arange(ks, dtype=torch.float32)
        mean = (ks - 1) / 2
        k1 = torch.exp(-(g - mean) ** 2 / (2 * sigma ** 2))
        kern = k1[:, None] * k1[None, :]
        kern = kern / torch.sum(kern)
        self.register_buffer('weight', kern.view(1, 1, ks, ks).repeat(channels, 1, 1, 1))
        self.groups, self.scale, self.int_inv_scale = channels, scale, int(1 / scale)


def make_coordinate_grid(spatial_size, type=None):
    h, w = spatial_size
    x = 2 * (torch.arange(w).float() / (w - 1)) - 1
    y = 2 * (torch.arange(h).float() / (h - 1)) - 1
    g = torch.stack([x[None, :].repeat(h, 1), y[:, None].repeat(1, w)], 2)
    return g.type(type) if type is not None else g


def module_state_key(mod):
    """Identity of everything a packed engine was built from: device, version counter and storage of every parameter
    AND buffer (BatchNorm running statistics are buffers).  In-place updates (`load_state_dict` through any parent
    module, `copy_`, optimiser steps) bump `_version`; `.to()` / `.cuda()` change `data_ptr`.  Engines are rebuilt
    when the key changes (same scheme as Unet3D.engine)."""
    ts = list(mod.parameters()) + list(mod.buffers())
    return tuple((t.device, t._version, t.data_ptr()) for t in ts)
