"""Small host-side helpers the reference demos import from `misc` (Logger, resize, conf2fig, grid2fig).
Rendering is outside the hot path; heavy optional deps (matplotlib, cv2) are imported lazily."""
import sys
import numpy as np
import torch


class Logger(object):
    """tee of a stream into a file (reference misc.py:83-93)"""

    def __init__(self, filename='default.log', stream=sys.stdout):
        self.terminal, self.log = stream, open(filename, 'w')

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        pass


def resize(im, desired_size, interpolation):
    """aspect-preserving resize + centred zero padding to a square (reference misc.py:96-110)"""
    import cv2
    h, w = im.shape[:2]
    ratio = float(desired_size) / max(h, w)
    nh, nw = int(h * ratio), int(w * ratio)
    im = cv2.resize(im, (nw, nh), interpolation=interpolation)
    dh, dw = desired_size - nh, desired_size - nw
    return cv2.copyMakeBorder(im, dh // 2, dh - dh // 2, dw // 2, dw - dw // 2, cv2.BORDER_CONSTANT, value=[0, 0, 0])


def conf2fig(conf, img_size=128):
    """(1,h,w) occlusion map -> uint8 (img_size, img_size) (reference misc.py:76-80: nearest upsample)"""
    c = torch.nn.functional.interpolate(conf.unsqueeze(0), size=img_size)[0, 0]
    return np.array(c.detach().cpu().numpy() * 255, dtype=np.uint8)


def grid2fig(warped_grid, grid_size=32, img_size=256):
    """draws the warped sampling grid over the identity grid; needs matplotlib (reference misc.py:44-63)"""
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    lin = np.linspace(-1, 1, grid_size)
    idx, idy = np.meshgrid(lin, lin)
    fig, ax = plt.subplots()
    for gx, gy, col in ((idx, idy, "lightgrey"), (warped_grid[..., 0], warped_grid[..., 1], "C0")):
        ax.plot(gx, gy, color=col, linewidth=0.8)
        ax.plot(np.transpose(gx), np.transpose(gy), color=col, linewidth=0.8)
    ax.axis("off")
    fig.tight_layout(pad=0)
    fig.set_size_inches(img_size / 100, img_size / 100)
    fig.set_dpi(100)
    fig.canvas.draw()
    out = np.asarray(fig.canvas.buffer_rgba())[:, :, :3].copy()
    plt.close(fig)
    return out
