"""Output stage of the demo scripts (demo/demo_mug.py:126-145), B200 side.

`render_panels` composes the uint8 5-panel frames [source | generated | warped | sampling-grid figure | confidence] of one
sample on the GPU (csrc/render.cu through the C-ABI `lfdm_render_panels`); `AsyncGifWriter` moves them to pinned host memory
on a side stream and encodes the GIF on a worker thread, so the next `sample_one_video` call overlaps the encode.  The
reference does all of this per frame on the host (numpy + a matplotlib figure + PIL paste, ~0.1 s per frame)."""
import ctypes as C
import queue
import threading
import torch
from ._lib import lib, ptr, stream, check

MUG_MEAN = (0.0, 0.0, 0.0)      # demo/demo_mug.py:28 MEAN


def render_panels(src_img, out_vid, warped_vid, vid_grid, vid_conf, index=0, mean=MUG_MEAN, line_width=1.0):
    """src_img (B,3,H,W); out_vid / warped_vid (B,3,F,H,W); vid_grid (B,2,F,h,w); vid_conf (B,1,F,h,w) fp32 CUDA tensors
    -> uint8 CUDA tensor (F, H, 5W, 3): the frames demo_mug.py pastes together for sample `index`.
    mean: per-channel mean in 0..255 units that sample_img adds back (demo MEAN)."""
    for t in (src_img, out_vid, warped_vid, vid_grid, vid_conf):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("render_panels works on fp32 CUDA tensors (the attributes FlowDiffusion.sample_one_video sets)")
    _, _, f, H, W = out_vid.shape
    h, w = vid_grid.shape[-2:]
    sl = lambda t: t[index].contiguous()
    src, ov, wv, g, cf = sl(src_img), sl(out_vid), sl(warped_vid), sl(vid_grid), sl(vid_conf)
    ws = torch.empty((((f * 4 + 63) // 64) * 64 + f * 2 * H * W,), dtype=torch.int32, device=out_vid.device)
    out = torch.empty((f, H, 5 * W, 3), dtype=torch.uint8, device=out_vid.device)
    m3 = (C.c_float * 3)(*[float(m) / 255.0 for m in mean])
    check(lib().lfdm_render_panels(ptr(src), ptr(ov), ptr(wv), ptr(g), ptr(cf), m3, f, H, W, h, w, float(line_width),
                                   ptr(ws), ptr(out), stream()), "lfdm_render_panels")
    return out


class AsyncGifWriter:
    """frames (F, H, W5, 3) uint8 CUDA tensor -> GIF file, without blocking the sampling stream:
    D2H copy into pinned memory on a side stream, encode (PIL) on a worker thread."""

    def __init__(self, duration_ms=100):
        self.duration_ms = duration_ms
        self._q = queue.Queue()
        self._copy_stream = None
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()
        self.errors = []

    def submit(self, frames, path):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=frames.device)
        host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
        ready = torch.cuda.Event()
        self._copy_stream.wait_stream(torch.cuda.current_stream(frames.device))
        with torch.cuda.stream(self._copy_stream):
            host.copy_(frames, non_blocking=True)
            frames.record_stream(self._copy_stream)
            ready.record(self._copy_stream)
        self._q.put((host, ready, path))

    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            host, ready, path = item
            try:
                ready.synchronize()
                from PIL import Image
                imgs = [Image.fromarray(host[i].numpy(), "RGB") for i in range(host.shape[0])]
                imgs[0].save(path, save_all=True, append_images=imgs[1:], duration=self.duration_ms, loop=0)
            except Exception as e:          # surfaced by close()
                self.errors.append((path, e))
            finally:
                self._q.task_done()

    def close(self):
        self._q.join()
        self._q.put(None)
        self._thread.join()
        if self.errors:
            raise RuntimeError(f"GIF encoding failed: {self.errors}")
