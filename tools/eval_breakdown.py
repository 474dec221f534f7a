#!/usr/bin/env python
"""One UNet evaluation at the bench geometry (MUG-128, batch 8): CUDA-event time of every conv launch (ops.PROFILE),
aggregated by layer shape.  Wrap it in `ncu --metrics gpu__time_duration.sum` for the full per-kernel launch list of
exactly one evaluation (see profiles/README.md)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cvpr23_lfdm_b200 as P  # noqa: E402
from cvpr23_lfdm_b200.engine import ops  # noqa: E402

B, FRAMES = int(os.environ.get("B", 8)), 40
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = P.FlowDiffusion(is_train=False, sampling_timesteps=1000, img_size=32, num_frames=FRAMES, timesteps=1000,
                        config_pth=os.path.join(ROOT, "config", "mug128.yaml"), pretrained_pth="").to(dev).eval()
eng = model.unet.engine()
img = torch.rand(B, 3, 128, 128, device=dev)
cond = torch.randn(B, 768, device=dev)
fea_conv = eng.prepare_fea(model.generator.compute_fea(img))
x = torch.randn(B, 3, FRAMES, 32, 32, device=dev)
ss = eng.scale_shift(torch.full((B,), 500, device=dev, dtype=torch.long), cond)
for _ in range(2):
    eng.forward_hoisted(x, fea_conv, ss)
torch.cuda.synchronize()
ops.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(40e-3 * 1.9e9))
e0.record()
eng.forward_hoisted(x, fea_conv, ss)
e1.record()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for name, engine, flops, a, b, _ in prof:
    k = (name, engine, round(flops / 1e9, 2))
    t = a.elapsed_time(b)
    n, s = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, s + t)
tot = sum(s for _, s in agg.values())
print(f"# one evaluation: {e0.elapsed_time(e1):.3f} ms, conv launches {len(prof)}, conv total {tot:.3f} ms")
print("| layer (shape key) | engine | GFLOP | launches | total ms | avg us | TFLOP/s |")
print("|---|---|---:|---:|---:|---:|---:|")
for (name, engine, gf), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {engine} | {gf} | {n} | {s:.3f} | {1e3 * s / n:.1f} | {gf * n / s:.1f} |")
