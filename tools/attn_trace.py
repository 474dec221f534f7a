#!/usr/bin/env python
"""Pipeline timeline of the fused temporal-attention kernel (CTA 0): run with LFDM_ATTN_TRACE=1.
Prints, per flat head index g, the clock64() stamps of every role's barrier crossings relative to the first stamp."""
import os
import sys
os.environ["LFDM_ATTN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cvpr23_lfdm_b200._lib import SB  # noqa: E402
from cvpr23_lfdm_b200.engine import ops  # noqa: E402

B, F, hw = 8, 40, 32
dev = torch.device("cuda:0")
p, heads, ch = hw * hw, 8, 64
m = B * F * p
x = torch.randn(m, ch, device=dev)
gamma = torch.randn(ch, device=dev)
wqkv, wout = torch.randn(768, ch, 1, 1, device=dev) / 8, torch.randn(ch, 256, 1, 1, device=dev) / 16
ang = torch.outer(torch.arange(F, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32)))
cs, sn, bias = ang.cos().contiguous(), ang.sin().contiguous(), torch.randn(heads, F, F, device=dev)
out, out_sb = torch.empty(m, ch, device=dev), SB(m, ch, dev)
wq_img, wo_img = ops.pack_fused_attention(wqkv, wout, heads)
trace = torch.zeros(32 * 64, dtype=torch.int64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for _ in range(3):
    trace.zero_()
    if "--cold" in sys.argv:
        flush.fill_(0.0)
    ops.attn_temporal_fused(x, gamma, wq_img, wo_img, None, cs, sn, bias, out, out_sb, B, F, p, heads, 1e-5, debug=trace)
torch.cuda.synchronize()
t = trace.cpu().reshape(32, 64)
t0 = int(t[t > 0].min())
names = {1: "M.qkv", 8: "A.enter", 9: "A.qkvfull", 10: "A.sfull(g-1)", 11: "A.done", 3: "M.qk",
         12: "C.enter", 13: "C.sfull", 14: "C.computed", 15: "C.Pfree", 16: "C.done", 5: "M.pv",
         17: "B.v.enter", 18: "B.qkvfull", 19: "B.v.done", 20: "B.pvdfull", 21: "B.o.done", 7: "M.out",
         23: "E.done(it)", 24: "LN.ready(it)", 25: "LN.xnempty(it)", 26: "LN.done(it)"}
order = [1, 8, 9, 10, 11, 3, 17, 18, 19, 12, 13, 14, 15, 16, 5, 20, 21, 7, 23, 24, 25, 26]
print("last stamp", int(t.max()) - t0)
print("g " + " ".join(f"{names[s]:>14s}" for s in order))
for g in range(24):
    print(f"{g:2d} " + " ".join(f"{(int(t[s, g]) - t0) if t[s, g] > 0 else -1:14d}" for s in order))
