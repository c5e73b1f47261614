"""forward attention: timing (interleaved repeats) + parity vs fp32 at a small shape.  usage: python tools/attn_fwd_ab.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
tag = sys.argv[1] if len(sys.argv) > 1 else ""
torch.manual_seed(1)
q, k, v = (torch.randn(2, 3, 1000, 64, device="cuda").bfloat16() for _ in range(3))
o, lse = ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
p = torch.softmax(q.float() @ k.float().transpose(-1, -2) * 0.125, -1)
ref = (p @ v.float()).transpose(1, 2).reshape(2, 1000, 192)
err = ((o.float() - ref).norm() / ref.norm()).item()
res, data = {}, {}
for H, N in [(12, 7168), (6, 20480), (12, 20480)]:
    data[(H, N)] = tuple(torch.randn(1, H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
for rep in range(2):
    for (H, N), (q, k, v) in data.items():
        for _ in range(3): ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
        e1.record(); torch.cuda.synchronize()
        res.setdefault((H, N), []).append(e0.elapsed_time(e1) / 10)
print(tag, f"err {err:.2e}", {f"H{H}_N{N}": [round(t, 4) for t in ts] for (H, N), ts in res.items()}, flush=True)
