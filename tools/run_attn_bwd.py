"""Run the flash-attention backward kernel alone (for ncu).  usage: python tools/run_attn_bwd.py [H] [N] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops

H = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20480
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
q, k, v = (torch.randn(H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
dout = torch.randn(N, H * 64, device="cuda").to(torch.bfloat16)
o, lse = ops.flash_attn_fwd(q[None], k[None], v[None], 0.125, return_lse=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"attn bwd H={H} N={N}: {ms:.3f} ms, {10.0*N*N*64*H/ms/1e9:.1f} TFLOP/s")
