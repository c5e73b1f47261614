"""The patch-embedding kernel alone at 512x512x320 -> 20480 x 768 (for ncu).  usage: python tools/run_patch.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
vol = torch.rand(1, 320, 512, 512, device="cuda")
w = (torch.randn(768, 4096, device="cuda") * 0.02).bfloat16()
b = torch.randn(768, device="cuda")
pos = ops.sincos_table(20480, 768, "cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    out = ops.patch_embed_fwd(vol, w, b, pos)
e1.record()
torch.cuda.synchronize()
print(f"patch_embed 512x512x320: {e0.elapsed_time(e1):.4f} ms")
