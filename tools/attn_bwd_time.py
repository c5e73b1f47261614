"""time the attention backward only (fused unless SMBV_ATTN_BWD_DETERMINISTIC=1).  usage: python tools/attn_bwd_time.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
for H, N in [(6, 20480), (12, 7168)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
    dout = torch.randn(N, H * 64, device="cuda").to(torch.bfloat16)
    o, lse = ops.flash_attn_fwd(q[None], k[None], v[None], 0.125, return_lse=True)
    for _ in range(2):
        ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125)
    e1.record()
    torch.cuda.synchronize()
    print(sys.argv[1] if len(sys.argv) > 1 else "", f"H{H} N{N}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
