"""Time the two backward kernels separately via CUDA events around the C call with knock-out variants (SMBV_DQ_KNOCK)."""
import os, sys
os.environ.setdefault("SMBV_DEV_HOOKS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
H, N = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
q, k, v = (torch.randn(H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
dout = torch.randn(N, H * 64, device="cuda").to(torch.bfloat16)
o, lse = ops.flash_attn_fwd(q[None], k[None], v[None], 0.125, return_lse=True)
for _ in range(3):
    ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125)
e1.record()
torch.cuda.synchronize()
print(f"knock={os.environ.get('SMBV_DQ_KNOCK','0')} H={H} N={N}: dkdv+dq {e0.elapsed_time(e1)/5:.3f} ms")
