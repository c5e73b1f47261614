"""For the record only (never on the product path): torch's own SDPA (cuDNN / flash backend) forward and backward at the model's
attention shapes on this box, next to our kernels.  usage: python tools/sdpa_reference_timing.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from smb_vision_b200 import ops
from tools.gpu_check import timeit

for H, N in [(12, 20480), (6, 20480), (12, 7168)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, H, N, 64, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
    do = torch.randn(1, H, N, 64, device="cuda", dtype=torch.bfloat16)
    fwd = timeit(lambda: F.scaled_dot_product_attention(q, k, v), iters=10)
    o = F.scaled_dot_product_attention(q, k, v)
    def fb():
        q.grad = k.grad = v.grad = None
        F.scaled_dot_product_attention(q, k, v).backward(do)
    both = timeit(fb, iters=10)
    qd, kd, vd = q.detach(), k.detach(), v.detach()
    out, lse = ops.flash_attn_fwd(qd, kd, vd, 0.125, return_lse=True)
    dout = do.transpose(1, 2).reshape(1, N, H * 64).contiguous()
    ours_f = timeit(lambda: ops.flash_attn_fwd(qd, kd, vd, 0.125, return_lse=True), iters=10)
    ours_b = timeit(lambda: ops.flash_attn_bwd(qd, kd, vd, out, dout, lse, 0.125), iters=10)
    fl = 4.0 * N * N * 64 * H
    print(f"H={H} N={N}: torch sdpa fwd {fwd:.3f} ms ({fl/fwd/1e9:.0f} TF/s), bwd {both-fwd:.3f} ms ({2.5*fl/(both-fwd)/1e9:.0f} TF/s-eq) | "
          f"ours fwd {ours_f:.3f} ms ({fl/ours_f/1e9:.0f}), bwd {ours_b:.3f} ms ({2.5*fl/ours_b/1e9:.0f})", flush=True)
