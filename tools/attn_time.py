"""forward + backward attention timings at the model shapes.  usage: python tools/attn_time.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
tag = sys.argv[1] if len(sys.argv) > 1 else ""
def timeit(fn, iters=8, warmup=3):
    for _ in range(warmup): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for B, H, N in [(1, 12, 7168), (1, 6, 20480), (1, 12, 20480), (4, 12, 1960), (1, 16, 9216), (1, 16, 2688)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(B, H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
    dout = torch.randn(B, N, H * 64, device="cuda").to(torch.bfloat16)
    o, lse = ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
    tf = timeit(lambda: ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True))
    tb = timeit(lambda: ops.flash_attn_bwd(q, k, v, o, dout, lse, 0.125))
    print(tag, f"B{B} H{H} N{N}: fwd {tf:.3f} ms ({4.0*N*N*64*H*B/tf/1e9:.0f} TF/s)  bwd {tb:.3f} ms ({8.0*N*N*64*H*B/tb/1e9:.0f} TF/s of 8N^2dH)", flush=True)
