"""Small-shape launches of the round-2 kernels for compute-sanitizer (memcheck): fused attention backward incl. a split last wave and a
ragged N, the k-way split forward + combine, the TMA GELU / dGELU epilogues, the D prep kernel.
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
dev = "cuda"
torch.manual_seed(0)
for B, H, N in [(1, 1, 130), (2, 2, 216), (1, 150, 300), (1, 75, 1000)]:   # 150 x 3 = 450 units, 75 x 8 = 600 units: split last waves
    q, k, v = (torch.randn(B, H, N, 64, device=dev).bfloat16() for _ in range(3))
    dout = torch.randn(B, N, H * 64, device=dev).bfloat16()
    out, lse = ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
    dq, dk, dv = ops.flash_attn_bwd(q, k, v, out, dout, lse, 0.125, deterministic=False)
    d2 = ops.flash_attn_bwd(q, k, v, out, dout, lse, 0.125, deterministic=True)
    torch.cuda.synchronize()
    e = [((a.float() - b.float()).norm() / b.float().norm()).item() for a, b in zip((dq, dk, dv), d2)]
    print(f"attn B{B} H{H} N{N}: fused vs two-kernel {e[0]:.1e} {e[1]:.1e} {e[2]:.1e}", flush=True)
M, N, K = 1000, 512, 256
a, w, bias = torch.randn(M, K, device=dev).bfloat16(), (torch.randn(N, K, device=dev) * 0.05).bfloat16(), torch.randn(N, device=dev)
pre = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
h = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
ops.gemm_ex(a, w, M, N, K, ops.EPI_GELU_BF16, h, bias=bias, aux=pre)             # GELU + saved pre-activation (two slab stores)
ref = a.float() @ w.float().t() + bias
print("gelu+pre", ((pre.float() - ref).norm() / ref.norm()).item(), ((h.float() - torch.nn.functional.gelu(ref)).norm() / ref.norm()).item(), flush=True)
dy = torch.randn(M, K, device=dev).bfloat16()
dh = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
ops.gemm_ex(dy, w, M, N, K, ops.EPI_DGELU_BF16, dh, w_layout=0, aux=pre)  # dGELU: pre-activation slab loaded by TMA
x = pre.float().requires_grad_(True)
torch.nn.functional.gelu(x).backward(dy.float() @ w.float().t())
print("dgelu", ((dh.float() - x.grad).norm() / x.grad.norm()).item(), flush=True)
torch.cuda.synchronize()
print("done")
