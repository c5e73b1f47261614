"""Run one GEMM shape alone (for ncu). usage: python tools/run_gemm.py M N K epi(bf16|gelu|resid|qkv)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4]); epi = sys.argv[4] if len(sys.argv) > 4 else "bf16"
a = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16); bias = torch.randn(N, device="cuda")
res = torch.zeros(M, N, device="cuda")
def run():
    if epi == "qkv": return ops.gemm(a, w, bias, ops.EPI_QKV_HEADS, heads=N // 192, tokens=M)
    if epi == "resid": return ops.gemm(a, w, bias, ops.EPI_RESID_F32, residual=res)
    if epi == "gelu": return ops.gemm(a, w, bias, ops.EPI_GELU_BF16)
    return ops.gemm(a, w, bias, ops.EPI_BF16)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"gemm {M}x{N}x{K} {epi}: {e0.elapsed_time(e1):.4f} ms {2*M*N*K/e0.elapsed_time(e1)/1e9:.1f} TF/s")
