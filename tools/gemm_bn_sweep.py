"""Time the forward GEMM shapes of the model with the N tile forced (SMBV_GEMM_BN=128|256) or chosen by the dispatcher.
usage: SMBV_GEMM_BN=128 python tools/gemm_bn_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
from tools.gpu_check import timeit

dev = "cuda"
tag = os.environ.get("SMBV_GEMM_BN", "auto")
for M, N, K, epi in [(20480, 2304, 768, "qkv"), (20480, 768, 768, "resid"), (20480, 3072, 768, "gelu"), (20480, 768, 3072, "resid"),
                     (7168, 2304, 768, "qkv"), (7168, 768, 768, "resid"), (7168, 3072, 768, "gelu"), (7168, 768, 3072, "resid"),
                     (20480, 1152, 384, "qkv"), (20480, 384, 384, "resid"), (20480, 1536, 384, "gelu"), (20480, 384, 1536, "resid"),
                     (13312, 4096, 384, "bf16")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    res = torch.zeros(M, N, device=dev) if epi == "resid" else None
    if epi == "qkv":
        fn = lambda: ops.gemm(a, w, bias, ops.EPI_QKV_HEADS, heads=N // 192, tokens=M)
    elif epi == "resid":
        fn = lambda: ops.gemm(a, w, bias, ops.EPI_RESID_F32, residual=res)
    elif epi == "gelu":
        fn = lambda: ops.gemm(a, w, bias, ops.EPI_GELU_BF16)
    else:
        fn = lambda: ops.gemm(a, w, bias, ops.EPI_BF16)
    ms = timeit(fn, iters=20)
    print(f"bn={tag} {M}x{N}x{K} {epi}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
