"""Launches every hot kernel ONCE at its benchmark shape, in a fixed order, for one `ncu --set full` capture:

    python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'flash_attn_(fwd2|bwd_)|gemm_bf16|patch_embed|normpix_loss|layernorm_fwd|rope3d|adamw' \
        -o gpurun_out/prof_r02_kernels python tools/ncu_targets.py

The order of the (kernel-name-matching) launches is written to gpurun_out/ncu_targets_order.json; tools/summarize_ncu.py joins
it with the report (launch i of the report = entry i of the list) and writes profiles/<tag>_ncu_kernels.md / _ncu_traffic.json.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle.mim_mask import OracleMaskGenerator  # mask restatement only (synthetic input)
from smb_vision_b200 import ops

dev = "cuda"
torch.manual_seed(0)
order = []


def bf(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(torch.bfloat16)


def note(kernel, shape, alg_bytes=None, alg_flops=None):
    order.append(dict(kernel=kernel, shape=shape, algorithmic_bytes=alg_bytes, algorithmic_flops=alg_flops))


# ---- attention forward / backward at the three model shapes
for H, N in [(12, 20480), (6, 20480), (12, 7168)]:
    q, k, v = bf(1, H, N, 64), bf(1, H, N, 64), bf(1, H, N, 64)
    o, lse = ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
    note("flash_attn_fwd2_kernel", f"H={H} N={N} d=64", 4 * H * N * 64 * 2, 4.0 * N * N * 64 * H)
    if (H, N) != (12, 20480):  # the two shapes of the training step
        do = bf(1, N, H * 64)
        ops.flash_attn_bwd(q, k, v, o, do, lse, 0.125, deterministic=False)
        # Q, K, V, dO read + dK, dV written (bf16) + the fp32 dQ accumulator written once; 8 N^2 d H = the four products of the stored-P backward
        note("flash_attn_bwd_fused_kernel", f"H={H} N={N} d=64", 6 * H * N * 64 * 2 + H * N * 64 * 4, 8.0 * N * N * 64 * H)
        ops.flash_attn_bwd(q, k, v, o, do, lse, 0.125, deterministic=True)
        note("flash_attn_bwd_dkdv_kernel", f"H={H} N={N} d=64", 6 * H * N * 64 * 2, 6.0 * N * N * 64 * H)
        note("flash_attn_bwd_dq_kernel", f"H={H} N={N} d=64", 5 * H * N * 64 * 2, 2.0 * N * N * 64 * H)
    del q, k, v, o, lse
torch.cuda.synchronize()

# ---- GEMMs with their fused epilogues (forward shapes of the encoder / decoder)
for M, N, K, epi in [(20480, 2304, 768, "qkv"), (20480, 768, 768, "resid"), (20480, 3072, 768, "gelu"), (20480, 768, 3072, "resid"),
                     (13312, 4096, 384, "bf16"), (7168, 3072, 768, "gelu"), (20480, 1536, 384, "gelu")]:
    a, w, bias = bf(M, K), bf(N, K, s=0.05), torch.randn(N, device=dev)
    if epi == "qkv":
        ops.gemm(a, w, bias, ops.EPI_QKV_HEADS, heads=N // 192, tokens=M)
        out_b = 2
    elif epi == "resid":
        ops.gemm(a, w, bias, ops.EPI_RESID_F32, residual=torch.zeros(M, N, device=dev))
        out_b = 4  # TMA reduce-add: the SM writes fp32, the L2 does the read-modify-write
    elif epi == "gelu":
        ops.gemm(a, w, bias, ops.EPI_GELU_BF16)
        out_b = 2
    else:
        ops.gemm(a, w, bias, ops.EPI_BF16)
        out_b = 2
    note("gemm_bf16_kernel", f"M={M} N={N} K={K} epilogue={epi}", (M * K + N * K) * 2 + M * N * out_b, 2.0 * M * N * K)
# ---- the two-transfer epilogues of the training step: fc1 with GELU + saved pre-activation, fc2 dgrad with dGELU (encoder / decoder shape)
for M, d, m in [(7168, 768, 3072), (20480, 384, 1536)]:
    x, w1, w2, bm = bf(M, d), bf(m, d, s=0.05), bf(d, m, s=0.05), torch.randn(m, device=dev)
    pre, f = torch.empty(M, m, device=dev, dtype=torch.bfloat16), torch.empty(M, m, device=dev, dtype=torch.bfloat16)
    ops.gemm_ex(x, w1, M, m, d, ops.EPI_GELU_BF16, f, bias=bm, aux=pre)
    note("gemm_bf16_kernel", f"M={M} N={m} K={d} epilogue=gelu+saved pre-activation", (M * d + m * d) * 2 + 2 * M * m * 2, 2.0 * M * m * d)
    ops.linear_dgrad(x, w2, aux=pre)
    note("gemm_bf16_kernel", f"M={M} N={m} K={d} dgrad, epilogue=dgelu", (M * d + m * d) * 2 + 2 * M * m * 2, 2.0 * M * m * d)
torch.cuda.synchronize()

# ---- patch embedding over the whole volume, loss, LayerNorm
vol = torch.rand(1, 320, 512, 512, device=dev)
wpe, bpe = (torch.randn(768, 4096, device=dev) * 0.02).bfloat16(), torch.randn(768, device=dev)
pos = ops.sincos_table(20480, 768, dev)
ops.patch_embed_fwd(vol, wpe, bpe, pos)
note("patch_embed_kernel", "512x512x320 fp32 volume -> 20480 x 768", vol.numel() * 4 + 768 * 4096 * 2 + 20480 * 768 * 4 * 2, 2.0 * 20480 * 4096 * 768)
np.random.seed(0)
mask = torch.from_numpy(OracleMaskGenerator(512, 320, 32, 16, 0.65)())[None]
_, midx, _, _ = ops.mask_index(mask.to(torch.uint8).to(dev))
logits = bf(1, 13312, 4096, s=0.3)
ops.normpix_loss(vol, midx, 13312, logits, True, 0)
note("normpix_loss", "13312 masked patches x 4096 voxels, loss + dlogits", 13312 * 4096 * (4 + 2 + 2))
x = torch.randn(20480, 768, device=dev)
ops.layernorm_fwd(x, torch.ones(768, device=dev), torch.zeros(768, device=dev), 1e-12)
note("layernorm_fwd_kernel", "20480 x 768 fp32 -> bf16", 20480 * 768 * 6)
torch.cuda.synchronize()

# ---- V-JEPA rotary kernel at ViT-L, fused AdamW over 97 M parameters
qk = bf(2, 1, 16, 20480, 64)
ops.rope3d_(qk, 32, max_pos=32)
note("rope3d", "ViT-L Q,K: 2 x 16 heads x 20480 x 64 bf16 in place", qk.numel() * 4)
torch.cuda.synchronize()
del qk, vol, logits
from smb_vision_b200 import _lib
import ctypes as C

n = 97161088 // 4 * 4
p, g, m, v2 = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v2.abs_()
pb = torch.empty(n, dtype=torch.bfloat16, device=dev)
starts, flags = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.uint8, device=dev)
_lib.call("smbv_adamw_step", ops._ptr(p), ops._ptr(pb), ops._ptr(g), ops._ptr(m), ops._ptr(v2), n, ops._ptr(starts), ops._ptr(flags), 1,
          5e-5, 0.9, 0.999, 1e-8, 0.01, 1, C.c_void_p(0), 0.0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
note("adamw_kernel", "97.16 M parameters", n * 30)
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(order, open("gpurun_out/ncu_targets_order.json", "w"), indent=1)
print(f"{len(order)} target launches")
