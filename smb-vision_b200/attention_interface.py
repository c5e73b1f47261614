"""Operator-level plug-in: the sm_100a flash-attention kernels behind transformers' ``AttentionInterface`` registry.

This is the one plug-in point the reference itself dispatches through
(``ALL_ATTENTION_FUNCTIONS[config._attn_implementation]``, reference modeling_videomae.py:270-289):

    import smb_vision_b200.attention_interface as ai
    ai.register()                                   # AttentionInterface.register("b200_tcgen05", ...)
    config._attn_implementation = "b200_tcgen05"      # the UNMODIFIED reference / upstream model now runs the tcgen05 kernel

Contract (probed on the unmodified reference model): ``fn(module, query [B,H,N,64], key, value, attention_mask=None, *,
is_causal=False, scaling=0.125, dropout=0.0) -> (attn_output [B,N,H,64] contiguous, None)``.  Forward and backward run
through the C ABI (``smbv_flash_attn_fwd_ex`` / ``smbv_flash_attn_bwd``); there is no fallback: unsupported arguments raise.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import SmbvError

# transformers routes every implementation name containing "flash" to its own flash-attention loader, so the name avoids it
NAME = "b200_tcgen05"
# head_dim 16 / 32 with at least this many tokens run on the tcgen05 kernels with Q, K, V zero-padded to 64 (scores and
# outputs are unchanged; twice the tensor flops of a native head_dim-32 kernel, still far faster than the fp32 CUDA-core
# small-head kernels, which are meant for the tiny configs): the V-JEPA predictor (384/12) at 20 480 tokens.
PAD_TO_64_MIN_TOKENS = 1024


class _FlashAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale: float):
        out, lse = ops.flash_attn_fwd(q, k, v, scale, return_lse=True)  # [B,N,H*64] bf16, [B,H,N] fp32
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        dout = dout.to(torch.bfloat16).contiguous()
        dq, dk, dv = ops.flash_attn_bwd(q, k, v, out, dout, lse, ctx.scale)
        return dq, dk, dv, None


class _SmallHeadAttn(torch.autograd.Function):
    """head_dim 8/16/32 (e.g. the V-JEPA predictor, 384/12 = 32) on the small-head kernels, head-major [B,H,N,D] strides."""

    @staticmethod
    def forward(ctx, q, k, v, scale: float):
        out, lse = ops.attn_small_fwd_strided(q, k, v, scale)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = ops.attn_small_bwd_strided(q, k, v, out, dout.to(torch.bfloat16).contiguous(), lse, ctx.scale)
        return dq, dk, dv, None


def b200_flash_attention(module, query, key, value, attention_mask=None, *, is_causal=False, scaling=None, dropout=0.0, **kwargs):
    """Non-causal, mask-free, dropout-free multi-head attention (eager_attention_forward semantics, reference
    modeling_videomae.py:196-223 and modeling_vjepa.py:174-201 — V-JEPA's RoPE attention dispatches through the same
    registry, :352-370, with q/k already rotated).  fp32 / fp16 inputs are computed with bf16 operands (fp32 accumulate) and cast back."""
    if attention_mask is not None:
        raise SmbvError("b200_tcgen05: attention_mask is not supported (VideoMAE never passes one, reference :284)")
    if is_causal:
        raise SmbvError("b200_tcgen05: causal attention is not implemented (reference passes is_causal=False, :285)")
    if dropout and getattr(module, "training", False):
        raise SmbvError("b200_tcgen05: attention dropout is not implemented (attention_probs_dropout_prob is 0.0 on this path)")
    B, H, N, D = query.shape
    if D not in (8, 16, 32, 64):
        raise SmbvError(f"b200_tcgen05: head_dim {D} is not implemented (64 = tcgen05 kernels: smb-vision-base 768/12, decoder 384/6, "
                        "V-JEPA ViT-L 1024/16; 8/16/32 = small-head kernels: tiny configs, V-JEPA predictor 384/12)")
    dt = query.dtype
    q, k, v = (t.to(torch.bfloat16).contiguous() for t in (query, key, value))
    scale = float(scaling) if scaling is not None else D ** -0.5
    grad = torch.is_grad_enabled() and (query.requires_grad or key.requires_grad or value.requires_grad)
    if D == 64:
        out = _FlashAttn.apply(q, k, v, scale) if grad else ops.flash_attn_fwd(q, k, v, scale)
    elif D in (16, 32) and N >= PAD_TO_64_MIN_TOKENS:
        q, k, v = (torch.nn.functional.pad(t, (0, 64 - D)) for t in (q, k, v))  # differentiable: autograd slices the gradients back
        out = _FlashAttn.apply(q, k, v, scale) if grad else ops.flash_attn_fwd(q, k, v, scale)
        return out.view(B, N, H, 64)[..., :D].to(dt).contiguous(), None
    else:
        out = _SmallHeadAttn.apply(q, k, v, scale) if grad else ops.attn_small_fwd_strided(q, k, v, scale)[0]
    return out.view(B, N, H, D).to(dt), None


def register(name: str = NAME) -> str:
    """Register the kernel with transformers' attention registry and return the name to put in
    ``config._attn_implementation`` (or pass as ``attn_implementation=`` to ``from_pretrained``, src/run_mim.py:345-357)."""
    from transformers import AttentionInterface

    AttentionInterface.register(name, b200_flash_attention)
    return name
