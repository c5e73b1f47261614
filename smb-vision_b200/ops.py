"""Tensor-level wrappers over the C-ABI (``include/smbv_b200.h``).

PyTorch is used only for device memory and streams: every wrapper checks dtype/device/contiguity, allocates
the outputs, and hands raw pointers plus the current CUDA stream to the library.  There is no fallback: a CPU
tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_F32, EPI_GELU_BF16, EPI_POS_GATHER_F32, EPI_QKV_HEADS, EPI_RESID_F32, GemmArgs,
                   SmbvError, call)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SmbvError(f"{name}: expected a CUDA tensor (smb_vision_b200 has no CPU path)")
    if t.dtype != dtype:
        raise SmbvError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise SmbvError(f"{name}: expected a contiguous tensor")
    return t


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def mask_upsample(coarse: torch.Tensor, scale: int) -> torch.Tensor:
    """uint8 [B,cz,cy,cx] -> uint8 [B, N] at patch resolution (src/dataloader/mim.py:66-69)."""
    _chk(coarse, torch.uint8, "coarse")
    B, cz, cy, cx = coarse.shape
    fine = torch.empty((B, cz * scale * cy * scale * cx * scale), dtype=torch.uint8, device=coarse.device)
    call("smbv_mask_upsample", _ptr(coarse), _ptr(fine), B, cz, cy, cx, scale, _stream())
    return fine


def mask_index(fine: torch.Tensor):
    """uint8 [B,N] -> (vis_idx, msk_idx, slot, counts), all int32; no host synchronisation."""
    _chk(fine, torch.uint8, "fine")
    B, N = fine.shape
    dev = fine.device
    vis = torch.empty((B, N), dtype=torch.int32, device=dev)
    msk = torch.empty((B, N), dtype=torch.int32, device=dev)
    slot = torch.empty((B, N), dtype=torch.int32, device=dev)
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    call("smbv_mask_index", _ptr(fine), B, N, _ptr(vis), _ptr(msk), _ptr(slot), _ptr(counts), _stream())
    return vis, msk, slot, counts


def sincos_table(n: int, d: int, device) -> torch.Tensor:
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    call("smbv_sincos_table", _ptr(out), n, d, _stream())
    return out


def patch_embed_fwd(volume, weight, bias, pos, fine=None, slot=None, n_out=None) -> torch.Tensor:
    """volume fp32 [B,T,H,W]; weight fp32 [D,4096]; returns fp32 [B, n_out, D]."""
    _chk(volume, torch.float32, "volume")
    _chk(weight, torch.float32, "weight")
    _chk(bias, torch.float32, "bias")
    _chk(pos, torch.float32, "pos")
    B, T, H, W = volume.shape
    D = weight.shape[0]
    N = (T // 16) * (H // 16) * (W // 16)
    if fine is not None:
        _chk(fine, torch.uint8, "fine")
        _chk(slot, torch.int32, "slot")
        if n_out is None:
            raise SmbvError("patch_embed_fwd: n_out (visible tokens per sample) is required with a mask")
    else:
        n_out = N
    out = torch.empty((B, n_out, D), dtype=torch.float32, device=volume.device)
    call("smbv_patch_embed_fwd", _ptr(volume), _ptr(weight), _ptr(bias), _ptr(pos), _ptr(fine), _ptr(slot),
         B, T, H, W, 16, D, n_out, _ptr(out), _stream())
    return out


def layernorm_fwd(x, gamma, beta, eps: float, save_stats: bool = False, out=None):
    """x fp32 [..., d] -> bf16 [..., d] (+ mean, rstd fp32 [M] when save_stats)."""
    _chk(x, torch.float32, "x")
    _chk(gamma, torch.float32, "gamma")
    _chk(beta, torch.float32, "beta")
    d = x.shape[-1]
    M = x.numel() // d
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if out is None else _chk(out, torch.bfloat16, "out")
    mean = rstd = None
    if save_stats:
        mean = torch.empty((M,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    call("smbv_layernorm_fwd", _ptr(x), _ptr(gamma), _ptr(beta), float(eps), M, d, _ptr(y), _ptr(mean), _ptr(rstd), _stream())
    return (y, mean, rstd) if save_stats else y


def gemm(a, w, bias=None, epilogue=EPI_BF16, out=None, residual=None, heads=0, tokens=0, pos=None, row_map=None):
    """C = a[M,K] @ w[N,K]^T with a fused epilogue; a, w bf16.  Returns the output tensor."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.bfloat16, "w")
    K = a.shape[-1]
    M = a.numel() // K
    N = w.shape[0]
    if w.shape[1] != K:
        raise SmbvError(f"gemm: K mismatch a[..,{K}] vs w[{N},{w.shape[1]}]")
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    dev = a.device
    if out is None:
        if epilogue in (EPI_BF16, EPI_GELU_BF16):
            out = torch.empty((*a.shape[:-1], N), dtype=torch.bfloat16, device=dev)
        elif epilogue == EPI_QKV_HEADS:
            out = torch.empty((3, M // tokens, heads, tokens, 64), dtype=torch.bfloat16, device=dev)
        elif epilogue == EPI_RESID_F32:
            out = residual  # in place on the residual stream
        else:
            out = torch.empty((*a.shape[:-1], N), dtype=torch.float32, device=dev)
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw = a.data_ptr(), K, w.data_ptr(), K
    g.M, g.N, g.K = M, N, K
    g.bias = 0 if bias is None else bias.data_ptr()
    g.epilogue = epilogue
    g.out, g.ldo = out.data_ptr(), N
    g.residual = 0
    if epilogue == EPI_RESID_F32:
        _chk(residual, torch.float32, "residual")
        _chk(out, torch.float32, "out")
        g.residual = residual.data_ptr()
    g.heads, g.tokens = heads, tokens
    g.pos, g.ldpos, g.row_map = 0, 0, 0
    if epilogue == EPI_POS_GATHER_F32:
        _chk(pos, torch.float32, "pos")
        _chk(row_map, torch.int32, "row_map")
        g.pos, g.ldpos, g.row_map = pos.data_ptr(), pos.shape[-1], row_map.data_ptr()
    call("smbv_gemm_bf16", C.byref(g), _stream())
    return out


def flash_attn_fwd(q, k, v, scale: float, return_lse: bool = False, v_kmajor: bool = False):
    """q,k bf16 [B,H,N,64]; v bf16 [B,H,N,64] (or V^T [B,H,64,N] when v_kmajor) -> out bf16 [B,N,H*64]."""
    _chk(q, torch.bfloat16, "q")
    _chk(k, torch.bfloat16, "k")
    _chk(v, torch.bfloat16, "v")
    B, H, N, D = q.shape
    if D != 64:
        raise SmbvError(f"flash_attn_fwd: head_dim {D} not supported (64 only)")
    out = torch.empty((B, N, H * 64), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=q.device) if return_lse else None
    call("smbv_flash_attn_fwd_ex", _ptr(q), _ptr(k), _ptr(v), B, H, N, float(scale), _ptr(out), _ptr(lse),
         1 if v_kmajor else 0, _stream())
    return (out, lse) if return_lse else out


def fill_mask_tokens(x_dec, mask_token, pos, msk_idx, n_vis: int) -> None:
    _chk(x_dec, torch.float32, "x_dec")
    _chk(mask_token, torch.float32, "mask_token")
    _chk(pos, torch.float32, "pos")
    _chk(msk_idx, torch.int32, "msk_idx")
    B, N, d = x_dec.shape
    call("smbv_fill_mask_tokens", _ptr(x_dec), _ptr(mask_token), _ptr(pos), _ptr(msk_idx), B, N, n_vis, d,
         msk_idx.shape[1], _stream())


def normpix_loss(volume, msk_idx, n_mask: int, logits, want_grad: bool, loss_kind: int = 0, patch: int = 16):
    """volume fp32 [B,T,H,W]; logits bf16 [B,n_mask,P^3] -> (loss fp32 [], dlogits bf16 or None)."""
    _chk(volume, torch.float32, "volume")
    _chk(msk_idx, torch.int32, "msk_idx")
    _chk(logits, torch.bfloat16, "logits")
    B, T, H, W = volume.shape
    dl = torch.empty_like(logits) if want_grad else None
    partial = torch.empty((B * n_mask,), dtype=torch.float32, device=volume.device)
    loss = torch.empty((1,), dtype=torch.float32, device=volume.device)
    call("smbv_normpix_loss", _ptr(volume), B, T, H, W, patch, _ptr(msk_idx), n_mask, msk_idx.shape[1], _ptr(logits),
         _ptr(dl), _ptr(partial), _ptr(loss), loss_kind, _stream())
    return loss[0], dl


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    _chk(src, torch.float32, "src")
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    call("smbv_cast_f32_bf16", _ptr(src), _ptr(dst), src.numel(), _stream())
    return dst
