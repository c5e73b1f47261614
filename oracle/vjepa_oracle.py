"""Oracle restatement of the reference V-JEPA2-3D ENCODER forward (test infrastructure only).

A plain-torch, CPU, functional restatement of ``/root/reference/src/models/vjepa/modeling_vjepa.py`` for the path the
embedding extraction and the momentum target encoder run (``VJEPA2Model.forward(..., skip_predictor=True)``,
:1071-1149; called at ``src/run_vjepa.py:128-135``): Conv3d tubelet embedding -> L x [LN, Q/K/V with bias, 3-axis rotary
embedding of Q and K, attention, proj, LN, MLP] -> final LN -> ``apply_masks``.  Every function cites the reference
lines it follows.  Pinned by ``tests/golden/vjepa_small64.npz``, produced by the reference module itself
(``oracle/make_golden_vjepa.py``).  Works in float32 or float64.

Never imported by the product package.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class VJepaOracleConfig:
    """Subset of the reference ``VJEPA2Config`` (configuration_vjepa.py:98-150) the encoder reads; defaults after
    ``src/run_vjepa.py:220-232`` (in_chans 1, tubelet_size = patch_size) on the ViT-L encoder."""

    crop_size: int = 512
    frames_per_clip: int = 320
    patch_size: int = 16
    tubelet_size: int = 16
    in_chans: int = 1
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    mlp_ratio: float = 4.0
    layer_norm_eps: float = 1e-6
    qkv_bias: bool = True
    pred_hidden_size: int = 384
    pred_num_attention_heads: int = 12
    pred_num_hidden_layers: int = 12
    pred_num_mask_tokens: int = 10
    pred_mlp_ratio: float = 4.0

    @property
    def grid_size(self):
        return self.crop_size // self.patch_size

    @property
    def grid_depth(self):
        return self.frames_per_clip // self.tubelet_size

    @property
    def num_patches(self):
        return self.grid_depth * self.grid_size * self.grid_size


SMALL64_VJEPA = dict(crop_size=64, frames_per_clip=48, patch_size=16, tubelet_size=16, in_chans=1, hidden_size=128,
                     num_hidden_layers=2, num_attention_heads=2, mlp_ratio=4.0)  # head_dim 64, 3x4x4 = 48 tokens
SMALL64_VJEPA_PRED = dict(pred_hidden_size=64, pred_num_attention_heads=2, pred_num_hidden_layers=1, pred_num_mask_tokens=2)  # head_dim 32


def position_ids(ids: torch.Tensor, grid_size: int):
    """modeling_vjepa.py:282-316: token id -> (frame, height, width) index."""
    tpf = grid_size * grid_size
    frame = ids // tpf
    rem = ids - tpf * frame
    height = rem // grid_size
    return frame, height, rem - grid_size * height


def rotate(x: torch.Tensor, pos: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """modeling_vjepa.py:204-228 on one segment: x [B,H,N,S], pos broadcastable to [B,H,N].  The S/2 angles are TILED
    twice (`.repeat(1,1,1,2)`) while adjacent elements are paired, so element e uses angle e mod (S/2).
    transpose=True applies the transposed linear map (what autograd gives for the backward pass)."""
    S = x.shape[-1]
    h = S // 2
    omega = torch.arange(h, dtype=x.dtype) / (S / 2.0)
    omega = 1.0 / 10000**omega
    freq = pos.to(x.dtype).unsqueeze(-1) * omega  # [..., N, h]
    cos = freq.cos().repeat(*([1] * (freq.dim() - 1)), 2)
    sin = freq.sin().repeat(*([1] * (freq.dim() - 1)), 2)
    even, odd = x[..., 0::2], x[..., 1::2]
    if not transpose:
        y = torch.stack((-odd, even), dim=-1).flatten(-2)  # y[2i] = -x[2i+1], y[2i+1] = x[2i]
        return x * cos + y * sin
    gs = x * sin  # transposed: g_x[2i] = g[2i] cos_2i + g[2i+1] sin_2i+1 ; g_x[2i+1] = g[2i+1] cos_2i+1 - g[2i] sin_2i
    z = torch.stack((gs[..., 1::2], -gs[..., 0::2]), dim=-1).flatten(-2)
    return x * cos + z


def rope3d(x: torch.Tensor, ids: torch.Tensor | None, grid_size: int, transpose: bool = False) -> torch.Tensor:
    """modeling_vjepa.py:318-336 apply_rotary_embeddings: x [B,H,N,D]; ids [B,N] (position_mask) or None = arange(N)."""
    B, H, N, D = x.shape
    S = 2 * ((D // 3) // 2)  # :272-274
    ids = torch.arange(N) if ids is None else ids.unsqueeze(1)  # :297-301 (masks repeated over heads)
    out, s = [], 0
    for pos in position_ids(ids, grid_size):
        out.append(rotate(x[..., s:s + S], pos, transpose))
        s += S
    if s < D:
        out.append(x[..., s:])
    return torch.cat(out, dim=-1)


def _layer(x, sd, pre, heads, eps, grid_size, ids=None):
    """VJEPA2Layer.forward, modeling_vjepa.py:455-485 (+ attention :338-371, MLP :423-426)."""
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    a = pre + "attention."
    q = F.linear(h, sd[a + "query.weight"], sd.get(a + "query.bias"))
    k = F.linear(h, sd[a + "key.weight"], sd.get(a + "key.bias"))
    v = F.linear(h, sd[a + "value.weight"], sd.get(a + "value.bias"))
    B, N, _ = q.shape
    hd = d // heads
    q, k, v = (t.view(B, N, heads, hd).transpose(1, 2) for t in (q, k, v))
    q, k = rope3d(q, ids, grid_size), rope3d(k, ids, grid_size)  # :346-348
    s = torch.matmul(q, k.transpose(-1, -2)) * (hd**-0.5)
    o = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(B, N, d)
    x = x + F.linear(o, sd[a + "proj.weight"], sd[a + "proj.bias"])
    h = F.layer_norm(x, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    h = F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def encoder_forward(sd, cfg: VJepaOracleConfig, pixel_values_videos: torch.Tensor) -> torch.Tensor:
    """VJEPA2Encoder.forward, modeling_vjepa.py:509-546: pixel_values_videos [B,T,C,H,W] -> last_hidden_state [B,N,d]."""
    w, b = sd["encoder.embeddings.patch_embeddings.proj_3d.weight"], sd["encoder.embeddings.patch_embeddings.proj_3d.bias"]
    x = pixel_values_videos.permute(0, 2, 1, 3, 4).to(w.dtype)  # :157-160
    st = (cfg.tubelet_size, cfg.patch_size, cfg.patch_size)
    h = F.conv3d(x, w, b, stride=st).flatten(2).transpose(1, 2)  # :130-132
    for i in range(cfg.num_hidden_layers):
        h = _layer(h, sd, f"encoder.layer.{i}.", cfg.num_attention_heads, cfg.layer_norm_eps, cfg.grid_size)
    return F.layer_norm(h, (cfg.hidden_size,), sd["encoder.layernorm.weight"], sd["encoder.layernorm.bias"], cfg.layer_norm_eps)


def apply_masks(t: torch.Tensor, masks) -> torch.Tensor:
    """modeling_vjepa.py:543-557: gather the rows listed in every mask [B,K], concatenated along the batch."""
    return torch.cat([torch.gather(t, 1, m.unsqueeze(-1).expand(-1, -1, t.shape[-1])) for m in masks], dim=0)


def param_shapes(cfg: VJepaOracleConfig) -> dict:
    d, m = cfg.hidden_size, int(cfg.hidden_size * cfg.mlp_ratio)
    s = {"encoder.embeddings.patch_embeddings.proj_3d.weight": (d, cfg.in_chans, cfg.tubelet_size, cfg.patch_size, cfg.patch_size),
         "encoder.embeddings.patch_embeddings.proj_3d.bias": (d,)}
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        for ln in ("norm1", "norm2"):
            s[p + ln + ".weight"] = (d,)
            s[p + ln + ".bias"] = (d,)
        for lin in ("query", "key", "value", "proj"):
            s[p + f"attention.{lin}.weight"] = (d, d)
            if lin == "proj" or cfg.qkv_bias:
                s[p + f"attention.{lin}.bias"] = (d,)
        s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (m, d), (m,)
        s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (d, m), (d,)
    s["encoder.layernorm.weight"], s["encoder.layernorm.bias"] = (d,), (d,)
    return s


def synthetic_state_dict(cfg: VJepaOracleConfig, seed: int = 4321) -> dict:
    """Seeded encoder weights.  The reference init (modeling_vjepa.py:1017-1041: trunc-normal 0.02, zero biases) gives
    near-uniform attention at this width, and an encoder whose output barely depends on the rotary embedding or the K
    bias (measured: 1e-4 relative) — useless as a parity fixture.  So Q/K weights are drawn with std 0.3 and V/proj with
    0.1 (peaky attention that carries weight in the residual stream: dropping the rotary step moves the output by 40 %,
    a +1 K bias by 17 %), biases and LayerNorm parameters are perturbed, the rest keeps std 0.02."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if "norm" in k:
            t = torch.full(shp, 1.0 if k.endswith("weight") else 0.0) + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith(("query.weight", "key.weight")):
            t = 0.3 * torch.randn(shp, generator=g)
        elif k.endswith(("value.weight", "attention.proj.weight")):
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        sd[k] = t.float()
    return sd


def synthetic_video(cfg: VJepaOracleConfig, batch: int = 1, seed: int = 9) -> torch.Tensor:
    """uniform [0,1) fp32 [B,T,1,H,W] (the ScaleIntensityRanged output range of the "vjepa" transform preset)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, cfg.frames_per_clip, cfg.in_chans, cfg.crop_size, cfg.crop_size, generator=g)


# ---- predictor (groundwork for the native predictor; not on the product path yet) ----
def predictor_forward(sd, cfg: VJepaOracleConfig, encoder_hidden_states: torch.Tensor, context_mask, target_mask, mask_index: int = 1):
    """VJEPA2Predictor.forward, modeling_vjepa.py:686-746 (+ VJEPA2PredictorEmbeddings.forward :589-631): context rows ->
    Linear to the predictor width; one learned mask token (index `mask_index % pred_num_mask_tokens`; the reference's
    default mask_index is 1) repeated for every target position; both concatenated, SORTED by token position, run through
    the predictor blocks with the sorted positions as rotary ids, LayerNorm, unsorted, target rows projected back."""
    p = "predictor."
    ctx = apply_masks(encoder_hidden_states, context_mask)  # :703
    B, n_ctx, _ = ctx.shape
    context = F.linear(ctx, sd[p + "embeddings.predictor_embeddings.weight"], sd[p + "embeddings.predictor_embeddings.bias"])  # :605
    token = sd[p + "embeddings.mask_tokens"][mask_index % cfg.pred_num_mask_tokens]  # [1, 1, pd]  :608-609
    n_tgt = torch.cat(target_mask, dim=0).shape[1]
    target = token.expand(B, n_tgt, -1)  # :615-617: repeat to max index + 1, then gather -> the same token everywhere
    h = torch.cat([context, target.to(context.dtype)], dim=1)  # :621
    masks = torch.cat([torch.cat(context_mask, dim=0), torch.cat(target_mask, dim=0)], dim=1)  # :624-626
    order = torch.argsort(masks, dim=1)  # :708
    masks = torch.gather(masks, 1, order)
    h = torch.gather(h, 1, order.unsqueeze(-1).expand(-1, -1, h.shape[-1]))  # :640-648
    for i in range(cfg.pred_num_hidden_layers):
        h = _layer(h, sd, f"{p}layer.{i}.", cfg.pred_num_attention_heads, cfg.layer_norm_eps, cfg.grid_size, ids=masks)
    h = F.layer_norm(h, (cfg.pred_hidden_size,), sd[p + "layernorm.weight"], sd[p + "layernorm.bias"], cfg.layer_norm_eps)  # :730
    inverse = torch.argsort(order, dim=1)  # :674-679
    h = torch.gather(h, 1, inverse.unsqueeze(-1).expand(-1, -1, h.shape[-1]))[:, n_ctx:]  # :732-733
    return F.linear(h, sd[p + "proj.weight"], sd[p + "proj.bias"])  # :735


def predictor_param_shapes(cfg: VJepaOracleConfig) -> dict:
    d, pd, m = cfg.hidden_size, cfg.pred_hidden_size, int(cfg.pred_hidden_size * cfg.pred_mlp_ratio)
    s = {"predictor.embeddings.mask_tokens": (cfg.pred_num_mask_tokens, 1, 1, pd),
         "predictor.embeddings.predictor_embeddings.weight": (pd, d), "predictor.embeddings.predictor_embeddings.bias": (pd,)}
    for i in range(cfg.pred_num_hidden_layers):
        p = f"predictor.layer.{i}."
        for ln in ("norm1", "norm2"):
            s[p + ln + ".weight"], s[p + ln + ".bias"] = (pd,), (pd,)
        for lin in ("query", "key", "value", "proj"):
            s[p + f"attention.{lin}.weight"] = (pd, pd)
            if lin == "proj" or cfg.qkv_bias:
                s[p + f"attention.{lin}.bias"] = (pd,)
        s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (m, pd), (m,)
        s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (pd, m), (pd,)
    s["predictor.layernorm.weight"], s["predictor.layernorm.bias"] = (pd,), (pd,)
    s["predictor.proj.weight"], s["predictor.proj.bias"] = (d, pd), (d,)
    return s


def synthetic_predictor_state_dict(cfg: VJepaOracleConfig, seed: int = 777) -> dict:
    """Seeded predictor weights in the style of `synthetic_state_dict` (peaky attention, perturbed biases / LayerNorms);
    the mask tokens — zero at the reference init — get std 0.5 so that the choice of token index is visible."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in predictor_param_shapes(cfg).items():
        if "norm" in k:
            t = torch.full(shp, 1.0 if k.endswith("weight") else 0.0) + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("mask_tokens"):
            t = 0.5 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith(("query.weight", "key.weight")):
            t = 0.4 * torch.randn(shp, generator=g)
        elif k.endswith(("value.weight", "attention.proj.weight")):
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.05 * torch.randn(shp, generator=g)
        sd[k] = t.float()
    return sd
