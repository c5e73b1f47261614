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
