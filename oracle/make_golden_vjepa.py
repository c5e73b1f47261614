"""Generate tests/golden/vjepa_small64.npz from the REFERENCE V-JEPA module itself (authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_vjepa

Imports ``/root/reference/src/models/vjepa/modeling_vjepa.py`` (it loads unmodified under transformers 5.x), loads the
seeded synthetic encoder weights of ``oracle/vjepa_oracle.py`` into ``VJEPA2Model`` and stores
(a) ``apply_rotary_embeddings`` outputs of one attention module for arange ids and for a position mask, and the gradient
    autograd sends back through it (the transposed map),
(b) ``model(x, context_mask, target_mask, skip_predictor=True)``: last_hidden_state / masked / target hidden states,
(c) the gradients autograd gives a selection of encoder parameters for a fixed linear loss on last_hidden_state,
as the fixtures that pin the oracle.  ``/root/reference`` does not exist on the GPU box; nothing else reads it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle.vjepa_oracle import (SMALL64_VJEPA, SMALL64_VJEPA_PRED, VJepaOracleConfig, synthetic_predictor_state_dict,
                                 synthetic_state_dict, synthetic_video)

GRAD_KEYS = ("encoder.embeddings.patch_embeddings.proj_3d.bias",  # (the 2 MB proj_3d.weight gradient is checked against the oracle only)
             "encoder.layer.0.norm1.weight", "encoder.layer.0.attention.query.weight", "encoder.layer.0.attention.key.weight",
             "encoder.layer.0.attention.key.bias", "encoder.layer.0.attention.value.bias", "encoder.layer.0.attention.proj.weight",
             "encoder.layer.1.attention.query.bias", "encoder.layer.1.mlp.fc1.weight", "encoder.layer.1.mlp.fc2.bias",
             "encoder.layer.1.norm2.bias", "encoder.layernorm.weight", "encoder.layernorm.bias")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/src")
    from models.vjepa.configuration_vjepa import VJEPA2Config
    from models.vjepa.modeling_vjepa import VJEPA2Model

    torch.set_num_threads(8)
    cfg = VJepaOracleConfig(**SMALL64_VJEPA, **SMALL64_VJEPA_PRED)
    hf = VJEPA2Config(patch_size=cfg.patch_size, crop_size=cfg.crop_size, frames_per_clip=cfg.frames_per_clip,
                      tubelet_size=cfg.tubelet_size, hidden_size=cfg.hidden_size, in_chans=cfg.in_chans,
                      num_attention_heads=cfg.num_attention_heads, num_hidden_layers=cfg.num_hidden_layers,
                      mlp_ratio=cfg.mlp_ratio, pred_hidden_size=64, pred_num_attention_heads=2, pred_num_hidden_layers=1,
                      pred_num_mask_tokens=2)
    hf._attn_implementation = "eager"
    model = VJEPA2Model(hf).eval()
    sd = synthetic_state_dict(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("predictor.") for k in missing), (missing, unexpected)

    store = {}
    # (a) rotary embedding of one attention module
    attn = model.encoder.layer[0].attention
    g = torch.Generator().manual_seed(3)
    B, H, N, D = 2, cfg.num_attention_heads, cfg.num_patches, cfg.hidden_size // cfg.num_attention_heads
    q = torch.randn(B, H, N, D, generator=g)
    up = torch.randn(B, H, N, D, generator=g)
    hidden = torch.zeros(B, N, cfg.hidden_size)
    perm = torch.stack([torch.randperm(N, generator=g)[: N // 2].sort().values for _ in range(B)])  # a position mask [B, N/2]
    for name, qq, masks in (("arange", q, None), ("masked", q[:, :, : N // 2].contiguous(), perm)):
        qq = qq.clone().requires_grad_(True)
        h_ = hidden if masks is None else hidden[:, : N // 2]
        out = attn.apply_rotary_embeddings(qq, attn.get_position_ids(h_, masks=masks))
        u = up if masks is None else up[:, :, : N // 2]
        (out * u).sum().backward()
        store[f"rope_{name}_in"] = qq.detach().numpy()
        store[f"rope_{name}_out"] = out.detach().numpy()
        store[f"rope_{name}_upstream"] = u.numpy()
        store[f"rope_{name}_grad"] = qq.grad.numpy()
    store["rope_mask_ids"] = perm.numpy()

    # (b) encoder forward with masks
    x = synthetic_video(cfg, 2)
    ctx = torch.stack([torch.randperm(N, generator=g)[:20].sort().values for _ in range(2)])
    tgt = torch.stack([torch.randperm(N, generator=g)[:12].sort().values for _ in range(2)])
    with torch.no_grad():
        o = model(pixel_values_videos=x, context_mask=[ctx], target_mask=[tgt], skip_predictor=True)
    store.update(context_mask=ctx.numpy(), target_mask=tgt.numpy(), last_hidden_state=o.last_hidden_state.numpy(),
                 masked_hidden_state=o.masked_hidden_state.numpy(), target_hidden_state=o.target_hidden_state.numpy())
    # (d) the full model with the predictor (seeded predictor weights, non-zero mask tokens): predictor output for the masks above
    with torch.no_grad():
        missing, unexpected = model.load_state_dict({**sd, **synthetic_predictor_state_dict(cfg)}, strict=True)
        full = model(pixel_values_videos=x, context_mask=[ctx], target_mask=[tgt])
    store["predictor_last_hidden_state"] = full.predictor_output.last_hidden_state.numpy()
    store["predictor_target_hidden_state"] = full.predictor_output.target_hidden_state.numpy()

    # (c) gradients of the encoder parameters for loss = <last_hidden_state, U> (U fixed, seeded)
    U = torch.randn(o.last_hidden_state.shape, generator=g)
    for p_ in model.parameters():
        p_.requires_grad_(True)
    out = model(pixel_values_videos=x, context_mask=[ctx], target_mask=[tgt], skip_predictor=True)
    (out.last_hidden_state * U).sum().backward()
    store["grad_upstream"] = U.numpy()
    for k_, p_ in model.named_parameters():
        if k_ in GRAD_KEYS:
            store["grad::" + k_] = p_.grad.numpy()
    print("last_hidden_state", tuple(o.last_hidden_state.shape), float(o.last_hidden_state.abs().mean()))
    os.makedirs(GOLD, exist_ok=True)
    import json

    with open(os.path.join(GOLD, "vjepa_small64_keys.json"), "w") as f:  # checkpoint ABI of the reference class at this config
        json.dump({"config": {**SMALL64_VJEPA, "pred_hidden_size": 64, "pred_num_attention_heads": 2, "pred_num_hidden_layers": 1,
                              "pred_num_mask_tokens": 2},
                   "state_dict": {k: list(v.shape) for k, v in model.state_dict().items()}}, f, indent=0)
    np.savez_compressed(os.path.join(GOLD, "vjepa_small64.npz"), **store)
    print("wrote", os.path.join(GOLD, "vjepa_small64.npz"))


if __name__ == "__main__":
    main()
