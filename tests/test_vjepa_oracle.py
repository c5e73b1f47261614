"""CPU: pins oracle/vjepa_oracle.py to the fixtures the reference V-JEPA module produced (oracle/make_golden_vjepa.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vjepa_oracle as vj


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "vjepa_small64.npz"))


@pytest.fixture(scope="module")
def cfg():
    return vj.VJepaOracleConfig(**vj.SMALL64_VJEPA)


@pytest.mark.parametrize("case", ["arange", "masked"])
def test_rope_matches_reference(gold, cfg, case):
    """apply_rotary_embeddings (modeling_vjepa.py:318-336) with arange ids and with a position mask; the reference's own
    autograd gradient pins the transposed map used by the backward kernel."""
    x = torch.from_numpy(gold[f"rope_{case}_in"])
    ids = None if case == "arange" else torch.from_numpy(gold["rope_mask_ids"])
    out = vj.rope3d(x, ids, cfg.grid_size)
    assert torch.allclose(out, torch.from_numpy(gold[f"rope_{case}_out"]), rtol=0, atol=2e-6)
    up = torch.from_numpy(gold[f"rope_{case}_upstream"])
    g = vj.rope3d(up, ids, cfg.grid_size, transpose=True)
    assert torch.allclose(g, torch.from_numpy(gold[f"rope_{case}_grad"]), rtol=0, atol=2e-6)


def test_rope_pairs_use_different_angles(cfg):
    """The reference tiles the angle vector instead of interleaving it: the map is NOT a rotation (norms change) — a
    'fixed' interleaved RoPE would pass a norm-preservation test and fail parity."""
    x = torch.randn(1, 2, 48, 64, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    y = vj.rope3d(x, None, cfg.grid_size)
    assert torch.equal(y[..., 60:], x[..., 60:])  # 64 - 3*20 tail elements pass through
    assert torch.equal(y[:, :, 0], x[:, :, 0])  # token 0: all positions 0 -> identity
    assert (y.norm(dim=-1) - x.norm(dim=-1)).abs().max() > 1e-3
    # <R x, g> == <x, R^T g>
    g = torch.randn_like(x)
    assert abs(float((y * g).sum() - (x * vj.rope3d(g, None, cfg.grid_size, transpose=True)).sum())) < 1e-9


def test_encoder_matches_reference(gold, cfg):
    sd = vj.synthetic_state_dict(cfg)
    x = vj.synthetic_video(cfg, 2)
    h = vj.encoder_forward(sd, cfg, x)
    ref = torch.from_numpy(gold["last_hidden_state"])
    assert h.shape == ref.shape
    assert torch.allclose(h, ref, rtol=0, atol=2e-5), float((h - ref).abs().max())
    ctx, tgt = torch.from_numpy(gold["context_mask"]), torch.from_numpy(gold["target_mask"])
    assert torch.allclose(vj.apply_masks(h, [ctx]), torch.from_numpy(gold["masked_hidden_state"]), rtol=0, atol=2e-5)
    assert torch.allclose(vj.apply_masks(h, [tgt]), torch.from_numpy(gold["target_hidden_state"]), rtol=0, atol=2e-5)


def test_encoder_float64_agrees(gold, cfg):
    sd = {k: v.double() for k, v in vj.synthetic_state_dict(cfg).items()}
    h = vj.encoder_forward(sd, cfg, vj.synthetic_video(cfg, 2).double())
    assert float((h.float() - torch.from_numpy(gold["last_hidden_state"])).abs().max()) < 1e-4  # fp32 rounding of the reference run


def test_encoder_gradients_match_reference(gold, cfg):
    """the oracle is differentiable torch: its autograd gradients for loss = <last_hidden_state, U> equal the reference
    model's (selection of parameters stored by make_golden_vjepa.py, incl. the K bias and the tubelet convolution)."""
    sd = {k: v.clone().requires_grad_(True) for k, v in vj.synthetic_state_dict(cfg).items()}
    h = vj.encoder_forward(sd, cfg, vj.synthetic_video(cfg, 2))
    (h * torch.from_numpy(gold["grad_upstream"])).sum().backward()
    keys = [k[len("grad::"):] for k in gold.files if k.startswith("grad::")]
    assert len(keys) >= 12
    for k in keys:
        ref = torch.from_numpy(gold["grad::" + k])
        rel = float((sd[k].grad - ref).norm() / ref.norm())
        assert rel <= 1e-4, (k, rel)


def test_checkpoint_abi_equals_the_reference_class(golden_dir):
    """B200VJEPA2Model has exactly the reference VJEPA2Model's state-dict keys and shapes (tests/golden/vjepa_small64_keys.json,
    written from the reference class), loads a reference-named checkpoint strictly, and accepts upstream's `proj` name for
    the tubelet convolution."""
    import json

    from transformers import VJEPA2Config

    from smb_vision_b200.vjepa import B200VJEPA2Model

    meta = json.load(open(os.path.join(golden_dir, "vjepa_small64_keys.json")))
    model = B200VJEPA2Model(VJEPA2Config(**meta["config"]))
    own = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert own == meta["state_dict"]
    sd = {k: torch.full(shape, 0.5) for k, shape in meta["state_dict"].items()}
    assert not any(model.load_state_dict(sd, strict=True))
    up = {k.replace("proj_3d", "proj"): v for k, v in sd.items()}
    assert not any(model.load_state_dict(up, strict=True))
    enc_only = B200VJEPA2Model(VJEPA2Config(**meta["config"]), with_predictor=False)
    assert set(enc_only.state_dict()) == {k for k in own if k.startswith("encoder.")}


def test_vjepa_module_error_behaviour_on_cpu():
    """argument errors are raised before anything touches the GPU (reference modeling_vjepa.py:1103-1104 for None)."""
    from transformers import VJEPA2Config

    from smb_vision_b200 import SmbvError
    from smb_vision_b200.vjepa import B200VJEPA2Model

    cfgd = dict(vj.SMALL64_VJEPA)
    cfgd.pop("mlp_ratio")
    m = B200VJEPA2Model(VJEPA2Config(**cfgd), with_predictor=False)
    with pytest.raises(ValueError, match="pixel_values_videos"):
        m(None, skip_predictor=True)
    with pytest.raises(ValueError, match="channel"):
        m(torch.zeros(1, 48, 3, 64, 64), skip_predictor=True)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 48, 1, 64, 64), context_head_mask=torch.ones(2), skip_predictor=True)
    with pytest.raises(SmbvError, match="predictor"):
        m(torch.zeros(1, 48, 1, 64, 64))  # built without the predictor
    with pytest.raises(SmbvError):  # no CPU path: the kernels need CUDA tensors
        m(torch.zeros(1, 48, 1, 64, 64), skip_predictor=True)
    with pytest.raises(ValueError, match="multiple of the number of attention heads"):
        B200VJEPA2Model(VJEPA2Config(**dict(cfgd, num_attention_heads=3)), with_predictor=False)


def test_predictor_matches_reference(gold):
    """groundwork for the native predictor: oracle restatement of VJEPA2Predictor.forward (mask-token choice, sort by
    position, rotary ids from the sorted position masks, unsort, projection) against the reference model's output."""
    cfg = vj.VJepaOracleConfig(**vj.SMALL64_VJEPA, **vj.SMALL64_VJEPA_PRED)
    sd = {**vj.synthetic_state_dict(cfg), **vj.synthetic_predictor_state_dict(cfg)}
    ctx, tgt = torch.from_numpy(gold["context_mask"]), torch.from_numpy(gold["target_mask"])
    enc = vj.encoder_forward(sd, cfg, vj.synthetic_video(cfg, 2))
    out = vj.predictor_forward(sd, cfg, enc, [ctx], [tgt])
    ref = torch.from_numpy(gold["predictor_last_hidden_state"])
    assert out.shape == ref.shape == (2, tgt.shape[1], cfg.hidden_size)
    assert float((out - ref).abs().max()) <= 5e-5 * max(1.0, float(ref.abs().max()))
    assert torch.allclose(vj.apply_masks(enc, [tgt]), torch.from_numpy(gold["predictor_target_hidden_state"]), rtol=0, atol=2e-5)
    # the fixture is sensitive to the things a port gets wrong: token index 0 instead of 1, no sorting
    wrong_token = vj.predictor_forward(sd, cfg, enc, [ctx], [tgt], mask_index=0)
    assert float((wrong_token - ref).abs().max()) > 1e-2
