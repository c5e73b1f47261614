"""CPU: the oracle restatement against fixtures produced by the reference itself
(oracle/make_golden.py) and the mask known-answer vectors of SURVEY.md §8c."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle.mim_mask import OracleMaskGenerator
from oracle import videomae_oracle as vo


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module", params=["tiny", "small64"])
def tiny(golden_dir, request):
    """reference-generated fixture: BASELINE configs[0] ("tiny") and the smallest head_dim-64 config ("small64")."""
    meta = json.load(open(os.path.join(golden_dir, f"{request.param}_mim.json")))
    gold = np.load(os.path.join(golden_dir, f"{request.param}_mim.npz"))
    cfg = vo.OracleConfig(**meta["config"])
    sd = vo.synthetic_state_dict(cfg, meta["weight_seed"])
    x = vo.synthetic_volume(cfg, 1, meta["volume_seed"])
    return meta, gold, cfg, sd, x


def test_mask_known_answers(golden_dir):
    for kat in json.load(open(os.path.join(golden_dir, "mask_kat.json"))):
        np.random.seed(kat["seed"])
        g = OracleMaskGenerator(kat["input_size"], kat["depth"], kat["mask_patch_size"], kat["model_patch_size"], kat["mask_ratio"])
        assert (g.token_count, g.mask_count) == (kat["cells"], kat["masked_cells"])
        coarse = g.coarse()
        fine = g.upsample(coarse, g.scale)
        assert sha16(coarse.astype(np.uint8)) == kat["sha_coarse"]
        assert sha16(fine.astype(np.uint8)) == kat["sha_fine"]
        assert fine.size == kat["n"] and int(fine.sum()) == kat["n_mask"]
        assert list(np.nonzero(fine)[0][:10]) == kat["first_masked"]


def test_mask_survey_hashes():
    # the hashes SURVEY.md §8c quotes, independent of the json
    np.random.seed(0)
    assert sha16(OracleMaskGenerator(96, 96, 32, 16, 0.65)().astype(np.uint8)) == "15d06e74fc2d00f7"
    np.random.seed(0)
    assert sha16(OracleMaskGenerator(512, 320, 32, 16, 0.65)().astype(np.uint8)) == "4a598b65ed9ea8db"


def test_mask_ctor_errors():
    # src/dataloader/mim.py:47-52
    for args in [(100, 96, 32, 16), (96, 100, 32, 16), (96, 96, 32, 12)]:
        with pytest.raises(ValueError):
            OracleMaskGenerator(*args, 0.5)


def test_synthetic_inputs_are_stable(tiny):
    meta, gold, cfg, sd, x = tiny
    assert sha16(x.numpy()) == meta["volume_sha"]
    assert sha16(np.concatenate([sd[k].numpy().ravel() for k in sd])) == meta["weights_sha"]
    np.random.seed(meta["mask_seed"])
    m = OracleMaskGenerator(cfg.image_size, cfg.num_frames, meta["mask_patch_size"], 16, meta["mask_ratio"])()
    assert np.array_equal(m[None], gold["mask"])


def test_oracle_matches_reference_forward(tiny):
    meta, gold, cfg, sd, x = tiny
    mask = torch.from_numpy(gold["mask"])
    with torch.no_grad():
        loss, logits, _ = vo.pretrain_forward(sd, cfg, x, mask)
        emb = vo.encoder(sd, cfg, x, None)
    assert abs(loss.item() - float(gold["loss"])) <= 2e-6 * float(gold["loss"])
    assert np.abs(logits.numpy() - gold["logits"]).max() <= 2e-5
    assert np.abs(emb.numpy() - gold["embeddings"]).max() <= 2e-5


def test_oracle_matches_reference_gradients(tiny):
    meta, gold, cfg, sd, x = tiny
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mask = torch.from_numpy(gold["mask"])
    loss, _, _ = vo.pretrain_forward(sd, cfg, x, mask)
    loss.backward()
    pairs = {
        "g_patch_w": "videomae.embeddings.patch_embeddings.projection.weight",
        "g_mask_token": "mask_token",
        "g_q_bias0": "videomae.encoder.layer.0.attention.attention.q_bias",
        "g_e2d": "encoder_to_decoder.weight",
        "g_head_b": "decoder.head.bias",
        "g_fc1_w_l1": "videomae.encoder.layer.1.intermediate.dense.weight",
    }
    for gk, pk in pairs.items():
        ref = gold[gk]
        got = sd[pk].grad.numpy()
        assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref) + 1e-12, gk
    norms = np.array([float(sd[k].grad.norm()) for k in meta["grad_keys"]])
    assert np.allclose(norms, gold["grad_norms"], rtol=1e-4, atol=1e-10)


def test_oracle_float64_agrees(tiny):
    meta, gold, cfg, sd, x = tiny
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        loss, logits, _ = vo.pretrain_forward(sd64, cfg, x.double(), torch.from_numpy(gold["mask"]))
    assert abs(loss.item() - float(gold["loss"])) < 1e-6
    assert np.abs(logits.numpy() - gold["logits"]).max() < 2e-5


def test_oracle_errors(tiny):
    meta, gold, cfg, sd, x = tiny
    with pytest.raises(ValueError):
        vo.embed(sd, cfg, x.repeat(1, 1, 3, 1, 1), None)  # channel mismatch, modeling_videomae.py:181-184
    with pytest.raises(ValueError):
        vo.embed(sd, cfg, x[..., :80], None)  # size mismatch :185-188
    with pytest.raises(ValueError):
        vo.pretrain_forward(sd, cfg, x, None)  # :807-808


def test_biased_variance_would_be_caught(tiny):
    """a biased-variance bug moves the loss by ~2.4e-4 rel; tolerance 1e-4 must catch it."""
    meta, gold, cfg, sd, x = tiny
    P = vo.patchify(x, cfg)
    lab_b = (P - P.mean(-1, keepdim=True)) / (P.var(-1, unbiased=False, keepdim=True).sqrt() + 1e-6)
    mask = torch.from_numpy(gold["mask"])
    lab_b = lab_b[mask].reshape(1, -1, lab_b.shape[-1])
    loss_b = torch.nn.functional.mse_loss(torch.from_numpy(gold["logits"]), lab_b).item()
    assert abs(loss_b - float(gold["loss"])) / float(gold["loss"]) > 1e-4


# ---- classification head with additional features (reference modeling_videomae.py:917-1023) ----
CLS = {"single": (3, torch.long), "multi": (3, torch.float32), "regression": (1, torch.float32)}


@pytest.mark.parametrize("name,cfgd", [("tiny", vo.TINY), ("small64", None)])
@pytest.mark.parametrize("ptype", list(CLS))
def test_oracle_classification_matches_reference(golden_dir, name, cfgd, ptype):
    if cfgd is None:
        from __graft_entry__ import SMALL64 as cfgd
    gold = np.load(os.path.join(golden_dir, f"{name}_cls.npz"))
    cfg = vo.OracleConfig(**cfgd)
    n_lab, ldt = CLS[ptype]
    feats = torch.from_numpy(gold["features"])
    labels = torch.from_numpy(gold[f"{ptype}_labels"]).to(ldt)
    sd = {k: v.clone().requires_grad_(True) for k, v in vo.synthetic_cls_state_dict(cfg, n_lab, feats.shape[1], 1234).items()}
    x = vo.synthetic_volume(cfg, feats.shape[0], 11)
    full = {"single": "single_label_classification", "multi": "multi_label_classification", "regression": "regression"}[ptype]
    loss, logits = vo.classify_forward(sd, cfg, x, feats, labels, n_lab, full)
    loss.backward()
    assert abs(loss.item() - float(gold[f"{ptype}_loss"])) <= 5e-6 * abs(float(gold[f"{ptype}_loss"]))
    assert np.abs(logits.detach().numpy() - gold[f"{ptype}_logits"]).max() <= 2e-5
    for gk, pk in {"g_classifier_w": "classifier.weight", "g_classifier_b": "classifier.bias", "g_fc_norm_w": "fc_norm.weight",
                   "g_fc_norm_b": "fc_norm.bias", "g_patch_b": "videomae.embeddings.patch_embeddings.projection.bias",
                   "g_qw0": "videomae.encoder.layer.0.attention.attention.query.weight"}.items():
        ref = gold[f"{ptype}_{gk}"]
        got = sd[pk].grad.numpy()
        assert np.linalg.norm(got - ref) <= 2e-4 * np.linalg.norm(ref) + 1e-12, gk


def test_oracle_classification_feature_size_error():
    cfg = vo.OracleConfig(**vo.TINY)
    sd = vo.synthetic_cls_state_dict(cfg, 3, 2, 1234)
    x = vo.synthetic_volume(cfg, 1, 11)
    with pytest.raises(ValueError):  # reference :983-986
        vo.classify_forward(sd, cfg, x, torch.zeros(1, 5), None, 3)


def test_oracle_config_variants_match_upstream():
    """qkv_bias=False and use_mean_pooling=False (final encoder LayerNorm, reference :517-520): the oracle against the upstream
    class the reference imports, run here (no fixture needed: both sides are evaluated in this test)."""
    import transformers

    import __graft_entry__ as ge

    cfgd = dict(vo.TINY, qkv_bias=False, use_mean_pooling=False)
    cfg = vo.OracleConfig(**cfgd)
    sd = vo.synthetic_state_dict(cfg, 1234)
    hc = ge.hf_config(cfgd)
    hc._attn_implementation = "eager"
    up = transformers.VideoMAEForPreTraining(hc).eval()
    up.load_state_dict(sd, strict=True)
    x = vo.synthetic_volume(cfg, 1, 7)
    np.random.seed(0)
    mask = torch.from_numpy(OracleMaskGenerator(96, 96, 32, 16, 0.65)())[None]
    with torch.no_grad():
        want = up(x, mask)
        emb_w = up.videomae(x).last_hidden_state
        loss, logits, _ = vo.pretrain_forward(sd, cfg, x, mask)
        emb = vo.encoder(sd, cfg, x, None)
    assert abs(loss.item() - want.loss.item()) <= 2e-6 * want.loss.item()
    assert (logits - want.logits).abs().max().item() <= 2e-5 and (emb - emb_w).abs().max().item() <= 2e-5


def test_simmim_variant_of_the_oracle_is_self_consistent():
    """`pretrain_forward_simmim` (north-star variant; no reference model implements it, so it is pinned only to the pieces it is
    made of): labels equal the MAE path's, the blend is `torch.where(mask, mask_token, emb)` — a token's logits do not depend on the
    voxels of MASKED patches — and gradient reaches the encoder mask token but not the decoder-width one."""
    cfg = vo.OracleConfig(**vo.TINY)
    sd = vo.synthetic_state_dict(cfg, 3)
    sd["videomae.embeddings.mask_token"] = 0.2 * torch.randn(1, 1, cfg.hidden_size, generator=torch.Generator().manual_seed(4))
    x = vo.synthetic_volume(cfg, 1, 5)
    np.random.seed(0)
    mask = torch.from_numpy(OracleMaskGenerator(96, 96, 32, 16, 0.65)())[None]
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss, logits, ex = vo.pretrain_forward_simmim(sdg, cfg, x, mask, "l1")
    _, _, ex_mae = vo.pretrain_forward(sd, cfg, x, mask)
    assert torch.equal(ex["labels"], ex_mae["labels"]) and logits.shape == (1, 144, 4096)
    loss.backward()
    assert float(sdg["videomae.embeddings.mask_token"].grad.abs().sum()) > 0 and sdg["mask_token"].grad is None
    # perturb the voxels of every masked patch: the logits must not move (those embeddings were replaced), the loss does
    P = vo.patchify(x, cfg).clone()
    P[mask] += 0.25
    ts, ps = cfg.tubelet_size, cfg.patch_size
    g = cfg.grid
    x2 = P.view(1, g[0], g[1], g[2], ts, ps, ps, 1).permute(0, 1, 4, 7, 2, 5, 3, 6).reshape(x.shape)
    assert torch.equal(vo.patchify(x2, cfg), P)
    with torch.no_grad():
        _, logits2, _ = vo.pretrain_forward_simmim(sd, cfg, x2, mask, "l1")
    assert torch.allclose(logits2, logits.detach(), atol=1e-6)
