"""Generate tests/golden/* from the REFERENCE ITSELF (run in the authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Imports (a) the vendored reference model ``/root/reference/src/models/videomae/modeling_videomae.py``
behind the two-line transformers-5.x compat shim described in SURVEY.md §8c and (b) upstream
``transformers.VideoMAEForPreTraining`` (what ``src/run_mim.py:19-20`` really imports), loads the
seeded synthetic weights into both, asserts (a) == (b), and stores their outputs on the tiny
config (BASELINE.json configs[0]) as the fixtures that pin ``oracle/videomae_oracle.py``.
``/root/reference`` does not exist on the GPU box; nothing else reads it.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

from oracle.mim_mask import OracleMaskGenerator
from oracle.videomae_oracle import TINY, OracleConfig, synthetic_state_dict, synthetic_volume

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load_reference_module():
    import transformers
    import transformers.pytorch_utils as pu
    from transformers import PreTrainedModel

    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        pu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    if not hasattr(PreTrainedModel, "get_head_mask"):
        PreTrainedModel.get_head_mask = lambda self, hm, n, *a: [None] * n
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/src")
    from models.videomae import modeling_videomae as ref  # noqa

    return ref, transformers


def _hf_config(transformers, cfg: OracleConfig, attn="eager"):
    c = transformers.VideoMAEConfig(
        hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
        num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size,
        decoder_hidden_size=cfg.decoder_hidden_size, decoder_num_hidden_layers=cfg.decoder_num_hidden_layers,
        decoder_num_attention_heads=cfg.decoder_num_attention_heads,
        decoder_intermediate_size=cfg.decoder_intermediate_size,
    )
    # src/run_mim.py:322-330
    c.update(dict(image_size=cfg.image_size, patch_size=cfg.patch_size, num_channels=cfg.num_channels,
                  num_frames=cfg.num_frames, tubelet_size=cfg.tubelet_size))
    c._attn_implementation = attn
    return c


def sha16(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def mask_kats():
    out = []
    for name, (size, depth) in {"tiny": (96, 96), "full": (512, 320)}.items():
        np.random.seed(0)
        g = OracleMaskGenerator(size, depth, 32, 16, 0.65)
        state = np.random.get_state()
        perm = np.random.permutation(g.token_count)[: g.mask_count]
        np.random.set_state(state)
        coarse = g.coarse()
        fine = g.upsample(coarse, g.scale)
        out.append(dict(name=name, input_size=size, depth=depth, mask_patch_size=32, model_patch_size=16,
                        mask_ratio=0.65, seed=0, cells=int(g.token_count), masked_cells=int(g.mask_count),
                        n=int(fine.size), n_mask=int(fine.sum()), perm_head=[int(v) for v in perm[:12]],
                        sha_coarse=sha16(coarse.astype(np.uint8)), sha_fine=sha16(fine.astype(np.uint8)),
                        first_masked=[int(v) for v in np.nonzero(fine)[0][:10]]))
    return out


def mim_golden(ref, transformers, name: str, cfgd: dict, check_upstream: bool = True):
    """loss / logits / embeddings / selected gradients of the REFERENCE model on config `cfgd` -> tests/golden/<name>_mim.*"""
    cfg = OracleConfig(**cfgd)
    sd = synthetic_state_dict(cfg, seed=1234, perturb=True)
    x = synthetic_volume(cfg, batch=1, seed=7)
    np.random.seed(0)
    mask = torch.from_numpy(OracleMaskGenerator(cfg.image_size, cfg.num_frames, 32, 16, 0.65)()).unsqueeze(0)

    hf = _hf_config(transformers, cfg)
    m_ref = ref.VideoMAEForPreTraining(hf).eval()
    m_up = transformers.VideoMAEForPreTraining(hf).eval()
    assert set(m_ref.state_dict().keys()) == set(sd.keys()), set(m_ref.state_dict().keys()) ^ set(sd.keys())
    m_ref.load_state_dict(sd, strict=True)
    m_up.load_state_dict(sd, strict=True)

    for p in m_ref.parameters():
        p.requires_grad_(True)
    out = m_ref(x, mask)
    out.loss.backward()
    with torch.no_grad():
        out_up = m_up(x, mask)
        emb = m_ref.videomae(x).last_hidden_state
        emb_up = m_up.videomae(x).last_hidden_state
    d_loss = abs(out.loss.item() - out_up.loss.item())
    d_logits = (out.logits - out_up.logits).abs().max().item()
    d_emb = (emb - emb_up).abs().max().item()
    print(f"[{name}] reference-vs-upstream: dloss={d_loss:.3e} dlogits={d_logits:.3e} demb={d_emb:.3e}")
    assert d_loss < 1e-6 and d_logits < 1e-5 and d_emb < 1e-5

    grads = {k: p.grad.detach().numpy() for k, p in m_ref.named_parameters()}
    gsel = {
        "g_patch_w": grads["videomae.embeddings.patch_embeddings.projection.weight"],
        "g_mask_token": grads["mask_token"],
        "g_q_bias0": grads["videomae.encoder.layer.0.attention.attention.q_bias"],
        "g_e2d": grads["encoder_to_decoder.weight"],
        "g_head_b": grads["decoder.head.bias"],
        "g_fc1_w_l1": grads["videomae.encoder.layer.1.intermediate.dense.weight"],
    }
    np.savez_compressed(
        os.path.join(GOLD, f"{name}_mim.npz"),
        loss=np.float64(out.loss.item()), logits=out.logits.detach().numpy().astype(np.float32),
        embeddings=emb.numpy().astype(np.float32), mask=mask.numpy(),
        grad_norms=np.array([float(np.linalg.norm(grads[k])) for k in sorted(grads)], dtype=np.float64),
        **gsel,
    )
    meta = dict(
        config=cfgd, weight_seed=1234, volume_seed=7, mask_seed=0, mask_ratio=0.65, mask_patch_size=32,
        volume_sha=sha16(x.numpy()), weights_sha=sha16(np.concatenate([sd[k].numpy().ravel() for k in sd])),
        loss=out.loss.item(), reference_vs_upstream=dict(dloss=d_loss, dlogits=d_logits, demb=d_emb),
        torch=torch.__version__, transformers=transformers.__version__, numpy=np.__version__,
        grad_keys=sorted(grads),
        source="reference /root/reference/src/models/videomae/modeling_videomae.py (fp32, eager attention, CPU)",
    )
    with open(os.path.join(GOLD, f"{name}_mim.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"[{name}] loss", out.loss.item())


CLS_CASES = {  # problem type -> (num_labels, labels)
    "single_label_classification": (3, torch.tensor([2, 0])),
    "multi_label_classification": (3, torch.tensor([[1.0, 0.0, 1.0], [0.0, 0.0, 1.0]])),
    "regression": (1, torch.tensor([0.7, -1.3])),
}


def cls_golden(ref, transformers, name: str, cfgd: dict):
    """VideoMAEForVideoClassification with additional features (BASELINE configs[3]; reference :917-1023), all three
    problem types of :995-1012 -> tests/golden/<name>_cls.npz"""
    from oracle.videomae_oracle import synthetic_cls_state_dict

    cfg = OracleConfig(**cfgd)
    n_feat, Bc = 2, 2
    xc = synthetic_volume(cfg, Bc, 11)
    feats = torch.randn(Bc, n_feat, generator=torch.Generator().manual_seed(5))  # age / sex style (src/run_classification.py:227-271)
    store = dict(features=feats.numpy())
    for ptype, (n_lab, labels) in CLS_CASES.items():
        hfc = _hf_config(transformers, cfg)
        hfc.num_labels = n_lab
        hfc.additional_features_size = n_feat
        hfc.problem_type = ptype
        m_cls = ref.VideoMAEForVideoClassification(hfc).eval()
        sdc = synthetic_cls_state_dict(cfg, n_lab, n_feat, 1234)
        assert set(m_cls.state_dict().keys()) == set(sdc.keys()), set(m_cls.state_dict().keys()) ^ set(sdc.keys())
        m_cls.load_state_dict(sdc, strict=True)
        for p_ in m_cls.parameters():
            p_.requires_grad_(True)
        oc = m_cls(xc, additional_features=feats, labels=labels)
        oc.loss.backward()
        gc = {k: p_.grad.detach().numpy() for k, p_ in m_cls.named_parameters()}
        t = ptype.split("_")[0]
        store.update({f"{t}_loss": np.float64(oc.loss.item()), f"{t}_logits": oc.logits.detach().numpy(), f"{t}_labels": labels.numpy(),
                      f"{t}_g_classifier_w": gc["classifier.weight"], f"{t}_g_classifier_b": gc["classifier.bias"],
                      f"{t}_g_fc_norm_w": gc["fc_norm.weight"], f"{t}_g_fc_norm_b": gc["fc_norm.bias"],
                      f"{t}_g_patch_b": gc["videomae.embeddings.patch_embeddings.projection.bias"],
                      f"{t}_g_qw0": gc["videomae.encoder.layer.0.attention.attention.query.weight"]})
        print(f"[{name}] cls {ptype} loss", oc.loss.item())
    np.savez_compressed(os.path.join(GOLD, f"{name}_cls.npz"), **store)


def main():
    from __graft_entry__ import SMALL64

    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    ref, transformers = _load_reference_module()
    mim_golden(ref, transformers, "tiny", TINY)        # BASELINE configs[0]
    mim_golden(ref, transformers, "small64", SMALL64)  # the smallest head_dim-64 config (what the GPU parity tests run)
    cls_golden(ref, transformers, "tiny", TINY)
    cls_golden(ref, transformers, "small64", SMALL64)
    with open(os.path.join(GOLD, "mask_kat.json"), "w") as f:
        json.dump(mask_kats(), f, indent=1)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
