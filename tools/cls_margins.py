"""Print the parity margins of the classification golden comparison, several times (flake hunting)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle import videomae_oracle as vo
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import CLS, _cls_model, frob, maxrel, DEV

gold = np.load("tests/golden/small64_cls.npz")
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for ptype in CLS:
        cfg, sd, model = _cls_model(ptype)
        feats = torch.from_numpy(gold["features"])
        labels = torch.from_numpy(gold[f"{ptype}_labels"]).to(CLS[ptype][2])
        x = vo.synthetic_volume(cfg, feats.shape[0], 11)
        junk = torch.full((1 << 22,), float("nan"), device=DEV); del junk  # poison the allocator's free blocks
        out = model(x.to(DEV), additional_features=feats, labels=labels)
        out.loss.backward()
        ref_loss = float(gold[f"{ptype}_loss"])
        params = dict(model.named_parameters())
        gs = {gk: frob(params[pk].grad, torch.from_numpy(gold[f"{ptype}_{gk}"])) for gk, pk in {
            "g_classifier_w": "classifier.weight", "g_classifier_b": "classifier.bias", "g_fc_norm_w": "fc_norm.weight",
            "g_fc_norm_b": "fc_norm.bias", "g_patch_b": "videomae.embeddings.patch_embeddings.projection.bias",
            "g_qw0": "videomae.encoder.layer.0.attention.attention.query.weight"}.items()}
        print(rep, ptype, f"loss_rel={abs(out.loss.item()-ref_loss)/abs(ref_loss):.2e} logits_maxrel={maxrel(out.logits, torch.from_numpy(gold[f'{ptype}_logits'])):.2e}",
              " ".join(f"{k}={v:.2e}" for k, v in gs.items()), flush=True)
