"""CPU: the V-JEPA host-side data entry points (smb_vision_b200.data.VJEPAMaskGenerator / vjepa_collate_fn) against
known answers produced by the reference class itself (oracle/make_golden_vjepa_masks.py) and the properties the
reference's own test asserts (tests/test_vjepa_transforms.py:104-107)."""
import hashlib
import json
import os
import random

import pytest
import torch

from smb_vision_b200.data import VJEPAMaskGenerator, vjepa_collate_fn


def digest(t):
    v = t.reshape(-1).to(torch.int64).contiguous()
    return dict(n=int(v.numel()), first=v[:8].tolist(), sha16=hashlib.sha256(v.numpy().tobytes()).hexdigest()[:16])


def test_masks_equal_the_reference_for_the_same_torch_seed(golden_dir):
    kats = json.load(open(os.path.join(golden_dir, "vjepa_mask_kat.json")))
    assert len(kats) >= 21
    for kat in kats:
        params = {k: tuple(v) if isinstance(v, list) else v for k, v in kat["params"].items()}
        torch.manual_seed(kat["seed"])
        gen = VJEPAMaskGenerator(**params)
        for want in kat["draws"]:  # two consecutive draws: the global RNG stream advances exactly as in the reference
            out = gen({})
            assert digest(out["context_mask"]) == want["context"], (kat["case"], kat["seed"])
            assert digest(out["target_mask"]) == want["target"], (kat["case"], kat["seed"])


@pytest.mark.parametrize("params", [
    dict(input_size=(224, 224, 16), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0), num_blocks=1),
    dict(input_size=(224, 224, 16), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0), num_blocks=3),
    dict(input_size=(224, 224, 16), patch_size=(16, 16, 16), pred_mask_scale=(0.4, 0.9), aspect_ratio=(0.3, 3.0), num_blocks=1),
    dict(input_size=(512, 512, 320), num_blocks=3),
])
def test_context_and_target_partition_the_grid(params):
    """reference tests/test_vjepa_transforms.py:104-107: the masks cover all patches and are disjoint."""
    torch.manual_seed(7)
    gen = VJEPAMaskGenerator(**params)
    total = gen.depth * gen.height * gen.width
    for _ in range(5):
        out = gen({"image": None})
        c, t = out["context_mask"].reshape(-1).tolist(), out["target_mask"].reshape(-1).tolist()
        assert len(c) + len(t) == total and not (set(c) & set(t)) and set(c) | set(t) == set(range(total))
        assert c == sorted(c) and t == sorted(t) and len(t) > 0


def test_options():
    torch.manual_seed(3)
    a = VJEPAMaskGenerator(input_size=96, patch_size=16, num_blocks=2)({})
    torch.manual_seed(3)
    b = VJEPAMaskGenerator(input_size=96, patch_size=16, num_blocks=2, inv_block=True)({})
    assert torch.equal(a["context_mask"], b["target_mask"]) and torch.equal(a["target_mask"], b["context_mask"])
    torch.manual_seed(3)
    c = VJEPAMaskGenerator(input_size=96, patch_size=16, num_blocks=2, max_keep=5)({})
    assert torch.equal(c["context_mask"], a["context_mask"][:5]) and torch.equal(c["target_mask"], a["target_mask"][:5])
    torch.manual_seed(3)
    d = VJEPAMaskGenerator(input_size=96, patch_size=16, num_blocks=2, full_complement=True)({})  # raises in the reference
    assert torch.equal(d["target_mask"], a["target_mask"]) and torch.equal(d["context_mask"], a["context_mask"])
    torch.manual_seed(3)
    e = VJEPAMaskGenerator(input_size=96, patch_size=16, num_blocks=2, pred_full_complement=True)({})
    assert torch.equal(e["context_mask"], a["context_mask"])


def test_collate_shares_one_examples_masks():
    """reference src/run_vjepa.py:144-160."""
    torch.manual_seed(0)
    gen = VJEPAMaskGenerator(input_size=(64, 64, 64), patch_size=(16, 16, 16), num_blocks=1)
    examples = [[gen({"image": torch.full((4, 1, 64, 64), float(i))})] for i in range(3)]  # nested single-element lists
    random.seed(5)
    pick = random.choice([0, 1, 2])
    random.seed(5)
    batch = vjepa_collate_fn(examples)
    assert batch["pixel_values_videos"].shape == (3, 4, 1, 64, 64)
    assert isinstance(batch["context_mask"], list) and len(batch["context_mask"]) == 1
    want = examples[pick][0]
    for b in range(3):
        assert torch.equal(batch["context_mask"][0][b], want["context_mask"])
        assert torch.equal(batch["target_mask"][0][b], want["target_mask"])
