"""Known-answer vectors for the V-JEPA mask generator, produced by the REFERENCE class itself (authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_vjepa_masks

``/root/reference/src/dataloader/transforms.py`` imports MONAI at module level (absent from this image);
``VJEPAMaskGenerator`` itself only uses torch, so the module is imported with a stub ``monai.transforms`` (empty
``Transform`` / ``MapTransform`` base classes, ``Compose`` = list holder) and the class is run unmodified.  Writes
tests/golden/vjepa_mask_kat.json: for every (parameter set, torch seed) and two consecutive draws, the count, first indices
and a hash of the context / target index lists.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import torch

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [  # the reference test's three parameter sets (tests/test_vjepa_transforms.py:116-147) + the "vjepa" preset (:257-263) + options
    dict(input_size=(224, 224, 16), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0), num_blocks=1),
    dict(input_size=(224, 224, 16), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0), num_blocks=3),
    dict(input_size=(224, 224, 16), patch_size=(16, 16, 16), pred_mask_scale=(0.4, 0.9), aspect_ratio=(0.3, 3.0), num_blocks=1),
    dict(input_size=(384, 384, 256), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0), num_blocks=3),
    dict(input_size=(512, 512, 320), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0), num_blocks=3),
    dict(input_size=(96, 96, 96), patch_size=(16, 16, 16), num_blocks=2, max_keep=50, inv_block=True),
    dict(input_size=128, patch_size=16, num_blocks=1),
]


def load_reference_class():
    class _Base:
        def __init__(self, *a, **k):
            pass

    stub = types.ModuleType("monai.transforms")
    for name in ("CenterSpatialCropd", "EnsureChannelFirstd", "LoadImaged", "Orientationd", "ScaleIntensityRanged", "Spacingd",
                 "SpatialPadd", "ToTensord", "Transform", "MapTransform", "Compose"):
        setattr(stub, name, type(name, (_Base,), {}))
    sys.modules.setdefault("monai", types.ModuleType("monai"))
    sys.modules["monai.transforms"] = stub
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/src")
    from dataloader.transforms import VJEPAMaskGenerator  # noqa

    return VJEPAMaskGenerator


def digest(t: torch.Tensor) -> dict:
    """count, first 8 indices and sha256[:16] of the int64 little-endian index list (full lists would be 1 MB of JSON)."""
    v = t.reshape(-1).to(torch.int64).contiguous()
    return dict(n=int(v.numel()), first=v[:8].tolist(), sha16=hashlib.sha256(v.numpy().tobytes()).hexdigest()[:16])


def main():
    ref = load_reference_class()
    kats = []
    for ci, params in enumerate(CASES):
        for seed in (0, 1, 2024):
            torch.manual_seed(seed)
            gen = ref(**params)
            out = [gen({}) for _ in range(2)]  # two consecutive draws: the global stream position matters too
            kats.append(dict(case=ci, params={k: list(v) if isinstance(v, tuple) else v for k, v in params.items()}, seed=seed,
                             draws=[{key: digest(o[key + "_mask"]) for key in ("context", "target")} for o in out]))
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, "vjepa_mask_kat.json"), "w") as f:
        json.dump(kats, f)
    print("wrote", len(kats), "known answers")


if __name__ == "__main__":
    main()
