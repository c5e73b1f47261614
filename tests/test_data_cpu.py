"""CPU: host-side data entry points of the path — the product MaskGenerator /
GenerateMask / collate_fn against the known-answer vectors and the oracle, and the numpy restatement of the
ScaleIntensityRanged / SpatialPadd / CenterSpatialCropd / PermuteImage tail against hand-checked cases."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po
from oracle.mim_mask import OracleMaskGenerator
from smb_vision_b200.data import GenerateMask, MaskGenerator, collate_fn


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_product_mask_generator_known_answers(golden_dir):
    for kat in json.load(open(os.path.join(golden_dir, "mask_kat.json"))):
        np.random.seed(kat["seed"])
        g = MaskGenerator(kat["input_size"], kat["depth"], kat["mask_patch_size"], kat["model_patch_size"], kat["mask_ratio"])
        m = g()
        assert m.dtype == torch.bool and m.shape == (kat["n"],) and int(m.sum()) == kat["n_mask"] == g.num_masked
        assert sha16(m.numpy().astype(np.uint8)) == kat["sha_fine"]
        assert torch.nonzero(m)[:10, 0].tolist() == kat["first_masked"]


def test_product_mask_generator_follows_the_reference_rng_stream():
    """same global-RNG consumption as the reference: consecutive calls == consecutive calls of the oracle restatement."""
    np.random.seed(3)
    a = MaskGenerator(96, 96, 32, 16, 0.65)
    got = [a().numpy() for _ in range(3)]
    np.random.seed(3)
    b = OracleMaskGenerator(96, 96, 32, 16, 0.65)
    want = [b() for _ in range(3)]
    assert all(np.array_equal(x, y) for x, y in zip(got, want))
    for args in [(100, 96, 32, 16), (96, 100, 32, 16), (96, 96, 32, 12)]:  # src/dataloader/mim.py:47-52
        with pytest.raises(ValueError):
            MaskGenerator(*args, 0.5)


def test_generate_mask_and_collate():
    np.random.seed(0)
    t = GenerateMask(96, 96, 32, 16, 0.65)
    exs = [[t({"image": torch.full((96, 1, 96, 96), float(i))})] for i in range(3)]  # nested single-element lists (run_mim.py:196-201)
    batch = collate_fn(exs)
    assert set(batch) == {"pixel_values", "bool_masked_pos"}
    assert batch["pixel_values"].shape == (3, 96, 1, 96, 96) and batch["bool_masked_pos"].shape == (3, 216)
    assert batch["bool_masked_pos"].dtype == torch.bool and batch["pixel_values"][2, 0, 0, 0, 0] == 2.0
    assert (batch["bool_masked_pos"].sum(1) == 144).all()


def test_preprocess_oracle_hand_checked_cases():
    # intensity: HU -1000 -> 0, 0 -> 0.5, 1000 -> 1, clipping outside
    v = po.scale_intensity_range(np.array([-2000, -1000, 0, 500, 1000, 3000], dtype=np.int16))
    assert v.dtype == np.float32 and np.array_equal(v, np.array([0, 0, 0.5, 0.75, 1, 1], dtype=np.float32))
    # symmetric pad: 5 -> 8 pads (1, 2); centre crop 9 -> 4 starts at 9//2 - 4//2 = 2
    a = np.arange(5, dtype=np.float32) + 1
    assert np.array_equal(po.spatial_pad(a, (8,)), np.array([0, 1, 2, 3, 4, 5, 0, 0], dtype=np.float32))
    b = np.arange(9, dtype=np.float32)
    assert np.array_equal(po.center_spatial_crop(b, (4,)), np.array([2, 3, 4, 5], dtype=np.float32))
    # more odd / even combinations, worked out by hand from MONAI's published rules (monai/transforms/croppad/array.py:
    # SpatialPad.compute_pad_width: width = max(target - size, 0), (width // 2, width - width // 2) -> the EXTRA voxel of an
    # odd width goes to the END; CenterSpatialCrop -> SpatialCrop(roi_center = size // 2, roi_size): start = center - roi // 2)
    assert np.array_equal(po.spatial_pad(np.arange(6, dtype=np.float32) + 1, (9,)), np.array([0, 1, 2, 3, 4, 5, 6, 0, 0], dtype=np.float32))  # width 3 -> (1, 2)
    assert np.array_equal(po.spatial_pad(np.arange(5, dtype=np.float32) + 1, (9,)), np.array([0, 0, 1, 2, 3, 4, 5, 0, 0], dtype=np.float32))  # width 4 -> (2, 2)
    assert np.array_equal(po.spatial_pad(np.arange(3, dtype=np.float32) + 1, (2,)), np.array([1, 2, 3], dtype=np.float32))                    # never shrinks
    assert np.array_equal(po.center_spatial_crop(np.arange(10, dtype=np.float32), (5,)), np.array([3, 4, 5, 6, 7], dtype=np.float32))         # 10 // 2 - 5 // 2 = 3
    assert np.array_equal(po.center_spatial_crop(np.arange(9, dtype=np.float32), (5,)), np.array([2, 3, 4, 5, 6], dtype=np.float32))           # 9 // 2 - 5 // 2 = 2
    assert np.array_equal(po.center_spatial_crop(np.arange(7, dtype=np.float32), (2,)), np.array([2, 3], dtype=np.float32))                    # 7 // 2 - 2 // 2 = 2
    assert np.array_equal(po.center_spatial_crop(np.arange(3, dtype=np.float32), (5,)), np.arange(3, dtype=np.float32))                        # roi larger than the axis: untouched
    # pad THEN crop on different axes of one volume (mim.py:161-170 order): X 5 -> 8 (pad 1, 2), Y 9 -> 4 (crop from 2), Z kept
    vol = np.arange(5 * 9 * 2, dtype=np.float32).reshape(5, 9, 2)
    pc = po.center_spatial_crop(po.spatial_pad(vol, (8, 4, 2)), (8, 4, 2))
    assert pc.shape == (8, 4, 2) and np.array_equal(pc[1:6], vol[:, 2:6]) and not pc[0].any() and not pc[6:].any()
    # full tail: layout [X,Y,Z] -> [Z,1,X,Y], identity geometry when sizes already match
    raw = (np.arange(4 * 4 * 2, dtype=np.float32).reshape(4, 4, 2) * 10 - 100)
    out = po.prepare_volume(raw, 4, 2)
    assert out.shape == (2, 1, 4, 4)
    assert out[1, 0, 2, 3] == po.scale_intensity_range(raw)[2, 3, 1]
