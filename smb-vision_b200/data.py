"""Host side of the data that enters the hot path.

* ``MaskGenerator`` / ``GenerateMask`` / ``collate_fn`` — same names, arguments, errors and RNG stream as the reference
  (``src/dataloader/mim.py:33-85``, ``src/run_mim.py:194-219``): the coarse cell permutation is drawn from numpy's legacy
  GLOBAL generator (``np.random.seed``), so for the same seed the masks are index-identical to the reference's.  The
  expansion to patch resolution and the visible/masked index lists are built on the GPU (``smbv_mask_upsample`` +
  ``smbv_mask_index``) by :meth:`MaskGenerator.device_batch`, without a host synchronisation.
* ``VolumePreprocessor`` — the tail of ``MIMDataset.train_transforms`` (``mim.py:154-170`` + ``PermuteImage`` :86-91):
  ScaleIntensityRanged -> SpatialPadd -> CenterSpatialCropd -> permute, as ONE kernel over the raw resampled volume
  (fp32, or the int16 HU a NIfTI CT stores: half the host->device bytes).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import ops
from ._lib import SmbvError


class MaskGenerator:
    """reference src/dataloader/mim.py:33-69."""

    def __init__(self, input_size=224, depth=96, mask_patch_size=32, model_patch_size=16, mask_ratio=0.6):
        self.input_size, self.depth = input_size, depth
        self.mask_patch_size, self.model_patch_size, self.mask_ratio = mask_patch_size, model_patch_size, mask_ratio
        if self.input_size % self.mask_patch_size != 0:
            raise ValueError("Input size must be divisible by mask patch size")
        if self.depth % self.mask_patch_size != 0:
            raise ValueError("Depth must be divisible by mask patch size")
        if self.mask_patch_size % self.model_patch_size != 0:
            raise ValueError("Mask patch size must be divisible by model patch size")
        self.rand_size = self.input_size // self.mask_patch_size
        self.rand_depth = self.depth // self.mask_patch_size
        self.scale = self.mask_patch_size // self.model_patch_size
        self.token_count = self.rand_size**2 * self.rand_depth
        self.mask_count = int(np.ceil(self.token_count * self.mask_ratio))
        self.num_patches = self.token_count * self.scale**3
        self.num_masked = self.mask_count * self.scale**3

    def coarse(self) -> np.ndarray:
        """uint8 [rand_depth, rand_size, rand_size]; consumes exactly one np.random.permutation (mim.py:62-66)."""
        mask_idx = np.random.permutation(self.token_count)[: self.mask_count]
        mask = np.zeros(self.token_count, dtype=np.uint8)
        mask[mask_idx] = 1
        return mask.reshape((self.rand_depth, self.rand_size, self.rand_size))

    def __call__(self) -> torch.Tensor:
        """bool [N] on the host, exactly what the reference returns (mim.py:67-69)."""
        m = self.coarse()
        m = m.repeat(self.scale, axis=0).repeat(self.scale, axis=1).repeat(self.scale, axis=2)
        return torch.from_numpy(m.reshape(-1)).bool()

    def device_batch(self, batch: int, device):
        """`batch` fresh masks as the model's mask pack (fine uint8 [B,N], vis_idx, msk_idx, slot, n_vis, n_mask): only the
        coarse cells (2.5 KB per volume at 512x512x320) cross PCIe; upsampling and index lists run on the GPU and the
        counts are known on the host, so nothing synchronises."""
        coarse = torch.from_numpy(np.stack([self.coarse() for _ in range(batch)]))
        fine = ops.mask_upsample(coarse.to(device, non_blocking=True), self.scale)
        vis, msk, slot, _ = ops.mask_index(fine)
        return fine, vis, msk, slot, self.num_patches - self.num_masked, self.num_masked


class GenerateMask:
    """reference src/dataloader/mim.py:72-85: attaches ``inputs["mask"]``."""

    def __init__(self, input_size=224, depth=96, mask_patch_size=32, model_patch_size=16, mask_ratio=0.75):
        self.mask_generator = MaskGenerator(input_size, depth, mask_patch_size, model_patch_size, mask_ratio)

    def __call__(self, inputs):
        inputs["mask"] = self.mask_generator()
        return inputs


def collate_fn(examples):
    """reference src/run_mim.py:194-219: unwrap single-element lists, stack -> {"pixel_values", "bool_masked_pos"}."""
    unpacked = []
    for ex in examples:
        while isinstance(ex, (list, tuple)) and len(ex) == 1:
            ex = ex[0]
        unpacked.append(ex)
    pixel_values = torch.stack([ex["image"] for ex in unpacked])
    masks = torch.stack([ex["mask"] for ex in unpacked])
    return {"pixel_values": pixel_values, "bool_masked_pos": masks}


class VolumePreprocessor:
    """ScaleIntensityRanged(a_min, a_max, b_min, b_max, clip) -> SpatialPadd((img, img, depth)) ->
    CenterSpatialCropd((img, img, depth)) -> PermuteImage, on the GPU (defaults = src/dataloader/mim.py:154-170).

    ``__call__(raw)``: raw = one resampled volume ``[X, Y, Z]`` or ``[1, X, Y, Z]`` (channel first, as after
    EnsureChannelFirstd/Spacingd), fp32 or int16, on the host (pinned memory makes the copy asynchronous) or already on
    the device.  Returns fp32 ``[depth, 1, img, img]`` on the device — one sample of ``pixel_values``."""

    def __init__(self, img_size: int, depth: int, a_min=-1000.0, a_max=1000.0, b_min=0.0, b_max=1.0, clip=True, device="cuda"):
        self.img_size, self.depth = int(img_size), int(depth)
        self.a_min, self.a_max, self.b_min, self.b_max, self.clip = float(a_min), float(a_max), float(b_min), float(b_max), bool(clip)
        self.device = torch.device(device)

    def __call__(self, raw: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if isinstance(raw, np.ndarray):
            raw = torch.from_numpy(raw)
        if raw.dim() == 4:
            if raw.shape[0] != 1:
                raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
            raw = raw[0]
        if raw.dim() != 3:
            raise ValueError(f"expected a [X,Y,Z] or [1,X,Y,Z] volume, got shape {tuple(raw.shape)}")
        if raw.dtype not in (torch.float32, torch.int16):
            raise SmbvError(f"VolumePreprocessor: dtype {raw.dtype} not supported (float32 or int16)")
        raw = raw.to(self.device, non_blocking=True).contiguous()
        res = ops.prepare_volume(raw, self.img_size, self.img_size, self.depth, self.a_min, self.a_max, self.b_min, self.b_max,
                                 self.clip, out=None if out is None else out.view(self.depth, self.img_size, self.img_size))
        return res.view(self.depth, 1, self.img_size, self.img_size)

    def batch(self, raws: Sequence[torch.Tensor]) -> torch.Tensor:
        """stack of `len(raws)` prepared volumes: fp32 [B, depth, 1, img, img] (= ``pixel_values``)."""
        out = torch.empty((len(raws), self.depth, 1, self.img_size, self.img_size), dtype=torch.float32, device=self.device)
        for b, r in enumerate(raws):
            self(r, out=out[b])
        return out
