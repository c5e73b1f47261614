"""Host side of the data that enters the hot path.

* ``MaskGenerator`` / ``GenerateMask`` / ``collate_fn`` — same names, arguments, errors and RNG stream as the reference
  (``src/dataloader/mim.py:33-85``, ``src/run_mim.py:194-219``): the coarse cell permutation is drawn from numpy's legacy
  GLOBAL generator (``np.random.seed``), so for the same seed the masks are index-identical to the reference's.  The
  expansion to patch resolution and the visible/masked index lists are built on the GPU (``smbv_mask_upsample`` +
  ``smbv_mask_index``) by :meth:`MaskGenerator.device_batch`, without a host synchronisation.
* ``VJEPAMaskGenerator`` / ``vjepa_collate_fn`` — the V-JEPA counterparts (``src/dataloader/transforms.py:96-217``,
  ``src/run_vjepa.py:144-160``), same torch RNG stream -> index-identical context / target lists.
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


class VJEPAMaskGenerator:
    """Context / target index lists for V-JEPA (reference src/dataloader/transforms.py:96-217, used by the "vjepa"
    transform preset :244-265 and consumed at src/run_vjepa.py:116-131): `num_blocks` random boxes of one sampled size are
    cut out of the (depth, height, width) patch grid; the patches left over are the context, the cut-out ones the target.

    Same arguments and the same RNG consumption as the reference — one draw from torch's GLOBAL generator seeds a local
    one (block scale, then aspect ratio), the box corners come from the global generator again (depth, height, width per
    block) — so after the same `torch.manual_seed` the index lists are identical to the reference's
    (tests/golden/vjepa_mask_kat.json).  Like the reference, `input_size` / `patch_size` are (depth, height, width) in the
    order GIVEN and indices are flattened in that order.  `full_complement` / `pred_full_complement` raise in the reference
    (`torch.tensor(set(...))`, :200-205); here they return the sorted complement they describe."""

    def __init__(self, input_size=(224, 224, 160), patch_size=(16, 16, 16), pred_mask_scale=(0.2, 0.8), aspect_ratio=(0.3, 3.0),
                 num_blocks=1, max_keep=None, inv_block=False, full_complement=False, pred_full_complement=False):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 3
        if not isinstance(patch_size, tuple):
            patch_size = (patch_size,) * 3
        self.input_size, self.patch_size = input_size, patch_size
        self.depth, self.height, self.width = (input_size[i] // patch_size[i] for i in range(3))
        self.pred_mask_scale, self.aspect_ratio, self.num_blocks = pred_mask_scale, aspect_ratio, num_blocks
        self.max_keep, self.inv_block = max_keep, inv_block
        self.full_complement, self.pred_full_complement = full_complement, pred_full_complement

    def _block_size(self, generator):
        import math

        lo, hi = self.pred_mask_scale
        n_keep = int(self.depth * self.height * self.width * (lo + torch.rand(1, generator=generator).item() * (hi - lo)))
        lo, hi = self.aspect_ratio
        ar = lo + torch.rand(1, generator=generator).item() * (hi - lo)
        inv = 1.0 / ar
        d = int(round(math.pow(n_keep * ar * inv, 1 / 3)))
        return min(d, self.depth), min(int(round(d * ar)), self.height), min(int(round(d * inv)), self.width)

    def __call__(self, data: dict) -> dict:
        gen = torch.Generator()
        gen.manual_seed(torch.randint(0, 2**32, (1,)).item())
        d, h, w = self._block_size(gen)
        keep = torch.ones((self.depth, self.height, self.width), dtype=torch.bool)
        for _ in range(self.num_blocks):
            z0 = int(torch.randint(0, self.depth - d + 1, (1,)))
            y0 = int(torch.randint(0, self.height - h + 1, (1,)))
            x0 = int(torch.randint(0, self.width - w + 1, (1,)))
            keep[z0:z0 + d, y0:y0 + h, x0:x0 + w] = False
        keep = keep.flatten()
        context = torch.nonzero(keep).squeeze()    # .squeeze() as the reference: a single index becomes a 0-d tensor
        target = torch.nonzero(~keep).squeeze()
        if self.full_complement or self.pred_full_complement:
            everything = torch.ones(keep.numel(), dtype=torch.bool)
            if self.full_complement:
                everything[context.reshape(-1)] = False
                target = torch.nonzero(everything).squeeze()
            else:
                everything[target.reshape(-1)] = False
                context = torch.nonzero(everything).squeeze()
        if self.max_keep is not None:
            context, target = context[: self.max_keep], target[: self.max_keep]
        data["context_mask"], data["target_mask"] = (target, context) if self.inv_block else (context, target)
        return data


def vjepa_collate_fn(examples):
    """reference src/run_vjepa.py:144-160: unwrap single-element lists, stack the volumes, and share the masks of ONE
    randomly chosen example (python's global `random`) across the batch -> {"pixel_values_videos", "context_mask": [..],
    "target_mask": [..]} (lists of one [B, K] tensor, the form VJEPA2Model.forward takes)."""
    import random

    unpacked = []
    for ex in examples:
        while isinstance(ex, (list, tuple)) and len(ex) == 1:
            ex = ex[0]
        unpacked.append(ex)
    pixel_values = torch.stack([ex["image"] for ex in unpacked])
    chosen = random.choice(unpacked)
    context = torch.stack([chosen["context_mask"] for _ in unpacked])
    target = torch.stack([chosen["target_mask"] for _ in unpacked])
    return {"pixel_values_videos": pixel_values, "context_mask": [context], "target_mask": [target]}


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
