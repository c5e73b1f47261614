"""Oracle restatement (numpy, fp32) of the tail of the reference data pipeline — TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/src/dataloader/mim.py:154-170`` (ScaleIntensityRanged, SpatialPadd, CenterSpatialCropd) and
``:86-91`` (PermuteImage).  The transforms themselves live in MONAI, a third-party dependency the reference does not
vendor or pin (``pyproject.toml:30``: ``"monai"``) and that is absent from this image, so their PUBLISHED algorithms are
restated here:

* ``ScaleIntensityRange.__call__``: ``img = (img - a_min) / (a_max - a_min)``; ``img = img * (b_max - b_min) + b_min``;
  ``clip(img, b_min, b_max)`` — fp32 tensor ops with Python-float scalars;
* ``SpatialPad(method="symmetric", mode="constant")``: per axis ``width = max(target - size, 0)``, pad
  ``(width // 2, width - width // 2)`` with 0 (after the intensity scaling, so padding is 0.0 in output units);
* ``CenterSpatialCrop``: ``center = size // 2``, ``start = max(center - roi // 2, 0)``, ``end = start + roi``.

PARITY UNPINNED for this file: there is no MONAI here to generate fixtures from and the reference has no test for it.
"""
from __future__ import annotations

import numpy as np


def scale_intensity_range(img: np.ndarray, a_min=-1000.0, a_max=1000.0, b_min=0.0, b_max=1.0, clip=True) -> np.ndarray:
    img = img.astype(np.float32)
    img = (img - np.float32(a_min)) / np.float32(a_max - a_min)
    img = img * np.float32(b_max - b_min) + np.float32(b_min)
    if clip:
        img = np.clip(img, np.float32(b_min), np.float32(b_max))
    return img.astype(np.float32)


def spatial_pad(img: np.ndarray, size) -> np.ndarray:
    pads = []
    for s, r in zip(img.shape, size):
        w = max(r - s, 0)
        pads.append((w // 2, w - w // 2))
    return np.pad(img, pads, mode="constant", constant_values=0)


def center_spatial_crop(img: np.ndarray, roi) -> np.ndarray:
    sl = []
    for s, r in zip(img.shape, roi):
        start = max(s // 2 - r // 2, 0)
        sl.append(slice(start, start + r))
    return img[tuple(sl)]


def prepare_volume(raw_xyz: np.ndarray, img_size: int, depth: int, **kw) -> np.ndarray:
    """[X,Y,Z] -> fp32 [depth, 1, img, img] (one sample of pixel_values): mim.py:154-170 then permute(3,0,1,2) of [1,X,Y,Z]."""
    v = scale_intensity_range(raw_xyz, **kw)
    v = spatial_pad(v, (img_size, img_size, depth))
    v = center_spatial_crop(v, (img_size, img_size, depth))
    return np.ascontiguousarray(v[None].transpose(3, 0, 1, 2))
