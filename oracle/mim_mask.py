"""Oracle restatement of the reference MIM mask generator (test infrastructure only).

Follows ``/root/reference/src/dataloader/mim.py:33-69`` (``MaskGenerator``), which
cannot be imported here because the module imports MONAI at the top
(``mim.py:7-22``).  It uses numpy's *legacy global* RNG, so results after
``np.random.seed(s)`` are stable across numpy versions; the known-answer vectors
in ``tests/golden/mask_kat.json`` pin it.
"""
from __future__ import annotations

import numpy as np


class OracleMaskGenerator:
    """Coarse random mask of ``mask_patch_size``^3 cells, repeated to patch resolution.

    reference: src/dataloader/mim.py:33-59 (ctor / checks), :61-69 (__call__).
    """

    def __init__(self, input_size=224, depth=96, mask_patch_size=32, model_patch_size=16, mask_ratio=0.6):
        if input_size % mask_patch_size != 0:  # mim.py:47-48
            raise ValueError("Input size must be divisible by mask patch size")
        if depth % mask_patch_size != 0:  # mim.py:49-50
            raise ValueError("Depth must be divisible by mask patch size")
        if mask_patch_size % model_patch_size != 0:  # mim.py:51-52
            raise ValueError("Mask patch size must be divisible by model patch size")
        self.rand_size = input_size // mask_patch_size  # mim.py:54
        self.rand_depth = depth // mask_patch_size  # mim.py:55
        self.scale = mask_patch_size // model_patch_size  # mim.py:56
        self.token_count = self.rand_size**2 * self.rand_depth  # mim.py:58
        self.mask_count = int(np.ceil(self.token_count * mask_ratio))  # mim.py:59

    def coarse(self) -> np.ndarray:
        """uint8[rand_depth, rand_size, rand_size]; consumes one np.random.permutation (mim.py:62-66)."""
        idx = np.random.permutation(self.token_count)[: self.mask_count]
        m = np.zeros(self.token_count, dtype=np.uint8)
        m[idx] = 1
        return m.reshape(self.rand_depth, self.rand_size, self.rand_size)

    @staticmethod
    def upsample(coarse: np.ndarray, scale: int) -> np.ndarray:
        """repeat x scale on the 3 axes, flatten z-major (mim.py:67-69)."""
        return coarse.repeat(scale, 0).repeat(scale, 1).repeat(scale, 2).reshape(-1)

    def __call__(self) -> np.ndarray:
        """bool[N] fine mask, token order (tz, ty, tx) z-major."""
        return self.upsample(self.coarse(), self.scale).astype(bool)
