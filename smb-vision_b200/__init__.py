"""smb_vision_b200 — B200-native (sm_100a) implementation of smb-vision's 3D-ViT MIM hot path.

Importable as ``smb_vision_b200`` (the directory is ``smb-vision_b200/``; a symlink provides the Python name).
Host code is Python/PyTorch (device memory, streams, torch.distributed); the compute is hand-written CUDA
behind the C ABI in ``include/smbv_b200.h`` (``lib/libsmbv_b200.so``).  No CPU / library fallback.
"""
from ._lib import LIB_PATH, SmbvError, load  # noqa: F401

__all__ = ["LIB_PATH", "SmbvError", "load", "B200VideoMAEModel", "B200VideoMAEForPreTraining",
           "B200VideoMAEForVideoClassification", "B200VJEPA2Model", "DataParallelStep"]


def __getattr__(name):  # lazy: importing the package must not import torch-heavy modules unless asked
    if name in ("B200VideoMAEModel", "B200VideoMAEForPreTraining", "B200VideoMAEForVideoClassification"):
        from . import modeling

        return getattr(modeling, name)
    if name == "B200VJEPA2Model":
        from .vjepa import B200VJEPA2Model

        return B200VJEPA2Model
    if name == "DataParallelStep":
        from .training import DataParallelStep

        return DataParallelStep
    raise AttributeError(name)
