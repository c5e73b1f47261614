"""Embedding output formats either side of inference — host-side, overlapped with GPU compute.

* ``.npy``  : ``np.save(<stem>.npy, last_hidden_state.cpu().numpy())`` — fp32 ``[1, N, d]`` (reference
  ``src/run_inference.py:89-96``, stem = file name without ``.nii``/``.gz``, :107-109);
* parquet  : one snappy file per uid under ``<save_dir>/model_id=<id>/<uid>.parquet`` with columns ``uid``,
  ``embedding`` (flattened fp32), ``embedding_shape``, ``model_id`` (reference
  ``scripts/inference/inspect/run_inspect.py:140-175``), and resume-by-existing-file (:33-50).

Writers run on a small thread pool (numpy / pyarrow release the GIL while writing), fed from the pinned host buffers
``EmbeddingRunner.embed_stream`` yields, so disk I/O overlaps the next volumes' H2D + compute + D2H.
"""
from __future__ import annotations

import os
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional, Set

import numpy as np
import torch


def npy_stem(image_path: str) -> str:
    """reference src/run_inference.py:107-108: ``Path(p).stem.replace(".nii", "")`` ("ct_01.nii.gz" -> "ct_01")."""
    base = os.path.basename(image_path)
    stem = base[: base.rfind(".")] if "." in base else base
    return stem.replace(".nii", "")


def processed_uids(save_dir: str) -> Set[str]:
    """uids that already have a parquet file under any ``model_id=*`` directory (run_inspect.py:33-41)."""
    done: Set[str] = set()
    if os.path.exists(save_dir):
        for model_dir in os.listdir(save_dir):
            model_path = os.path.join(save_dir, model_dir)
            if os.path.isdir(model_path):
                for f in os.listdir(model_path):
                    if f.endswith(".parquet"):
                        done.add(f.replace(".parquet", ""))
    return done


def unprocessed_files(image_dir: str, save_dir: str) -> List[Dict[str, str]]:
    """``[{"image": path, "uid": uid}]`` for every ``*.nii.gz`` in image_dir without a parquet yet (run_inspect.py:43-50)."""
    done = processed_uids(save_dir)
    files = []
    for filename in os.listdir(image_dir):
        if filename.endswith(".nii.gz"):
            uid = filename.replace(".nii.gz", "")
            if uid not in done:
                files.append({"image": os.path.join(image_dir, filename), "uid": uid})
    return files


def write_npy(path: str, emb: np.ndarray) -> str:
    np.save(path, emb)  # run_inference.py:91-92 (np.save appends ".npy" when missing, like the reference call)
    return path if path.endswith(".npy") else path + ".npy"


def write_parquet(save_dir: str, uid: str, emb: np.ndarray, model_id: str) -> str:
    """Same file the reference's ``pd.DataFrame({...}).to_parquet(compression="snappy")`` writes (run_inspect.py:150-172):
    readable by ``pd.read_parquet`` into the same columns/values; built with pyarrow straight from the numpy buffer
    (no per-element Python objects)."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    if emb.ndim == 3 and emb.shape[0] == 1:  # `.squeeze(0)` of run_inspect.py:150
        emb = emb[0]
    flat = np.ascontiguousarray(emb, dtype=np.float32).reshape(-1)
    table = pa.table({
        "uid": pa.array([uid], type=pa.string()),
        "embedding": pa.ListArray.from_arrays(pa.array([0, flat.size], type=pa.int32()), pa.array(flat)),
        "embedding_shape": pa.array([list(emb.shape)], type=pa.list_(pa.int64())),
        "model_id": pa.array([model_id], type=pa.string()),
    })
    model_dir = os.path.join(save_dir, f"model_id={model_id}")
    os.makedirs(model_dir, exist_ok=True)
    out = os.path.join(model_dir, f"{uid}.parquet")
    tmp = out + ".tmp"
    pq.write_table(table, tmp, compression="snappy")
    os.replace(tmp, out)  # a crash never leaves a half-written file that the resume scan would count as done
    return out


class EmbeddingWriter:
    """Asynchronous writer: ``submit(uid_or_path, emb)`` copies the (pinned, soon to be reused) host tensor and hands it
    to a worker; ``close()`` waits and re-raises the first failure.  fmt = "parquet" (run_inspect) or "npy" (run_inference)."""

    def __init__(self, save_dir: str, fmt: str = "parquet", model_id: Optional[str] = None, workers: int = 2, max_pending: int = 4):
        if fmt not in ("parquet", "npy"):
            raise ValueError(f"unknown format {fmt!r}")
        if fmt == "parquet" and not model_id:
            raise ValueError("parquet output needs model_id (the model_id=<id>/ directory of run_inspect.py:164)")
        self.save_dir, self.fmt, self.model_id = save_dir, fmt, model_id
        os.makedirs(save_dir, exist_ok=True)
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending: List[Future] = []
        self.max_pending = max_pending
        self.written: List[str] = []

    def _drain(self, keep: int) -> None:
        while len(self.pending) > keep:
            self.written.append(self.pending.pop(0).result())

    def submit(self, key: str, emb: torch.Tensor) -> None:
        arr = emb.detach().to("cpu", torch.float32).numpy().copy()  # `.float().cpu().numpy()` of the reference
        if self.fmt == "npy":
            fut = self.pool.submit(write_npy, os.path.join(self.save_dir, npy_stem(key) + ".npy"), arr)
        else:
            fut = self.pool.submit(write_parquet, self.save_dir, key, arr, self.model_id)
        self.pending.append(fut)
        self._drain(self.max_pending)

    def write_stream(self, keys: Iterable[str], embeddings: Iterable[torch.Tensor]) -> List[str]:
        for k, e in zip(keys, embeddings):
            self.submit(k, e)
        return self.close()

    def close(self) -> List[str]:
        self._drain(0)
        self.pool.shutdown(wait=True)
        return self.written
