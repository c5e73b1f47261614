"""Host-side multi-GPU logic (one process per GPU, torch.distributed): volume sharding for embedding inference (no
collective) and the bucketed gradient reducer used by the data-parallel MIM step.

Reference strategy: inference = contiguous per-GPU chunks of the dataset list, each process its own model copy
(scripts/inference/inspect/run_inspect.py:206-241); training = DDP gradient all-reduce.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_volumes(items: Sequence, rank: int, world: int) -> List:
    """Contiguous chunk of `items` for `rank` (run_inspect.py:218-221: chunk = ceil(len / n_gpus))."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    chunk = (len(items) + world - 1) // world
    return list(items[rank * chunk:(rank + 1) * chunk])


class BucketReducer:
    """Mean-all-reduce of contiguous slices ("buckets") of one flat fp32 gradient buffer, launched asynchronously as
    backward completes each bucket and folded back at `finish()`.

    wire_dtype bf16 halves the NVLink bytes (north star); fp32 keeps exact sums (what the reference's DDP sends).
    `cast_down(src_f32, dst_wire)` / `cast_up(src_wire, dst_f32, scale)` default to torch ops so the logic is testable
    on CPU with gloo; the CUDA path passes the library's cast kernels.
    """

    def __init__(self, flat: torch.Tensor, bounds: Sequence[int], group=None, wire_dtype=torch.bfloat16, cast_down=None, cast_up=None):
        self.flat, self.bounds, self.group, self.wire_dtype = flat, list(bounds), group, wire_dtype
        if group is False:  # explicit "no communication" (independent replicas: bench.py's no-collective comparison run)
            self.world, self.group = 1, None
        else:
            self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.wire = torch.empty(flat.numel(), dtype=wire_dtype, device=flat.device) if (self.world > 1 and wire_dtype != flat.dtype) else None
        self.cast_down = cast_down or (lambda s, d: d.copy_(s))
        self.cast_up = cast_up or (lambda s, d, scale: d.copy_(s.to(d.dtype) * scale))
        self.pending: List[Tuple[object, int, int]] = []
        # overlap=True: every bucket is exchanged as soon as backward has produced it (NCCL's CTAs run beside the backward kernels);
        # False (SMBV_DP_OVERLAP=0): ONE all-reduce over the whole buffer after backward.  The backward kernels are persistent
        # one-CTA-per-SM grids: every SM NCCL holds delays one of their CTAs — and with it the whole kernel — by the exchange's
        # duration, so the overlapped form is not automatically the faster one (measured: DESIGN.md section 7).
        self.overlap = os.environ.get("SMBV_DP_OVERLAP", "1") != "0"
        # a bucket whose exchange was launched `fold_lag` buckets ago (one transformer block of backward each: ~0.7 ms against a
        # ~0.1 ms exchange) is folded back (wait + cast-up / scale) from reduce_bucket itself, so that finish() — the non-overlapped
        # tail of the step — only has the last `fold_lag` buckets left instead of all of them
        self.fold_lag = 2

    def _fold(self, entry) -> None:
        work, lo, hi = entry
        work.wait()
        if self.wire is not None:
            self.cast_up(self.wire[lo:hi], self.flat[lo:hi], 1.0 / self.world)
        else:
            self.flat[lo:hi].mul_(1.0 / self.world)

    def reduce_bucket(self, i: int) -> None:
        lo, hi = self.bounds[i], self.bounds[i + 1]
        if self.world == 1 or hi == lo:
            return
        if not self.overlap:  # one exchange over the whole buffer at finish(): see __init__
            return
        while len(self.pending) >= self.fold_lag:
            self._fold(self.pending.pop(0))
        if self.wire is not None:
            w = self.wire[lo:hi]
            self.cast_down(self.flat[lo:hi], w)
        else:
            w = self.flat[lo:hi]
        self.pending.append((dist.all_reduce(w, group=self.group, async_op=True), lo, hi))

    def finish(self) -> None:
        if self.world > 1 and not self.overlap:
            lo, hi = self.bounds[0], self.bounds[-1]
            if self.wire is not None:
                w = self.wire[lo:hi]
                self.cast_down(self.flat[lo:hi], w)
            else:
                w = self.flat[lo:hi]
            self.pending.append((dist.all_reduce(w, group=self.group, async_op=True), lo, hi))
        for entry in self.pending:
            self._fold(entry)
        self.pending.clear()
