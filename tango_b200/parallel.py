"""Multi-GPU plumbing: embarrassingly-parallel prompt sharding, one process per GPU.

The reference's inference is single-process / single-GPU (tango.py:10); samples are independent (GroupNorm and
LayerNorm are per-sample), so the only exchanges are a one-time weight broadcast from rank 0 and the final gather of
int16 waveforms. Both go through torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests); there is
no per-step collective, hence nothing to fuse into a kernel.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank() -> int:
    return dist.get_rank() if is_dist() else 0


def world_size() -> int:
    return dist.get_world_size() if is_dist() else 1


def shard_range(n: int, r: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n items: rank r gets [lo, hi); the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def shard_rows(t: torch.Tensor, r: int, world: int) -> torch.Tensor:
    """Rows of a full-batch tensor that belong to rank r (used to slice full-batch noise so that results do not
    depend on the GPU count, SURVEY.md §7 'RNG contract')."""
    lo, hi = shard_range(t.shape[0], r, world)
    return t[lo:hi]


def broadcast_state_dict(sd: Dict[str, torch.Tensor], src: int = 0, device=None) -> Dict[str, torch.Tensor]:
    """One-time weight broadcast from `src` (rank 0 loads the checkpoint, the others receive it over NCCL/NVLink).
    Every rank must pass a dict with the same keys/shapes (non-src contents are overwritten)."""
    if not is_dist() or world_size() == 1:
        return sd
    out = {}
    for k in sorted(sd):
        t = sd[k].to(device) if device is not None else sd[k]
        t = t.contiguous()
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def gather_waves(waves: Sequence[np.ndarray], dst: int = 0) -> List[np.ndarray]:
    """Gather per-rank lists of int16 waveforms on `dst` in rank order (other ranks get their own list back)."""
    if not is_dist() or world_size() == 1:
        return list(waves)
    local = [np.asarray(w) for w in waves]
    bucket = [None] * world_size() if rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if rank() != dst:
        return local
    out: List[np.ndarray] = []
    for part in bucket:
        out += list(part)
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Max-reduce a host float (e.g. an elapsed time) over all ranks."""
    if not is_dist() or world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not is_dist() or world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
