"""Multi-GPU plumbing: embarrassingly-parallel prompt sharding, one process per GPU.

The reference's inference is single-process / single-GPU (tango.py:10); samples are independent (GroupNorm and
LayerNorm are per-sample), so the only exchanges are a one-time weight broadcast from rank 0 and the final all-gather of
int16 waveforms. Both go through torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests); there is
no per-step collective, hence nothing to fuse into a kernel.
"""
from __future__ import annotations

from typing import Dict, Tuple

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


def allgather_waves(wave: np.ndarray, device=None) -> np.ndarray:
    """All-gather of per-rank int16 waveform blocks [n_r, L] in rank order -> [sum n_r, L] on every rank: one NCCL
    all_gather of a device int16 tensor over NVLink (gloo / CPU tensors in the CPU tests). Ranks may hold different
    (even zero) counts: blocks are padded to the largest count, the counts travel in a first tiny all_gather."""
    if not is_dist() or world_size() == 1:
        return np.asarray(wave)
    world = world_size()
    dev = torch.device(device) if device is not None and dist.get_backend() != "gloo" else torch.device("cpu")
    w = torch.from_numpy(np.ascontiguousarray(wave)).to(dev)
    meta = torch.tensor([w.shape[0], w.shape[1] if w.dim() == 2 else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    counts = [int(m[0]) for m in metas]
    length = max(int(m[1]) for m in metas)
    nmax = max(counts)
    pad = torch.zeros(nmax, length, dtype=torch.int16, device=dev)
    if w.numel():
        pad[:w.shape[0]] = w
    raw = pad.view(torch.uint8)                      # bytes travel: gloo (CPU tests) has no int16 collectives
    parts = [torch.zeros_like(raw) for _ in range(world)]
    dist.all_gather(parts, raw)
    return torch.cat([p.view(torch.int16)[:c] for p, c in zip(parts, counts)], 0).cpu().numpy()


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
