"""Data-parallel plumbing for the grounding path: images are independent, so each rank takes a contiguous shard of the
batch and runs the whole path locally -- there is NO collective on the hot path.  The only exchange is the gather of
the packed results (u8 masks, scores, IoU predictions, depth) for scoring, mirroring the reference's per-rank
``DistributedSampler`` evaluation + scalar ``all_reduce`` (evaluation_walkgpt.py:393-402, 956-958)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(images: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets: Sequence[int], rank: int, world: int):
    """Slice (images [N,...], seg_hidden [sum S, H], seg_offsets [N+1]) to this rank's images and their [SEG] rows."""
    n = images.shape[0]
    assert len(seg_offsets) == n + 1
    lo, hi = shard_range(n, rank, world)
    p0, p1 = int(seg_offsets[lo]), int(seg_offsets[hi])
    local_offsets = [int(seg_offsets[i]) - p0 for i in range(lo, hi + 1)]
    return images[lo:hi], seg_hidden[p0:p1], local_offsets


def gather_results(local: Dict[str, torch.Tensor], keys: Sequence[str] = ("masks", "scores", "iou", "depth")) -> Dict[str, torch.Tensor]:
    """all_gather of per-prompt results with ragged first dimensions (prompt counts differ per rank).
    Works with NCCL (device tensors, NVLink) and gloo (CPU tensors, used by the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: local[k] for k in keys if k in local}
    world = dist.get_world_size()
    some = next(local[k] for k in keys if k in local)
    n_local = torch.tensor([some.shape[0]], dtype=torch.int64, device=some.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    out: Dict[str, torch.Tensor] = {}
    for k in keys:
        if k not in local:
            continue
        t = local[k]
        pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        out[k] = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return out
