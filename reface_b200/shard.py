"""Host-side multi-GPU plumbing: the path shards by independent swap pairs -- contiguous slices of the
batch per rank, ONE broadcast of the flat checkpoint at start-up, results gathered by rank order.  No collective
runs inside the DDIM loop."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of `total` items for `rank` (first `total % world` ranks get one more)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int):
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def video_segments(n_frames: int, chunk: int, rank: int, world: int):
    """Video mode (BASELINE configs[4]; scripts/inference_swap_video.py:560-724): frames are streamed in contiguous
    chunks of `chunk` frames, chunk k goes to rank k % world; every frame keeps its index (the reference's
    `segment_id`) so that results can be put back in order (inference_swap_video.py:710-713).
    Returns this rank's list of (first_frame, last_frame_exclusive)."""
    chunks = [(lo, min(n_frames, lo + chunk)) for lo in range(0, n_frames, chunk)]
    return [c for k, c in enumerate(chunks) if k % world == rank]


def broadcast_checkpoint(flat: torch.Tensor, src: int = 0):
    """One collective for the whole (flat fp32) checkpoint; NCCL over NVLink on GPUs, gloo in the CPU tests."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src)
    return flat


def gather_batch(local: torch.Tensor, total: int):
    """All ranks' output slices concatenated in rank order (sizes follow shard_range)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    bufs = [torch.empty((hi - lo, *local.shape[1:]), dtype=local.dtype, device=local.device) for lo, hi in sizes]
    dist.all_gather(bufs, local) if len({hi - lo for lo, hi in sizes}) == 1 else _uneven_gather(bufs, local, sizes)
    return torch.cat(bufs, 0)


def _uneven_gather(bufs, local, sizes):
    rank = dist.get_rank()
    for r, b in enumerate(bufs):
        if r == rank:
            b.copy_(local)
        dist.broadcast(b, r)
