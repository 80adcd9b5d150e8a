"""Window-granular sharding across ranks (one process per GPU) with a host-side ordered gather.

The reference shards at *file* granularity (CallConsensusInterface.py:138-140) and merges results
implicitly when stitch globs every ``*.hdf``.  For a window tensor already in memory (benchmarks,
the 3 M-window synthetic contig set) windows are independent, so the index space is split in
contiguous, balanced ranges; the only communication is the final gather of uint8 labels
(2 KB/window) to rank 0 over the process group (gloo or nccl) -- no collective on the compute path.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_windows, world_size, rank):
    """Contiguous balanced split: the first (n % world) ranks get one extra window."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    base, extra = divmod(int(n_windows), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_labels(local_base, local_rle, n_windows, group=None):
    """Ordered gather of per-rank label arrays [n_local, T] uint8 on rank 0.

    Returns (base [n_windows, T], rle [n_windows, T]) numpy arrays on rank 0, (None, None) elsewhere.
    Works on CPU tensors (gloo) or CUDA tensors (nccl)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return _np(local_base), _np(local_rle)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local_base, local_rle = torch.as_tensor(local_base), torch.as_tensor(local_rle)
    seq = local_base.shape[1]
    device = local_base.device
    max_local = max(shard_bounds(n_windows, world, r)[1] - shard_bounds(n_windows, world, r)[0] for r in range(world))
    packed = torch.zeros((2, max_local, seq), dtype=torch.uint8, device=device)
    packed[0, : local_base.shape[0]] = local_base
    packed[1, : local_rle.shape[0]] = local_rle
    out = [torch.empty_like(packed) for _ in range(world)] if rank == 0 else None
    if dist.get_backend(group) == "nccl":
        gathered = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(gathered, packed, group=group)
        out = gathered if rank == 0 else None
    else:
        dist.gather(packed, out, dst=0, group=group)
    if rank != 0:
        return None, None
    base = np.empty((n_windows, seq), np.uint8)
    rle = np.empty((n_windows, seq), np.uint8)
    for r in range(world):
        s, e = shard_bounds(n_windows, world, r)
        base[s:e] = out[r][0, : e - s].cpu().numpy()
        rle[s:e] = out[r][1, : e - s].cpu().numpy()
    return base, rle


def predict_sharded(predict_fn, images, group=None):
    """Run predict_fn(images[start:end]) -> (base, rle) on this rank's window range and gather on rank 0."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, end = shard_bounds(len(images), world, rank)
    base, rle = predict_fn(images[start:end])
    return gather_labels(base, rle, len(images), group)


def _np(x):
    return x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
