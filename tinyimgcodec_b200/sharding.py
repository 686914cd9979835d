"""Sharding a batch by image across the GPUs of one box: images are independent, so
each rank encodes a contiguous range and no collective touches the data path."""


def partition(n_items, world_size):
    """Contiguous ranges [lo, hi) per rank, sizes differing by at most one (what bench.py uses)."""
    return [(r * n_items // world_size, (r + 1) * n_items // world_size) for r in range(world_size)]


def balanced_partition(pixel_counts, world_size):
    """Contiguous ranges balancing the PIXEL count per rank for ragged batches: boundary r is placed
    where the running pixel total first reaches r/world of the whole."""
    total = sum(pixel_counts)
    bounds, acc, idx = [0], 0, 0
    n = len(pixel_counts)
    for r in range(1, world_size):
        target = total * r / world_size
        while idx < n and acc + pixel_counts[idx] / 2 <= target:
            acc += pixel_counts[idx]
            idx += 1
        bounds.append(idx)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def compress_sharded(images, quality, rank, world_size, compress_batch_fn, group=None):
    """Each rank encodes its range with `compress_batch_fn(list_of_images, quality) -> list[bytes]`;
    rank 0 receives every stream in image order (only compressed bytes travel: gather_object).
    Returns the full list on rank 0 and None elsewhere."""
    import torch.distributed as dist
    ranges = balanced_partition([int(im.shape[0]) * int(im.shape[1]) for im in images], world_size)
    lo, hi = ranges[rank]
    mine = compress_batch_fn(images[lo:hi], quality) if hi > lo else []
    if world_size == 1:
        return mine
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = []
    for part in gathered:
        out.extend(part)
    return out
