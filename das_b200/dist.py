"""Batch sharding across the GPUs of one box.

Images are independent (the reference loops per image, das_head.py:666), so every rank decodes its
own contiguous shard with no traffic inside the path; only the final small pose lists are gathered
with ONE all-gather of each rank's output block (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU
tests of this host logic).  Reference analogue: mmdet ``collect_results_gpu`` after
``multi_gpu_test`` (tools/test.py:201-206), which all-gathers pickled python objects instead.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` images for `rank`; the first total % world ranks get one more."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_list(items: Sequence, world: int, rank: int) -> List:
    lo, hi = shard_bounds(len(items), world, rank)
    return list(items[lo:hi])


def padded_shard_size(total: int, world: int) -> int:
    """Images per rank when `total` is not a multiple of `world`: every rank's plan is built for ceil(total / world)
    images (ranks with a shorter shard pad it, e.g. by repeating their last image) so that all output blocks have the
    same byte size and one all_gather_into_tensor moves them; the padding rows are dropped after the gather."""
    return -(-total // world)


def pad_shard(items: Sequence, size: int) -> List:
    """Pad a rank's shard (list of per-image things) to `size` entries by repeating the last one."""
    items = list(items)
    assert 0 < len(items) <= size, (len(items), size)
    return items + [items[-1]] * (size - len(items))


def gather_blocks(block: torch.Tensor, world: int, out: torch.Tensor = None, check_sizes: bool = True) -> torch.Tensor:
    """All-gather equally sized uint8 output blocks -> [world, nbytes] on every rank.  Blocks of different byte size
    (plans built for different batch sizes) would make the collective hang or corrupt memory, so the sizes are
    compared first (one 8-byte all-gather) and a mismatch raises on every rank; see padded_shard_size."""
    if world == 1:
        if out is None:
            out = torch.empty((1, block.numel()), dtype=block.dtype, device=block.device)
        out[0].copy_(block)
        return out
    if check_sizes:
        n = torch.tensor([block.numel()], dtype=torch.int64, device=block.device)
        sizes = torch.empty(world, dtype=torch.int64, device=block.device)
        dist.all_gather_into_tensor(sizes, n)
        sizes = sizes.tolist()
        if len(set(sizes)) != 1:
            raise ValueError(f"gather_blocks: per-rank output blocks differ in size {sizes}; build every rank's plan for "
                             f"padded_shard_size(total, world) images and drop the padding after the gather")
    if out is None:
        out = torch.empty((world, block.numel()), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out.view(-1), block)
    return out


def gather_counts(n_real: int, world: int, device) -> List[int]:
    """Number of real (non-padding) images in every rank's block."""
    if world == 1:
        return [int(n_real)]
    t = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(t, torch.tensor([int(n_real)], dtype=torch.int64, device=device))
    return [int(x) for x in t.tolist()]


def merge_results(per_rank_results: Sequence[Sequence[dict]]) -> List[dict]:
    """Concatenate per-rank result lists in rank order (= original image order for contiguous shards)."""
    merged = []
    for r in per_rank_results:
        merged.extend(r)
    return merged


# ---- drop-in for the CLI flow: mmdet multi_gpu_test + collect_results (reference tools/test.py:201-206) ------------
def sampler_indices(size: int, world: int, rank: int) -> List[int]:
    """Dataset indices a rank sees under the reference's test DistributedSampler (shuffle=False): the index list is
    padded by wrapping to a multiple of `world`, rank r takes r, r+world, ..."""
    total = -(-size // world) * world
    idx = list(range(size)) + list(range(total - size))
    return idx[rank:total:world]


def interleave_results(per_rank_results: Sequence[Sequence[dict]], size: int) -> List[dict]:
    """Re-order per-rank result lists into dataset order exactly like mmdet's collect_results:
    zip the parts, flatten, drop the sampler padding."""
    ordered = []
    for group in zip(*per_rank_results):
        ordered.extend(group)
    return ordered[:size]


def collect_results(plan, local_results_block: torch.Tensor, local_metas: Sequence[dict], world: int, rank: int,
                    size: int, all_metas: Sequence[Sequence[dict]] = None, n_real: int = None, order: str = "sampler"):
    """Pickle-free replacement of collect_results_gpu for ONE decoded batch per rank: ONE all-gather of the packed output
    blocks, then rank 0 rebuilds the per-image dicts of every rank and puts them into dataset order.

    Every rank's `plan` must have the same batch size (padded_shard_size); `n_real` = how many leading images of this
    rank's block are real (default: all), the rest is padding and is dropped.  `size` = number of images of the whole
    batch (all ranks).  order = 'sampler': ranks hold images r, r+world, ... like the reference's test DistributedSampler
    (results are interleaved and the wrap-around padding cut, mmdet collect_results); 'contiguous': rank r holds
    shard_bounds(size, world, r).  `all_metas[r]` are rank r's img_metas (filenames); if omitted, `local_metas` is used
    for every rank's file names."""
    n_real = plan.batch if n_real is None else int(n_real)
    assert 0 <= n_real <= plan.batch
    gathered = gather_blocks(local_results_block, world)
    counts = gather_counts(n_real, world, local_results_block.device)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        views = plan.views_of_block(gathered[r])
        metas = list(all_metas[r] if all_metas is not None else local_metas)
        metas = pad_shard(metas, plan.batch) if len(metas) < plan.batch else metas[:plan.batch]
        parts.append(plan.results(metas, src=views)[:counts[r]])
    if order == "contiguous":
        merged = merge_results(parts)
        assert len(merged) == size, (len(merged), size)
        return merged
    # sampler order: every rank decoded ceil(size / world) images (the sampler wraps around), interleave and cut
    return interleave_results(parts, size)
