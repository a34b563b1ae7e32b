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


def gather_blocks(block: torch.Tensor, world: int, out: torch.Tensor = None) -> torch.Tensor:
    """All-gather equally sized uint8 output blocks -> [world, nbytes] on every rank."""
    if out is None:
        out = torch.empty((world, block.numel()), dtype=block.dtype, device=block.device)
    if world == 1:
        out[0].copy_(block)
        return out
    dist.all_gather_into_tensor(out.view(-1), block)
    return out


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
                    size: int, all_metas: Sequence[Sequence[dict]] = None):
    """Pickle-free replacement of collect_results_gpu: ONE all-gather of the packed output blocks, then rank 0 rebuilds
    the per-image dicts of every rank and interleaves them into dataset order.  `all_metas[r]` are rank r's img_metas
    (filenames); if omitted only rank-local metas are known and file names are taken from `local_metas` on rank 0 only."""
    gathered = gather_blocks(local_results_block, world)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        views = plan.views_of_block(gathered[r])
        metas = all_metas[r] if all_metas is not None else local_metas
        parts.append(plan.results(metas, src=views))
    return interleave_results(parts, size)
