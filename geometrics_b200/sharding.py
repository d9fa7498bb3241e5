"""Batch-axis sharding of the hot path over one process per GPU.

Every kernel of the path treats batch items independently (chamfer_distance.cu:18,
tri_distance.cu:111, utils.py:607-612), so rank r of W simply owns items [r*B/W, (r+1)*B/W): no halo
and no exchange during the searches.  The only communication is ONE all-reduce of the partial loss
sums (two or three fp32 scalars) -- ``utils.all_reduce_sums`` -- after which every rank divides by
the global element count.  The reference has no distributed code at all; this module is new.
"""
from __future__ import annotations

import torch


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous, balanced split: the first ``batch % world`` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's slice of a (B, ...) tensor (a view)."""
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def global_means(partial_sums: torch.Tensor, local_counts, group=None) -> torch.Tensor:
    """means[i] = (sum over ranks of partial_sums[i]) / (sum over ranks of local_counts[i]).

    ``local_counts`` are this rank's element counts (python numbers); they ride in the same
    all-reduce as the sums, so ragged shards (B not divisible by W) still give the exact global mean
    ``torch.mean`` over all B*N rows would (utils.py:416-417), up to fp32 summation order."""
    from .utils import all_reduce_sums

    counts = torch.tensor([float(c) for c in local_counts], dtype=partial_sums.dtype, device=partial_sums.device)
    packed = all_reduce_sums(torch.cat((partial_sums, counts)), group)
    k = partial_sums.numel()
    return packed[:k] / packed[k:]
