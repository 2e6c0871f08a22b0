"""Batch sharding of the loss path across ranks (one process per GPU).

Every term of the view-synthesis loss is per batch item (per pixel with a 3x3 halo inside one
image); K, inv_K, T and the masks are per item.  The only cross-item operations are the masked mean
(trainer.py:1113-1114) and the `.mean()`s, and under the reference's DDP those are RANK-LOCAL: each
process normalises by its own shard (`batch_size` is per process, trainer.py:215-218) and DDP averages
the parameter gradients.  So the path needs no data-path collective; this module only holds

* `shard_batch`      the per-rank slice of a global batch dict (what accelerate's DataLoaderShard does),
* `global_loss_stats` an optional 16-byte all-reduce of (sum r*mask, sum mask) per scale for logging or
  for checking a sharded run against a single-process run (not reference behaviour),
* `allreduce_mean_`  the gradient averaging DDP performs, for callers that do not wrap the model in DDP.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d (the reference drops ragged "
                         "batches, trainer.py:215-218)" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(tensors: dict, global_batch: int, rank: int, world: int) -> dict:
    """Slices every tensor whose leading dimension is the global batch."""
    lo, hi = shard_range(global_batch, rank, world)
    out = {}
    for k, v in tensors.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == global_batch:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def global_loss_stats(sums: torch.Tensor, batch: int, num_scales: int, group=None) -> torch.Tensor:
    """(S,2) tensor of [sum(r*mask), sum(mask)] over all ranks, from the per-rank `sums` vector of
    ppea_vsl_forward (include/ppea_vsl.h: row stride 8 + 4*batch)."""
    stride = 8 + 4 * batch
    local = sums.view(num_scales, stride)[:, :2].clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def global_reproj_loss(stats: torch.Tensor) -> torch.Tensor:
    """Globally-normalised masked mean per scale (what a single process over the whole batch computes)."""
    return stats[:, 0] / (stats[:, 1] + 1e-7)


def allreduce_mean_(tensors, group=None):
    """In-place average over ranks (DDP's gradient reduction) of a list of tensors, as one flat bucket."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensors
    world = dist.get_world_size(group)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors
