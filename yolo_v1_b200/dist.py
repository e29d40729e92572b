"""Batch sharding of the hot path across the GPUs of one node.

Cells and images are independent, so the batch is cut into contiguous shards, one per rank (one process per
GPU); every rank runs the kernels on its own shard.  The reference has no distributed code at all (a single
`nn.DataParallel(net, device_ids=[0])`, train.py:80); what DDP would do around its loss module is what
happens here: each rank evaluates the loss module on its shard (so the `[:2]` rule of v1Loss.py:101 applies
to the first two objects of EACH shard, exactly as W independent reference calls would), then ONE
`all_reduce(sum)` of the 5-float terms vector over NCCL/NVLink gives the job-wide terms.  Gradients need
no exchange (d loss / d pred is local); decode + NMS needs no collective.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "all_reduce_terms", "sharded_loss"]


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced partition: the first (n_items % world_size) ranks take one extra item.
    Returns (start, stop)."""
    if world_size <= 0 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError("bad shard request: n=%d rank=%d world=%d" % (n_items, rank, world_size))
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_reduce_terms(terms, group=None, average=True):
    """Sum (or mean) the per-rank loss terms [5] over the ranks, in place, on the tensor's own stream.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()):
        return terms
    dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group)
    if average:
        terms /= dist.get_world_size(group)
    return terms


def sharded_loss(loss_module, pred_shard, target_shard, group=None, average=True):
    """loss_module(pred_shard, target_shard) on this rank's shard, then the terms all-reduce.
    Returns (local_loss -- call .backward() on it --, global_terms float32[5])."""
    local = loss_module(pred_shard, target_shard)
    terms = loss_module.last_terms.clone()
    return local, all_reduce_terms(terms, group=group, average=average)
