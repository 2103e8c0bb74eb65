"""Data-parallel plumbing: the one collective on the hot path is the gradient all-reduce (mean) the reference gets
from apex DistributedDataParallel (structure/model.py:27-34; per-rank batch = global / world,
data/data_loader.py:40-43).  One flat bucket (33 MB for CRNN) over NCCL / NVLink; CUDA-graph capturable."""
import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def shard_range(n_items, rank, world):
    """[lo, hi) of this rank's slice of a global batch (reference: batch_size // world per rank)."""
    per = n_items // world
    return rank * per, (rank + 1) * per


def allreduce_mean_grads_(params, group=None):
    """In place: p.grad <- mean over ranks of p.grad, for every parameter that has a gradient."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = _flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    for g, s in zip(grads, _unflatten_dense_tensors(flat, grads)):
        g.copy_(s)
