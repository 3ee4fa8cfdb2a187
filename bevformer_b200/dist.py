"""Data-parallel plumbing for the encoder path.

The path shards along the batch axis only (the reference trains 1 sample per GPU under DDP,
projects/mmdet3d_plugin/bevformer/apis/mmdet_train.py:70-90), so the single exchange per step is the
average of the parameter gradients.  The encoder has 4.94 M parameters: they travel as ONE flat bucket
through one all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) -- at 9.9 MB in bf16 the
collective is latency-sized, so bucketing for overlap would only add launches.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def average_gradients_flat(params, world_size: int | None = None, group=None) -> None:
    """In-place average of ``p.grad`` over the process group, using a single all-reduce.
    Parameters without a gradient on this rank are an error (every rank must contribute the same
    bucket layout)."""
    params = [p for p in params]
    missing = [i for i, p in enumerate(params) if p.grad is None]
    if missing:
        raise RuntimeError(f"average_gradients_flat: parameters {missing[:8]} have no gradient")
    if world_size is None:
        world_size = dist.get_world_size(group)
    if not params or world_size == 1:
        return
    grads = [p.grad for p in params]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    flat.div_(world_size)
    torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(flat, grads))
