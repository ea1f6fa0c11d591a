"""Data-parallel plumbing: cells are sharded over ranks (one process per GPU), the
parameters and BatchNorm buffers are replicated, and the per-rank gradients are summed with ONE
all-reduce of the flat gradient buffer per step (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The per-rank loss rides in the tail slot of the same buffer, so the step issues a single
collective.  Defined semantics: R-rank result == sum_r reference(shard_r) - every rank
normalises with ITS shard's BatchNorm statistics and class means, exactly like R independent
reference steps whose gradients are added."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def dp_world_size(group=None) -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def allreduce_grads_and_loss(flat_grads: torch.Tensor, n_param_floats: int, loss: Optional[torch.Tensor] = None,
                             group=None) -> None:
    """In place: flat_grads[:n] <- sum over ranks; flat_grads[n] <- sum over ranks of `loss`.

    ``flat_grads`` has at least ``n_param_floats + 1`` elements (ModelRuntime allocates 32 spare
    floats)."""
    if flat_grads.numel() <= n_param_floats:
        raise ValueError("flat gradient buffer has no tail slot for the loss")
    if loss is not None:
        flat_grads[n_param_floats] = loss.to(flat_grads.dtype)
    if dp_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Rank r takes rows [r*B, (r+1)*B) of a global minibatch of R*B cells."""
    per = n_rows // world
    return slice(rank * per, (rank + 1) * per)
