"""Data-parallel plumbing: cells are sharded over ranks (one process per GPU), the
parameters and BatchNorm buffers are replicated, and the per-rank gradients are summed with ONE
all-reduce of the flat gradient buffer per step (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The per-rank loss rides in the tail slot of the same buffer, so the step issues a single
collective.  Defined semantics: R-rank result == sum_r reference(shard_r) - every rank
normalises with ITS shard's BatchNorm statistics and class means, exactly like R independent
reference steps whose gradients are added."""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


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


class _RawCuda:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerAllReduce:
    """The step's gradient all-reduce as ONE kernel over NVLink peer memory (``drv_peer_allreduce``).

    Allocates this rank's buffer of ``n_floats`` floats through the C ABI, exchanges CUDA IPC handles over the
    process group and maps every other rank's buffer.  ``self.data`` is the local buffer as a torch tensor (the
    runtime re-bases its flat gradient buffer onto it); ``all_reduce()`` sums it in place over the ranks, on the
    current stream.  Single-node groups of 2, 4 or 8 ranks; ``create`` returns None (caller keeps NCCL) otherwise
    or when the IPC set-up fails."""

    def __init__(self, n_floats: int, device: torch.device, group=None):
        self.lib = _lib.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_floats = int(n_floats)
        self.device = torch.device(device)
        header = self.lib.drv_comm_header_bytes()
        base = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        opened, bases, err = [], [], None
        with torch.cuda.device(self.device):
            # Every step below is collective: a rank that fails keeps taking part and reports the failure, so that
            # ALL ranks fall back to NCCL together (a rank in ncclAllReduce next to ranks in the peer kernel hangs).
            rc = self.lib.drv_comm_alloc(4 * ((self.n_floats + 3) // 4 * 4), C.byref(base), handle)
            if rc != 0:
                err = f"drv_comm_alloc failed ({rc})"
            handles = [None] * self.world
            dist.all_gather_object(handles, (bytes(handle) if err is None else None, os.uname().nodename), group=group)
            if err is None and any(h[0] is None for h in handles):
                err = "a rank could not allocate its buffer"
            if err is None and len({h[1] for h in handles}) != 1:
                err = "ranks on different nodes"
            if err is None:
                for j, (h, _) in enumerate(handles):
                    if j == self.rank:
                        bases.append(base.value)
                        continue
                    peer = C.c_void_p()
                    rc = self.lib.drv_comm_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(peer))
                    if rc != 0:
                        err = f"drv_comm_open failed ({rc})"
                        break
                    opened.append(peer.value)
                    bases.append(peer.value)
            # agreement + barrier: every rank has mapped every buffer before the first kernel touches them
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                for p in opened:
                    self.lib.drv_comm_close(C.c_void_p(p))
                if base.value:
                    self.lib.drv_comm_free(base)
                raise RuntimeError(err or "another rank could not set up the peer buffers")
        self._own = base.value
        self._bases = (C.c_void_p * self.world)(*bases)
        self.data = torch.as_tensor(_RawCuda(base.value + header, self.n_floats), device=self.device)

    @classmethod
    def create(cls, n_floats: int, device, group=None) -> Optional["PeerAllReduce"]:
        if os.environ.get("DRV_PEER_ALLREDUCE", "1") == "0":
            return None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(group) != "nccl":
            return None
        if dist.get_world_size(group) not in (2, 4, 8):
            return None
        try:
            return cls(n_floats, device, group)
        except Exception as e:  # IPC not permitted, ranks on several nodes, ...: NCCL does the job
            warnings.warn(f"drvish_b200: peer-memory all-reduce unavailable ({e}); using NCCL")
            return None

    def bind(self) -> None:
        """Let the calling thread's ``drv_elbo_step`` calls launch the collective themselves (DRV_FLAG_FUSED_ALLREDUCE)."""
        _lib.check(self.lib.drv_comm_bind(self._bases, self.rank, self.world, self.n_floats), "drv_comm_bind")

    def all_reduce(self) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.drv_peer_allreduce(self._bases, self.rank, self.world, self.n_floats, C.c_void_p(st)),
                   "drv_peer_allreduce")
