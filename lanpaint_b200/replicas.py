"""Multi-GPU: independent replicas, one process per GPU.

The Langevin path has no cross-sample quantity, so requests shard across ranks with
NO data-path collective.  `torch.distributed` (NCCL over NVLink on the GPU box, gloo in
the CPU tests) is used for exactly three things: the one-time broadcast of the denoiser
weights from rank 0 (north_star), the barrier around a timed region with the
max-over-ranks reduction of its duration, and optionally collecting results on rank 0.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_requests: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n mod world) ranks take one extra request."""
    if world <= 0 or not (0 <= rank < world) or n_requests < 0:
        raise ValueError((n_requests, world, rank))
    base, extra = divmod(n_requests, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_requests: int, world: int) -> List[int]:
    return [b - a for a, b in (shard_bounds(n_requests, world, r) for r in range(world))]


class ReplicaGroup:
    """Thin wrapper over the default process group (or a single-process stand-in)."""

    def __init__(self, backend: Optional[str] = None, device: Optional[torch.device] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self.owns_group = False
        if self.world > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if (device is not None and device.type == "cuda") else "gloo")
            kw = {"device_id": device} if backend == "nccl" and device is not None else {}
            dist.init_process_group(backend, **kw)
            self.owns_group = True

    # -- weights: the only payload that ever crosses NVLink ---------------------------------
    def broadcast_weights(self, tensors: Iterable[torch.Tensor], src: int = 0) -> int:
        """In-place broadcast of every tensor from `src`; returns the number of bytes moved."""
        moved = 0
        for t in tensors:
            if self.world > 1:
                dist.broadcast(t, src=src)
            moved += t.numel() * t.element_size()
        return moved

    # -- requests ---------------------------------------------------------------------------
    def my_slice(self, n_requests: int) -> slice:
        a, b = shard_bounds(n_requests, self.world, self.rank)
        return slice(a, b)

    def gather_results(self, local: torch.Tensor, n_requests: int, dst: int = 0) -> Optional[torch.Tensor]:
        """Concatenate the per-rank result batches in request order on `dst` (None elsewhere).
        Host-side collection: a ComfyUI LATENT lives on the intermediate (CPU) device anyway."""
        if self.world == 1:
            return local
        sizes = shard_sizes(n_requests, self.world)
        if self.rank == dst:
            bufs = [local.new_empty((s,) + tuple(local.shape[1:])) for s in sizes]
            bufs[dst] = local
            reqs = [dist.irecv(bufs[r], src=r) for r in range(self.world) if r != dst and sizes[r] > 0]
            for q in reqs:
                q.wait()
            return torch.cat(bufs, dim=0)
        if local.shape[0] > 0:
            dist.send(local, dst=dst)
        return None

    # -- timing -----------------------------------------------------------------------------
    def barrier(self):
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def max_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return float(value)
        t = torch.tensor([value], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return float(value)
        t = torch.tensor([value], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.owns_group and dist.is_initialized():
            dist.destroy_process_group()
