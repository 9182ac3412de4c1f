"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL/NVLink.

The path shards by independent units: environments for the rollout (no communication) and
graphs of a minibatch for the train step, where the exchanges are exactly
  (1) a 4-float all-reduce of the label counts (ratio-of-sums denominators must be global), and
  (2) ONE all-reduce of the packed [grad_cbf | grad_actor | stats] buffer per optimizer step.
The reference has no multi-device code at all -- this is new work, not a port.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise from torchrun's RANK / LOCAL_RANK / WORLD_SIZE.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the C-ABI kernels launch on the CURRENT device: bind it to this rank's GPU whatever the backend / world size
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal shard [lo, hi) of n units (np.array_split convention)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t)
    return t


def allreduce_max_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
