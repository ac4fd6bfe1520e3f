"""Data-parallel plumbing for the update step (not in the reference).

One process per GPU.  The batch shards over ENVIRONMENTS: rank r owns env columns
``[r*N/R, (r+1)*N/R)`` for all T steps, so the TD(lambda) scan is shard-local.  Every minibatch
step takes ``mb_local`` rows from a rank-local permutation; each rank's losses are normalised by
``world_size * mb_local`` so that the SUM all-reduce of the flat gradient arena is the gradient of
the global mean loss; clip + Adam then run redundantly (and identically) on every rank.
``torch.distributed`` is only used for rendezvous, the NCCL unique-id broadcast and the
once-per-update metric reduction; the per-step gradient all-reduce is issued by the C library on
its own NCCL communicator inside the captured CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping

import torch
import torch.distributed as dist

from . import _lib as L

SUM_SLOTS = [i for i, k in enumerate(L.METRIC_NAMES)
             if k not in ("target_max", "target_min", "critic_grad_norm", "actor_grad_norm", "temperature", "lagrangian")]
MAX_SLOTS = [L.METRIC_NAMES.index("target_max")]
MIN_SLOTS = [L.METRIC_NAMES.index("target_min")]


def shard_envs(batch: Mapping[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """``[T, N, ...]`` rollout -> this rank's env columns (contiguous)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[1]
        if n % world:
            raise ValueError(f"num_envs={n} is not divisible by world_size={world}")
        per = n // world
        out[k] = v[:, rank * per:(rank + 1) * per].contiguous()
    return out


def local_permutations(seed: int, num_epochs: int, rows_local: int, rank: int, device=None) -> torch.Tensor:
    """Rank-local shuffles, seeded by (seed, epoch, rank): ``[num_epochs, rows_local]`` int32."""
    perms = []
    for e in range(num_epochs):
        g = torch.Generator().manual_seed((seed * 1_000_003 + e) * 4099 + rank)
        perms.append(torch.randperm(rows_local, generator=g))
    p = torch.stack(perms).to(torch.int32)
    return p.to(device) if device is not None else p


def attach_nccl(engine, rank: int, world: int, communicators: int = 2):
    """Create the library's own NCCL communicators: rank 0 makes each unique id, torch.distributed broadcasts it.
    The first communicator serves the critic chain, the optional second one the actor chain (two chains may not
    share a communicator; with only one, the engine runs the two chains serially under data parallelism)."""
    dev = engine.device if dist.get_backend() == "nccl" else torch.device("cpu")
    for _ in range(communicators):
        raw = C.create_string_buffer(128)
        if rank == 0:
            L.check(engine.lib.reppo_nccl_unique_id(engine.ctx, raw), engine.ctx, "reppo_nccl_unique_id")
        t = torch.tensor(list(raw.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        ident = bytes(t.cpu().tolist())
        L.check(engine.lib.reppo_nccl_init(engine.ctx, ident, rank, world), engine.ctx, "reppo_nccl_init")


def reduce_metrics(m: torch.Tensor) -> torch.Tensor:
    """Per-rank metric vectors (partial sums already divided by world*mb) -> global metrics on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return m
    m = m.clone()
    s, mx, mn = m[..., SUM_SLOTS].contiguous(), m[..., MAX_SLOTS].contiguous(), m[..., MIN_SLOTS].contiguous()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    m[..., SUM_SLOTS], m[..., MAX_SLOTS], m[..., MIN_SLOTS] = s, mx, mn
    return m
