"""Data-parallel plumbing for the update step (not in the reference).

One process per GPU.  The batch shards over ENVIRONMENTS: rank r owns env columns
``[r*N/R, (r+1)*N/R)`` for all T steps, so the TD(lambda) scan is shard-local.  Every minibatch
step takes ``mb_local`` rows from a rank-local permutation; each rank's losses are normalised by
``world_size * mb_local`` so that the SUM all-reduce of the flat gradient arena is the gradient of
the global mean loss; clip + Adam then run redundantly (and identically) on every rank.
``torch.distributed`` is only used for rendezvous, the NCCL unique-id broadcast and the
once-per-update metric reduction; the per-step gradient all-reduce is issued by the C library inside
the captured CUDA graph: fused into the clip+Adam kernel over NVSwitch multicast memory
(``attach_multicast``), or as ``ncclAllReduce`` on the library's own communicators (``attach_nccl``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping

import torch
import torch.distributed as dist

from . import _lib as L

# NCCL environment that measured best for this path's all-reduces (4.57 MB + 1.12 MB per step, latency-bound); it must be in the
# environment before the process creates its first communicator (NCCL reads it once): bench.py sets it as a default.
RECOMMENDED_NCCL_ENV = {"NCCL_PROTO": "LL"}

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


def attach_multicast(engine, group=None) -> bool:
    """Move the engine's gradient arenas into one symmetric allocation (CUDA VMM + NVSwitch multicast object, through
    ``torch.distributed._symmetric_memory``) and register its mappings with the library: from then on the per-step
    gradient all-reduce runs INSIDE the clip+Adam launch (``multimem.ld_reduce`` / ``multimem.st``, csrc/optim.cu) and no
    NCCL kernel sits in the step chains.  Returns False (engine untouched, NCCL keeps serving) when the box has no
    multicast support or the world size does not divide 16."""
    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world < 2 or 16 % world or dist.get_backend(group) != "nccl":
        return False
    try:
        import torch.distributed._symmetric_memory as symm_mem
    except ImportError:
        return False
    nc, na = engine.critic.grads.numel(), engine.actor.grads.numel()     # both padded to a multiple of 4 floats
    flag_floats = L.MC_FLAG_BYTES // 4
    try:
        buf = symm_mem.empty(nc + na + flag_floats, dtype=torch.float32, device=engine.device)
        hdl = symm_mem.rendezvous(buf, group.group_name)
        mc_ptr = int(hdl.multicast_ptr)
    except Exception:  # noqa: BLE001 -- no VMM / fabric support on this box: stay on NCCL
        return False
    ok = torch.tensor([1 if mc_ptr else 0], device=engine.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if not int(ok.item()):
        return False
    buf.zero_()
    buf[:nc].copy_(engine.critic.grads)
    buf[nc:nc + na].copy_(engine.actor.grads)
    torch.cuda.synchronize(engine.device)
    dist.barrier(group)          # every rank's flags are zero before anyone can signal
    m = L.Multicast()
    m.local_base, m.mc_base = buf.data_ptr(), mc_ptr
    for r in range(world):
        m.peer_base[r] = int(hdl.buffer_ptrs[r])
    m.critic_grads_off, m.actor_grads_off, m.flags_off, m.flags_bytes = 0, nc * 4, (nc + na) * 4, L.MC_FLAG_BYTES
    m.rank, m.world_size = rank, world
    L.check(engine.lib.reppo_dp_attach_multicast(engine.ctx, C.byref(m)), engine.ctx, "reppo_dp_attach_multicast")
    engine.critic.grads, engine.actor.grads = buf[:nc], buf[nc:nc + na]
    engine._symmetric = (buf, hdl)             # keeps the allocation and its mappings alive
    for cb in getattr(engine, "_grads_moved", []):
        cb()
    return True


def dp_timed_out(engine) -> bool:
    """True when a cross-GPU wait of the fused all-reduce gave up (a rank died): the update's results are invalid."""
    v = C.c_int32(0)
    L.check(engine.lib.reppo_dp_status(engine.ctx, C.byref(v)), engine.ctx, "reppo_dp_status")
    return bool(v.value)


def mc_phase_summary(engine):
    """Median duration (us) of the phases of the fused all-reduce per chain, from the stamp ring (REPPO_MC_STAMPS=1 must be
    in the environment when ``attach_multicast`` runs): wait at the first twin barrier, reduce + fence, wait at the second
    twin barrier, norm + clip + Adam."""
    import numpy as np
    words = 8 + 8192 * 8
    out = np.zeros(words, dtype=np.uint64)
    L.check(engine.lib.reppo_dp_stamps(engine.ctx, out.ctypes.data, words), engine.ctx, "reppo_dp_stamps")
    n = int(min(out[0], 8192))
    rec = out[8:8 + n * 8].reshape(n, 8).astype(np.int64)
    res = {"records": n}
    for chain, name in ((0, "critic"), (1, "actor")):
        r = rec[(rec[:, 0] >> 32) == chain]
        if len(r) == 0:
            continue
        t = r[:, 1:7].astype(np.float64) / 1e3
        med = lambda x: float(np.median(x))
        res[name] = {"entry_to_block_arrived_us": med(t[:, 1] - t[:, 0]), "barrier_a_wait_us": med(t[:, 2] - t[:, 1]),
                     "reduce_and_fence_us": med(t[:, 3] - t[:, 2]), "barrier_b_wait_us": med(t[:, 4] - t[:, 3]),
                     "norm_clip_adam_us": med(t[:, 5] - t[:, 4]), "block_total_us": med(t[:, 5] - t[:, 0])}
    return res


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
