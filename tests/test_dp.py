"""Data-parallel path.  CPU (gloo, world_size 2): the sharding scheme's arithmetic and the host
plumbing -- per-rank losses normalised by world*mb_local + SUM all-reduce == gradient of the global
mean loss; metric reduction; env sharding of the scan.  GPU (needs >= 2 devices, NCCL): two ranks
with half-minibatches reproduce the single-GPU full-minibatch update."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reppo_oracle as O

KW = dict(num_steps=8, num_envs=64, num_mini_batches=4, num_epochs=1, critic_hidden_dim=64, actor_hidden_dim=64,
          num_bins=51, vmin=0.0, vmax=20.0, obs_dim=5, critic_obs_dim=5, action_dim=2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reppo_b200 import dist as D
        from reppo_b200 import _lib as L
        torch.set_num_threads(1)
        hp = O.Hyper(**KW)
        sem = O.JAX
        cp, ap = O.init_params(hp, sem, seed=0)
        batch = O.synthetic_rollout(hp, seed=0, terminate=True)
        # (1) the scan shards over env columns with no exchange
        full_t = O.compute_targets({k: v.clone() for k, v in batch.items()}, hp, sem)
        loc = D.shard_envs(batch, rank, world)
        loc_t = O.compute_targets({k: v.clone() for k, v in loc.items()}, hp, sem)
        per = hp.num_envs // world
        assert torch.equal(loc_t, full_t[:, rank * per:(rank + 1) * per])
        # (2) gradient: sum over ranks of grad(mean-over-local-rows / world) == grad(mean over all rows)
        flat_full = O.flatten_batch(batch, full_t)
        flat_loc = O.flatten_batch(loc, loc_t)
        S_loc = flat_loc["target"].numel()
        mb_loc = S_loc // hp.num_mini_batches
        perm = D.local_permutations(7, 1, S_loc, rank)[0].long()
        idx = perm[:mb_loc]
        _, m_loc, g_loc = O._value_and_grad(lambda p: tuple(x if i else x / world for i, x in
                                                            enumerate(O.critic_loss(p, O.take(flat_loc, idx), hp, sem))), cp)
        g_sum = {k: v.clone() for k, v in g_loc.items()}
        for v in g_sum.values():
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        # the equivalent global minibatch = the union of both ranks' local rows
        rows = []
        for r in range(world):
            pr = D.local_permutations(7, 1, S_loc, r)[0].long()[:mb_loc]
            t, n = pr // per, pr % per
            rows.append(t * hp.num_envs + r * per + n)
        gidx = torch.cat(rows)
        _, m_full, g_full = O._value_and_grad(lambda p: O.critic_loss(p, O.take(flat_full, gidx), hp, sem), cp)
        for k in g_full:
            torch.testing.assert_close(g_sum[k], g_full[k], rtol=2e-5, atol=1e-7)
        # (3) metric reduction: partial sums add up, max/min reduce as max/min
        m = torch.zeros(L.NUM_METRICS)
        m[L.METRIC_NAMES.index("critic_loss")] = m_loc["loss"] / world
        m[L.METRIC_NAMES.index("target_max")] = m_loc["target_max"]
        m[L.METRIC_NAMES.index("target_min")] = m_loc["target_min"]
        red = D.reduce_metrics(m)
        assert abs(red[L.METRIC_NAMES.index("critic_loss")].item() - m_full["loss"].item()) < 1e-5
        assert red[L.METRIC_NAMES.index("target_max")].item() == m_full["target_max"].item()
        assert red[L.METRIC_NAMES.index("target_min")].item() == m_full["target_min"].item()
        # (4) the multicast exchange needs NCCL + NVSwitch multicast memory: on any other backend the attach declines before it
        # touches the engine, and the caller stays on the collectives it already has
        assert D.attach_multicast(object()) is False
        assert D.RECOMMENDED_NCCL_ENV == {"NCCL_PROTO": "LL"}
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_dp_arithmetic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


# ---- GPU, 2 ranks over NCCL -------------------------------------------------------------------------
def _nccl_worker(rank, world, port, out, gemm_mode, allreduce="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if allreduce == "mc":
        os.environ["REPPO_ALLREDUCE"] = "mc"   # both exchanges get attached below; the knob (read once by the library) picks
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from reppo_b200 import dist as D
        from reppo_b200.engine import EngineConfig, HyperParams, ReppoEngine
        hp = O.Hyper(**dict(KW, num_epochs=2))
        sem = O.JAX
        cp, ap = O.init_params(hp, sem, seed=0)
        batch = O.synthetic_rollout(hp, seed=0, terminate=True)
        per = hp.num_envs // world
        loc = D.shard_envs(batch, rank, world)
        S_loc = hp.num_steps * per
        mb_loc = S_loc // hp.num_mini_batches
        eps_pi_g, eps_kl_g = O.synthetic_noise(hp)            # global noise [E, mb, A]; rank slices its rows
        perms = D.local_permutations(7, hp.num_epochs, S_loc, rank)
        ecfg = EngineConfig(obs_dim=5, action_dim=2, critic_hidden_dim=64, actor_hidden_dim=64, num_bins=51, vmin=0.0,
                            vmax=20.0, semantics="jax", gemm_mode=gemm_mode, max_rows=hp.minibatch_size)
        eng = ReppoEngine(ecfg)
        eng.load_params(cp, ap)
        D.attach_nccl(eng, rank, world)
        if allreduce == "mc":
            # the all-reduce fused into clip+Adam over NVSwitch multicast memory (csrc/optim.cu) instead of ncclAllReduce
            assert D.attach_multicast(eng), "no multicast-capable symmetric memory on this box"
        dev = {k: v.cuda() for k, v in loc.items()}
        tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], dev["importance_weight"],
                            hp.gamma, hp.lmbda)
        flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
        flat["target"] = tgt.reshape(-1)
        e1 = eps_pi_g[:, rank * mb_loc:(rank + 1) * mb_loc].contiguous().cuda()
        e2 = eps_kl_g[:, :, rank * mb_loc:(rank + 1) * mb_loc].contiguous().cuda()
        hyp = HyperParams(world_size=world)
        m, sm = eng.update(eng.make_batch(flat), perms.cuda(), e1, e2, hyp, hp.num_mini_batches, use_graph=True)
        torch.cuda.synchronize()
        assert not D.dp_timed_out(eng)
        sm = D.reduce_metrics(sm)
        # single-process reference: the oracle on the equivalent global minibatches
        rows = []
        for r in range(world):
            pr = D.local_permutations(7, hp.num_epochs, S_loc, r).long()
            t, n = pr // per, pr % per
            rows.append((t * hp.num_envs + r * per + n).reshape(hp.num_epochs, hp.num_mini_batches, mb_loc))
        gperm = torch.cat(rows, dim=2).reshape(hp.num_epochs, -1)     # minibatch j = [rank0 rows | rank1 rows]
        ts2, _, _, traces = O.learn_step(O.TrainState.create(cp, ap), {k: v.clone() for k, v in batch.items()}, gperm,
                                         eps_pi_g, eps_kl_g, hp, sem, trace=True)
        tol = 2e-5 if gemm_mode == "fp32" else 2e-2
        from reppo_b200 import _lib as L
        c = L.METRIC_NAMES.index("critic_loss")
        assert abs(sm[0, c].item() - traces[0]["critic/loss"].item()) <= tol * abs(traces[0]["critic/loss"].item())
        for k in cp:
            err = (eng.critic.view(eng.critic.params, k).cpu() - ts2.critic[k]).abs().max().item()
            assert err <= 2 * hp.lr * len(traces) + 1e-6, (k, err)
        # every rank holds identical parameters
        p = eng.critic.params.clone()
        p0 = p.clone()
        dist.broadcast(p0, 0)
        assert torch.equal(p, p0)
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("allreduce", ["nccl", "mc"])
@pytest.mark.parametrize("gemm_mode", ["fp32", "bf16"])
def test_dp_two_ranks_match_single_process_oracle(gemm_mode, allreduce):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, gemm_mode, allreduce)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"
