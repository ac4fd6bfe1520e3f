#!/usr/bin/env python
"""bench.py -- REPPO update samples/s on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2]

A "step" is ONE full REPPO update of one rollout batch: the TD(lambda) scan plus
num_epochs x num_mini_batches critic+actor minibatch steps (src/jaxrl/reppo.py:417-673).  The
workload at N=1 is C2 = CheetahRun-shaped mjx_dmc_large_data (T=128, N=1024 envs, S=131072
samples, 64 minibatches x 8 epochs, mb=2048, obs 17 / act 6, H=512, 151 bins).  Under torchrun
every rank owns its own 1024 env columns (weak scaling over environments); gradients are
all-reduced with NCCL each minibatch step.  Rank 0 prints ONE JSON line.

value     : samples/s with the rollout already resident in HBM (CUDA events around K updates).
e2e       : the same through the public API (reppo_b200.torch_api.reppo_update) with HOST pinned
            buffers: H2D of the rollout and D2H of the metrics inside the timed region.
roofline  : the dominant kernel family (the MLP GEMMs): algorithmic FLOPs / measured kernel time.
cpu_baseline / --impl reference : the oracle port of the reference update timed on the host cores
            (a bounded sample of minibatch steps, extrapolated to the full update).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (env group, experiment override, obs_dim, action_dim, terminate, reward_scaling, extra hyper overrides)
    "C1": ("mjx_dmc", "mjx_dmc_small_data", 5, 1, False, 1.0, {}),
    "C2": ("mjx_dmc", "mjx_dmc_large_data", 17, 6, False, 1.0, {}),
    "C3": ("brax", "default", 27, 8, True, 0.1, {}),
    "C4": ("humanoid_brax", "default", 244, 17, True, 0.1, {}),
    "C5": ("mjx_dmc", "mjx_dmc_large_data", 17, 6, False, 1.0,
           {"critic_hidden_dim": 1024, "actor_hidden_dim": 1024, "num_envs": 2048, "num_mini_batches": 16}),
}


_RESULT_OUT = None
_T0 = time.time()


def stage(msg):
    """Progress line on stderr (rank-tagged): which leg is running when a run is slow or stuck."""
    sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - _T0:6.1f}s] {msg}\n")
    sys.stderr.flush()


def arm_watchdog():
    """A stuck rank (a collective whose peer never arrives) must end the run, not hold the box: after
    REPPO_BENCH_WATCHDOG_S seconds (default 1500) every thread's Python stack goes to stderr and the process exits."""
    import faulthandler
    limit = float(os.environ.get("REPPO_BENCH_WATCHDOG_S", "1500"))
    if limit > 0:
        faulthandler.dump_traceback_later(limit, exit=True)
    import signal
    faulthandler.register(signal.SIGUSR1, all_threads=True)   # `kill -USR1 <pid>`: where is it right now?


def emit(obj):
    """The one JSON line of this run, on the real stdout."""
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_traffic():
    """dram bytes per launch of the dominant kernel family from the newest committed `ncu --set full` capture."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d.get("dram_bytes_per_launch_mean"), os.path.basename(files[-1])


def make_cfg(name, semantics, more=()):
    from reppo_b200.config import compose
    env, ov, D, A, term, rs, extra = CONFIGS[name]
    overrides = [f"env={env}", f"experiment_overrides={ov}", f"b200.semantics={semantics}"]
    overrides += [f"hyperparameters.{k}={v}" for k, v in extra.items()] + list(more)
    cfg = compose(overrides)
    return cfg, D, A, term, rs


def oracle_hyper(cfg, D, A):
    from oracle import reppo_oracle as O
    h = cfg.hyperparameters
    return O.Hyper(num_steps=h.num_steps, num_envs=h.num_envs, num_mini_batches=h.num_mini_batches, num_epochs=h.num_epochs,
                   gamma=h.gamma, lmbda=h.lmbda, lr=h.lr, max_grad_norm=h.max_grad_norm, critic_hidden_dim=h.critic_hidden_dim,
                   actor_hidden_dim=h.actor_hidden_dim, vmin=float(h.vmin), vmax=float(h.vmax), num_bins=h.num_bins,
                   kl_bound=h.kl_bound, obs_dim=D, critic_obs_dim=D, action_dim=A, actor_min_std=h.actor_min_std)


def gemm_flops_per_update(hp):
    """Algorithmic FLOPs of the Linear layers: critic step 6(f+h+p), actor step
    2(3a + a + 2(f+h)) per sample visit."""
    H, Ha, B, D, A = hp.critic_hidden_dim, hp.actor_hidden_dim, hp.num_bins, hp.obs_dim, hp.action_dim
    f = (hp.critic_obs_dim + A) * H + H * H
    h = H * H + H * B
    p = H * H + H * (H + 1)
    a = D * Ha + Ha * Ha + 2 * A * Ha
    per_visit = 6 * (f + h + p) + 2 * (3 * a + a + 2 * (f + h))
    return per_visit * hp.batch_size * hp.num_epochs


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = next((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


class CpuArm:
    """The reference update on the host cores, as a bounded sample of the workload.

    kind "reference": the reference's OWN Torch closures (make_postprocess_fn / make_critic_update_fn /
    make_actor_update_fn, src/torchrl/reppo.py:173-360, stock code path, its own noise draws) imported from
    /root/reference or from the byte-compiled oracle/_ref that oracle/build_ref.py ships to the GPU box.
    kind "port": oracle/reppo_oracle.py (only when neither exists).
    One sample = compute_gve over the whole rollout + `m_steps` consecutive critic+actor minibatch steps, i.e. the fraction
    f = m_steps / (num_epochs * num_mini_batches) of one update; value = f * S / wall time of the sample (nothing is
    extrapolated: ms_per_step is the measured wall time of the sample)."""

    def __init__(self, hp, sem_name, m_steps):
        import torch
        from oracle import ref_torch_shim as RB
        from oracle import reppo_oracle as O
        self.hp, self.O, self.m = hp, O, min(m_steps, hp.num_epochs * hp.num_mini_batches)
        self.cores = os.cpu_count() or 1
        self.kind = "reference" if RB.available() else "port"
        self.sem_name = "torch" if self.kind == "reference" else sem_name
        sem = O.TORCH if self.sem_name == "torch" else O.JAX
        self.sem = sem
        cp, ap = O.init_params(hp, sem, seed=0)
        self.batch = O.synthetic_rollout(hp, seed=0)
        g = torch.Generator().manual_seed(5)
        self.perm = torch.randperm(hp.batch_size, generator=g)
        if self.kind == "reference":
            ref = RB.load_reference()   # importing the reference sets OMP_NUM_THREADS=1 (reppo.py:37); the shim restores it
            self.ts = RB.build_train_state(ref, hp, cp, ap)
            cfg = RB.make_cfg(hp)
            self.post = ref.reppo.make_postprocess_fn(cfg, None)
            self.upd_c = ref.reppo.make_critic_update_fn(cfg, self.ts)
            self.upd_a = ref.reppo.make_actor_update_fn(cfg, self.ts)
            self.transition = RB.to_ref_transition(self.batch)
        else:
            self.ts = O.TrainState.create(cp, ap)
            self.eps = O.synthetic_noise(hp)
        torch.set_num_threads(self.cores)
        self.threads = torch.get_num_threads()

    def run(self):
        """One sample; returns its wall time in seconds."""
        O, hp = self.O, self.hp
        mbs, n_mb = hp.minibatch_size, hp.num_mini_batches
        t0 = time.perf_counter()
        if self.kind == "reference":
            data = self.post(self.ts, self.transition)          # compute_gve + flatten (:173-221)
            for j in range(self.m):
                mini = data[self.perm[(j % n_mb) * mbs:(j % n_mb + 1) * mbs]]
                self.upd_c(mini)
                self.upd_a(mini)
        else:
            target = O.compute_targets({k: v.clone() for k, v in self.batch.items()}, hp, self.sem)
            flat = O.flatten_batch(self.batch, target)
            for j in range(self.m):
                mb = O.take(flat, self.perm[(j % n_mb) * mbs:(j % n_mb + 1) * mbs])
                self.ts, _, _ = O.critic_step(self.ts, mb, hp, self.sem)
                self.ts, _, _ = O.actor_step(self.ts, mb, self.eps[0][0], self.eps[1][0], hp, self.sem)
        return time.perf_counter() - t0

    def describe(self, wall):
        hp = self.hp
        frac = self.m / (hp.num_epochs * hp.num_mini_batches)
        what = ("the reference's own Torch closures (src/torchrl/reppo.py make_postprocess_fn / make_critic_update_fn / "
                "make_actor_update_fn, stock code path, byte-compiled copy oracle/_ref)" if self.kind == "reference"
                else f"oracle/reppo_oracle.py ({self.sem_name} semantics)")
        return {"value": frac * hp.batch_size / wall, "unit": "samples/s", "cores": self.threads, "kind": self.kind,
                "sample": f"compute_gve over the whole rollout (S={hp.batch_size}) + {self.m} consecutive critic+actor minibatch "
                          f"steps (mb={hp.minibatch_size}) = {frac:.4f} of one update, {wall:.2f} s wall, fp32 torch CPU, {what}; "
                          f"value = {frac:.4f} * S / wall"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the update on the host cores (JAX is not installable in
    this image, so it is the reference's Torch closures; see CpuArm).  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, D, A, term, rs = make_cfg(args.config, args.semantics)
    hp = oracle_hyper(cfg, D, A)
    arm = CpuArm(hp, args.semantics, m_steps=hp.num_mini_batches)   # one epoch's worth of minibatch steps per sample
    walls = []
    for i in range(args.warmup + args.steps):
        w = arm.run()
        if i >= args.warmup:
            walls.append(w)
    wall = sum(walls) / len(walls)
    base = arm.describe(wall)
    line = {
        "impl": "reference", "metric": "reppo_update_samples_per_s", "value": base["value"], "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, cfg, D, A, 1), "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "full_update_ms_at_this_rate": hp.batch_size / base["value"] * 1e3,
        "note": "jax/flax/optax are not installable in this image: the reference arm runs the reference's Torch update path "
                "(kind=reference) when oracle/_ref or /root/reference is present, else the oracle port (kind=port)",
    }
    emit(line)


ALLREDUCE = {"mode": "none"}


def dp_defaults():
    from reppo_b200 import dist as dp_plumbing
    return dp_plumbing.RECOMMENDED_NCCL_ENV


def _attach_multicast(eng, dp_plumbing):
    """Data-parallel gradient exchange: ncclAllReduce inside the graph (default: it measured faster at 2 and 8 GPUs), or with
    REPPO_ALLREDUCE=mc the multicast all-reduce fused into clip+Adam when the box has NVSwitch multicast memory."""
    ALLREDUCE["mode"] = "nccl (ncclAllReduce per chain and step, inside the graph)"
    if not os.environ.get("REPPO_ALLREDUCE", "").lower().startswith("m"):
        return
    if dp_plumbing.attach_multicast(eng):
        ALLREDUCE["mode"] = "multicast (multimem.ld_reduce/st fused into the clip+Adam launch, no NCCL kernel in the step chains)"


def workload_config(name, cfg, D, A, world):
    h = cfg.hyperparameters
    S = h.num_steps * h.num_envs
    return {"workload": f"{name}: REPPO update, {CONFIGS[name][0]} / {CONFIGS[name][1]}-shaped synthetic rollout, "
                        f"T={h.num_steps} x N={h.num_envs} envs per GPU (S={S}), {h.num_mini_batches} minibatches x "
                        f"{h.num_epochs} epochs (mb={S // h.num_mini_batches}/GPU), obs {D} / act {A}, H={h.critic_hidden_dim}, "
                        f"{h.num_bins} bins",
            "samples_per_gpu": S, "minibatch_per_gpu": S // h.num_mini_batches, "optimizer_steps": h.num_epochs * h.num_mini_batches,
            "parallelism": f"dp{world} over environments, gradient all-reduce: {ALLREDUCE['mode']}" if world > 1 else "single GPU",
            "l2": "rollout batch (292 MB at C2) exceeds the 126 MB L2; weights/Adam state are L2-resident by design"}


def measure_config(name, semantics, gemm, world, rank, dev, steps, warmup, strong=False, overrides=()):
    """One more workload at this GPU count: builds its own engine + synthetic rollout, times `steps` full updates (CUDA events,
    max over ranks) and frees everything again.  strong=True: ONE named batch whose env columns are split over the ranks
    (N / world envs and minibatch / world rows per rank) instead of one named batch per rank."""
    import torch
    import torch.distributed as dist
    from oracle import reppo_oracle as O
    from reppo_b200 import torch_api as T
    from reppo_b200.engine import HyperParams
    cfg, D, A, term, rs = make_cfg(name, semantics, list(overrides))
    if gemm in ("bf16", "tf32"):
        cfg.platform.amp_dtype = gemm
    h = cfg.hyperparameters
    hp_g = oracle_hyper(cfg, D, A)                      # the named (global) batch
    if strong:
        if h.num_envs % world:
            return {"skipped": f"num_envs={h.num_envs} not divisible by {world}"}
        h.num_envs = h.num_envs // world
    hp = oracle_hyper(cfg, D, A)                        # this rank's share
    S, n_mb, E = hp.batch_size, hp.num_mini_batches, hp.num_epochs
    mb = S // n_mb
    sem = O.JAX if semantics == "jax" else O.TORCH
    if strong:
        batch = O.synthetic_rollout(hp_g, seed=0, terminate=term, reward_scaling=rs, num_envs=hp.num_envs,
                                    env_offset=rank * hp.num_envs)
    else:
        batch = O.synthetic_rollout(hp, seed=rank, terminate=term, reward_scaling=rs)
    cp, ap_ = O.init_params(hp, sem, seed=0)
    ts = T.make_train_state(cfg, D, D, A, device=dev)
    eng = ts.engine
    eng.load_params(cp, ap_)
    if world > 1:
        from reppo_b200 import dist as dp_plumbing
        dp_plumbing.attach_nccl(eng, rank, world)
        _attach_multicast(eng, dp_plumbing)
        stage(f"  {name}: communicators up ({ALLREDUCE['mode']})")
    hyp = HyperParams(gamma=h.gamma, lmbda=h.lmbda, lr=h.lr, max_grad_norm=h.max_grad_norm, kl_bound=h.kl_bound,
                      ent_target_mult=h.ent_target_mult, aux_loss_mult=h.aux_loss_mult, actor_min_std=h.actor_min_std,
                      actor_kl_clip_mode=h.actor_kl_clip_mode, world_size=world)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    devb = {k: v.to(dev) for k, v in batch.items()}
    del batch
    perms = torch.stack([torch.randperm(S, device=dev, generator=g) for _ in range(E)]).to(torch.int32)
    eps_pi = torch.randn(E, mb, A, device=dev, generator=g)
    eps_kl = torch.randn(E, 16, mb, A, device=dev, generator=g)
    target = torch.empty(hp.num_steps, hp.num_envs, device=dev)
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in devb.items()}
    flat["target"] = target.reshape(-1)
    cb = eng.make_batch(flat)
    step_metrics = torch.empty(E * n_mb, 24, device=dev)

    def one_update():
        # the update's own random draws are part of it (learn_step shuffles and samples inside, src/jaxrl/reppo.py:535-557,645-647)
        for e in range(E):
            perms[e].copy_(torch.randperm(S, device=dev, generator=g))
        eps_pi.normal_(generator=g)
        eps_kl.normal_(generator=g)
        eng.td_lambda(devb["soft_reward"], devb["value"], devb["done"], devb["truncated"], devb["importance_weight"],
                      h.gamma, h.lmbda, out=target)
        eng.sync_actor_old()
        eng.update(cb, perms, eps_pi, eps_kl, hyp, n_mb, use_graph=True, step_metrics=step_metrics)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        one_update()
    barrier()
    stage(f"  {name}: warm-up done")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        one_update()
    ev1.record()
    barrier()
    stage(f"  {name}: timed region done")
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    fm = eng.metrics_dict(eng.metrics)
    out = {"value": S * world / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "dtype": {"fp32": "f32", "bf16": "bf16", "tf32": "tf32"}[gemm], "scaling": "strong" if strong else "weak",
           "in_update_tflops": gemm_flops_per_update(hp) * world / (ms * 1e-3) / 1e12,
           "workload": workload_config(name, cfg, D, A, world)["workload"],
           "finite": bool(all(math.isfinite(fm[k]) for k in ("critic_loss", "actor_loss", "kl")))}
    if world > 1:
        from reppo_b200 import dist as dp_plumbing
        out["allreduce"] = ALLREDUCE["mode"]
        out["dp_timed_out"] = dp_plumbing.dp_timed_out(eng)
        eng.lib.reppo_nccl_finalize(eng.ctx)
    del eng, ts, devb, flat, cb, perms, eps_pi, eps_kl, target, step_metrics
    torch.cuda.empty_cache()
    return out


def dp_check(world, rank, dev, gemm):
    """world > 1, outside every timed region: two epochs x two minibatch steps at the benchmarked step shape (global
    minibatch 2048 rows, H = 512, obs 17 / act 6) with the env columns split over the ranks and the gradients all-reduced by
    the library; rank 0 replays the rank-concatenated minibatches through the CPU oracle (tests/test_dp.py does the same on
    2 GPUs).  Checks: step-0 critic loss / gradient norm / actor loss against the oracle, every rank bit-identical
    parameters after the update."""
    import torch
    import torch.distributed as dist
    from oracle import reppo_oracle as O
    from reppo_b200 import _lib as L
    from reppo_b200 import dist as D_
    from reppo_b200.engine import EngineConfig, HyperParams, ReppoEngine
    hp = O.Hyper(num_steps=4, num_envs=1024, num_mini_batches=2, num_epochs=2, critic_hidden_dim=512, actor_hidden_dim=512,
                 num_bins=151, vmin=0.0, vmax=150.0, obs_dim=17, critic_obs_dim=17, action_dim=6)
    if hp.num_envs % world or hp.minibatch_size % world:
        return {"ok": None, "skipped": f"1024 envs not divisible by {world}"}
    sem = O.JAX
    cp, ap = O.init_params(hp, sem, seed=0)
    batch = O.synthetic_rollout(hp, seed=0, terminate=True)
    per = hp.num_envs // world
    loc = D_.shard_envs(batch, rank, world)
    S_loc = hp.num_steps * per
    mb_loc = S_loc // hp.num_mini_batches
    eps_pi_g, eps_kl_g = O.synthetic_noise(hp)
    perms = D_.local_permutations(7, hp.num_epochs, S_loc, rank)
    eng = ReppoEngine(EngineConfig(obs_dim=17, action_dim=6, semantics="jax", gemm_mode=gemm, max_rows=mb_loc,
                                   max_batch_rows=S_loc), device=dev)
    eng.load_params(cp, ap)
    D_.attach_nccl(eng, rank, world)
    _attach_multicast(eng, D_)
    devb = {k: v.to(dev) for k, v in loc.items()}
    tgt = eng.td_lambda(devb["soft_reward"], devb["value"], devb["done"], devb["truncated"], devb["importance_weight"],
                        hp.gamma, hp.lmbda)
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in devb.items()}
    flat["target"] = tgt.reshape(-1)
    e1 = eps_pi_g[:, rank * mb_loc:(rank + 1) * mb_loc].contiguous().to(dev)
    e2 = eps_kl_g[:, :, rank * mb_loc:(rank + 1) * mb_loc].contiguous().to(dev)
    m, sm = eng.update(eng.make_batch(flat), perms.to(dev), e1, e2, HyperParams(world_size=world), hp.num_mini_batches,
                       use_graph=True)
    torch.cuda.synchronize()
    stage("  dp_check: update done")
    sm = D_.reduce_metrics(sm)
    same = True
    for net in (eng.critic, eng.actor):
        p0 = net.params.clone()
        dist.broadcast(p0, 0)
        flag = torch.tensor([1.0 if torch.equal(net.params, p0) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = same and bool(flag.item() == 1.0)
    res = None
    if rank == 0:
        rows = []
        for r in range(world):
            pr = D_.local_permutations(7, hp.num_epochs, S_loc, r).long()
            t, n = pr // per, pr % per
            rows.append((t * hp.num_envs + r * per + n).reshape(hp.num_epochs, hp.num_mini_batches, mb_loc))
        gperm = torch.cat(rows, dim=2).reshape(hp.num_epochs, -1)
        ts2, _, _, traces = O.learn_step(O.TrainState.create(cp, ap), {k: v.clone() for k, v in batch.items()}, gperm,
                                         eps_pi_g, eps_kl_g, hp, sem, trace=True)
        col = {k: i for i, k in enumerate(L.METRIC_NAMES)}
        tol = 2e-5 if gemm == "fp32" else 2e-2
        rel = lambda a, b: abs(a - b) / (abs(b) + 1e-12)
        smc = sm.cpu()
        e_loss = rel(smc[0, col["critic_loss"]].item(), traces[0]["critic/loss"].item())
        e_gn = rel(smc[0, col["critic_grad_norm"]].item(), traces[0]["critic/grad_norm"].item())
        e_al = rel(smc[0, col["actor_loss"]].item(), traces[0]["actor/loss"].item())
        perr = max((eng.critic.view(eng.critic.params, k).cpu() - v).abs().max().item() for k, v in ts2.critic.items())
        ok = same and e_loss <= tol and e_gn <= 5 * tol and e_al <= 10 * tol and perr <= 2 * hp.lr * len(traces) + 1e-6
        res = {"ok": bool(ok), "ranks_bit_identical_params": same, "step0_critic_loss_rel_err": e_loss,
               "step0_critic_grad_norm_rel_err": e_gn, "step0_actor_loss_rel_err": e_al, "max_param_err": perr,
               "param_bound": 2 * hp.lr * len(traces), "tolerance": tol, "steps": len(traces),
               "shape": f"global minibatch {hp.minibatch_size} rows ({mb_loc} per rank), H=512, obs 17 / act 6, {gemm} GEMMs, "
                        f"graph + two-chain pipeline + gradient all-reduce inside the graph: {ALLREDUCE['mode']}",
               "oracle": "oracle/reppo_oracle.py learn_step on the rank-concatenated minibatches (CPU, rank 0)"}
    eng.lib.reppo_nccl_finalize(eng.ctx)
    del eng
    torch.cuda.empty_cache()
    return res



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--semantics", default="jax", choices=["jax", "torch"])
    ap.add_argument("--gemm", default=os.environ.get("REPPO_GEMM", "bf16"), choices=["fp32", "bf16", "tf32"],
                    help="bf16 = tcgen05 tensor-core path (bf16 operands, fp32 accumulate; the product); fp32 = FFMA reference-precision mode")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--override", action="append", default=[], metavar="KEY=VALUE",
                    help="extra hydra-style override for experiments, e.g. hyperparameters.num_epochs=4 (not the named workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the large-shape GEMM leg (C5 sweep shapes)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra workload legs (the other BASELINE.json configurations, the fp32 leg, strong scaling, dp_check)")
    ap.add_argument("--profile-only", action="store_true", help="warm-up + one update, no timing legs (for ncu)")
    ap.add_argument("--timeline", default=None, metavar="CSV",
                    help="warm-up, then ONE update under the CUPTI kernel tracer (torch.profiler): per-kernel in-situ start / "
                         "duration / stream written to CSV, summarised by tools/timeline_summary.py; no timing legs")
    args = ap.parse_args()
    # stdout carries exactly one JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL prints
    # its version banner there when NCCL_DEBUG=VERSION and ignores NCCL_DEBUG_FILE at that level), so fd 1 is pointed at
    # stderr for the whole run and the result line goes to a private duplicate of the original stdout.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    arm_watchdog()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from oracle import reppo_oracle as O  # synthetic inputs + cpu_baseline leg only
    from reppo_b200 import torch_api as T
    from reppo_b200.engine import HyperParams

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    if world > 1:
        # the per-step gradient all-reduces are ~1-5 MB and latency-bound: NCCL's LL protocol measured 2 % faster per update than
        # its own choice at 2 and at 8 GPUs (115.2 -> 112.8 ms, 124.6 -> 122.1 ms); LL128 and a channel cap measured slower
        os.environ.setdefault("NCCL_PROTO", dp_defaults()["NCCL_PROTO"])
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    peaks = load_peaks()

    cfg, D, A, term, rs = make_cfg(args.config, args.semantics, args.override)
    if args.gemm in ("bf16", "tf32"):
        cfg.platform.amp_dtype = args.gemm
    h = cfg.hyperparameters
    hp = oracle_hyper(cfg, D, A)
    S, n_mb, E = hp.batch_size, hp.num_mini_batches, hp.num_epochs
    mb = S // n_mb
    sem = O.JAX if args.semantics == "jax" else O.TORCH

    # ---- synthetic rollout (host, seeded per rank), networks -------------------------------------------
    batch = O.synthetic_rollout(hp, seed=rank, terminate=term, reward_scaling=rs)
    cp, ap_ = O.init_params(hp, sem, seed=0)
    ts = T.make_train_state(cfg, D, D, A, device=dev)
    eng = ts.engine
    eng.load_params(cp, ap_)
    if world > 1:
        from reppo_b200 import dist as dp_plumbing
        dp_plumbing.attach_nccl(eng, rank, world)   # two communicators: critic chain + actor chain
        _attach_multicast(eng, dp_plumbing)         # ... superseded by the multicast all-reduce inside clip+Adam when the box has it
    hyp = HyperParams(gamma=h.gamma, lmbda=h.lmbda, lr=h.lr, max_grad_norm=h.max_grad_norm, kl_bound=h.kl_bound,
                      ent_target_mult=h.ent_target_mult, aux_loss_mult=h.aux_loss_mult, actor_min_std=h.actor_min_std,
                      actor_kl_clip_mode=h.actor_kl_clip_mode, world_size=world)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    devb = {k: v.to(dev) for k, v in batch.items()}
    perms = torch.stack([torch.randperm(S, device=dev, generator=g) for _ in range(E)]).to(torch.int32)
    eps_pi = torch.randn(E, mb, A, device=dev, generator=g)
    eps_kl = torch.randn(E, 16, mb, A, device=dev, generator=g)
    target = torch.empty(hp.num_steps, hp.num_envs, device=dev)
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in devb.items()}
    flat["target"] = target.reshape(-1)
    cb = eng.make_batch(flat)
    step_metrics = torch.empty(E * n_mb, 24, device=dev)
    use_graph = not args.no_graph

    def one_update():
        # the update's own random draws are inside the timed region (learn_step shuffles and samples inside,
        # src/jaxrl/reppo.py:535-557,645-647); same buffers every time, so the captured graph is replayed
        for e in range(E):
            perms[e].copy_(torch.randperm(S, device=dev, generator=g))
        eps_pi.normal_(generator=g)
        eps_kl.normal_(generator=g)
        eng.td_lambda(devb["soft_reward"], devb["value"], devb["done"], devb["truncated"], devb["importance_weight"],
                      h.gamma, h.lmbda, out=target)
        eng.sync_actor_old()
        eng.update(cb, perms, eps_pi, eps_kl, hyp, n_mb, use_graph=use_graph, step_metrics=step_metrics)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage(f"main workload {args.config}: warm-up")
    for _ in range(args.warmup):
        one_update()
    barrier()
    stage("main workload: timed region")
    if args.profile_only:
        one_update()
        torch.cuda.synchronize()
        return
    if args.timeline:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one_update()
            torch.cuda.synchronize()
        rows = []
        for ev in prof.profiler.kineto_results.events():
            if "cuda" in str(ev.device_type()).lower():
                rows.append((ev.start_ns() / 1e3, ev.duration_ns() / 1e3, ev.device_resource_id(), ev.name()))
        rows.sort()
        import csv
        with open(args.timeline, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["start_us", "dur_us", "stream", "name"])
            t0_ = rows[0][0] if rows else 0
            for st_, du_, dv_, nm_ in rows:
                w.writerow([round(st_ - t0_, 3), round(du_, 3), dv_, nm_[:100]])
        emit({"timeline": args.timeline, "kernels": len(rows)})
        return
    launches_per_update = eng.launch_count() + 2  # + scan + actor_old copy
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        one_update()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    final_metrics = eng.metrics_dict(eng.metrics)
    mc_phases = None
    if world > 1 and os.environ.get("REPPO_MC_STAMPS") and ALLREDUCE["mode"].startswith("multicast"):
        from reppo_b200 import dist as dp_plumbing
        mc_phases = dp_plumbing.mc_phase_summary(eng)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = S * world / (ms / 1e3)

    # ---- e2e: public API, host pinned buffers, H2D + D2H inside the timed region ---------------------------
    host = {"observations": batch["obs"], "critic_observations": batch["critic_obs"], "actions": batch["action"],
            "rewards": batch["soft_reward"].unsqueeze(-1), "raw_rewards": batch["reward"].unsqueeze(-1),
            "next_embeddings": batch["next_emb"], "next_values": batch["value"].unsqueeze(-1),
            "dones": batch["done"].unsqueeze(-1), "truncations": batch["truncated"].unsqueeze(-1),
            "importance_weight": batch["importance_weight"].unsqueeze(-1)}
    stage("e2e leg")
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    # Input pipeline of the public API: the H2D copy of the NEXT rollout (prefetch_next) is issued inside every call and
    # overlaps that call's kernels (copy stream, second staging set).  Every timed step therefore still contains one full
    # 292 MB H2D and one metrics D2H; the clock stops only after the last copy has completed (barrier = device sync).
    T.reppo_prefetch(ts, host, cfg, generator=g)
    for _ in range(2):
        T.reppo_update(ts, host, cfg, generator=g, use_graph=use_graph, world_size=world, prefetch_next=host)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        logs = T.reppo_update(ts, host, cfg, generator=g, use_graph=use_graph, world_size=world, prefetch_next=host)  # floats: D2H read inside
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e = {"value": S * world / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 24 * 4,
           "ms_per_step": e2e_ms, "pipeline": "H2D of rollout k+1 overlaps update k (torch_api.reppo_update prefetch_next)"}

    # ---- the other named workloads, the reference-precision leg, strong scaling, DP numerics (all ranks take part) -----
    extra_cfgs, fp32_leg, tf32_leg, strong_leg, dpc = None, None, None, None, None
    if not args.no_extra:
        # (the main workload's engine stays alive for the roofline legs below)
        extra_cfgs = {}
        for name in sorted(CONFIGS):
            if name == args.config:
                continue
            stage(f"extra workload {name}")
            extra_cfgs[name] = measure_config(name, args.semantics, args.gemm, world, rank, dev, steps=3, warmup=2)
        if args.gemm != "fp32":
            stage("fp32 leg")
            fp32_leg = measure_config(args.config, args.semantics, "fp32", world, rank, dev, steps=2, warmup=1)
            fp32_leg["note"] = ("reference-precision mode (fp32 FFMA GEMMs, src/jaxrl/__init__.py:3 'highest'); the headline value is "
                                "bf16 operands / fp32 accumulate like the reference Torch path's 'medium' (src/torchrl/reppo.py:35)")
        if args.gemm != "tf32":
            stage("tf32 leg")
            tf32_leg = measure_config(args.config, args.semantics, "tf32", world, rank, dev, steps=2, warmup=1)
            tf32_leg["dtype"] = "tf32"
            tf32_leg["note"] = ("reference-precision mode on tensor cores: tcgen05.mma.kind::tf32 on the fp32 activations and an aligned "
                                "fp32 copy of the weights (10-bit mantissa products, fp32 accumulation; what "
                                "torch.set_float32_matmul_precision('high') asks for); activation operands with ragged leading dimensions "
                                "(obs + act inputs, the 151-bin / H+1 / 2A head gradients) are re-laid into row-padded scratch first")
        if world > 1:
            stage("strong-scaling leg")
            strong_leg = measure_config(args.config, args.semantics, args.gemm, world, rank, dev, steps=3, warmup=2, strong=True)
            stage("dp_check")
            dpc = dp_check(world, rank, dev, args.gemm)
        stage("extra legs done")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline legs (rank 0, N=1 semantics) ----------------------------------------------------------------
    def time_fn(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # GEMM family: every distinct Linear GEMM of one critic+actor step, timed in isolation with CUDA events
    Hc, Ha, B, K0 = hp.critic_hidden_dim, hp.actor_hidden_dim, hp.num_bins, hp.critic_obs_dim + A
    P = Hc + (1 if args.semantics == "jax" else 0)
    crit_layers = [(K0, Hc), (Hc, Hc), (Hc, Hc), (Hc, B), (Hc, Hc), (Hc, P)]   # feature x2, head x2, pred x2
    act_layers = [(D, Ha), (Ha, Ha), (Ha, 2 * A)]
    # (op, in, out, count per critic+actor step): critic step fwd+dgrad+wgrad; actor step: actor fwd/old fwd/bwd, critic fwd+dgrad
    plan = {}
    def add(op, i, o, n=1):
        plan[(op, i, o)] = plan.get((op, i, o), 0) + n
    for i, o in crit_layers:
        add(0, i, o); add(2, i, o)
    for i, o in crit_layers[1:]:
        add(1, i, o)
    for i, o in crit_layers[:4]:
        add(0, i, o); add(1, i, o)
    for i, o in act_layers:
        add(0, i, o, 2); add(2, i, o)
    for i, o in act_layers[1:]:
        add(1, i, o)
    gemm_ms, gemm_flops, l2_bytes = 0.0, 0.0, 0.0
    top = None
    for (op, i, o), n in plan.items():
        if op == 0:
            a_, b_ = torch.randn(mb, i, device=dev), torch.randn(o, i, device=dev)
        elif op == 1:
            a_, b_ = torch.randn(mb, o, device=dev), torch.randn(o, i, device=dev)
        else:
            a_, b_ = torch.randn(mb, o, device=dev), torch.randn(mb, i, device=dev)
        out = eng.linear(op, a_, b_)     # stages the bf16 operands (tensor-core mode)
        flag = 0x100 if args.gemm == "bf16" else 0   # then time the GEMM kernel alone
        # (wgrad accumulates into `out` with atomics: the values grow, the work per launch is unchanged)
        # 50 back-to-back launches are captured into a CUDA graph so that the host launch path (ctypes + tensor-map
        # encode, ~6 us per call) is not what the CUDA events measure
        reps = 50
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                for _ in range(reps):
                    eng.lib.reppo_linear(eng.ctx, op | flag, mb, i, o, a_.data_ptr(), b_.data_ptr(), out.data_ptr(), None,
                                         eng._stream())
        torch.cuda.current_stream().wait_stream(side)
        t_ms = time_fn(lambda: gr.replay(), 5) / reps
        fl = 2.0 * mb * i * o
        gemm_ms += n * t_ms
        gemm_flops += n * fl
        if top is None or n * t_ms > top[0]:
            top = (n * t_ms, op, i, o, t_ms, fl)
        # bf16 operand bytes the kernel's CTAs pull from L2 with 128 x 64 tiles: every 64-column tile re-reads its A slab,
        # every 128-row tile its B slab (the L2 -> SM fabric, not the tensor pipe, bounds these short-K GEMMs: DESIGN.md 7)
        m_, n_, k_ = (mb, o, i) if op == 0 else ((mb, i, o) if op == 1 else (o, i, mb))
        op_bytes = 2.0 * k_ * (m_ * math.ceil(n_ / 64) + n_ * math.ceil(m_ / 128))
        l2_bytes += n * op_bytes
    n_steps = E * n_mb
    peak_tf = peaks["bf16_tflops_sustained"]
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12
    traffic, traffic_src = load_traffic()
    roofline = {
        "bound": "tensor", "kernel": "gemm_f32_kernel (fp32 FFMA)" if args.gemm == "fp32" else "gemm_bf16_tc_kernel (tcgen05)",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
        "traffic_note": f"mean dram__bytes_read+write per GEMM launch inside the update, profiles/{traffic_src} (operands are "
                        "L2-resident: below the algorithmic 2.5-8.5 MB per launch)" if traffic is not None else None,
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
        "method": "every distinct Linear GEMM of one critic+actor minibatch step timed alone: 50 back-to-back launches "
                  "replayed as a CUDA graph between CUDA events, weighted by its launches per step",
        "gemm_ms_per_update": gemm_ms * n_steps, "gemm_share_of_update": gemm_ms * n_steps / ms,
        "top_gemm": {"op": ["fwd", "dgrad", "wgrad"][top[1]], "in": top[2], "out": top[3], "rows": mb, "ms": top[4],
                     "tflops": top[5] / (top[4] * 1e-3) / 1e12},
        "in_update_tflops": gemm_flops_per_update(hp) / (ms * 1e-3) / 1e12,
        "l2_operand_view": {"operand_gb_per_step": l2_bytes / 1e9, "achieved_gbs": l2_bytes / (gemm_ms * 1e-3) / 1e9,
                            "fabric_gbs": 6300 * 1.965, "frac_of_fabric": l2_bytes / (gemm_ms * 1e-3) / 1e9 / (6300 * 1.965),
                            "note": "bytes the GEMM CTAs pull from L2 per critic+actor step (128 x 64 tiles) / the same isolated "
                                    "GEMM time, against the chip-wide L2 -> SM fabric (B300 guide: LTS cap ~6300 B/clk, here at "
                                    "1965 MHz; 73 GB/s/SM measured in the 83%-tensor-active pair GEMM).  The 2048 x 512 x 512 "
                                    "launches reach ~5.2 TB/s in their main loop; launch / first-load latency and the thin layers "
                                    "pull the step average down"},
    }

    # TD(lambda) scan bandwidth on a batch far larger than L2 (the named shape moves 3 MB: launch-latency bound)
    Tn, Nn = 128, 1 << 20
    big = [torch.rand(Tn, Nn, device=dev) for _ in range(2)] + [-torch.rand(Tn, Nn, device=dev)] + \
          [torch.zeros(Tn, Nn, device=dev), torch.zeros(Tn, Nn, device=dev)]
    outb = torch.empty(Tn, Nn, device=dev)
    scan_ms = time_fn(lambda: eng.td_lambda(big[0], big[1], big[3], big[4], big[2], 0.99, 0.95, convention="jax", out=outb), 10)
    scan_bytes = 24.0 * Tn * Nn
    scan = {"kernel": "td_lambda_kernel<4,4,jax,iw>", "bound": "hbm", "achieved": scan_bytes / (scan_ms * 1e-3) / 1e9,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": scan_bytes / (scan_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "shape": [Tn, Nn], "ms": scan_ms, "algorithmic_bytes": scan_bytes,
            "named_shape_latency_ms": time_fn(lambda: eng.td_lambda(devb["soft_reward"], devb["value"], devb["done"],
                                                                    devb["truncated"], devb["importance_weight"], h.gamma,
                                                                    h.lmbda, out=target), 50)}
    del big, outb

    # clip+Adam bandwidth (28 B / parameter)
    import ctypes as C
    ns, hs = eng.critic.c_state(), hyp.c_struct()
    adam_ms = time_fn(lambda: eng.lib.reppo_clip_adam(eng.ctx, C.byref(ns), eng.critic.count, C.byref(hs), None, eng._stream()), 50)
    adam = {"kernel": "clip_adam_kernel (two launches: per-block partial sums of squares, then global norm + clip + Adam + bf16 shadow refresh)", "bound": "hbm (L2-resident at this size)",
            "achieved": 28.0 * eng.critic.count / (adam_ms * 1e-3) / 1e9, "unit": "GB/s", "ms": adam_ms, "params": eng.critic.count}

    # rollout-side inference: one policy step + one bootstrap step over this GPU's environments
    n_env = hp.num_envs
    r_obs = torch.randn(n_env, D, device=dev)
    r_eps = torch.randn(n_env, A, device=dev)
    r_rew = torch.rand(n_env, device=dev)
    def rollout_step():
        eng.policy_act(r_obs, r_eps, clip=0.999, min_std=h.actor_min_std)
        eng.bootstrap(r_obs, r_obs, r_eps, r_rew, h.gamma, clip=1.0 - 1e-6, min_std=h.actor_min_std)
    rs_ms = time_fn(rollout_step, 50)
    nrm = T.EmpiricalNormalization(D, device=dev)
    big_obs = torch.randn(1 << 20, D, device=dev)
    nrm.train()
    nrm_ms = time_fn(lambda: nrm(big_obs), 20)
    rollout = {"kernels": "policy_sample_kernel + MLP forwards + value_head_kernel + soft_reward_kernel", "envs": n_env,
               "ms_per_env_step": rs_ms, "env_steps_per_s": n_env / (rs_ms * 1e-3),
               "obs_normalize": {"rows": 1 << 20, "dim": D, "ms": nrm_ms, "bound": "hbm",
                                 "achieved": (2 * 4.0 * (1 << 20) * D + 4.0 * (1 << 20) * D) / (nrm_ms * 1e-3) / 1e9, "unit": "GB/s",
                                 "peak": peaks["hbm_gbs"],
                                 "frac": (3 * 4.0 * (1 << 20) * D) / (nrm_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                 "algorithmic_bytes": "2 reads (one-pass statistics, apply) + 1 write of the batch"}}
    del big_obs

    # loss / sampling kernels at a row count where they move more than L2 holds (S = 1M rows): HBM GB/s against the measured
    # copy bandwidth (algorithmic bytes per row: SURVEY 8(d))
    loss_kernels = None
    if not args.no_large:
        R_ = 1 << 20
        Bn, Pn, Hn = hp.num_bins, Hc + (1 if args.semantics == "jax" else 0), Hc
        probes = []
        def probe(name, which, ins, out, scratch, bytes_per_row):
            t_ms = time_fn(lambda: eng.kernel_probe(which, ins[0], ins[1], ins[2], ins[3], out, scratch, hyp), 10)
            gbs = bytes_per_row * R_ / (t_ms * 1e-3) / 1e9
            probes.append({"kernel": name, "rows": R_, "ms": t_ms, "algorithmic_bytes_per_row": bytes_per_row, "achieved": gbs,
                           "unit": "GB/s", "peak": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"]})
        x_ = torch.randn(R_, Bn, device=dev); o_ = torch.empty(R_, Bn, device=dev)
        tg_ = float(h.vmin) + (float(h.vmax) - float(h.vmin)) * torch.rand(R_, device=dev)
        probe("critic_ce_kernel", 0, (x_, tg_, torch.zeros(R_, device=dev), torch.zeros(Bn, device=dev)), o_,
              torch.zeros(24 + 2 * Bn, device=dev), 2.0 * Bn * 4 + 12)
        del x_, o_, tg_
        x_ = torch.randn(R_, Pn, device=dev); ne_ = torch.randn(R_, Hn, device=dev); o_ = torch.empty(R_, Pn, device=dev)
        probe("critic_aux_kernel", 1, (x_, ne_, torch.rand(R_, device=dev), torch.zeros(R_, device=dev)), o_,
              torch.zeros(24 + Pn, device=dev), (2.0 * Pn + Hn) * 4 + 8)
        del x_, ne_, o_
        x_ = 0.3 * torch.randn(R_, 2 * A, device=dev); xo_ = 0.3 * torch.randn(R_, 2 * A, device=dev)
        probe("actor_sample_kernel", 2, (x_, xo_, torch.randn(R_, A, device=dev), torch.randn(16, R_, A, device=dev)),
              torch.empty(R_, A, device=dev), torch.zeros(R_ * (2 + 2 * A), device=dev), (2 * 2 * A + A + 16 * A + A + 2 + 2 * A) * 4.0)
        del x_, xo_
        torch.cuda.empty_cache()
        loss_kernels = {"bound": "hbm", "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})", "kernels": probes}

    # the large Linear shapes of the wide-net / 256k-sample sweep (BASELINE.json configs[4]): persistent
    # CTA-pair tcgen05 kernel (gemm_bf16_tc2.cu), each GEMM timed alone against the measured bf16 BURST peak
    gemm_large = None
    if args.gemm == "bf16" and not args.no_large:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import gemm_bench
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
            else {"bf16_tflops": peaks["bf16_tflops"]}
        rows_l = []
        for rows_, H_ in ((16384, 1024), (32768, 2048)):
            rows_l += gemm_bench.measure(rows_, H_, reps=20, check=False, peaks=pk)
        gemm_large = {"kernel": "gemm_bf16_tc2_kernel (tcgen05 cta_group::2, persistent)", "bound": "tensor", "unit": "TFLOP/s",
                      "peak": pk["bf16_tflops"], "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)",
                      "shapes": rows_l, "best_frac": max(r["frac_burst"] for r in rows_l)}

    cpu = None
    if not args.no_cpu_baseline:
        arm = CpuArm(hp, args.semantics, m_steps=4 * hp.num_mini_batches)   # ~10 s of CPU work at C2
        arm.run_warm = arm.m
        arm.m = 2
        arm.run()                      # warm-up (allocator, thread pool)
        arm.m = arm.run_warm
        cpu = arm.describe(arm.run())

    line = {
        "metric": "reppo_update_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.gemm == "fp32" else "bf16", "data": "synthetic",
        "config": workload_config(args.config, cfg, D, A, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_update * args.steps),
        "launches_per_update": int(launches_per_update), "cuda_graph": use_graph, "semantics": args.semantics,
        "sample_visits_per_s": value * E, "roofline": roofline, "gemm_large": gemm_large, "loss_kernels": loss_kernels, "scan": scan, "adam": adam, "rollout": rollout,
        "cpu_baseline": cpu,
        "configs": extra_cfgs, "fp32": fp32_leg, "tf32": tf32_leg, "strong": strong_leg, "dp_check": dpc,
        "knobs": {k: os.environ[k] for k in os.environ if k.startswith("REPPO_")},
        "allreduce": ALLREDUCE["mode"] if world > 1 else None,
        "nccl_env": {k: os.environ[k] for k in os.environ if k.startswith("NCCL_")} if world > 1 else None,
        "mc_phases": mc_phases,
        "final_metrics": {k: final_metrics[k] for k in ("critic_loss", "actor_loss", "kl", "entropy")},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
