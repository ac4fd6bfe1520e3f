"""The Torch trainer entry point: ``main`` of src/torchrl/reppo.py:437-730 over this package's engine.

Same stages, same order, same config surface (``reppo.yaml`` + ``env`` / ``platform`` / ``experiment_overrides`` groups):

    make_envs -> normalisers -> Actor / Critic / old_actor -> collect_fn -> postprocess_fn (TD(lambda)) ->
    num_epochs x num_mini_batches x (update_critic, update_actor) -> old_actor <- actor -> logs / evaluate

The update stage runs either through the three reference closures, minibatch by minibatch (``fused=False``: the literal loop of
src/torchrl/reppo.py:614-632), or -- the default -- through ``torch_api.reppo_update``, which runs the same arithmetic for
the whole rollout as one captured CUDA graph.  Environments are anything with the interface of the reference's wrappers
(src/env_utils/torch_wrappers/*.py: ``num_envs``, ``num_obs``, ``num_actions``, ``asymmetric_obs``, ``max_episode_steps``,
``reset()``, ``step(actions)``); simulators hand their buffers over as CUDA tensors (DLPack inside the wrappers), so the
rollout never leaves the device.  ``env.type = synthetic`` is a device-resident stand-in (no simulator is installable in
this image) that keeps the whole loop testable.

Out of scope, as in DESIGN.md section 8: wandb, tqdm, AMP grad scaling, ``torch.compile`` (all no-ops or absent here).
"""
from __future__ import annotations

import random
import time
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import torch_api as T
from .config import DictConfig, compose, validate_update_config


# ---- environments ---------------------------------------------------------------------------------------------------
class SyntheticEnv:
    """A device-resident vector environment with the reference wrappers' interface.

    State ``x [N, n_obs]`` follows ``x' = tanh(x W_x + a W_a) + 0.01 noise``; the reward prefers small states and small
    actions; an episode terminates when ``max |x| > 0.98`` and is truncated after ``max_episode_steps``; finished
    environments reset in place (auto-reset, like the MJX / Brax wrappers).  Every draw comes from the environment's own
    generator, so two environments built with one seed produce identical rollouts for identical actions."""

    def __init__(self, num_envs: int, n_obs: int, n_act: int, device, seed: int = 0, max_episode_steps: int = 64,
                 asymmetric_obs: bool = False, n_critic_obs: Optional[int] = None):
        self.num_envs, self.num_obs, self.num_actions = int(num_envs), int(n_obs), int(n_act)
        self.asymmetric_obs = bool(asymmetric_obs)
        self.num_privileged_obs = int(n_critic_obs or n_obs) if asymmetric_obs else 0
        self.max_episode_steps = int(max_episode_steps)
        self.device = torch.device(device)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed))
        g = torch.Generator().manual_seed(int(seed) + 1)
        self._wx = (torch.randn(n_obs, n_obs, generator=g) * (0.9 / n_obs ** 0.5)).to(self.device)
        self._wa = (torch.randn(n_act, n_obs, generator=g) * (0.5 / n_act ** 0.5)).to(self.device)
        self._wc = (torch.randn(n_obs, self.num_privileged_obs, generator=g) / n_obs ** 0.5).to(self.device) if asymmetric_obs else None
        self._x = torch.zeros(self.num_envs, n_obs, device=self.device)
        self._t = torch.zeros(self.num_envs, device=self.device)

    def _fresh(self, n):
        return 0.5 * torch.randn(n, self.num_obs, device=self.device, generator=self._gen)

    def _critic(self, x):
        return torch.tanh(x @ self._wc)

    def reset(self, **_):
        self._x = self._fresh(self.num_envs)
        self._t.zero_()
        return self._x.clone(), {}

    def reset_with_critic_obs(self):
        obs, _ = self.reset()
        return obs, self._critic(obs)

    def step(self, actions):
        a = actions.to(self.device, torch.float32)
        x = torch.tanh(self._x @ self._wx + a @ self._wa) + 0.01 * torch.randn(self._x.shape, device=self.device, generator=self._gen)
        self._t += 1
        rewards = 1.0 - x.pow(2).mean(-1) - 0.1 * a.pow(2).mean(-1)
        terminated = x.abs().amax(-1) > 0.98
        truncated = (self._t >= self.max_episode_steps) & ~terminated
        finished = terminated | truncated
        infos = {"final_observation": x.clone()}
        x = torch.where(finished.unsqueeze(-1), self._fresh(self.num_envs), x)
        self._t = torch.where(finished, torch.zeros_like(self._t), self._t)
        self._x = x
        if self.asymmetric_obs:
            infos["observations"] = {"critic": self._critic(x)}
        return x.clone(), rewards, terminated, truncated, infos


def make_envs(cfg: DictConfig, device, seed: Optional[int] = None):
    """``make_envs`` of src/torchrl/envs.py:5-97.  The simulator-backed types import the reference's own wrapper modules
    (they are environment code, out of scope here) and fail with the import error when the simulator is absent."""
    kind = cfg.env.type
    h = cfg.hyperparameters
    if kind == "synthetic":
        e = cfg.env
        mk = lambda s: SyntheticEnv(int(h.num_envs), int(e.get("num_obs", 17)), int(e.get("num_actions", 6)), device, seed=s,
                                    max_episode_steps=int(e.get("max_episode_steps", 64)),
                                    asymmetric_obs=bool(e.get("asymmetric_obs", False)),
                                    n_critic_obs=e.get("num_privileged_obs", None))
        return mk(seed or 0), mk((seed or 0) + 10_000)
    if kind in ("humanoid_bench", "isaaclab", "mjx", "maniskill", "mtbench"):
        try:
            from src.torchrl.envs import make_envs as _ref_make_envs   # the reference checkout on sys.path
        except ImportError as e:
            raise ImportError(f"env.type={kind} needs the reference's environment wrappers (src/torchrl/envs.py) and the "
                              f"simulator they wrap on sys.path: {e}") from e
        return _ref_make_envs(cfg, device, seed)
    raise ValueError(f"Unknown environment type: {kind}. Supported types are 'humanoid_bench', 'isaaclab', 'maniskill', "
                     f"'mjx' and 'synthetic'.")


# ---- evaluation (src/torchrl/reppo.py:362-418) ----------------------------------------------------------------------
def make_evaluate_fn(cfg: DictConfig, eval_envs):
    @torch.no_grad()
    def evaluate(train_state: T.TrainState, stochastic_eval: bool = False):
        eng = T._engine_for(cfg, train_state)
        dev = eng.device
        train_state.normalizer.eval()
        n = eval_envs.num_envs
        episode_returns = torch.zeros(n, device=dev)
        episode_lengths = torch.zeros(n, device=dev)
        done_masks = torch.zeros(n, dtype=torch.bool, device=dev)
        if cfg.env.type == "isaaclab" or cfg.env.get("asymmetric_obs", False):
            obs, _ = eval_envs.reset(random_start_init=False)
        else:
            obs, _ = eval_envs.reset()
        A = eng.cfg.action_dim
        infos = {}
        for _ in range(int(eval_envs.max_episode_steps)):
            obs = train_state.normalizer(obs)
            eps = torch.randn(n, A, device=dev) if stochastic_eval else torch.zeros(n, A, device=dev)
            # deterministic action = tanh(mean): the policy kernel with zero noise
            actions, _, _ = eng.policy_act(obs, eps, clip=0.999, min_std=float(train_state.actor.min_std))
            next_obs, rewards, dones, _, infos = eval_envs.step(actions)
            rewards = rewards.to(dev, torch.float32)
            episode_returns = torch.where(~done_masks, episode_returns + rewards, episode_returns)
            episode_lengths = torch.where(~done_masks, episode_lengths + 1, episode_lengths)
            done_masks = torch.logical_or(done_masks, dones.to(dev).bool())
            if done_masks.all():
                break
            obs = next_obs
        train_state.normalizer.train()
        if cfg.env.type == "maniskill":
            info = {"info_return": infos["log_info"]["return"].mean(), "episode_len": infos["log_info"]["episode_len"].float().mean(),
                    "success": infos["log_info"]["success"].float().mean(), "return": episode_returns.mean().item()}
        else:
            info = {}
        return episode_returns.mean().item(), episode_lengths.mean().item(), info

    return evaluate


def configure_platform(cfg: DictConfig) -> DictConfig:
    """src/torchrl/reppo.py:421-434.  There is one platform here: a CUDA device; AMP is the engine's bf16 operand mode."""
    if not torch.cuda.is_available():
        raise RuntimeError("reppo_b200 has no CPU path: a CUDA device is required")
    p = cfg.setdefault("platform", DictConfig()) if isinstance(cfg, dict) else cfg.platform
    p["amp_device"] = "cuda"
    return cfg


_LOG_KEYS = (("critic/qf_loss", "qf_loss"), ("critic/qf_max", "qf_max"), ("critic/qf_min", "qf_min"), ("critic/qf_mean", "qf_mean"),
             ("critic/embedding_loss", "embedding_loss"), ("critic/critic_grad_norm", "critic_grad_norm"),
             ("actor/actor_loss", "actor_loss"), ("actor/actor_grad_norm", "actor_grad_norm"), ("actor/kl", "kl"),
             ("actor/entropy", "entropy"), ("actor/temperature", "temperature"), ("actor/lagrangian", "lagrangian"),
             ("actor/entropy_loss", "entropy_loss"), ("actor/lagrangian_loss", "lagrangian_loss"))


def _scalar(v):
    return float(v.mean().item()) if isinstance(v, torch.Tensor) else float(v)


def main(cfg: Optional[DictConfig] = None, overrides: Sequence[str] = (), *, envs=None, eval_envs=None,
         log_fn: Optional[Callable[[dict, int], None]] = None, fused: bool = True, max_updates: Optional[int] = None,
         return_state: bool = False, checkpoint_path: Optional[str] = None, save_path: Optional[str] = None):
    """Train.  ``cfg``: a composed config (``config.compose``) or None to compose ``overrides`` like the hydra CLI.
    ``envs`` / ``eval_envs``: pre-built environments (else ``make_envs``).  ``log_fn(logs, frame)`` stands where the
    reference calls ``wandb.log``.  ``checkpoint_path`` resumes from a file in the reference's checkpoint layout (the branch the
    reference leaves as a TODO, src/torchrl/reppo.py:579-594), ``save_path`` writes one at the end (``torch_api.save_params`` /
    ``load_params``, src/torchrl/reppo_util.py:723-753).  Returns the list of per-update log dicts (and the final TrainState on
    request)."""
    if cfg is None:
        # the Torch entry point does not merge experiment_overrides into hyperparameters (src/torchrl/reppo.py never does)
        cfg = compose(list(overrides), merge_overrides=False)
    cfg = configure_platform(cfg)
    validate_update_config(cfg)
    h = cfg.hyperparameters
    num_batches = int(h.num_mini_batches)
    batch_size = int(h.num_envs) * int(h.num_steps) // num_batches
    seed = int(cfg.get("seed", 0))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    device = torch.device(f"cuda:{int(cfg.platform.get('device_rank', 0))}")
    torch.cuda.set_device(device)
    if envs is None:
        envs, eval_envs = make_envs(cfg, device, seed)
    if eval_envs is None:
        eval_envs = envs

    n_act = envs.num_actions
    n_obs = envs.num_obs if isinstance(envs.num_obs, int) else envs.num_obs[0]
    if envs.asymmetric_obs:
        n_critic_obs = envs.num_privileged_obs if isinstance(envs.num_privileged_obs, int) else envs.num_privileged_obs[0]
    else:
        n_critic_obs = n_obs

    # Actor / Critic / old_actor / optimizers bound to one engine (src/torchrl/reppo.py:497-531)
    train_state = T.make_train_state(cfg, n_obs, n_critic_obs, n_act, device=device)
    if h.normalize_env:
        train_state.normalizer = T.EmpiricalNormalization(shape=n_obs, device=device)
        train_state.critic_normalizer = T.EmpiricalNormalization(shape=n_critic_obs, device=device)
    else:
        train_state.normalizer, train_state.critic_normalizer = nn.Identity(), nn.Identity()
    if envs.asymmetric_obs:
        obs, critic_obs = envs.reset_with_critic_obs()
        critic_obs = torch.as_tensor(critic_obs, device=device, dtype=torch.float)
    else:
        obs, _ = envs.reset()
        critic_obs = obs
    train_state.obs, train_state.critic_obs = obs, critic_obs

    collect_fn = T.make_collect_fn(cfg, envs)
    evaluate = make_evaluate_fn(cfg, eval_envs)
    if not fused:
        postprocess_fn = T.make_postprocess_fn(cfg, envs)
        update_critic = T.make_critic_update_fn(cfg, train_state)
        update_actor = T.make_actor_update_fn(cfg, train_state)

    global_step = 0
    if checkpoint_path:
        global_step = int(T.load_params(checkpoint_path, train_state.actor, train_state.critic, train_state.normalizer,
                                        train_state.critic_normalizer, old_actor=train_state.old_actor))
    total_env_steps = int(h.total_time_steps) // (int(h.num_envs) * int(h.num_steps)) + 1
    if max_updates is not None:
        total_env_steps = min(total_env_steps, int(max_updates))
    eval_interval = total_env_steps // max(int(h.num_eval), 1) if int(h.get("num_eval", 0)) > 0 else 0
    stochastic_eval = cfg.env.get("stochastic_eval", False)
    start_time, measure_burnin = None, 0
    frames = int(h.num_envs) * int(h.num_steps)
    history: List[dict] = []

    while global_step < total_env_steps:
        if start_time is None and global_step >= int(cfg.get("measure_burnin", 0)):
            start_time = time.time()
            measure_burnin = global_step

        train_state, transition, infos = collect_fn(train_state)
        if fused:
            logs_dict = T.reppo_update(train_state, transition, cfg)
            rewards_batch = transition["rewards"].mean()
        else:
            data = postprocess_fn(train_state, transition)
            for _ in range(int(h.num_epochs)):
                indices = torch.randperm(frames, device=device)
                data = data[indices].contiguous()
                for j in range(num_batches):
                    mini_batch = data[j * batch_size:(j + 1) * batch_size]
                    logs_dict = {**update_critic(mini_batch), **update_actor(mini_batch)}
            T._engine_for(cfg, train_state).sync_actor_old()   # old_actor <- actor (src/torchrl/reppo.py:631-632)
            rewards_batch = data["rewards"].mean()

        if start_time is not None:
            torch.cuda.synchronize(device)
            speed = frames * (global_step - measure_burnin) / max(time.time() - start_time, 1e-9)
            logs = {k: _scalar(logs_dict[src]) for k, src in _LOG_KEYS}
            logs["train/rewards_batch"] = _scalar(rewards_batch)
            if cfg.env.type == "maniskill":
                logs.update({
                    "train/return": torch.stack([i["log_info"]["return"] for i in infos]).mean().item(),
                    "train/episode_len": torch.stack([i["log_info"]["episode_len"] for i in infos]).float().mean().item(),
                    "train/success": torch.stack([i["log_info"]["success"] for i in infos]).float().mean().item()})
            if eval_interval > 0 and global_step % eval_interval == 0:
                if stochastic_eval:
                    _, _, stoch_info = evaluate(train_state, stochastic_eval=True)
                    ret, length, eval_info = evaluate(train_state)
                    eval_info = {**eval_info, **{f"stoch/{k}": v for k, v in stoch_info.items()}}
                else:
                    ret, length, eval_info = evaluate(train_state)
                if cfg.env.type in ("humanoid_bench", "isaaclab", "mtbench"):
                    train_state.obs, _ = envs.reset()   # these simulators share the training environments with evaluation
                logs["eval/avg_return"], logs["eval/avg_length"] = ret, length
                for k, v in eval_info.items():
                    logs[f"eval/{k}"] = _scalar(v) if isinstance(v, (torch.Tensor, np.ndarray)) else v
            logs["speed"], logs["frame"] = speed, global_step * frames
            history.append(logs)
            if log_fn is not None:
                log_fn(logs, global_step * frames)
        global_step += 1
    if save_path:
        T.save_params(global_step, train_state.actor, train_state.critic, None, train_state.normalizer, train_state.critic_normalizer,
                      {"env": dict(cfg.env), "hyperparameters": dict(h)}, save_path)
    return (history, train_state) if return_state else history


if __name__ == "__main__":
    import json
    import sys
    for entry in main(overrides=sys.argv[1:]):
        print(json.dumps(entry))
