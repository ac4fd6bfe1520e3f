"""Dump fixtures from the REAL JAX ``learn_step`` (src/jaxrl/reppo.py:417-673).  TEST INFRASTRUCTURE; needs JAX.

The build image has no jax / flax / optax / distrax, so this script has NOT been run there: the JAX-semantics constants of
``oracle/reppo_oracle.py`` stay "parity unpinned"  until somebody runs it where the reference itself
runs (CPU backend is enough):

    cd <reference checkout> && JAX_PLATFORMS=cpu python <this repo>/oracle/gen_golden_jax.py --out <this repo>/tests/golden

and commits the ``jax_<case>.npz`` files it writes.  ``tests/test_jax_fixture_hook.py`` picks up every such file and holds
the oracle (``-m "not gpu"``) and the CUDA path (``-m gpu``) against it.  Format: ``oracle/jax_fixture.py``.

What it does, per case:
  1. builds the networks exactly like ``make_init`` (:178-301) and takes the REAL ``learn_step`` closure out of
     ``make_train_fn`` (:304-673; it is a free variable of ``train_fn``), with a shape-only stand-in environment;
  2. feeds it a seeded synthetic ``Transition`` batch (numpy generator, so the inputs do not depend on this repository);
  3. re-derives the permutations and the standard-normal draws ``learn_step`` made: same ``jax.random.split`` sequence
     (:645-647, :664-669), same distrax sampling call on a unit Normal (``sample_and_log_prob(seed=key)`` and
     ``sample_and_log_prob(sample_shape=(16,), seed=key)``, :535 / :557, with the epoch key AFTER the shuffle split --
     ``key`` is late-bound inside ``minibatch_update``);
  4. writes inputs, recorded randomness and outputs: post-update parameters and the metrics of the last epoch (the
     TD(lambda) targets are internal to ``learn_step``; they are covered through ``metrics/target_values``, the value
     loss and every parameter, all of which are functions of them).
Nothing of the reference is copied: it is imported and executed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

CASES = {
    "small": dict(num_steps=8, num_envs=64, num_mini_batches=4, num_epochs=2, critic_hidden_dim=128, actor_hidden_dim=128,
                  num_bins=51, vmin=0.0, vmax=20.0, obs_dim=5, action_dim=2, actor_min_std=0.05),
    "medium": dict(num_steps=8, num_envs=128, num_mini_batches=4, num_epochs=2, critic_hidden_dim=256, actor_hidden_dim=256,
                   num_bins=151, vmin=0.0, vmax=150.0, obs_dim=17, action_dim=6, actor_min_std=0.05),
    "full_mode": dict(num_steps=4, num_envs=64, num_mini_batches=2, num_epochs=1, critic_hidden_dim=128, actor_hidden_dim=128,
                      num_bins=51, vmin=-10.0, vmax=10.0, obs_dim=7, action_dim=3, actor_min_std=0.05,
                      actor_kl_clip_mode="full", kl_bound=0.05),
}
# config/reppo.yaml defaults for everything a case does not set
DEFAULTS = dict(lr=3e-4, gamma=0.99, total_time_steps=50_000_000, lmbda=0.95, lmbda_min=0.5, max_grad_norm=0.5, normalize_env=False,
                polyak=1.0, exploration_noise_min=1.0, exploration_noise_max=1.0, exploration_base_envs=0, ent_start=0.01,
                ent_target_mult=0.5, kl_start=0.01, hl_gauss=True, kl_bound=0.1, aux_loss_mult=1.0, use_critic_norm=True,
                num_critic_encoder_layers=2, num_critic_head_layers=2, num_critic_pred_layers=2, use_actor_norm=True,
                num_actor_layers=3, reduce_kl=True, reverse_kl=False, anneal_lr=False, actor_kl_clip_mode="clipped")
ORACLE_HYPER = ("num_steps", "num_envs", "num_mini_batches", "num_epochs", "gamma", "lmbda", "lr", "max_grad_norm", "critic_hidden_dim",
                "actor_hidden_dim", "vmin", "vmax", "num_bins", "use_critic_norm", "use_actor_norm", "num_critic_encoder_layers",
                "num_critic_head_layers", "num_critic_pred_layers", "num_actor_layers", "actor_min_std", "kl_start", "kl_bound",
                "reduce_kl", "update_kl_lagrangian", "actor_kl_clip_mode", "ent_start", "ent_target_mult",
                "update_entropy_lagrangian", "aux_loss_mult")


def synthetic_batch(c, obs_dim, act_dim, seed=0):
    """A rollout-shaped batch with terminations, truncations and non-trivial importance weights."""
    g = np.random.default_rng(seed)
    T, N, H = c["num_steps"], c["num_envs"], c["critic_hidden_dim"]
    R = c["vmax"] - c["vmin"]
    f = lambda *s: g.standard_normal(s).astype(np.float32)
    obs = f(T, N, obs_dim)
    trunc = (g.random((T, N)) < 0.01).astype(np.float32)
    done = np.clip((g.random((T, N)) < 0.02).astype(np.float32) + trunc, 0, 1)
    reward = g.random((T, N)).astype(np.float32)
    return dict(obs=obs, critic_obs=obs.copy(), action=np.clip(np.tanh(f(T, N, act_dim)), -0.999, 0.999),
                reward=reward, soft_reward=(reward - c.get("gamma", 0.99) * 0.01 * (f(T, N) - act_dim / 2)).astype(np.float32),
                next_emb=f(T, N, H), value=(c["vmin"] + R * (0.05 + 0.55 * g.random((T, N)))).astype(np.float32),
                done=done, truncated=trunc, importance_weight=np.log(0.5 + 0.5 * g.random((T, N))).astype(np.float32))


class _Space:
    def __init__(self, n):
        self.shape = (n,)


class _ShapeEnv:
    """What ``make_init`` / ``make_train_fn`` ask of the environment before any step: its spaces."""

    def __init__(self, obs_dim, act_dim):
        self._o, self._a = obs_dim, act_dim

    def observation_space(self, params=None):
        return (_Space(self._o), _Space(self._o))

    def action_space(self, params=None):
        return _Space(self._a)


def run_case(name, case, out_dir):
    import distrax
    import jax
    import jax.numpy as jnp
    import optax
    from flax import nnx
    from src.jaxrl import reppo as R
    from src.networks.jax_models import CategoricalCriticNetwork, SACActorNetworks

    c = {**DEFAULTS, **case}
    obs_dim, act_dim = c.pop("obs_dim"), c.pop("action_dim")
    cfg = R.ReppoConfig(**c)
    env = _ShapeEnv(obs_dim, act_dim)
    train_fn = R.make_train_fn(cfg, env)
    learn_step = dict(zip(train_fn.__code__.co_freevars, (cell.cell_contents for cell in train_fn.__closure__)))["learn_step"]

    # networks + optimizers as in make_init (:183-268)
    model_key = jax.random.PRNGKey(0)
    mk_actor = lambda: SACActorNetworks(obs_dim=obs_dim, action_dim=act_dim, hidden_dim=cfg.actor_hidden_dim, ent_start=cfg.ent_start,
                                        kl_start=cfg.kl_start, use_norm=cfg.use_actor_norm, layers=cfg.num_actor_layers,
                                        use_skip=cfg.use_actor_skip, rngs=nnx.Rngs(model_key))
    actor, actor_target = mk_actor(), mk_actor()
    critic = CategoricalCriticNetwork(obs_dim=obs_dim, action_dim=act_dim, hidden_dim=cfg.critic_hidden_dim, num_bins=cfg.num_bins,
                                      vmin=cfg.vmin, vmax=cfg.vmax, use_norm=cfg.use_critic_norm,
                                      encoder_layers=cfg.num_critic_encoder_layers, use_simplical_embedding=cfg.use_simplical_embedding,
                                      head_layers=cfg.num_critic_head_layers, pred_layers=cfg.num_critic_pred_layers,
                                      use_skip=cfg.use_critic_skip, rngs=nnx.Rngs(model_key))
    tx = lambda: optax.chain(optax.clip_by_global_norm(cfg.max_grad_norm), optax.adam(cfg.lr))
    mk = lambda m, t: nnx.TrainState.create(graphdef=nnx.graphdef(m), params=nnx.state(m), tx=t)
    state = R.SACTrainState(actor=mk(actor, tx()), actor_target=mk(actor_target, optax.set_to_zero()), critic=mk(critic, tx()),
                            iteration=0, time_steps=0, last_env_state=None, last_obs=None, last_critic_obs=None)

    b = synthetic_batch({**c, "gamma": cfg.gamma}, obs_dim, act_dim)
    batch = R.Transition(info={}, **{k: jnp.asarray(v) for k, v in b.items()})
    key = jax.random.PRNGKey(1234)
    new_state, metrics = jax.jit(learn_step)(key, state, batch)

    # the randomness learn_step drew from `key` (:645-647, :664-669; late-bound noise key, :535 / :557)
    T, N, E, n_mb = cfg.num_steps, cfg.num_envs, cfg.num_epochs, cfg.num_mini_batches
    S, mb = T * N, T * N // n_mb
    _, train_key = jax.random.split(key)
    perms, eps_pi, eps_kl = [], [], []
    unit = distrax.Normal(loc=jnp.zeros((mb, act_dim)), scale=jnp.ones((mb, act_dim)))
    for ek in jax.random.split(train_key, E):
        k, shuffle_key = jax.random.split(ek)
        perms.append(np.asarray(jax.random.permutation(shuffle_key, S)))
        eps_pi.append(np.asarray(unit.sample_and_log_prob(seed=k)[0]))
        eps_kl.append(np.asarray(unit.sample_and_log_prob(sample_shape=(16,), seed=k)[0]))

    arrays = {"hyper": np.array(json.dumps({**{k: getattr(cfg, k) for k in ORACLE_HYPER}, "obs_dim": obs_dim, "critic_obs_dim": obs_dim,
                                            "action_dim": act_dim, "vmin": float(cfg.vmin), "vmax": float(cfg.vmax)}))}
    arrays.update({f"batch/{k}": v for k, v in b.items()})
    arrays.update(perms=np.stack(perms).astype(np.int64), eps_pi=np.stack(eps_pi), eps_kl=np.stack(eps_kl))

    def put(group, params):
        flat = params.flat_state() if hasattr(params, "flat_state") else nnx.traversals.flatten_mapping(nnx.to_pure_dict(params))
        items = flat.items() if hasattr(flat, "items") else flat
        for path, leaf in items:
            val = getattr(leaf, "value", leaf)
            arrays[f"{group}/" + "/".join(str(p) for p in path)] = np.asarray(val)
    put("in/critic", state.critic.params), put("in/actor", state.actor.params)
    put("out/critic", new_state.critic.params), put("out/actor", new_state.actor.params)
    for k, v in metrics.items():
        arrays[f"out/metrics/{k}"] = np.asarray(jnp.mean(v), dtype=np.float64)
    path = os.path.join(out_dir, f"jax_{name}.npz")
    np.savez(path, **arrays)
    print("wrote", path)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--reference", default=os.getcwd(), help="reference checkout (must contain src/jaxrl/reppo.py)")
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    a = ap.parse_args()
    sys.path.insert(0, a.reference)
    os.makedirs(a.out, exist_ok=True)
    for n in a.cases:
        run_case(n, dict(CASES[n]), a.out)
