"""Hook for fixtures of the REAL JAX ``learn_step`` (src/jaxrl/reppo.py:417-673): every ``tests/golden/jax_*.npz`` written by
``oracle/gen_golden_jax.py`` (needs a JAX machine; none has produced one yet -> "parity unpinned") is held
against the oracle on the CPU and against the CUDA path on the GPU.  While no such file exists the same code runs on a file
the ORACLE wrote in the same format, which pins nothing but keeps loader, name mapping and comparison exercised."""
import glob
import os

import pytest
import torch

from oracle import jax_fixture as F
from oracle import reppo_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REAL = sorted(glob.glob(os.path.join(GOLDEN, "jax_*.npz")))
# JAX metric name (src/jaxrl/reppo.py:514-524, 609-622; actor's `loss` shadows the critic's in the merged dict) -> oracle metric
METRICS = {"value_loss": "critic/value_loss", "q": "critic/q", "loss": "actor/loss", "kl": "actor/kl", "temp": "actor/temperature",
           "lagrangian": "actor/lagrangian", "target_values": "critic/target_mean", "abs_pred_action": "actor/abs_pred_action"}


def _self_fixture(tmp_path):
    hp = O.Hyper(num_steps=4, num_envs=32, num_mini_batches=2, num_epochs=2, critic_hidden_dim=64, actor_hidden_dim=64,
                 num_bins=21, vmin=0.0, vmax=10.0, obs_dim=5, critic_obs_dim=5, action_dim=2, actor_min_std=0.05,
                 num_critic_encoder_layers=3)
    path = os.path.join(str(tmp_path), "jax_selftest.npz")
    F.dump_from_oracle(path, hp)
    return path


def _oracle_run(hp, g):
    ts = O.TrainState.create(g["critic"], g["actor"])
    return O.learn_step(ts, {k: v.clone() for k, v in g["batch"].items()}, g["perms"], g["eps_pi"], g["eps_kl"], hp, O.JAX)


def _check_oracle(path, real):
    hp, g = F.load(path)
    ts, metrics, target, _ = _oracle_run(hp, g)
    # fp32 on two different CPU stacks (XLA vs ATen): losses 1e-4 relative, parameters within a few Adam ulps of lr
    for jk, ok in METRICS.items():
        if real and jk in g["metrics"]:
            assert float(metrics[ok]) == pytest.approx(g["metrics"][jk], rel=2e-4, abs=1e-5), jk
    if not real:
        for k, v in g["metrics"].items():
            assert float(metrics[k]) == pytest.approx(v, rel=1e-6, abs=1e-7), k
        assert torch.equal(target, g["target_values"])
    tol = 0.0 if not real else 2.0 * hp.lr
    for name, got, want in (("critic", ts.critic, g["out_critic"]), ("actor", ts.actor, g["out_actor"])):
        for k, v in want.items():
            err = (got[k].reshape(v.shape) - v).abs()
            assert err.max().item() <= tol, f"{name}.{k}: max |delta| {err.max().item():.3e}"
            if real:
                assert (err <= 1e-5).float().mean().item() >= 0.999, f"{name}.{k}: more than 0.1 % of entries off by > 1e-5"


def test_fixture_format_round_trip(tmp_path):
    path = _self_fixture(tmp_path)
    hp, g = F.load(path)
    assert set(g["critic"]) == set(O.critic_param_shapes(hp, O.JAX)) and set(g["actor"]) == set(O.actor_param_shapes(hp))
    assert g["critic"]["feature_module.net.1.0.weight"].shape == (64, 64)          # kernel [in, out] came back as [out, in]
    assert F.oracle_key("feature_module.norm.scale.value", hp) is None              # FCNN.norm is never applied (jax_models.py:100)
    with pytest.raises(KeyError, match="unrecognised nnx parameter path"):
        F.oracle_key("feature_module/unknown/kernel", hp)
    _check_oracle(path, real=False)


@pytest.mark.skipif(not REAL, reason="no tests/golden/jax_*.npz: the reference's JAX path has not been run (oracle/gen_golden_jax.py)")
@pytest.mark.parametrize("path", REAL, ids=[os.path.basename(p) for p in REAL])
def test_oracle_matches_real_jax_learn_step(path):
    _check_oracle(path, real=True)


@pytest.mark.gpu
def test_cuda_update_matches_fixture(tmp_path):
    """The CUDA path (fp32 GEMMs) on every real fixture, or on the oracle-written stand-in when there is none."""
    from reppo_b200.engine import EngineConfig, HyperParams, ReppoEngine
    for path in (REAL or [_self_fixture(tmp_path)]):
        hp, g = F.load(path)
        eng = ReppoEngine(EngineConfig(obs_dim=hp.obs_dim, action_dim=hp.action_dim, critic_hidden_dim=hp.critic_hidden_dim,
                                       actor_hidden_dim=hp.actor_hidden_dim, num_bins=hp.num_bins, vmin=hp.vmin, vmax=hp.vmax,
                                       num_critic_encoder_layers=hp.num_critic_encoder_layers, semantics="jax",
                                       max_rows=hp.minibatch_size))
        eng.load_params(g["critic"], g["actor"])
        dev = {k: v.cuda() for k, v in g["batch"].items()}
        tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], dev["importance_weight"], hp.gamma, hp.lmbda)
        flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
        flat["target"] = tgt.reshape(-1)
        hyp = HyperParams(**{k: getattr(hp, k) for k in ("gamma", "lmbda", "lr", "max_grad_norm", "kl_bound", "ent_target_mult",
                                                         "aux_loss_mult", "actor_min_std", "actor_kl_clip_mode", "reduce_kl",
                                                         "update_entropy_lagrangian", "update_kl_lagrangian")})
        eng.update(eng.make_batch(flat), g["perms"].to(torch.int32).cuda(), g["eps_pi"].cuda(), g["eps_kl"].cuda(), hyp,
                   hp.num_mini_batches, use_graph=False)
        torch.cuda.synchronize()
        for net, want in ((eng.critic, g["out_critic"]), (eng.actor, g["out_actor"])):
            for k, v in want.items():
                err = (net.view(net.params, k).cpu().reshape(v.shape) - v).abs()
                assert err.max().item() <= 2.0 * hp.lr * hp.num_epochs * hp.num_mini_batches, k
                assert (err <= 2e-5).float().mean().item() >= 0.995, k
