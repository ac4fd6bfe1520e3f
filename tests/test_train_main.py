"""The Torch trainer entry point (reppo_b200/train.py = ``main`` of src/torchrl/reppo.py:437-730) driven end to end against a
device-resident synthetic environment with the reference wrappers' interface (src/env_utils/torch_wrappers/*.py)."""
import math

import pytest
import torch

from reppo_b200.config import compose


def _small_cfg(extra=()):
    return compose(["env=synthetic", "hyperparameters.num_envs=64", "hyperparameters.num_steps=8",
                    "hyperparameters.num_mini_batches=4", "hyperparameters.num_epochs=2", "hyperparameters.critic_hidden_dim=128",
                    "hyperparameters.actor_hidden_dim=128", "hyperparameters.num_bins=51", "hyperparameters.total_time_steps=4096",
                    "hyperparameters.num_eval=2", "measure_burnin=0", "env.num_obs=5", "env.num_actions=2",
                    "env.max_episode_steps=16", *extra], merge_overrides=False)


# ---- CPU: the stand-in environment and the env factory ---------------------------------------------------------------
def test_synthetic_env_interface_and_determinism():
    from reppo_b200.train import SyntheticEnv
    a, b = (SyntheticEnv(8, 5, 2, "cpu", seed=3, max_episode_steps=4) for _ in range(2))
    oa, _ = a.reset()
    ob, _ = b.reset()
    assert torch.equal(oa, ob) and oa.shape == (8, 5)
    assert (a.num_envs, a.num_obs, a.num_actions, a.asymmetric_obs, a.max_episode_steps) == (8, 5, 2, False, 4)
    g = torch.Generator().manual_seed(0)
    saw_trunc = False
    for _ in range(9):
        act = torch.tanh(torch.randn(8, 2, generator=g))
        ra, rb = a.step(act), b.step(act)
        for x, y in zip(ra[:4], rb[:4]):
            assert torch.equal(x, y)
        nobs, rew, term, trunc, info = ra
        assert nobs.shape == (8, 5) and rew.shape == (8,) and term.dtype == torch.bool and trunc.dtype == torch.bool
        assert not (term & trunc).any() and "final_observation" in info
        saw_trunc |= bool(trunc.any())
    assert saw_trunc                                     # max_episode_steps = 4 truncates within 9 steps
    c = SyntheticEnv(4, 5, 2, "cpu", asymmetric_obs=True, n_critic_obs=7)
    obs, cobs = c.reset_with_critic_obs()
    assert cobs.shape == (4, 7) and c.num_privileged_obs == 7
    assert c.step(torch.zeros(4, 2))[4]["observations"]["critic"].shape == (4, 7)


def test_make_envs_dispatch():
    from reppo_b200.train import make_envs
    envs, eval_envs = make_envs(_small_cfg(), torch.device("cpu"), seed=1)
    assert envs.num_envs == 64 and eval_envs is not envs and envs.num_obs == 5 and envs.num_actions == 2
    with pytest.raises(ValueError, match="Unknown environment type"):
        make_envs(_small_cfg(["env.type=atari"]), torch.device("cpu"))
    with pytest.raises(ImportError, match="reference's environment wrappers"):
        make_envs(_small_cfg(["env.type=mjx"]), torch.device("cpu"))   # no simulator in this image


def test_main_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from reppo_b200.train import main
    with pytest.raises(RuntimeError, match="no CPU path"):
        main(_small_cfg())


# ---- GPU: the loop ----------------------------------------------------------------------------------------------------
_KEYS = ("critic/qf_loss", "critic/qf_max", "critic/qf_min", "critic/qf_mean", "critic/embedding_loss", "critic/critic_grad_norm",
         "actor/actor_loss", "actor/actor_grad_norm", "actor/kl", "actor/entropy", "actor/temperature", "actor/lagrangian",
         "actor/entropy_loss", "actor/lagrangian_loss", "train/rewards_batch", "speed", "frame")


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_main_runs_reference_loop(fused):
    from reppo_b200.train import main
    seen = []
    history, ts = main(_small_cfg(), fused=fused, log_fn=lambda logs, frame: seen.append(frame), return_state=True)
    # total_time_steps // (num_envs * num_steps) + 1 = 9 iterations (src/torchrl/reppo.py:597-601), one log per iteration
    assert len(history) == 9 and seen == [i * 64 * 8 for i in range(9)]
    for logs in history:
        for k in _KEYS:
            assert k in logs and math.isfinite(logs[k]), k
    evals = [h for h in history if "eval/avg_return" in h]
    assert len(evals) >= 2 and all(0 < h["eval/avg_length"] <= 16 for h in evals)
    # the update really moved the networks, and old_actor <- actor holds at the end of every iteration (:631-632)
    assert torch.equal(ts.engine.actor_old, ts.engine.actor.params)
    assert history[0]["critic/qf_loss"] != history[-1]["critic/qf_loss"]
    assert ts.normalizer.count.item() > 1   # normalize_env=true: the running statistics saw the rollouts


@pytest.mark.gpu
def test_main_fused_and_literal_loops_agree_on_first_update():
    """Same seed, same environment: until the first optimizer step both loops see the same rollout, so the first logged
    qf statistics of the TD(lambda) targets (a function of the rollout + initial networks only) agree."""
    from reppo_b200.train import main
    a = main(_small_cfg(), fused=True, max_updates=1)
    b = main(_small_cfg(), fused=False, max_updates=1)
    assert a[0]["train/rewards_batch"] == pytest.approx(b[0]["train/rewards_batch"], rel=1e-5)


@pytest.mark.gpu
def test_main_learns_on_synthetic_env():
    """Not a parity test: a sanity check that the loop as a whole optimises (critic loss falls over updates)."""
    from reppo_b200.train import main
    h = main(_small_cfg(["hyperparameters.total_time_steps=20480"]))
    first = sum(x["critic/qf_loss"] for x in h[:3]) / 3
    last = sum(x["critic/qf_loss"] for x in h[-3:]) / 3
    assert last < first


@pytest.mark.gpu
def test_main_checkpoint_resume(tmp_path):
    """save_path / checkpoint_path: a run that stops after 3 updates and is resumed holds the saved networks and normaliser
    statistics and continues from the saved step (the loop the reference leaves as a TODO, src/torchrl/reppo.py:579-594)."""
    from reppo_b200.train import main
    path = str(tmp_path / "ckpt.pt")
    h1, ts1 = main(_small_cfg(), max_updates=3, save_path=path, return_state=True)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert ck["global_step"] == 3 and set(ck) >= {"actor_state_dict", "qnet_state_dict", "obs_normalizer_state", "args"}
    for k, v in ts1.actor.state_dict().items():
        assert torch.equal(ck["actor_state_dict"][k], v.cpu()), k
    h2, ts2 = main(_small_cfg(), max_updates=5, checkpoint_path=path, return_state=True)
    assert len(h1) == 3 and len(h2) == 2 and h2[0]["frame"] == 3 * 64 * 8      # resumed at update 3
    assert ts2.normalizer.count.item() > ts1.normalizer.count.item()
