"""Slab programs (reppo_b200/csrc/slab.cu: a whole critic / actor chain of a step in one persistent cluster kernel)
(opt-in: REPPO_SLAB=1 at context creation) against (a) the default engine (separate tcgen05 GEMM + row kernels, identical bf16 operand rounding)
and (b) the fp32 CPU oracle's autograd gradients.

Tolerances.  (a) both paths round the same operands to bf16 and accumulate in fp32, so they differ only by fp32
summation order and by the rare bf16 rounding flip that follows from it: max-abs error <= 1e-2 of the tensor's
max-abs gradient, cosine >= 0.9999, losses 1e-3 relative.  (b) the bf16-mode tolerances of test_gpu_parity.py.
"""
import math
import os

import pytest
import torch

from oracle import reppo_oracle as O
from tests.test_gpu_parity import _cos, _engine, _grad_close, _hyper, _setup

pytestmark = pytest.mark.gpu

# 3 slabs (the last one ragged), 8-CTA clusters, CheetahRun dims
MID = dict(num_steps=6, num_envs=150, num_mini_batches=3, num_epochs=1, critic_hidden_dim=512, actor_hidden_dim=512,
           num_bins=151, vmin=0.0, vmax=150.0, obs_dim=17, critic_obs_dim=17, action_dim=6)
# critic wider than the policy (idle CTAs in the policy's normalised layers), LayerNorm, asymmetric observations
MIXED = dict(num_steps=5, num_envs=64, num_mini_batches=2, num_epochs=1, critic_hidden_dim=256, actor_hidden_dim=128,
             num_bins=51, vmin=-10.0, vmax=10.0, obs_dim=11, critic_obs_dim=23, action_dim=3, norm_type="layer")
# humanoid-like wide inputs (K0 = 261 -> 5 k-blocks), 3-layer heads
WIDE_IN = dict(num_steps=4, num_envs=96, num_mini_batches=2, num_epochs=1, critic_hidden_dim=128, actor_hidden_dim=192,
               num_bins=101, vmin=0.0, vmax=200.0, obs_dim=244, critic_obs_dim=244, action_dim=17,
               num_critic_head_layers=3, num_actor_layers=4)
CASES = [(O.JAX, MID), (O.TORCH, MID), (O.JAX, MIXED), (O.TORCH, dict(MIXED, norm_type="rms", actor_min_std=0.05)),
         (O.JAX, WIDE_IN), (O.TORCH, dict(MID, use_critic_norm=False, use_actor_norm=False, partial_reset=True))]


def _make(hp, sem, slab):
    if slab:
        os.environ["REPPO_SLAB"] = "1"
    else:
        os.environ.pop("REPPO_SLAB", None)
    try:
        return _engine(hp, sem, gemm_mode="bf16")
    finally:
        os.environ.pop("REPPO_SLAB", None)


@pytest.mark.parametrize("sem,kw", CASES)
def test_slab_matches_kernel_chain_and_oracle(sem, kw):
    hp = O.Hyper(**kw)
    cp, ap, ap_old, flat, eps_pi, eps_kl, perms = _setup(hp, sem)
    hyp = _hyper(hp)
    idx = perms[0, :hp.minibatch_size]
    out = {}
    for slab in (True, False):
        eng = _make(hp, sem, slab)
        eng.load_params(cp, ap)
        for k in eng.actor.layout:
            eng.actor.view(eng.actor_old, k).copy_(ap_old[k])
        batch = eng.make_batch(flat)
        mc = eng.metrics_dict(eng.critic_grad(batch, idx, hyp))
        gc = {k: eng.critic.view(eng.critic.grads, k).detach().cpu().clone() for k in eng.critic.layout}
        ma = eng.metrics_dict(eng.actor_grad(batch, idx, eps_pi[0], eps_kl[0], hyp))
        ga = {k: eng.actor.view(eng.actor.grads, k).detach().cpu().clone() for k in eng.actor.layout}
        torch.cuda.synchronize()
        out[slab] = (mc, gc, ma, ga)
    (mc1, gc1, ma1, ga1), (mc0, gc0, ma0, ga0) = out[True], out[False]
    for k in ("critic_loss", "ce_mean", "aux_mean", "q_mean", "target_mean", "target_max", "target_min"):
        assert math.isclose(mc1[k], mc0[k], rel_tol=1e-3, abs_tol=1e-5), (k, mc1[k], mc0[k])
    for k in ("actor_loss", "policy_loss", "kl", "entropy", "temperature", "lagrangian", "entropy_loss", "lagrangian_loss",
              "actor_q_mean", "abs_pred_action"):
        assert math.isclose(ma1[k], ma0[k], rel_tol=2e-3, abs_tol=2e-5), (k, ma1[k], ma0[k])
    for name, g1, g0 in (("critic", gc1, gc0), ("actor", ga1, ga0)):
        for k in g1:
            _grad_close(f"{name}.{k} (slab vs chain)", g1[k], g0[k], rel=1e-2)
            if g0[k].numel() > 8 and g0[k].abs().max() > 0:
                assert _cos(g1[k], g0[k]) >= 0.9999, (name, k)
    # against the oracle's autograd (bf16-mode tolerances)
    mb = O.take(flat, idx)
    loss, om, grads = O._value_and_grad(lambda p: O.critic_loss(p, mb, hp, sem), cp)
    assert math.isclose(mc1["critic_loss"], loss.item(), rel_tol=1e-2)
    for k, gref in grads.items():
        _grad_close(f"critic.{k}", gc1[k], gref, rel=4e-2)
        if gref.numel() > 8:
            assert _cos(gc1[k], gref) >= 0.999, k
    loss, om, grads = O._value_and_grad(lambda p: O.actor_loss(p, ap_old, cp, mb, eps_pi[0], eps_kl[0], hp, sem), ap)
    assert math.isclose(ma1["actor_loss"], loss.item(), rel_tol=2e-2, abs_tol=2e-3)
    for k, gref in grads.items():
        _grad_close(f"actor.{k}", ga1[k], gref, rel=8e-2)   # un-normalised nets amplify bf16 rounding
        if gref.numel() > 8:
            assert _cos(ga1[k], gref) >= 0.998, k


@pytest.mark.parametrize("sem,kw", [(O.JAX, MID), (O.JAX, MIXED), (O.TORCH, dict(MID, use_critic_norm=False, use_actor_norm=False))])
def test_fused_dgrad_norm_backward_matches_kernel_chain(sem, kw):
    """REPPO_FUSE_BWD=1: dgrad + norm / swish backward in one tcgen05 cluster kernel (gemm_bf16_tc_bwd.cu) against the
    default dgrad GEMM + row kernel pair."""
    hp = O.Hyper(**kw)
    cp, ap, ap_old, flat, eps_pi, eps_kl, perms = _setup(hp, sem)
    hyp = _hyper(hp)
    idx = perms[0, :hp.minibatch_size]
    out = {}
    for fused in (True, False):
        if fused:
            os.environ["REPPO_FUSE_BWD"] = "1"
        eng = _engine(hp, sem, gemm_mode="bf16")
        eng.load_params(cp, ap)
        for k in eng.actor.layout:
            eng.actor.view(eng.actor_old, k).copy_(ap_old[k])
        batch = eng.make_batch(flat)
        eng.critic_grad(batch, idx, hyp)
        gc = {k: eng.critic.view(eng.critic.grads, k).detach().cpu().clone() for k in eng.critic.layout}
        eng.actor_grad(batch, idx, eps_pi[0], eps_kl[0], hyp)
        ga = {k: eng.actor.view(eng.actor.grads, k).detach().cpu().clone() for k in eng.actor.layout}
        launches = eng.launch_count()
        torch.cuda.synchronize()
        out[fused] = (gc, ga, launches)
        os.environ.pop("REPPO_FUSE_BWD", None)
    assert out[True][2] < out[False][2], "the fused path must launch fewer kernels"
    for name, i in (("critic", 0), ("actor", 1)):
        for k in out[True][i]:
            _grad_close(f"{name}.{k} (fused vs chain)", out[True][i][k], out[False][i][k], rel=1e-2)
    mb = O.take(flat, idx)
    _, _, gref_c = O._value_and_grad(lambda p: O.critic_loss(p, mb, hp, sem), cp)
    _, _, gref_a = O._value_and_grad(lambda p: O.actor_loss(p, ap_old, cp, mb, eps_pi[0], eps_kl[0], hp, sem), ap)
    for k, g in gref_c.items():
        _grad_close(f"critic.{k}", out[True][0][k], g, rel=4e-2)
    for k, g in gref_a.items():
        _grad_close(f"actor.{k}", out[True][1][k], g, rel=8e-2)


def test_slab_full_update_pipeline():
    """Whole update (graph, two-chain pipeline) with slab programs against the kernel chain: same losses step by step."""
    hp = O.Hyper(**dict(MID, num_epochs=2))
    sem = O.JAX
    cp, ap = O.init_params(hp, sem, seed=0)
    batch = O.synthetic_rollout(hp, seed=0, terminate=True)
    perms = O.synthetic_perms(hp)
    eps_pi, eps_kl = O.synthetic_noise(hp)
    res = {}
    for slab in (True, False):
        eng = _make(hp, sem, slab)
        eng.load_params(cp, ap)
        dev = {k: v.cuda() for k, v in batch.items()}
        tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], dev["importance_weight"],
                            hp.gamma, hp.lmbda)
        flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
        flat["target"] = tgt.reshape(-1)
        for _ in range(2):   # second call replays the captured graph
            eng.load_params(cp, ap)
            eng.reset_optimizer()
            m, sm = eng.update(eng.make_batch(flat), perms.to(torch.int32).cuda(), eps_pi.cuda(), eps_kl.cuda(), _hyper(hp),
                               hp.num_mini_batches, use_graph=True)
        torch.cuda.synchronize()
        res[slab] = (sm.cpu().clone(), {k: eng.critic.view(eng.critic.params, k).cpu().clone() for k in eng.critic.layout},
                     {k: eng.actor.view(eng.actor.params, k).cpu().clone() for k in eng.actor.layout})
    from reppo_b200 import _lib as L
    col = {k: i for i, k in enumerate(L.METRIC_NAMES)}
    sm1, sm0 = res[True][0], res[False][0]
    steps = hp.num_epochs * hp.num_mini_batches
    for i in range(steps):
        for k in ("critic_loss", "actor_loss", "kl", "entropy"):
            a, b = sm1[i, col[k]].item(), sm0[i, col[k]].item()
            assert math.isclose(a, b, rel_tol=3e-2, abs_tol=3e-3), (i, k, a, b)
    for k in res[True][1]:
        err = (res[True][1][k] - res[False][1][k]).abs().max().item()
        assert err <= 2 * hp.lr * steps + 1e-6, (k, err)
