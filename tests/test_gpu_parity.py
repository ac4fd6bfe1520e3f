"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
reference-generated golden fixtures.  Run on the B200 box with ``-m gpu``.

Tolerances (fp32 GEMM mode).  The scan is bit-exact where the reference arithmetic is exactly
reproducible (Torch convention; JAX convention with zero log importance weights) and within
2 ulp-per-step drift otherwise.  Forward activations / losses: rtol 2e-5.  Gradients:
max-abs error <= 2e-5 of the tensor's max-abs gradient (+1e-8).  Post-step parameters: Adam makes
|delta| ~ lr regardless of gradient size, so a sign flip of a near-zero gradient costs 2*lr; we
require <= 1e-5 absolute on >= 99.9 % of the entries and <= 2*lr*steps everywhere.
"""
import math
from dataclasses import replace

import pytest
import torch

from oracle import reppo_oracle as O
from tests.golden_util import load_case

pytestmark = pytest.mark.gpu


def _engine(hp: O.Hyper, sem: O.Semantics, max_rows=None, gemm_mode="fp32", max_batch_rows=0):
    from reppo_b200.engine import EngineConfig, ReppoEngine
    cfg = EngineConfig(
        obs_dim=hp.obs_dim, action_dim=hp.action_dim, critic_obs_dim=hp.critic_obs_dim,
        critic_hidden_dim=hp.critic_hidden_dim, actor_hidden_dim=hp.actor_hidden_dim, num_bins=hp.num_bins,
        vmin=hp.vmin, vmax=hp.vmax, num_critic_encoder_layers=hp.num_critic_encoder_layers,
        num_critic_head_layers=hp.num_critic_head_layers, num_critic_pred_layers=hp.num_critic_pred_layers,
        num_actor_layers=hp.num_actor_layers, use_critic_norm=hp.use_critic_norm, use_actor_norm=hp.use_actor_norm,
        norm_type=hp.norm_type, semantics=sem.name, gemm_mode=gemm_mode,
        max_rows=max_rows or hp.minibatch_size, max_batch_rows=max_batch_rows)
    return ReppoEngine(cfg)


def _hyper(hp: O.Hyper):
    from reppo_b200.engine import HyperParams
    return HyperParams(gamma=hp.gamma, lmbda=hp.lmbda, lr=hp.lr, max_grad_norm=hp.max_grad_norm, kl_bound=hp.kl_bound,
                       ent_target_mult=hp.ent_target_mult, aux_loss_mult=hp.aux_loss_mult, actor_min_std=hp.actor_min_std,
                       actor_kl_clip_mode=hp.actor_kl_clip_mode, reduce_kl=hp.reduce_kl,
                       update_entropy_lagrangian=hp.update_entropy_lagrangian,
                       update_kl_lagrangian=hp.update_kl_lagrangian, partial_reset=hp.partial_reset,
                       reverse_kl=getattr(hp, "reverse_kl", False))


def _grad_close(name, got, want, rel=2e-5):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= rel * scale + 1e-8, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


SMALL = dict(num_steps=8, num_envs=64, num_mini_batches=4, num_epochs=2, critic_hidden_dim=128, actor_hidden_dim=128,
             num_bins=51, vmin=0.0, vmax=20.0, obs_dim=5, critic_obs_dim=5, action_dim=2)
RAGGED = dict(num_steps=6, num_envs=50, num_mini_batches=3, num_epochs=1, critic_hidden_dim=96, actor_hidden_dim=160,
              num_bins=151, vmin=-10.0, vmax=10.0, obs_dim=17, critic_obs_dim=23, action_dim=6)


# ---------------------------------------------------------------------------------------------
# K1: TD(lambda)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,N", [(1, 1), (2, 3), (32, 1024), (128, 1027), (5, 4096 * 128)])
@pytest.mark.parametrize("conv", ["jax", "torch"])
def test_td_lambda_bit_exact(T, N, conv):
    hp = O.Hyper(num_steps=T, num_envs=N, critic_hidden_dim=1, obs_dim=1, critic_obs_dim=1, action_dim=2)
    b = O.synthetic_rollout(hp, seed=3, terminate=True)
    eng = _engine(O.Hyper(**SMALL), O.JAX)
    dev = {k: b[k].cuda() for k in ("soft_reward", "value", "done", "truncated")}
    out = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], None, hp.gamma, hp.lmbda,
                        convention=conv)
    if conv == "jax":
        want = O.td_lambda_jax(b["soft_reward"], b["value"], b["done"], b["truncated"], None, hp.gamma, hp.lmbda)
    else:
        tr = b["truncated"].clone()
        want = O.td_lambda_torch(b["soft_reward"], b["done"], tr, b["value"], hp.gamma, hp.lmbda)
        assert torch.equal(dev["truncated"].cpu(), tr)  # row T-1 forced to 1 in place (torchrl/reppo.py:178)
    assert torch.equal(out.cpu(), want)


def test_td_lambda_importance_weights():
    hp = O.Hyper(num_steps=64, num_envs=2048, critic_hidden_dim=1, obs_dim=1, critic_obs_dim=1, action_dim=2)
    b = O.synthetic_rollout(hp, seed=5, terminate=True, iw_variant=True)
    eng = _engine(O.Hyper(**SMALL), O.JAX)
    d = {k: v.cuda() for k, v in b.items() if k in ("soft_reward", "value", "done", "truncated", "importance_weight")}
    out = eng.td_lambda(d["soft_reward"], d["value"], d["done"], d["truncated"], d["importance_weight"], hp.gamma, hp.lmbda,
                        convention="jax").cpu()
    want64 = O.td_lambda_jax(*(b[k].double() for k in ("soft_reward", "value", "done", "truncated", "importance_weight")),
                             hp.gamma, hp.lmbda)
    want32 = O.td_lambda_jax(b["soft_reward"], b["value"], b["done"], b["truncated"], b["importance_weight"], hp.gamma, hp.lmbda)
    err_gpu = (out.double() - want64).abs().max().item()
    err_cpu = (want32.double() - want64).abs().max().item()
    assert err_gpu <= 4 * err_cpu + 1e-5, (err_gpu, err_cpu)
    torch.testing.assert_close(out, want32, rtol=2e-6, atol=2e-5)


@pytest.mark.parametrize("name", ["torch_small", "torch_medium", "torch_full_mode"])
def test_td_lambda_golden(name):
    hp, g = load_case(name)
    b = g["inputs"]["batch"]
    eng = _engine(hp, O.TORCH)
    tr = b["truncated"].clone().cuda()
    out = eng.td_lambda(b["soft_reward"].cuda(), b["value"].cuda(), b["done"].cuda(), tr, None, hp.gamma, hp.lmbda,
                        convention="torch")
    assert torch.equal(out.cpu(), g["gve"])
    assert tr[-1].min().item() == 1.0


def test_td_lambda_properties_full_size():
    """Size-independent properties at a bandwidth-test size (T=128, N=262144)."""
    T, N = 128, 262144
    g = torch.Generator(device="cuda").manual_seed(0)
    r = torch.rand(T, N, device="cuda", generator=g)
    v = torch.rand(T, N, device="cuda", generator=g) * 50
    z = torch.zeros(T, N, device="cuda")
    eng = _engine(O.Hyper(**SMALL), O.JAX)
    # (1) lambda = 0: one-step target r + gamma * v (exactly, for both conventions)
    for conv in ("jax", "torch"):
        out = eng.td_lambda(r, v, z, z.clone(), None, 0.99, 0.0, convention=conv)
        assert torch.equal(out, r + torch.tensor(0.99, dtype=torch.float32) * v)
    # (2) everything truncated: same one-step target regardless of lambda
    out = eng.td_lambda(r, v, z, torch.ones_like(z), None, 0.99, 0.95, convention="torch")
    assert torch.equal(out, r + torch.tensor(0.99, dtype=torch.float32) * v)
    # (3) done everywhere: target = reward
    out = eng.td_lambda(r, v, torch.ones_like(z), z.clone(), None, 0.99, 0.95, convention="torch")
    assert torch.equal(out[:-1], r[:-1])
    # (4) linearity in (reward, value) up to rounding
    a = eng.td_lambda(r, v, z, z.clone(), None, 0.99, 0.95, convention="jax")
    b2 = eng.td_lambda(2 * r, 2 * v, z, z.clone(), None, 0.99, 0.95, convention="jax")
    assert torch.equal(b2, 2 * a)  # scaling by 2 is exact in binary floating point
    # (5) columns are independent: a column-permuted input gives the column-permuted output
    perm = torch.randperm(N, device="cuda", generator=g)
    c = eng.td_lambda(r[:, perm].contiguous(), v[:, perm].contiguous(), z, z.clone(), None, 0.99, 0.95, convention="jax")
    assert torch.equal(c, a[:, perm])


# ---------------------------------------------------------------------------------------------
# network forwards
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sem,kw", [(O.JAX, SMALL), (O.TORCH, SMALL), (O.JAX, RAGGED),
                                    (O.JAX, dict(RAGGED, norm_type="layer")),
                                    (O.TORCH, dict(SMALL, use_critic_norm=False, use_actor_norm=False)),
                                    (O.JAX, dict(SMALL, num_critic_encoder_layers=3, num_critic_head_layers=3,
                                                 num_critic_pred_layers=4, num_actor_layers=2))])
def test_forward_matches_oracle(sem, kw):
    hp = O.Hyper(**kw)
    cp, ap = O.init_params(hp, sem, seed=1)
    if hp.norm_type == "layer":  # non-trivial affine
        for d in (cp, ap):
            for k in d:
                if k.endswith(".1.weight") or k.endswith(".1.bias"):
                    d[k] = d[k] + 0.1 * torch.randn_like(d[k])
    eng = _engine(hp, sem, max_rows=300)
    eng.load_params(cp, ap)
    g = torch.Generator().manual_seed(0)
    rows = 257
    co = torch.randn(rows, hp.critic_obs_dim, generator=g)
    ob = torch.randn(rows, hp.obs_dim, generator=g)
    ac = torch.tanh(torch.randn(rows, hp.action_dim, generator=g))
    value, logits, pred, feat = eng.critic_forward(co, ac)
    v0, l0, p0, f0 = O.critic_forward(cp, co, ac, hp, sem)
    torch.testing.assert_close(feat.cpu(), f0, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(logits.cpu(), l0, rtol=2e-5, atol=5e-5)
    torch.testing.assert_close(pred.cpu(), p0, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(value.cpu(), v0, rtol=2e-5, atol=1e-4)
    mean, std = eng.actor_forward(ob, min_std=hp.actor_min_std)
    m0, s0 = O.actor_forward(ap, ob, hp, sem)
    torch.testing.assert_close(mean.cpu(), m0, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(std.cpu(), s0, rtol=2e-5, atol=2e-5)


def test_forward_golden_reference_modules():
    hp, g = load_case("torch_small")
    inp = g["inputs"]
    eng = _engine(hp, O.TORCH)
    eng.load_params(inp["critic"], inp["actor"])
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in inp["batch"].items()}
    idx = inp["perms"][0, :hp.minibatch_size]
    value, logits, pred, feat = eng.critic_forward(flat["critic_obs"][idx], flat["action"][idx])
    torch.testing.assert_close(feat.cpu(), g["fwd"]["features"], rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(logits.cpu(), g["fwd"]["logits"], rtol=2e-5, atol=5e-5)
    torch.testing.assert_close(pred.cpu(), g["fwd"]["pred"], rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(value.cpu(), g["fwd"]["value"], rtol=2e-5, atol=1e-4)
    mean, std = eng.actor_forward(flat["obs"][idx], min_std=hp.actor_min_std)
    torch.testing.assert_close(mean.cpu(), g["fwd"]["mean"], rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(std.cpu(), g["fwd"]["std"], rtol=2e-5, atol=2e-5)


# ---------------------------------------------------------------------------------------------
# gradients of one minibatch (hand-derived CUDA backward vs autograd of the oracle)
# ---------------------------------------------------------------------------------------------
def _setup(hp, sem, seed=0, perturb_old=True, terminate=True):
    cp, ap = O.init_params(hp, sem, seed=seed)
    batch = O.synthetic_rollout(hp, seed=seed, terminate=terminate)
    target = O.compute_targets({k: v.clone() for k, v in batch.items()}, hp, sem)
    if sem.td_convention == "torch":
        batch["truncated"][-1] = 1.0
    flat = O.flatten_batch(batch, target)
    eps_pi, eps_kl = O.synthetic_noise(hp)
    perms = O.synthetic_perms(hp)
    g = torch.Generator().manual_seed(11)
    ap_old = {k: (v + 0.05 * torch.randn(v.shape, generator=g) if perturb_old and k.startswith("model.") else v.clone())
              for k, v in ap.items()}
    return cp, ap, ap_old, flat, eps_pi, eps_kl, perms


def _f64(d):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


def _actor_grads_close(eng, hp, sem, ap, ap_old, cp, mb, e1, e2, grads, rel):
    """Actor gradients against the oracle.  The reference evaluates the KL samples as atanh(clip(tanh(x))) in fp32, which
    loses precision as |x| grows (its gradients sit up to ~3e-4 of their scale away from a float64 evaluation of the same
    formulas); the kernel uses the exact identity clamp(x, +-atanh(c)) (losses.cu).  So the kernel is held to `rel` against
    the FLOAT64 oracle, and to `rel` + the fp32 oracle's own distance from float64 against the fp32 oracle."""
    # the reference clips fp32 tensors with a Python scalar: the clip constant it actually uses is the fp32-rounded one
    sem64 = replace(sem, action_clip=float(torch.tensor(sem.action_clip, dtype=torch.float32)))
    _, om64, grads64 = O._value_and_grad(
        lambda p: O.actor_loss(p, _f64(ap_old), _f64(cp), _f64(mb), e1.double(), e2.double(), hp, sem64), _f64(ap))
    for k, gref in grads.items():
        got = eng.actor.view(eng.actor.grads, k)
        scale = grads64[k].abs().max().item()
        noise = (gref.double() - grads64[k]).abs().max().item() / (scale + 1e-30)
        _grad_close(f"actor.{k} (vs float64 oracle)", got, grads64[k], rel=rel)
        _grad_close(f"actor.{k} (vs fp32 oracle)", got, gref, rel=rel + 1.5 * noise)
    return om64


GRAD_CASES = [
    (O.JAX, SMALL), (O.TORCH, SMALL), (O.JAX, RAGGED), (O.TORCH, dict(RAGGED, actor_min_std=0.05)),
    (O.JAX, dict(RAGGED, norm_type="layer")),
    (O.JAX, dict(SMALL, actor_kl_clip_mode="full", kl_bound=0.02)),
    (O.JAX, dict(SMALL, actor_kl_clip_mode="value")),
    (O.JAX, dict(SMALL, kl_bound=0.004, reduce_kl=True)),
    (O.JAX, dict(SMALL, update_entropy_lagrangian=False, update_kl_lagrangian=False, aux_loss_mult=0.3)),
    (O.TORCH, dict(SMALL, partial_reset=True, use_critic_norm=False, use_actor_norm=False)),
    # cfg.reverse_kl (src/jaxrl/reppo.py:543-553): KL(pi || pi_old) from samples of the current policy, every clip mode
    (O.JAX, dict(SMALL, reverse_kl=True)), (O.JAX, dict(RAGGED, reverse_kl=True, actor_kl_clip_mode="full", kl_bound=0.02)),
    (O.JAX, dict(SMALL, reverse_kl=True, kl_bound=0.004)),
]


@pytest.mark.parametrize("sem,kw", GRAD_CASES)
def test_critic_and_actor_gradients(sem, kw):
    hp = O.Hyper(**kw)
    cp, ap, ap_old, flat, eps_pi, eps_kl, perms = _setup(hp, sem)
    eng = _engine(hp, sem)
    eng.load_params(cp, ap)
    for k in eng.actor.layout:
        eng.actor.view(eng.actor_old, k).copy_(ap_old[k])
    hyp = _hyper(hp)
    batch = eng.make_batch(flat)
    idx = perms[0, :hp.minibatch_size]
    mb = O.take(flat, idx)

    m = eng.metrics_dict(eng.critic_grad(batch, idx, hyp))
    loss, om, grads = O._value_and_grad(lambda p: O.critic_loss(p, mb, hp, sem), cp)
    assert math.isclose(m["critic_loss"], loss.item(), rel_tol=2e-5, abs_tol=1e-6)
    assert math.isclose(m["ce_mean"], om["ce_mean"].item(), rel_tol=2e-5, abs_tol=1e-6)
    assert math.isclose(m["aux_mean"], om["aux_mean"].item(), rel_tol=2e-5, abs_tol=1e-6)
    assert math.isclose(m["value_loss"], om["value_loss"].item(), rel_tol=1e-4, abs_tol=1e-5)
    assert math.isclose(m["q_mean"], om["q"].item(), rel_tol=1e-4, abs_tol=1e-5)
    assert math.isclose(m["target_mean"], om["target_mean"].item(), rel_tol=1e-5, abs_tol=1e-6)
    assert m["target_max"] == om["target_max"].item() and m["target_min"] == om["target_min"].item()
    for k, gref in grads.items():
        _grad_close(f"critic.{k}", eng.critic.view(eng.critic.grads, k), gref)

    m = eng.metrics_dict(eng.actor_grad(batch, idx, eps_pi[0], eps_kl[0], hyp))
    loss, om, grads = O._value_and_grad(lambda p: O.actor_loss(p, ap_old, cp, mb, eps_pi[0], eps_kl[0], hp, sem), ap)
    assert math.isclose(m["actor_loss"], loss.item(), rel_tol=5e-5, abs_tol=2e-6)
    assert math.isclose(m["policy_loss"], om["policy_loss"].item(), rel_tol=5e-5, abs_tol=2e-6)
    assert math.isclose(m["kl"], om["kl"].item(), rel_tol=2e-4, abs_tol=2e-6)
    assert math.isclose(m["entropy"], om["entropy"].item(), rel_tol=2e-5, abs_tol=2e-6)
    assert math.isclose(m["temperature"], om["temperature"].item(), rel_tol=1e-6)
    assert math.isclose(m["lagrangian"], om["lagrangian"].item(), rel_tol=1e-6)
    assert math.isclose(m["entropy_loss"], om["entropy_loss"].item(), rel_tol=2e-5, abs_tol=1e-6)
    assert math.isclose(m["lagrangian_loss"], om["lagrangian_loss"].item(), rel_tol=2e-4, abs_tol=1e-6)
    assert math.isclose(m["actor_q_mean"], om["q"].item(), rel_tol=1e-4, abs_tol=1e-5)
    assert math.isclose(m["abs_pred_action"], om["abs_pred_action"].item(), rel_tol=2e-5)
    om64 = _actor_grads_close(eng, hp, sem, ap, ap_old, cp, mb, eps_pi[0], eps_kl[0], grads, rel=5e-5)
    assert math.isclose(m["kl"], om64["kl"].item(), rel_tol=5e-5, abs_tol=2e-6)


# ---------------------------------------------------------------------------------------------
# optimizer + full steps
# ---------------------------------------------------------------------------------------------
def _param_close(name, got, want, lr, steps):
    got, want = got.detach().cpu().double().reshape(-1), want.detach().cpu().double().reshape(-1)
    err = (got - want).abs()
    frac_bad = (err > 1e-5).double().mean().item()
    assert frac_bad <= 1e-3, f"{name}: {frac_bad:.2e} of entries differ by > 1e-5"
    assert err.max().item() <= 2 * lr * steps + 1e-6, f"{name}: max err {err.max().item():.3e}"


@pytest.mark.parametrize("sem", [O.JAX, O.TORCH])
def test_clip_adam_matches_oracle(sem):
    hp = O.Hyper(**SMALL)
    cp, ap = O.init_params(hp, sem, seed=2)
    eng = _engine(hp, sem)
    eng.load_params(cp, ap)
    hyp = _hyper(hp)
    st = O.AdamState.zeros_like(cp)
    g = torch.Generator().manual_seed(0)
    params = cp
    import ctypes as C
    from reppo_b200 import _lib as L
    for it, scale in enumerate([5.0, 1e-3, 0.2]):  # clipped, unclipped, mixed
        grads = {k: scale * torch.randn(v.shape, generator=g) for k, v in cp.items()}
        for k, v in grads.items():
            eng.critic.view(eng.critic.grads, k).copy_(v)
        ns, h = eng.critic.c_state(), hyp.c_struct()
        gn = torch.zeros(1, device="cuda")
        L.check(eng.lib.reppo_clip_adam(eng.ctx, C.byref(ns), eng.critic.count, C.byref(h), gn.data_ptr(), eng._stream()),
                eng.ctx, "clip_adam")
        params, st, gn0 = O.clip_and_adam(params, grads, st, hp.lr, hp.max_grad_norm, sem)
        assert math.isclose(gn.item(), gn0.item(), rel_tol=1e-5)
        assert int(eng.critic.step.item()) == it + 1
        for k in cp:
            torch.testing.assert_close(eng.critic.view(eng.critic.params, k).cpu(), params[k], rtol=0, atol=2e-6)
            torch.testing.assert_close(eng.critic.view(eng.critic.adam_m, k).cpu(), st.m[k], rtol=1e-5, atol=1e-9)
            torch.testing.assert_close(eng.critic.view(eng.critic.adam_v, k).cpu(), st.v[k], rtol=1e-5, atol=1e-12)


def test_clip_adam_linear_lr_schedule():
    """cfg.anneal_lr (src/jaxrl/reppo.py:239-244): optax.linear_schedule(lr, 0, T) evaluated on the device at the optimizer
    step counter -- update i (0-based) uses lr * (1 - min(i, T) / T); past T the parameters stop moving."""
    import ctypes as C
    from dataclasses import replace as dc_replace
    from reppo_b200 import _lib as L
    hp = O.Hyper(**SMALL)
    sem = O.JAX
    cp, ap = O.init_params(hp, sem, seed=4)
    eng = _engine(hp, sem)
    eng.load_params(cp, ap)
    T_anneal = 3
    hyp = dc_replace(_hyper(hp), lr_anneal_steps=T_anneal)
    st = O.AdamState.zeros_like(cp)
    g = torch.Generator().manual_seed(1)
    params = cp
    for it in range(5):
        grads = {k: 0.3 * torch.randn(v.shape, generator=g) for k, v in cp.items()}
        for k, v in grads.items():
            eng.critic.view(eng.critic.grads, k).copy_(v)
        ns, h = eng.critic.c_state(), hyp.c_struct()
        L.check(eng.lib.reppo_clip_adam(eng.ctx, C.byref(ns), eng.critic.count, C.byref(h), None, eng._stream()), eng.ctx, "clip_adam")
        lr_i = hp.lr * (1.0 - min(it, T_anneal) / T_anneal)
        prev = params
        params, st, _ = O.clip_and_adam(params, grads, st, lr_i, hp.max_grad_norm, sem)
        for k in cp:
            torch.testing.assert_close(eng.critic.view(eng.critic.params, k).cpu(), params[k], rtol=0, atol=2e-6)
            if it >= T_anneal:
                assert torch.equal(params[k], prev[k])


@pytest.mark.parametrize("sem,kw,use_graph", [(O.JAX, SMALL, False), (O.JAX, SMALL, True), (O.TORCH, SMALL, False),
                                              (O.JAX, RAGGED, True), (O.JAX, dict(SMALL, reverse_kl=True), True),
                                              (O.JAX, dict(RAGGED, reverse_kl=True, actor_kl_clip_mode="full"), False)])
def test_full_update_matches_oracle(sem, kw, use_graph):
    hp = O.Hyper(**kw)
    cp, ap = O.init_params(hp, sem, seed=0)
    batch = O.synthetic_rollout(hp, seed=0, terminate=True)
    perms = O.synthetic_perms(hp)
    eps_pi, eps_kl = O.synthetic_noise(hp)
    ts = O.TrainState.create(cp, ap)
    ts2, ometrics, target, traces = O.learn_step(ts, {k: v.clone() for k, v in batch.items()}, perms, eps_pi, eps_kl, hp, sem,
                                                 trace=True)

    eng = _engine(hp, sem)
    eng.load_params(cp, ap)
    dev = {k: v.cuda() for k, v in batch.items()}
    tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], dev["importance_weight"],
                        hp.gamma, hp.lmbda)
    assert torch.equal(tgt.cpu(), target)
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
    flat["target"] = tgt.reshape(-1)
    b = eng.make_batch(flat)
    m, sm = eng.update(b, perms.to(torch.int32).cuda(), eps_pi.cuda(), eps_kl.cuda(), _hyper(hp), hp.num_mini_batches,
                       use_graph=use_graph)
    torch.cuda.synchronize()
    sm = sm.cpu()
    from reppo_b200 import _lib as L
    col = {k: i for i, k in enumerate(L.METRIC_NAMES)}
    steps = len(traces)
    assert sm.shape[0] == steps
    # step 0 is an exact-gradient comparison; later steps inherit Adam's sign sensitivity
    for i, t in enumerate(traces):
        tol = 2e-5 if i == 0 else 5e-3
        assert math.isclose(sm[i, col["critic_loss"]].item(), t["critic/loss"].item(), rel_tol=tol, abs_tol=1e-5), i
        assert math.isclose(sm[i, col["critic_grad_norm"]].item(), t["critic/grad_norm"].item(), rel_tol=10 * tol, abs_tol=1e-5), i
        assert math.isclose(sm[i, col["actor_loss"]].item(), t["actor/loss"].item(), rel_tol=10 * tol, abs_tol=1e-4), i
        assert math.isclose(sm[i, col["kl"]].item(), t["actor/kl"].item(), rel_tol=20 * tol, abs_tol=1e-5), i
        assert math.isclose(sm[i, col["entropy"]].item(), t["actor/entropy"].item(), rel_tol=10 * tol, abs_tol=1e-4), i
    for k in cp:
        _param_close(f"critic.{k}", eng.critic.view(eng.critic.params, k), ts2.critic[k], hp.lr, steps)
    for k in ap:
        _param_close(f"actor.{k}", eng.actor.view(eng.actor.params, k), ts2.actor[k], hp.lr, steps)
    # last-epoch mean metrics (src/jaxrl/reppo.py:660,671)
    md = eng.metrics_dict(m)
    assert math.isclose(md["critic_loss"], ometrics["critic/loss"].item(), rel_tol=5e-3, abs_tol=1e-5)
    assert math.isclose(md["kl"], ometrics["actor/kl"].item(), rel_tol=5e-2, abs_tol=1e-5)
    assert eng.launch_count() > 0


@pytest.mark.parametrize("name", ["torch_small", "torch_full_mode", "torch_medium"])
def test_update_against_reference_golden(name):
    """Per-step logs and final parameters of the reference's own Torch closures (tests/golden)."""
    hp, g = load_case(name)
    inp = g["inputs"]
    eng = _engine(hp, O.TORCH)
    eng.load_params(inp["critic"], inp["actor"])
    dev = {k: v.clone().cuda() for k, v in inp["batch"].items()}
    tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], None, hp.gamma, hp.lmbda,
                        convention="torch")
    assert torch.equal(tgt.cpu(), g["gve"])
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
    flat["target"] = tgt.reshape(-1)
    b = eng.make_batch(flat)
    m, sm = eng.update(b, inp["perms"].to(torch.int32).cuda(), inp["eps_pi"].cuda(), inp["eps_kl"].cuda(), _hyper(hp),
                       hp.num_mini_batches)
    torch.cuda.synchronize()
    sm = sm.cpu()
    from reppo_b200 import _lib as L
    col = {k: i for i, k in enumerate(L.METRIC_NAMES)}
    logs = g["logs"]
    for i, l in enumerate(logs):
        tol = 2e-5 if i == 0 else 5e-3
        assert math.isclose(sm[i, col["critic_loss"]].item(), l["qf_loss"].item(), rel_tol=tol, abs_tol=1e-5), i
        assert math.isclose(sm[i, col["aux_mean"]].item(), l["embedding_loss"].item(), rel_tol=tol, abs_tol=1e-5), i
        assert math.isclose(sm[i, col["critic_grad_norm"]].item(), l["critic_grad_norm"].item(), rel_tol=10 * tol, abs_tol=1e-5), i
        assert sm[i, col["target_max"]].item() == l["qf_max"].item() and sm[i, col["target_min"]].item() == l["qf_min"].item()
        assert math.isclose(sm[i, col["actor_loss"]].item(), l["actor_loss"].item(), rel_tol=10 * tol, abs_tol=1e-4), i
        assert math.isclose(sm[i, col["kl"]].item(), l["kl"].mean().item(), rel_tol=20 * tol, abs_tol=1e-5), i
        assert math.isclose(sm[i, col["entropy"]].item(), l["entropy"].mean().item(), rel_tol=10 * tol, abs_tol=1e-4), i
        assert math.isclose(sm[i, col["temperature"]].item(), l["temperature"].item(), rel_tol=1e-4), i
        assert math.isclose(sm[i, col["lagrangian"]].item(), l["lagrangian"].item(), rel_tol=1e-4), i
    if "final" in g:
        for k, v in g["final"]["critic"].items():
            _param_close(f"critic.{k}", eng.critic.view(eng.critic.params, k), v, hp.lr, len(logs))
        for k, v in g["final"]["actor"].items():
            _param_close(f"actor.{k}", eng.actor.view(eng.actor.params, k), v, hp.lr, len(logs))
    else:
        for net, e in (("critic", eng.critic), ("actor", eng.actor)):
            for k, dg in g["final_digest"][net].items():
                _param_close(f"{net}.{k}", e.view(e.params, k).reshape(-1)[:64], dg["head"], hp.lr, len(logs))


# ---------------------------------------------------------------------------------------------
# REPPO_GEMM_BF16: tcgen05 GEMMs on bf16 operands, fp32 accumulation, fp32 everything else.
# Tolerances (stated): forward activations 2e-2 of their max-abs; losses 1e-2 relative;
# gradients: max-abs error <= 4e-2 of the tensor's max-abs gradient and cosine similarity >= 0.999;
# after a full update: losses within 5 % of the fp32 oracle trace (Adam amplifies sign noise).
# ---------------------------------------------------------------------------------------------
def _cos(a, b):
    a, b = a.detach().cpu().double().reshape(-1), b.detach().cpu().double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("sem,kw", [(O.JAX, SMALL), (O.TORCH, RAGGED), (O.JAX, dict(RAGGED, norm_type="layer"))])
def test_bf16_mode_gradients(sem, kw):
    hp = O.Hyper(**kw)
    cp, ap, ap_old, flat, eps_pi, eps_kl, perms = _setup(hp, sem)
    eng = _engine(hp, sem, gemm_mode="bf16")
    eng.load_params(cp, ap)
    for k in eng.actor.layout:
        eng.actor.view(eng.actor_old, k).copy_(ap_old[k])
    hyp = _hyper(hp)
    batch = eng.make_batch(flat)
    idx = perms[0, :hp.minibatch_size]
    mb = O.take(flat, idx)
    m = eng.metrics_dict(eng.critic_grad(batch, idx, hyp))
    loss, om, grads = O._value_and_grad(lambda p: O.critic_loss(p, mb, hp, sem), cp)
    assert math.isclose(m["critic_loss"], loss.item(), rel_tol=1e-2)
    assert math.isclose(m["q_mean"], om["q"].item(), rel_tol=2e-2, abs_tol=1e-3)
    for k, gref in grads.items():
        got = eng.critic.view(eng.critic.grads, k)
        _grad_close(f"critic.{k}", got, gref, rel=4e-2)
        if gref.numel() > 8:
            assert _cos(got, gref) >= 0.999, k
    m = eng.metrics_dict(eng.actor_grad(batch, idx, eps_pi[0], eps_kl[0], hyp))
    loss, om, grads = O._value_and_grad(lambda p: O.actor_loss(p, ap_old, cp, mb, eps_pi[0], eps_kl[0], hp, sem), ap)
    assert math.isclose(m["actor_loss"], loss.item(), rel_tol=2e-2, abs_tol=2e-3)
    assert math.isclose(m["kl"], om["kl"].item(), rel_tol=5e-2, abs_tol=1e-4)
    assert math.isclose(m["entropy"], om["entropy"].item(), rel_tol=1e-2, abs_tol=1e-3)
    for k, gref in grads.items():
        got = eng.actor.view(eng.actor.grads, k)
        _grad_close(f"actor.{k}", got, gref, rel=6e-2)
        if gref.numel() > 8:
            assert _cos(got, gref) >= 0.998, k


@pytest.mark.parametrize("use_graph", [False, True])
def test_bf16_mode_full_update(use_graph):
    hp = O.Hyper(**SMALL)
    sem = O.JAX
    cp, ap = O.init_params(hp, sem, seed=0)
    batch = O.synthetic_rollout(hp, seed=0, terminate=True)
    perms = O.synthetic_perms(hp)
    eps_pi, eps_kl = O.synthetic_noise(hp)
    ts2, ometrics, target, traces = O.learn_step(O.TrainState.create(cp, ap), {k: v.clone() for k, v in batch.items()}, perms,
                                                 eps_pi, eps_kl, hp, sem, trace=True)
    eng = _engine(hp, sem, gemm_mode="bf16")
    eng.load_params(cp, ap)
    dev = {k: v.cuda() for k, v in batch.items()}
    tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], dev["importance_weight"], hp.gamma, hp.lmbda)
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
    flat["target"] = tgt.reshape(-1)
    m, sm = eng.update(eng.make_batch(flat), perms.to(torch.int32).cuda(), eps_pi.cuda(), eps_kl.cuda(), _hyper(hp),
                       hp.num_mini_batches, use_graph=use_graph)
    torch.cuda.synchronize()
    sm = sm.cpu()
    from reppo_b200 import _lib as L
    col = {k: i for i, k in enumerate(L.METRIC_NAMES)}
    for i, t in enumerate(traces):
        assert math.isclose(sm[i, col["critic_loss"]].item(), t["critic/loss"].item(), rel_tol=5e-2), i
        assert math.isclose(sm[i, col["entropy"]].item(), t["actor/entropy"].item(), rel_tol=5e-2, abs_tol=5e-3), i
    for k in cp:
        err = (eng.critic.view(eng.critic.params, k).cpu() - ts2.critic[k]).abs().max().item()
        assert err <= 2 * hp.lr * len(traces) + 1e-6, (k, err)


# ---------------------------------------------------------------------------------------------
# BASELINE.json's named configurations at their real layer shapes, reduced T x N:
# gradients of one minibatch in both GEMM modes against the oracle's autograd.
# ---------------------------------------------------------------------------------------------
_RED = dict(num_steps=4, num_envs=64, num_mini_batches=2, num_epochs=1)
NAMED = {
    "C1_cartpole": dict(_RED, obs_dim=5, critic_obs_dim=5, action_dim=1, critic_hidden_dim=512, actor_hidden_dim=512,
                        num_bins=151, vmin=0.0, vmax=150.0),
    "C2_cheetah": dict(_RED, obs_dim=17, critic_obs_dim=17, action_dim=6, critic_hidden_dim=512, actor_hidden_dim=512,
                       num_bins=151, vmin=0.0, vmax=150.0),
    "C3_ant": dict(_RED, obs_dim=27, critic_obs_dim=27, action_dim=8, critic_hidden_dim=512, actor_hidden_dim=512,
                   num_bins=151, vmin=0.0, vmax=150.0),
    "C4_humanoid": dict(_RED, obs_dim=244, critic_obs_dim=244, action_dim=17, critic_hidden_dim=512, actor_hidden_dim=512,
                        num_bins=151, vmin=0.0, vmax=200.0),
    "C5_wide": dict(_RED, obs_dim=17, critic_obs_dim=17, action_dim=6, critic_hidden_dim=1024, actor_hidden_dim=1024,
                    num_bins=151, vmin=0.0, vmax=150.0),
    "C5_very_wide": dict(_RED, obs_dim=17, critic_obs_dim=17, action_dim=6, critic_hidden_dim=2048, actor_hidden_dim=2048,
                         num_bins=151, vmin=0.0, vmax=150.0),
}


@pytest.mark.parametrize("name", sorted(NAMED))
@pytest.mark.parametrize("gemm_mode", ["fp32", "bf16", "tf32"])
def test_named_config_shapes(name, gemm_mode):
    hp = O.Hyper(**NAMED[name])
    sem = O.JAX
    terminate = name in ("C3_ant", "C4_humanoid")
    cp, ap, ap_old, flat, eps_pi, eps_kl, perms = _setup(hp, sem, terminate=terminate)
    eng = _engine(hp, sem, gemm_mode=gemm_mode)
    eng.load_params(cp, ap)
    for k in eng.actor.layout:
        eng.actor.view(eng.actor_old, k).copy_(ap_old[k])
    hyp = _hyper(hp)
    batch = eng.make_batch(flat)
    idx = perms[0, :hp.minibatch_size]
    mb = O.take(flat, idx)
    # tf32: 10-bit mantissa products on the fp32 operands -- an order of magnitude inside the bf16 tolerances
    rel_c, rel_a, cos_min, ltol = {"fp32": (2e-5, 5e-5, 0.999999, 5e-5), "bf16": (4e-2, 8e-2, 0.998, 2e-2),
                                   "tf32": (5e-3, 1e-2, 0.99995, 2e-3)}[gemm_mode]
    m = eng.metrics_dict(eng.critic_grad(batch, idx, hyp))
    loss, _, grads = O._value_and_grad(lambda p: O.critic_loss(p, mb, hp, sem), cp)
    assert math.isclose(m["critic_loss"], loss.item(), rel_tol=ltol, abs_tol=1e-6)
    for k, gref in grads.items():
        got = eng.critic.view(eng.critic.grads, k)
        _grad_close(f"critic.{k}", got, gref, rel=rel_c)
        if gref.numel() > 8:
            assert _cos(got, gref) >= cos_min, k
    m = eng.metrics_dict(eng.actor_grad(batch, idx, eps_pi[0], eps_kl[0], hyp))
    loss, _, grads = O._value_and_grad(lambda p: O.actor_loss(p, ap_old, cp, mb, eps_pi[0], eps_kl[0], hp, sem), ap)
    assert math.isclose(m["actor_loss"], loss.item(), rel_tol=max(ltol, 1e-4),
                        abs_tol={"bf16": 2e-3, "tf32": 2e-4, "fp32": 2e-6}[gemm_mode])
    if gemm_mode == "fp32":
        _actor_grads_close(eng, hp, sem, ap, ap_old, cp, mb, eps_pi[0], eps_kl[0], grads, rel=rel_a)
    for k, gref in grads.items():
        got = eng.actor.view(eng.actor.grads, k)
        if gemm_mode != "fp32":
            _grad_close(f"actor.{k}", got, gref, rel=rel_a)
        if gref.numel() > 8:
            assert _cos(got, gref) >= (cos_min if gemm_mode != "fp32" else 0.99999), k


# ---------------------------------------------------------------------------------------------
# Minibatches large enough for the persistent CTA-pair GEMM (gemm_bf16_tc2.cu: >= 56 pair tiles), the unfused
# norm path it selects and the wide-row norm backward (H = 1024): one minibatch's gradients against the oracle.
# ---------------------------------------------------------------------------------------------
LARGE_MB = {
    "mb16384_H256": dict(num_steps=8, num_envs=2048, num_mini_batches=1, num_epochs=1, obs_dim=17, critic_obs_dim=17,
                         action_dim=6, critic_hidden_dim=256, actor_hidden_dim=256, num_bins=151, vmin=0.0, vmax=150.0),
    "mb4096_H1024": dict(num_steps=4, num_envs=1024, num_mini_batches=1, num_epochs=1, obs_dim=17, critic_obs_dim=17,
                         action_dim=6, critic_hidden_dim=1024, actor_hidden_dim=1024, num_bins=151, vmin=0.0, vmax=150.0),
}


@pytest.mark.parametrize("name", sorted(LARGE_MB))
def test_large_minibatch_pair_gemm_path(name):
    hp = O.Hyper(**LARGE_MB[name])
    sem = O.JAX
    cp, ap, ap_old, flat, eps_pi, eps_kl, perms = _setup(hp, sem, terminate=False)
    eng = _engine(hp, sem, gemm_mode="bf16")
    eng.load_params(cp, ap)
    for k in eng.actor.layout:
        eng.actor.view(eng.actor_old, k).copy_(ap_old[k])
    hyp = _hyper(hp)
    batch = eng.make_batch(flat)
    idx = perms[0, :hp.minibatch_size]
    mb = O.take(flat, idx)
    m = eng.metrics_dict(eng.critic_grad(batch, idx, hyp))
    loss, _, grads = O._value_and_grad(lambda p: O.critic_loss(p, mb, hp, sem), cp)
    assert math.isclose(m["critic_loss"], loss.item(), rel_tol=2e-2, abs_tol=1e-6)
    for k, gref in grads.items():
        got = eng.critic.view(eng.critic.grads, k)
        _grad_close(f"critic.{k}", got, gref, rel=4e-2)
        if gref.numel() > 8:
            assert _cos(got, gref) >= 0.998, k
    m = eng.metrics_dict(eng.actor_grad(batch, idx, eps_pi[0], eps_kl[0], hyp))
    loss, _, grads = O._value_and_grad(lambda p: O.actor_loss(p, ap_old, cp, mb, eps_pi[0], eps_kl[0], hp, sem), ap)
    assert math.isclose(m["actor_loss"], loss.item(), rel_tol=2e-2, abs_tol=2e-3)
    for k, gref in grads.items():
        got = eng.actor.view(eng.actor.grads, k)
        _grad_close(f"actor.{k}", got, gref, rel=8e-2)
        if gref.numel() > 8:
            assert _cos(got, gref) >= 0.998, k


# ---------------------------------------------------------------------------------------------
# The benchmarked configurations at their REAL step shape (minibatch rows, hidden width, first-layer K), through the
# path bench.py times: reppo_update with use_graph (two-chain pipeline, critic-snapshot ping-pong, per-epoch
# pre-gather), 2 epochs x 4 minibatches (C5: 1 x 2), against the oracle's learn_step trace.
#   fp32 GEMM mode: step 0 losses rtol 2e-5, later steps 5e-3 (Adam sign sensitivity), parameters as _param_close.
#   bf16 GEMM mode: every step's critic loss / entropy within 5 %, kl within 25 % (+1e-4), step-0 critic loss within 1 %;
#   post-update parameters: every entry within 2 lr steps AND mean |delta| <= 0.35 lr steps (an entry whose gradient
#   sign flips under bf16 rounding moves by 2 lr per step; the mean bounds how many do).
# ---------------------------------------------------------------------------------------------
_C2 = dict(obs_dim=17, critic_obs_dim=17, action_dim=6, critic_hidden_dim=512, actor_hidden_dim=512, num_bins=151, vmin=0.0,
           vmax=150.0)
FULL_SIZE = {
    "C2_mb2048": (dict(_C2, num_steps=8, num_envs=1024, num_mini_batches=4, num_epochs=2), True),
    "C3_mb1024": (dict(_C2, obs_dim=27, critic_obs_dim=27, action_dim=8, num_steps=4, num_envs=1024, num_mini_batches=4,
                       num_epochs=2), True),
    "C4_mb2048_K261": (dict(_C2, obs_dim=244, critic_obs_dim=244, action_dim=17, vmax=200.0, num_steps=8, num_envs=1024,
                            num_mini_batches=4, num_epochs=2), True),
    "C5_mb16384_H1024": (dict(_C2, critic_hidden_dim=1024, actor_hidden_dim=1024, num_steps=16, num_envs=2048,
                              num_mini_batches=2, num_epochs=1), False),
}
FULL_CASES = [("C2_mb2048", "jax", "fp32"), ("C2_mb2048", "jax", "bf16"), ("C2_mb2048", "torch", "bf16"),
              ("C2_mb2048", "torch", "fp32"), ("C3_mb1024", "jax", "bf16"), ("C4_mb2048_K261", "jax", "bf16"),
              ("C4_mb2048_K261", "jax", "fp32"), ("C5_mb16384_H1024", "jax", "bf16"), ("C2_mb2048", "jax", "tf32"),
              ("C2_mb2048", "torch", "tf32")]


@pytest.mark.parametrize("name,sem_name,gemm_mode", FULL_CASES)
def test_full_size_update_matches_oracle(name, sem_name, gemm_mode):
    kw, terminate = FULL_SIZE[name]
    hp = O.Hyper(**kw)
    sem = O.JAX if sem_name == "jax" else O.TORCH
    cp, ap = O.init_params(hp, sem, seed=3)
    batch = O.synthetic_rollout(hp, seed=3, terminate=terminate)
    perms = O.synthetic_perms(hp)
    eps_pi, eps_kl = O.synthetic_noise(hp)
    ts2, ometrics, target, traces = O.learn_step(O.TrainState.create(cp, ap), {k: v.clone() for k, v in batch.items()}, perms,
                                                 eps_pi, eps_kl, hp, sem, trace=True)
    eng = _engine(hp, sem, gemm_mode=gemm_mode, max_batch_rows=hp.num_steps * hp.num_envs)
    eng.load_params(cp, ap)
    dev = {k: v.cuda() for k, v in batch.items()}
    tgt = eng.td_lambda(dev["soft_reward"], dev["value"], dev["done"], dev["truncated"], dev["importance_weight"], hp.gamma, hp.lmbda)
    if sem_name == "torch":
        assert torch.equal(tgt.cpu(), target)
    else:
        torch.testing.assert_close(tgt.cpu(), target, rtol=1e-5, atol=1e-5)
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in dev.items()}
    flat["target"] = tgt.reshape(-1)
    m, sm = eng.update(eng.make_batch(flat), perms.to(torch.int32).cuda(), eps_pi.cuda(), eps_kl.cuda(), _hyper(hp),
                       hp.num_mini_batches, use_graph=True)
    torch.cuda.synchronize()
    sm = sm.cpu()
    from reppo_b200 import _lib as L
    col = {k: i for i, k in enumerate(L.METRIC_NAMES)}
    steps = len(traces)
    assert sm.shape[0] == steps and torch.isfinite(sm[:, :20]).all()
    for i, t in enumerate(traces):
        got = {k: sm[i, col[k]].item() for k in ("critic_loss", "critic_grad_norm", "actor_loss", "kl", "entropy")}
        if gemm_mode == "fp32":
            tol = 2e-5 if i == 0 else 5e-3
            assert math.isclose(got["critic_loss"], t["critic/loss"].item(), rel_tol=tol, abs_tol=1e-5), (i, got)
            assert math.isclose(got["critic_grad_norm"], t["critic/grad_norm"].item(), rel_tol=10 * tol, abs_tol=1e-5), (i, got)
            assert math.isclose(got["actor_loss"], t["actor/loss"].item(), rel_tol=10 * tol, abs_tol=1e-4), (i, got)
            assert math.isclose(got["kl"], t["actor/kl"].item(), rel_tol=20 * tol, abs_tol=1e-5), (i, got)
            assert math.isclose(got["entropy"], t["actor/entropy"].item(), rel_tol=10 * tol, abs_tol=1e-4), (i, got)
        else:
            if gemm_mode == "tf32" and i == 0:
                assert math.isclose(got["critic_loss"], t["critic/loss"].item(), rel_tol=1e-3), (i, got)
                assert math.isclose(got["critic_grad_norm"], t["critic/grad_norm"].item(), rel_tol=1e-2), (i, got)
            assert math.isclose(got["critic_loss"], t["critic/loss"].item(), rel_tol=1e-2 if i == 0 else 5e-2), (i, got)
            assert math.isclose(got["entropy"], t["actor/entropy"].item(), rel_tol=5e-2, abs_tol=5e-3), (i, got)
            assert math.isclose(got["kl"], t["actor/kl"].item(), rel_tol=0.25, abs_tol=1e-4), (i, got)
    for net, e, ref in (("critic", eng.critic, ts2.critic), ("actor", eng.actor, ts2.actor)):
        tot, cnt = 0.0, 0
        for k, v in ref.items():
            got = e.view(e.params, k)
            if gemm_mode == "fp32":
                _param_close(f"{net}.{k}", got, v, hp.lr, steps)
            d = (got.cpu().double() - v.double()).abs()
            assert d.max().item() <= 2 * hp.lr * steps + 1e-6, (net, k, d.max().item())
            tot += d.sum().item(); cnt += d.numel()
        assert tot / cnt <= 0.35 * hp.lr * steps, (net, tot / cnt / (hp.lr * steps))
    assert eng.launch_count() > 0


def test_kernel_probe_cross_entropy_matches_oracle():
    """reppo_kernel_probe(0): the fused HL-Gauss histogram + softmax cross-entropy kernel on free-standing buffers against
    autograd through the oracle's restatement (the entry point bench.py's `loss_kernels` roofline leg times)."""
    import torch.nn.functional as F
    hp = O.Hyper(**SMALL)
    sem = O.JAX
    eng = _engine(hp, sem)
    rows, B = 777, hp.num_bins
    g = torch.Generator().manual_seed(5)
    head = torch.randn(rows, B, generator=g)
    zero = torch.rand(B, generator=g) * 0.1
    target = hp.vmin + (hp.vmax - hp.vmin) * torch.rand(rows, generator=g)
    trunc = (torch.rand(rows, generator=g) < 0.1).float()
    h = head.clone().requires_grad_(True)
    logits = h + sem.logit_bias_scale * zero
    ce = -(O.hl_gauss_jax(target, B, hp.vmin, hp.vmax) * F.log_softmax(logits, dim=-1)).sum(-1)
    ((1.0 - trunc) * ce).mean().backward()
    out = torch.empty(rows, B, device="cuda")
    scratch = torch.zeros(24 + 2 * B, device="cuda")
    eng.kernel_probe(0, head.cuda(), target.cuda(), trunc.cuda(), zero.cuda(), out, scratch, _hyper(hp))
    torch.testing.assert_close(out.cpu(), h.grad, rtol=2e-4, atol=2e-7)


def test_kernel_probe_aux_regression_matches_autograd():
    """reppo_kernel_probe(1): next-embedding (+ reward, JAX: slot 0 of the prediction head) regression and d pred on
    free-standing buffers against autograd of the restated loss term (src/jaxrl/reppo.py:491-507)."""
    hp = O.Hyper(**SMALL)
    sem = O.JAX
    eng = _engine(hp, sem)
    rows, H = 513, hp.critic_hidden_dim
    P = H + 1
    g = torch.Generator().manual_seed(11)
    pred = torch.randn(rows, P, generator=g)
    nxt = torch.randn(rows, H, generator=g)
    rew = torch.rand(rows, generator=g)
    done = (torch.rand(rows, generator=g) < 0.2).float()
    p = pred.clone().requires_grad_(True)
    sq = torch.cat([(p[:, 1:] - nxt) ** 2, (p[:, :1] - rew.reshape(-1, 1)) ** 2], dim=-1)
    aux = ((1.0 - done).reshape(-1, 1) * sq).mean(-1)
    # the probe passes `done` as the truncation mask too (mask = 1 - done)
    ((1.0 - done) * hp.aux_loss_mult * aux).mean().backward()
    out = torch.empty(rows, P, device="cuda")
    scratch = torch.zeros(24 + P, device="cuda")
    eng.kernel_probe(1, pred.cuda(), nxt.cuda(), rew.cuda(), done.cuda(), out, scratch, _hyper(hp))
    torch.testing.assert_close(out.cpu(), p.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("rows,in_f,out_f", [(256, 8, 128), (14336, 256, 128), (257, 1024, 256)])
def test_pair_kernel_boundary_shapes(rows, in_f, out_f):
    """Shapes on either side of the CTA-pair kernel's eligibility threshold (56 pair tiles: 14336 x 128 is exactly 56
    128-wide tiles; one pair tile and a ragged 257-row case fall back to the one-tile-per-CTA kernel)."""
    from tests.test_gemm_tc import _bf, _engine as _geng
    eng = _geng(max_rows=16384, H=1024)
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, in_f, device="cuda", generator=g)
    w = torch.randn(out_f, in_f, device="cuda", generator=g) / math.sqrt(in_f)
    y = eng.linear(0, x, w)
    torch.testing.assert_close(y, _bf(x) @ _bf(w).T, rtol=1e-4, atol=2e-4 * math.sqrt(in_f))
