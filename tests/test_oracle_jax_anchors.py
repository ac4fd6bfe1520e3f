"""Third-party anchors for the JAX-semantics constants of the oracle.

The reference's JAX update cannot run here (jax / flax / optax / distrax are absent), so the JAX side of
``oracle/reppo_oracle.py`` is pinned only to the documented upstream definitions.  These tests tie each of those
definitions to an INDEPENDENT implementation that is installed: torch.nn / torch.optim / torch.distributions and scipy
compute the same published formulas that flax.nnx.RMSNorm / LayerNorm, optax.adam / clip_by_global_norm /
softmax_cross_entropy, distrax.Transformed(Normal, Tanh) and jax.scipy.special.erf document.  They do not replace
fixtures recorded from the real JAX run; they catch a mis-restated constant or formula."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import reppo_oracle as O


def test_rmsnorm_and_layernorm_match_torch_modules():
    # flax.nnx.RMSNorm(epsilon=1e-6): x * rsqrt(mean(x^2) + eps) * scale ; nnx.LayerNorm(epsilon=1e-6, use_fast_variance)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 96, generator=g, dtype=torch.float64) * 3 + 0.5
    w, b = torch.rand(96, generator=g, dtype=torch.float64) + 0.5, torch.randn(96, generator=g, dtype=torch.float64)
    p = {"blk.1.weight": w, "blk.1.bias": b}
    rms = torch.nn.RMSNorm(96, eps=1e-6, dtype=torch.float64)
    rms.weight.data.copy_(w)
    got = O._norm(x, p, "blk", O.Hyper(norm_type="rms"), O.JAX)
    torch.testing.assert_close(got, rms(x), rtol=1e-12, atol=1e-12)
    ln = torch.nn.LayerNorm(96, eps=1e-6, dtype=torch.float64)
    ln.weight.data.copy_(w)
    ln.bias.data.copy_(b)
    got = O._norm(x, p, "blk", O.Hyper(norm_type="layer"), O.JAX)
    torch.testing.assert_close(got, ln(x), rtol=1e-9, atol=1e-9)   # fast variance E[x^2] - E[x]^2 vs two-pass


def test_tanh_gaussian_log_prob_matches_torch_distributions():
    # distrax.Transformed(Normal(mu, std), Tanh()).sample_and_log_prob: base log-prob at the pre-tanh sample minus
    # Tanh.forward_log_det_jacobian(x) = 2 (log 2 - x - softplus(-2x))
    from torch.distributions import Normal, TransformedDistribution
    from torch.distributions.transforms import TanhTransform
    g = torch.Generator().manual_seed(1)
    mu = torch.randn(128, 6, generator=g, dtype=torch.float64)
    std = torch.rand(128, 6, generator=g, dtype=torch.float64) + 0.1
    x = mu + std * torch.randn(128, 6, generator=g, dtype=torch.float64)
    dist = TransformedDistribution(Normal(mu, std), TanhTransform(cache_size=1))
    a = torch.tanh(x)
    dist.transforms[0]._cached_x_y = (x, a)   # evaluate the Jacobian at the pre-tanh sample, as distrax does
    want = Normal(mu, std).log_prob(x) - dist.transforms[0].log_abs_det_jacobian(x, a)
    got = O._normal_logpdf(x, mu, std) - O._fldj(x)
    torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)
    # and the closed form of the Jacobian itself: log(1 - tanh(x)^2)
    small = x.clamp(-3, 3)
    torch.testing.assert_close(O._fldj(small), torch.log1p(-torch.tanh(small) ** 2), rtol=1e-9, atol=1e-9)


def test_reverse_kl_matches_torch_distributions_and_closed_form():
    """cfg.reverse_kl (src/jaxrl/reppo.py:543-553): mean over samples a ~ pi of log pi(a) - log pi_old(a).  Without clipping it
    is an unbiased estimate of KL(N(mu, std) || N(mu_o, std_o)) (the tanh Jacobians cancel); the oracle's value is held against
    torch.distributions on the same samples, and its sample mean against the Gaussian closed form."""
    from torch.distributions import Normal, TransformedDistribution, kl_divergence
    from torch.distributions.transforms import TanhTransform
    hp = O.Hyper(num_steps=4, num_envs=64, num_mini_batches=1, num_epochs=1, critic_hidden_dim=32, actor_hidden_dim=32, num_bins=11,
                 vmin=0.0, vmax=5.0, obs_dim=3, critic_obs_dim=3, action_dim=2, reverse_kl=True, actor_kl_clip_mode="full")
    sem = O.JAX
    cp, ap = O.init_params(hp, sem, seed=0)
    ap_old = {k: v + 0.05 * torch.randn(v.shape, generator=torch.Generator().manual_seed(5)) for k, v in ap.items()}
    batch = O.synthetic_rollout(hp, seed=0)
    flat = O.flatten_batch(batch, O.compute_targets({k: v.clone() for k, v in batch.items()}, hp, sem))
    eps_pi, eps_kl = O.synthetic_noise(hp)
    _, m = O.actor_loss(ap, ap_old, cp, flat, eps_pi[0], eps_kl[0], hp, sem)
    mu, std = O.actor_forward(ap, flat["obs"], hp, sem)
    mu_o, std_o = O.actor_forward(ap_old, flat["obs"], hp, sem)
    # float64 restatement through torch.distributions: log pi at the sample itself (distrax sample_and_log_prob, before the
    # clip), log pi_old at the clipped action (old_pi.log_prob(pi_action) after jnp.clip)
    d = lambda t: t.double()
    x = d(mu).unsqueeze(0) + d(std).unsqueeze(0) * d(eps_kl[0])
    pi = TransformedDistribution(Normal(d(mu), d(std)), TanhTransform())
    old = TransformedDistribution(Normal(d(mu_o), d(std_o)), TanhTransform())
    lp_pi = Normal(d(mu), d(std)).log_prob(x) - torch.log1p(-torch.tanh(x) ** 2)
    a_c = torch.tanh(x).clamp(-sem.action_clip, sem.action_clip)
    want = (lp_pi.sum(-1) - old.log_prob(a_c).sum(-1)).mean(0).mean()
    assert abs(m["kl"].item() - want.item()) <= 2e-4 * max(1.0, abs(want.item()))   # the oracle runs in fp32
    unclipped = (torch.tanh(x).abs() <= sem.action_clip).all(-1)
    torch.testing.assert_close(pi.log_prob(torch.tanh(x))[unclipped.unsqueeze(-1).expand_as(x)], lp_pi[unclipped.unsqueeze(-1).expand_as(x)],
                               rtol=1e-6, atol=1e-6)
    # without the clip the estimate is unbiased for the Gaussian closed form (the tanh Jacobians cancel)
    closed = kl_divergence(Normal(mu, std), Normal(mu_o, std_o)).sum(-1).mean()
    free = (lp_pi.sum(-1) - (Normal(d(mu_o), d(std_o)).log_prob(x) - torch.log1p(-torch.tanh(x) ** 2)).sum(-1)).mean()
    assert abs(free.item() - closed.item()) <= 0.25 * closed.item()                   # 16 x 256 samples
    assert m["kl"].item() > 0


def test_optax_adam_chain_matches_torch_adam():
    # optax.chain(clip_by_global_norm(c), adam(lr, b1=.9, b2=.999, eps=1e-8, eps_root=0)) : g <- g * min(1, c / ||g||),
    # bias-corrected moments, p <- p - lr * m_hat / (sqrt(v_hat) + eps).  torch.optim.Adam implements the same update.
    g = torch.Generator().manual_seed(2)
    params = {"a": torch.randn(17, 5, generator=g, dtype=torch.float64), "b": torch.randn(9, generator=g, dtype=torch.float64)}
    tp = [torch.nn.Parameter(v.clone()) for v in params.values()]
    opt = torch.optim.Adam(tp, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    st = O.AdamState.zeros_like(params)
    for step in range(5):
        grads = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) * (3.0 if step % 2 else 0.01) for k, v in params.items()}
        gn = math.sqrt(sum(float((x ** 2).sum()) for x in grads.values()))
        scale = 1.0 if gn < 0.5 else 0.5 / gn            # optax.clip_by_global_norm(0.5)
        for p_, g_ in zip(tp, grads.values()):
            p_.grad = g_ * scale
        opt.step()
        params, st, got_gn = O.clip_and_adam(params, grads, st, 3e-4, 0.5, O.JAX)
        assert math.isclose(float(got_gn), gn, rel_tol=1e-12)
        for p_, v in zip(tp, params.values()):
            torch.testing.assert_close(v, p_.data, rtol=1e-10, atol=1e-12)


def test_hl_gauss_matches_scipy_erf_and_soft_cross_entropy():
    # jaxrl/utils.py:43-59: cdf = erf((edges - clip(x)) / (sqrt(2) sigma)), probs = diff(cdf) / (cdf[-1] - cdf[0]);
    # optax.softmax_cross_entropy(logits, probs) = -sum(probs * log_softmax(logits))
    from scipy.special import erf
    hp = O.Hyper()
    B, vmin, vmax = hp.num_bins, hp.vmin, hp.vmax
    x = np.array([vmin - 5.0, vmin, 0.3 * vmax, 0.77 * vmax, vmax, vmax + 9.0])
    bw = (vmax - vmin) / (B - 1)
    sigma = 0.75 * bw
    edges = np.linspace(vmin - bw / 2, vmax + bw / 2, B + 1)
    cdf = erf((edges[None, :] - np.clip(x, vmin, vmax)[:, None]) / (math.sqrt(2.0) * sigma))
    want = (cdf[:, 1:] - cdf[:, :-1]) / (cdf[:, -1:] - cdf[:, :1])
    got = O.hl_gauss_jax(torch.tensor(x, dtype=torch.float64), B, vmin, vmax)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-12)
    logits = torch.randn(6, B, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    ce = -(got * F.log_softmax(logits, dim=-1)).sum(-1)
    torch.testing.assert_close(ce, F.cross_entropy(logits, got, reduction="none"), rtol=1e-12, atol=1e-12)


def test_td_lambda_jax_matches_direct_recursion_in_numpy():
    # compute_nstep_lambda, src/jaxrl/reppo.py:422-450, restated a second time as a plain numpy loop over (t, n)
    T, N = 9, 5
    rng = np.random.default_rng(4)
    sr, val = rng.normal(size=(T, N)), rng.normal(size=(T, N))
    done = (rng.random((T, N)) < 0.2).astype(np.float64)
    trunc = (rng.random((T, N)) < 0.15).astype(np.float64)
    iw = np.log(rng.uniform(0.5, 1.0, size=(T, N)))
    gamma, lam = 0.97, 0.9
    got = O.td_lambda_jax(*(torch.tensor(a) for a in (sr, val, done, trunc, iw)), gamma, lam).numpy()
    want = np.zeros((T, N))
    for n in range(N):
        lam_ret, trunc_next, w_next = val[T - 1, n], 1.0, 0.0     # scan carry: (value[-1], truncated = 1, iw = 0)
        for t in range(T - 1, -1, -1):
            lam_t = lam * math.exp(w_next)
            lam_sum = lam_t * lam_ret + (1.0 - lam_t) * val[t, n]
            # jnp.where(truncated, value, (1 - done) * lambda_sum): the truncation branch is not masked by done
            lam_ret = sr[t, n] + gamma * (val[t, n] if trunc_next != 0 else (1.0 - done[t, n]) * lam_sum)
            want[t, n] = lam_ret
            trunc_next, w_next = trunc[t, n], iw[t, n]
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
