"""Generate the committed golden fixtures under ``tests/golden/`` by RUNNING THE REFERENCE.

Run in the build container only (``/root/reference`` present):

    python -m oracle.gen_golden

It imports the reference's own Torch closures through ``oracle/ref_torch_shim.py`` and
records, for seeded synthetic rollouts and injected weights / permutation / noise:

* ``gve`` -- ``compute_gve`` output (src/torchrl/reppo.py:175-189),
* per-step logs of ``update_critic`` / ``update_actor`` (:275-283, :348-358),
* post-update parameters of ``Actor`` / ``Critic`` (full for the small case, digests otherwise),
* ``hl_gauss`` and ``Critic`` / ``Actor`` forward outputs on fixed inputs.

Test infrastructure only; the reference is never copied, only executed.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import reppo_oracle as O  # noqa: E402
from oracle import ref_torch_shim as RB  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (Hyper kwargs, terminate, store_full_params)
    "torch_small": (dict(num_steps=8, num_envs=64, num_mini_batches=4, num_epochs=2,
                         critic_hidden_dim=128, actor_hidden_dim=128, num_bins=51, vmin=0.0, vmax=20.0,
                         obs_dim=5, critic_obs_dim=5, action_dim=2, actor_min_std=0.0), True, True),
    "torch_medium": (dict(num_steps=8, num_envs=128, num_mini_batches=4, num_epochs=2,
                          critic_hidden_dim=256, actor_hidden_dim=256, num_bins=151, vmin=0.0, vmax=150.0,
                          obs_dim=17, critic_obs_dim=17, action_dim=6, actor_min_std=0.0), False, False),
    "torch_full_mode": (dict(num_steps=4, num_envs=64, num_mini_batches=2, num_epochs=1,
                             critic_hidden_dim=128, actor_hidden_dim=128, num_bins=51, vmin=-10.0, vmax=10.0,
                             obs_dim=7, critic_obs_dim=7, action_dim=3, actor_min_std=0.05,
                             actor_kl_clip_mode="full", kl_bound=0.05), True, True),
}


def digest(t: torch.Tensor):
    f = t.detach().double().reshape(-1)
    return {"head": t.detach().reshape(-1)[:64].clone(), "sum": f.sum(), "abs_sum": f.abs().sum(),
            "sq_sum": (f * f).sum(), "shape": tuple(t.shape)}


def gen_case(ref, name, kwargs, terminate, full):
    hp = O.Hyper(**kwargs)
    cp, ap = O.init_params(hp, O.TORCH, seed=0)
    batch = O.synthetic_rollout(hp, seed=0, terminate=terminate)
    perms = O.synthetic_perms(hp)
    eps_pi, eps_kl = O.synthetic_noise(hp)
    ts = RB.build_train_state(ref, hp, cp, ap)
    data, logs = RB.run_update(ref, ts, hp, {k: v.clone() for k, v in batch.items()}, perms, eps_pi, eps_kl)
    # forward goldens on the first minibatch with the INITIAL weights
    ts0 = RB.build_train_state(ref, hp, cp, ap)
    idx = perms[0, :hp.minibatch_size]
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in batch.items()}
    with torch.no_grad():
        value, logits, pred, feat = ts0.critic(flat["critic_obs"][idx], flat["action"][idx])
        pi, tanh_mean, temp, beta = ts0.actor(flat["obs"][idx])
    out = {
        "hyper": kwargs, "terminate": terminate, "semantics": "torch",
        "inputs": {"critic": cp, "actor": ap, "batch": batch, "perms": perms, "eps_pi": eps_pi, "eps_kl": eps_kl},
        "gve": data["gve"].reshape(hp.num_steps, hp.num_envs).clone(),
        "logs": [{k: (v.clone() if v.dim() == 0 else v.clone()) for k, v in l.items()} for l in logs],
        "fwd": {"value": value, "logits": logits, "pred": pred, "features": feat,
                "mean": pi.base_dist.loc.clone(), "std": pi.base_dist.scale.clone(), "tanh_mean": tanh_mean},
    }
    cs, as_ = ts.critic.state_dict(), ts.actor.state_dict()
    if full:
        out["final"] = {"critic": {k: v.clone() for k, v in cs.items()}, "actor": {k: v.clone() for k, v in as_.items()}}
    else:
        out["final_digest"] = {"critic": {k: digest(v) for k, v in cs.items()},
                               "actor": {k: digest(v) for k, v in as_.items()}}
        out["inputs"]["batch"] = {k: v for k, v in batch.items() if k != "next_emb"}
        out["inputs"]["next_emb_digest"] = digest(batch["next_emb"])
        out["inputs"]["critic"] = None  # regenerated from seed 0 by init_params; digests pin them
        out["inputs"]["actor"] = None
        out["inputs"]["param_digest"] = {"critic": {k: digest(v) for k, v in cp.items()},
                                          "actor": {k: digest(v) for k, v in ap.items()}}
        # per-row vectors in the logs are small; keep
    torch.save(out, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: {len(logs)} steps, qf_loss[0]={logs[0]['qf_loss'].item():.6f} "
          f"actor_loss[-1]={logs[-1]['actor_loss'].item():.6f}")


def gen_subfunctions(ref):
    xs = torch.tensor([0.0, 1.0, 74.3, 150.0, 200.0, -5.0, 12.345, 149.99])
    out = {"x": xs, "hl_151_0_150": ref.util.hl_gauss(xs.unsqueeze(-1), 0, 150, 151)}
    xs2 = torch.tensor([-12.0, -10.0, 0.0, 0.07, 9.5, 11.0])
    out["x2"] = xs2
    out["hl_51_m10_10"] = ref.util.hl_gauss(xs2.unsqueeze(-1), -10, 10, 51)
    torch.save(out, os.path.join(OUT, "torch_hl_gauss.pt"))
    print("hl_gauss goldens written")



def gen_rollout(ref):
    """Rollout-side goldens from the reference modules: the policy / bootstrap arithmetic of collect_fn
    (src/torchrl/reppo.py:100-158) with injected noise, and EmpiricalNormalization (reppo_util.py:394-471)."""
    kwargs = CASES["torch_small"][0]
    hp = O.Hyper(**kwargs)
    cp, ap = O.init_params(hp, O.TORCH, seed=0)
    ts0 = RB.build_train_state(ref, hp, cp, ap)
    g = torch.Generator().manual_seed(5)
    n = 96
    obs = torch.randn(n, hp.obs_dim, generator=g)
    nobs = torch.randn(n, hp.obs_dim, generator=g)
    ncobs = torch.randn(n, hp.critic_obs_dim, generator=g)
    eps, eps2 = torch.randn(n, hp.action_dim, generator=g), torch.randn(n, hp.action_dim, generator=g) * 1.5
    reward = torch.rand(n, generator=g)
    with torch.no_grad():
        pi, _, _, _ = ts0.actor(obs)
        a = torch.tanh(pi.base_dist.loc + pi.base_dist.scale * eps)                 # pi.sample() with the noise injected
        lp = pi.log_prob(a.clip(-0.999, 0.999)).sum(-1)                              # :146
        npi, _, temp, _ = ts0.actor(nobs)
        na = torch.tanh(npi.base_dist.loc + npi.base_dist.scale * eps2)
        nlp = npi.log_prob(na.clip(-1 + 1e-6, 1 - 1e-6)).sum(-1)                     # :132-134
        nv, _, _, nemb = ts0.critic(ncobs, na)                                       # :135
        soft = reward - hp.gamma * nlp * temp                                        # :136-138
    out = {"hyper": kwargs, "obs": obs, "next_obs": nobs, "next_critic_obs": ncobs, "eps": eps, "eps_next": eps2, "reward": reward,
           "action": a, "log_prob": lp, "next_action": na, "next_log_prob": nlp, "next_value": nv, "next_emb": nemb,
           "soft_reward": soft}
    # normaliser: three training-mode calls, then one eval call
    norm = ref.util.EmpiricalNormalization(shape=7, device="cpu")
    xs = [torch.randn(33, 7, generator=g) * 3 + 1, torch.randn(64, 7, generator=g) * 0.5 - 2, torch.randn(5, 7, generator=g)]
    ys, states = [], []
    norm.train()
    for x in xs:
        ys.append(norm(x).clone())
        states.append({"mean": norm._mean.clone(), "var": norm._var.clone(), "std": norm._std.clone(), "count": int(norm.count)})
    norm.eval()
    ys.append(norm(xs[0]).clone())
    out["norm"] = {"xs": xs, "ys": ys, "states": states, "keys": list(norm.state_dict().keys())}
    out["state_dict_keys"] = {"actor": list(ts0.actor.state_dict().keys()), "critic": list(ts0.critic.state_dict().keys())}
    torch.save(out, os.path.join(OUT, "torch_rollout.pt"))
    print(f"rollout goldens written: logp[0]={lp[0].item():.6f} next_value[0]={nv[0].item():.6f}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order for the fixtures
    ref = RB.load_reference()
    for name, (kw, term, full) in CASES.items():
        gen_case(ref, name, kw, term, full)
    gen_subfunctions(ref)
    gen_rollout(ref)


if __name__ == "__main__":
    main()
