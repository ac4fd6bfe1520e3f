"""Oracle A -- CPU restatement of REPPO's update step (TEST INFRASTRUCTURE ONLY).

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The shipped path (``reppo_b200``) never routes through it.

It restates, in plain PyTorch on the CPU (fp32 or fp64), what the reference computes
for one on-policy rollout batch:

* JAX semantics (the parity target) -- ``src/jaxrl/reppo.py:417-673`` (``learn_step``),
  optimizer ``:246-257``, networks ``src/networks/jax_models.py:35-49,57-124,211-383``,
  ``hl_gauss`` ``src/jaxrl/utils.py:43-59``.
* Torch semantics (the API kept) -- ``src/torchrl/reppo.py:173-360,613-632``, networks
  ``src/networks/torch_models.py:71-142,209-335``, ``hl_gauss``
  ``src/torchrl/reppo_util.py:756-777``.

Parity status.  The Torch-semantics mode is PINNED: ``oracle/gen_golden.py`` runs the
reference's own Torch closures (imported from /root/reference through stub modules) and
``tests/test_oracle_golden.py`` checks this restatement against those outputs.  The
JAX-semantics constants (RMSNorm eps 1e-6, logit bias 40.0, clip 1-1e-4, min_std 0.1,
reward-aux head, carry-shifted truncation / importance weights) are PARITY UNPINNED:
jax/flax/optax/distrax are not installable in this image, so they follow the reference
source text and the published semantics of jax==0.5.2 / flax>=0.10.3 / optax>=0.2.4 /
distrax>=0.1.5 only.

Gradients here come from ``torch.autograd``; the CUDA path uses hand-derived
backward kernels, so agreement is an independent check of that derivation.

Parameter layout: dicts keyed by the reference Torch ``state_dict`` names
(``src/networks/torch_models.py:209-335``), weights ``[out, in]``.  In JAX mode the last
pred layer has ``H+1`` outputs (slot 0 = reward, ``jax_models.py:318-320``).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
KL_SAMPLES = 16  # src/jaxrl/reppo.py:557, src/torchrl/reppo.py:308


# ----------------------------------------------------------------------------------------
# semantics + hyper-parameters
# ----------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Semantics:
    """Constants on which the three reference implementations disagree."""

    name: str
    norm_eps: float          # RMSNorm eps: flax 1e-6 | torch finfo(f32).eps
    logit_bias_scale: float  # 40.0 jax_models.py:298 | 40.9 torch_models.py:282
    action_clip: float       # 1-1e-4 jaxrl/reppo.py:558 | 1-1e-6 torchrl/reppo.py:301,308
    clip_policy_sample: bool  # torch clips the rsample before log_prob (:301)
    old_logp_at_clipped: bool  # torch: old_pi.log_prob(clipped) (:309); jax: at raw x'
    force_min_std: Optional[float]  # jax ignores cfg.actor_min_std -> 0.1 (jax_models.py:336)
    reward_head: bool        # pred head H+1 with reward slot (jax) | H (torch)
    aux_mask_done: bool      # aux masked by (1-done) (jax :499) | only by loss mask (torch)
    td_convention: str       # "jax" (:422-450) | "torch" (:175-189)
    adamw_weight_decay: float  # 0 (optax.adam) | 0.01 (torch AdamW default)
    clip_eps: float          # optax: none | torch clip_grad_norm_: 1e-6


JAX = Semantics("jax", 1e-6, 40.0, 1.0 - 1e-4, False, False, 0.1, True, True, "jax", 0.0, 0.0)
TORCH = Semantics("torch", float(torch.finfo(torch.float32).eps), 40.9, 1.0 - 1e-6, True, True,
                  None, False, False, "torch", 0.01, 1e-6)


@dataclass
class Hyper:
    """The slice of ``config/reppo.yaml`` the update consumes (defaults :16-70)."""

    num_steps: int = 128
    num_envs: int = 1024
    num_mini_batches: int = 128
    num_epochs: int = 4
    gamma: float = 0.99
    lmbda: float = 0.95
    lr: float = 3e-4
    max_grad_norm: float = 0.5
    critic_hidden_dim: int = 512
    actor_hidden_dim: int = 512
    vmin: float = 0.0
    vmax: float = 150.0
    num_bins: int = 151
    use_critic_norm: bool = True
    use_actor_norm: bool = True
    num_critic_encoder_layers: int = 2
    num_critic_head_layers: int = 2
    num_critic_pred_layers: int = 2
    num_actor_layers: int = 3
    actor_min_std: float = 0.0
    kl_start: float = 0.01
    kl_bound: float = 0.1
    reduce_kl: bool = True
    reverse_kl: bool = False  # jaxrl/reppo.py:543-553 (JAX update only)
    update_kl_lagrangian: bool = True
    actor_kl_clip_mode: str = "clipped"
    ent_start: float = 0.01
    ent_target_mult: float = 0.5
    update_entropy_lagrangian: bool = True
    aux_loss_mult: float = 1.0
    norm_type: str = "rms"  # "rms" (jax_models.py:46) | "layer" (reppo_mj_playground.py:65-72)
    partial_reset: bool = False  # cfg.env.get("partial_reset") torchrl/reppo.py:237
    obs_dim: int = 17
    critic_obs_dim: int = 17
    action_dim: int = 6

    @property
    def batch_size(self) -> int:
        return self.num_steps * self.num_envs

    @property
    def minibatch_size(self) -> int:
        return self.batch_size // self.num_mini_batches

    def min_std(self, sem: Semantics) -> float:
        return sem.force_min_std if sem.force_min_std is not None else self.actor_min_std


# ----------------------------------------------------------------------------------------
# parameter construction
# ----------------------------------------------------------------------------------------
def _fcnn_keys(prefix: str, layers: int, input_activation: bool) -> List[str]:
    """Sequential indices of ``torch_models.FCNN`` (:97-139): an input activation
    occupies slot 0 and shifts every later block by one."""
    off = 1 if input_activation else 0
    return [f"{prefix}.net.{off + i}" for i in range(layers)]


def critic_param_shapes(hp: Hyper, sem: Semantics) -> Dict[str, tuple]:
    H, B = hp.critic_hidden_dim, hp.num_bins
    K0 = hp.critic_obs_dim + hp.action_dim
    pred_out = H + 1 if sem.reward_head else H
    shapes: Dict[str, tuple] = {"zero_dist": (B,)}

    def fcnn(prefix, in_f, out_f, layers, inp_act, use_norm):
        blocks = _fcnn_keys(prefix, layers, inp_act)
        dims = [in_f] + [H] * (layers - 1) + [out_f]
        for i, blk in enumerate(blocks):
            shapes[f"{blk}.0.weight"] = (dims[i + 1], dims[i])
            shapes[f"{blk}.0.bias"] = (dims[i + 1],)
            if use_norm and i < layers - 1:
                shapes[f"{blk}.1.weight"] = (dims[i + 1],)
                if hp.norm_type == "layer":
                    shapes[f"{blk}.1.bias"] = (dims[i + 1],)

    fcnn("feature_module", K0, H, hp.num_critic_encoder_layers, False, hp.use_critic_norm)
    fcnn("critic_module", H, B, hp.num_critic_head_layers, True, hp.use_critic_norm)
    fcnn("pred_module", H, pred_out, hp.num_critic_pred_layers, True, hp.use_critic_norm)
    return shapes


def actor_param_shapes(hp: Hyper) -> Dict[str, tuple]:
    Ha = hp.actor_hidden_dim
    shapes: Dict[str, tuple] = {"log_temp": (), "log_lagrange": ()}
    blocks = _fcnn_keys("model", hp.num_actor_layers, False)
    dims = [hp.obs_dim] + [Ha] * (hp.num_actor_layers - 1) + [2 * hp.action_dim]
    for i, blk in enumerate(blocks):
        shapes[f"{blk}.0.weight"] = (dims[i + 1], dims[i])
        shapes[f"{blk}.0.bias"] = (dims[i + 1],)
        if hp.use_actor_norm and i < hp.num_actor_layers - 1:
            shapes[f"{blk}.1.weight"] = (dims[i + 1],)
            if hp.norm_type == "layer":
                shapes[f"{blk}.1.bias"] = (dims[i + 1],)
    return shapes


def init_params(hp: Hyper, sem: Semantics, seed: int = 0, dtype=torch.float32):
    """Random-init weights with the reference's value distribution.

    JAX mode: lecun-normal kernels (truncated normal, sigma = sqrt(1/fan_in)/0.87962566),
    zero biases (flax ``nnx.Linear`` defaults).  Torch mode: ``nn.Linear`` default
    U(+-1/sqrt(fan_in)) for weight and bias.  Norm scales 1, ``zero_dist = hl_gauss(0)``,
    ``log_temp = log(ent_start)``, ``log_lagrange = log(kl_start)``
    (jax_models.py:288-290,354-357; torch_models.py:268-276,313-318).
    Parity tests inject the same tensors on both sides, so the init only sets scale.
    """
    g = torch.Generator().manual_seed(seed)

    def fill(shapes):
        out = {}
        for k, shp in shapes.items():
            if k.endswith(".0.weight"):
                fan_in = shp[1]
                if sem.name == "jax":
                    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
                    w = torch.empty(shp, dtype=torch.float64)
                    torch.nn.init.trunc_normal_(w, 0.0, 1.0, -2.0, 2.0, generator=g)
                    out[k] = (w * std).to(dtype)
                else:
                    bound = 1.0 / math.sqrt(fan_in)
                    out[k] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
            elif k.endswith(".0.bias"):
                if sem.name == "jax":
                    out[k] = torch.zeros(shp, dtype=dtype)
                else:
                    # fan_in of the matching weight
                    fan_in = shapes[k[:-4] + "weight"][1]
                    bound = 1.0 / math.sqrt(fan_in)
                    out[k] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
            elif k.endswith(".1.weight"):
                out[k] = torch.ones(shp, dtype=dtype)
            elif k.endswith(".1.bias"):
                out[k] = torch.zeros(shp, dtype=dtype)
        return out

    critic = fill(critic_param_shapes(hp, sem))
    hl = hl_gauss_jax if sem.name == "jax" else hl_gauss_torch
    critic["zero_dist"] = hl(torch.zeros(1, dtype=dtype), hp.num_bins, hp.vmin, hp.vmax).reshape(-1)
    actor = fill(actor_param_shapes(hp))
    actor["log_temp"] = torch.tensor(math.log(hp.ent_start), dtype=dtype)
    actor["log_lagrange"] = torch.tensor(math.log(hp.kl_start), dtype=dtype)
    # keep reference ordering (zero_dist / log_* first)
    critic = {k: critic[k] for k in critic_param_shapes(hp, sem)}
    actor = {k: actor[k] for k in actor_param_shapes(hp)}
    return critic, actor


# ----------------------------------------------------------------------------------------
# A.1 soft TD(lambda) targets
# ----------------------------------------------------------------------------------------
def td_lambda_jax(soft_reward, value, done, truncated, importance_weight, gamma, lmbda):
    """``compute_nstep_lambda`` + reverse scan, src/jaxrl/reppo.py:422-450.

    All inputs ``[T, N]``.  The truncation flag and the (log-space) importance weight used
    at step t are those of step t+1 (the carry is updated after use, :434-439); the scan is
    seeded with ``value[T-1]``, ``trunc=1``, ``iw=0`` (:444-446).
    """
    T = soft_reward.shape[0]
    G = value[-1].clone()
    trunc_next = torch.ones_like(truncated[0])
    iw_next = torch.zeros_like(soft_reward[0])
    out = torch.empty_like(soft_reward)
    for t in range(T - 1, -1, -1):
        w = torch.exp(iw_next) * lmbda
        lambda_sum = w * G + (1 - w) * value[t]
        delta = gamma * torch.where(trunc_next != 0, value[t], (1.0 - done[t]) * lambda_sum)
        G = soft_reward[t] + delta
        trunc_next = truncated[t]
        iw_next = importance_weight[t] if importance_weight is not None else iw_next
        out[t] = G
    return out


def td_lambda_torch(rewards, dones, truncated, next_values, gamma, lmbda):
    """``compute_gve``, src/torchrl/reppo.py:175-189.  Uses ``truncated[t]`` itself, no
    importance weights, seed 0, and WRITES ``truncated[-1] = 1`` into its argument (:178)."""
    T = rewards.shape[0]
    truncated[-1] = 1.0
    last = torch.zeros_like(rewards[0])
    out = torch.empty_like(rewards)
    for t in range(T - 1, -1, -1):
        lambda_sum = lmbda * last + (1.0 - lmbda) * next_values[t]
        delta = gamma * torch.where(truncated[t] != 0, next_values[t], (1.0 - dones[t]) * lambda_sum)
        last = rewards[t] + delta
        out[t] = last
    return out


# ----------------------------------------------------------------------------------------
# A.2 HL-Gauss
# ----------------------------------------------------------------------------------------
def _support_edges(num_bins, vmin, vmax, dtype):
    bw = (vmax - vmin) / (num_bins - 1)
    # the reference builds the edges in float32 (utils.py:48-50 dtype=jnp.float32;
    # torch.linspace default dtype) -- keep that, then widen if asked
    edges = torch.linspace(vmin - bw / 2, vmax + bw / 2, num_bins + 1, dtype=torch.float32)
    return edges.to(dtype), bw


def hl_gauss_jax(x: Tensor, num_bins: int, vmin: float, vmax: float) -> Tensor:
    """src/jaxrl/utils.py:43-59 with epsilon=0, vmapped over the batch (reppo.py:480-482)."""
    xc = torch.clamp(x, vmin, vmax)
    edges, bw = _support_edges(num_bins, vmin, vmax, x.dtype)
    sigma = bw * 0.75
    sqrt2 = math.sqrt(2.0) if x.dtype == torch.float64 else float(torch.sqrt(torch.tensor(2.0)))
    cdf = torch.erf((edges.unsqueeze(0) - xc.reshape(-1, 1)) / (sqrt2 * sigma))
    z = cdf[:, -1:] - cdf[:, :1]
    return (cdf[:, 1:] - cdf[:, :-1]) / z


def hl_gauss_torch(x: Tensor, num_bins: int, vmin: float, vmax: float) -> Tensor:
    """src/torchrl/reppo_util.py:756-777 (extra 1e-6 in both denominators)."""
    xc = torch.clamp(x, vmin, vmax)
    edges, bw = _support_edges(num_bins, vmin, vmax, x.dtype)
    sigma = bw * 0.75
    denom = torch.sqrt(torch.tensor(2.0, dtype=x.dtype)) * sigma + 1e-6
    cdf = torch.erf((edges.unsqueeze(0) - xc.reshape(-1, 1)) / denom)
    z = cdf[:, -1:] - cdf[:, :1]
    return (cdf[:, 1:] - cdf[:, :-1]) / (z + 1e-6)


# ----------------------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------------------
def _norm(x, p, blk, hp: Hyper, sem: Semantics):
    w = p[f"{blk}.1.weight"]
    if hp.norm_type == "rms":
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + sem.norm_eps) * w
    # flax LayerNorm: fast variance E[x^2]-E[x]^2, eps 1e-6 (reppo_mj_playground.py:65-72)
    mean = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mean * mean
    return (x - mean) * torch.rsqrt(var + 1e-6) * w + p[f"{blk}.1.bias"]


def fcnn(x, p, prefix, layers, input_activation, use_norm, hp, sem):
    """``FCNN.__call__`` jax_models.py:109-124 == ``torch_models.FCNN`` :82-142 for
    layers >= 2: [swish] -> (Linear, norm, swish) x (layers-1) -> Linear."""
    if layers < 2:
        raise ValueError("layers == 1 has divergent JAX/Torch semantics (SURVEY App. B)")
    if input_activation:
        x = F.silu(x)
    blocks = _fcnn_keys(prefix, layers, input_activation)
    for i, blk in enumerate(blocks):
        x = F.linear(x, p[f"{blk}.0.weight"], p[f"{blk}.0.bias"])
        if i < layers - 1:
            if use_norm:
                x = _norm(x, p, blk, hp, sem)
            x = F.silu(x)
    return x


def support(hp: Hyper, dtype):
    return torch.linspace(hp.vmin, hp.vmax, hp.num_bins, dtype=torch.float32).to(dtype)


def critic_features(p, critic_obs, action, hp, sem):
    x = torch.cat([critic_obs, action], dim=-1)
    return fcnn(x, p, "feature_module", hp.num_critic_encoder_layers, False, hp.use_critic_norm, hp, sem)


def critic_logits(p, feat, hp, sem):
    head = fcnn(feat, p, "critic_module", hp.num_critic_head_layers, True, hp.use_critic_norm, hp, sem)
    return head + sem.logit_bias_scale * p["zero_dist"]


def critic_forward(p, critic_obs, action, hp, sem):
    """returns (value, logits, pred, features) like ``Critic.forward`` torch_models.py:278-285;
    in JAX mode ``pred`` is ``[mb, H+1]`` with the reward in slot 0 (jax_models.py:312-323)."""
    feat = critic_features(p, critic_obs, action, hp, sem)
    logits = critic_logits(p, feat, hp, sem)
    value = torch.softmax(logits, dim=-1) @ support(hp, logits.dtype)
    pred = fcnn(feat, p, "pred_module", hp.num_critic_pred_layers, True, hp.use_critic_norm, hp, sem)
    return value, logits, pred, feat


def critic_value(p, critic_obs, action, hp, sem):
    feat = critic_features(p, critic_obs, action, hp, sem)
    logits = critic_logits(p, feat, hp, sem)
    return torch.softmax(logits, dim=-1) @ support(hp, logits.dtype)


def actor_forward(p, obs, hp, sem):
    out = fcnn(obs, p, "model", hp.num_actor_layers, False, hp.use_actor_norm, hp, sem)
    mu, log_std = torch.split(out, hp.action_dim, dim=-1)
    std = torch.exp(log_std) + hp.min_std(sem)
    return mu, std


LOG2 = math.log(2.0)
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _fldj(x):
    """Tanh forward log-det-Jacobian, distrax / torch TanhTransform (torch_models.py:52-55)."""
    return 2.0 * (LOG2 - x - F.softplus(-2.0 * x))


def _normal_logpdf(x, mu, std):
    return -0.5 * ((x - mu) / std) ** 2 - torch.log(std) - HALF_LOG_2PI



# ----------------------------------------------------------------------------------------
# rollout side: policy step, bootstrap, observation normalisation
# ----------------------------------------------------------------------------------------
def policy_act(pa, obs, eps, hp: Hyper, sem: Semantics, clip: float, scale: Optional[Tensor] = None):
    """collect_fn src/torchrl/reppo.py:104-109,146 / step_env src/jaxrl/reppo.py:347-364.
    Returns (action, log_prob of the clipped action under the unscaled policy, the same under the scaled policy);
    clip <= 0: distrax ``sample_and_log_prob`` (log-prob at the pre-tanh sample, src/jaxrl/reppo.py:367-369)."""
    mu, std = actor_forward(pa, obs, hp, sem)
    s = std if scale is None else std * scale.reshape(-1, 1)   # jax_models.py:366
    x = mu + s * eps
    a = torch.tanh(x)
    if clip > 0:
        u = torch.atanh(a.clip(-clip, clip))
        lp = (_normal_logpdf(u, mu, std) - _fldj(u)).sum(-1)
        lps = (_normal_logpdf(u, mu, s) - _fldj(u)).sum(-1)
    else:
        lp = lps = (_normal_logpdf(x, mu, s) - _fldj(x)).sum(-1)
    return a, lp, lps


def bootstrap(pa, pc, next_obs, next_critic_obs, eps, reward, hp: Hyper, sem: Semantics, clip: float):
    """src/torchrl/reppo.py:117-138 (clip 1 - 1e-6) / src/jaxrl/reppo.py:367-374 (clip <= 0)."""
    a, lp, _ = policy_act(pa, next_obs, eps, hp, sem, clip)
    value, _, _, feat = critic_forward(pc, next_critic_obs, a, hp, sem)
    soft = reward - hp.gamma * lp * torch.exp(pa["log_temp"])
    return {"next_value": value, "next_emb": feat, "soft_reward": soft, "next_action": a, "next_log_prob": lp}


class EmpiricalNormalization:
    """Restatement of src/torchrl/reppo_util.py:394-471 (running mean / var by Chan's parallel algorithm)."""

    def __init__(self, dim: int, eps: float = 1e-2, until: Optional[int] = None, dtype=torch.float32):
        self.eps, self.until = eps, until
        self.mean = torch.zeros(1, dim, dtype=dtype)
        self.var = torch.ones(1, dim, dtype=dtype)
        self.std = torch.ones(1, dim, dtype=dtype)
        self.count = 0

    def update(self, x: Tensor):
        x = x.reshape(-1, x.shape[-1])
        if self.until is not None and self.count >= self.until:
            return
        b = x.shape[0]
        bm = x.mean(0, keepdim=True)
        new = self.count + b
        self.mean = self.mean + (b / new) * (bm - self.mean)                      # :446-447
        if self.count > 0:
            bv = ((x - bm) ** 2).mean(0, keepdim=True)                             # :452
            d2 = bm - self.mean                                                    # the UPDATED mean (:453)
            m2 = self.var * self.count + bv * b + d2 ** 2 * (self.count * b / new)
            self.var = m2 / new
        else:
            self.var = ((x - self.mean) ** 2).mean(0, keepdim=True)                # :461-462
        self.std = torch.sqrt(self.var)
        self.count = new

    def __call__(self, x: Tensor, update: bool = True, center: bool = True) -> Tensor:
        if update:
            self.update(x)
        return (x - self.mean) / (self.std + self.eps) if center else x / (self.std + self.eps)

# ----------------------------------------------------------------------------------------
# A.3 critic loss
# ----------------------------------------------------------------------------------------
def critic_loss(p, mb: Dict[str, Tensor], hp: Hyper, sem: Semantics):
    """JAX ``critic_loss_fn`` reppo.py:474-524 / Torch critic ``update`` :227-264.

    ``mb`` keys: critic_obs, action, target, next_emb, reward (raw), done, truncated."""
    value, logits, pred, _ = critic_forward(p, mb["critic_obs"], mb["action"], hp, sem)
    hl = hl_gauss_jax if sem.name == "jax" else hl_gauss_torch
    target_cat = hl(mb["target"], hp.num_bins, hp.vmin, hp.vmax)
    ce = -(target_cat * F.log_softmax(logits, dim=-1)).sum(-1)
    if hp.partial_reset and sem.name == "torch":
        mask = torch.ones_like(mb["truncated"])
    else:
        mask = 1.0 - mb["truncated"]
    if sem.reward_head:
        sq = torch.cat([(pred[:, 1:] - mb["next_emb"]) ** 2,
                        (pred[:, :1] - mb["reward"].reshape(-1, 1)) ** 2], dim=-1)
        aux = ((1.0 - mb["done"]).reshape(-1, 1) * sq).mean(-1)
        loss = (mask * (ce + hp.aux_loss_mult * aux)).mean()
        aux_mean = aux.mean()
    else:
        aux_row = ((pred - mb["next_emb"]) ** 2).mean(-1)
        qf = (mask * ce).mean()
        aux_mean = (mask * aux_row).mean()           # "embedding_loss" torchrl/reppo.py:255-262
        loss = qf + hp.aux_loss_mult * aux_mean
    metrics = {
        "loss": loss.detach(),
        "ce_mean": ce.mean().detach(),
        "aux_mean": aux_mean.detach(),
        "value_loss": ((value - mb["target"]) ** 2).mean().detach(),
        "q": value.mean().detach(),
        "target_mean": mb["target"].mean().detach(),
        "target_max": mb["target"].max().detach(),
        "target_min": mb["target"].min().detach(),
    }
    return loss, metrics


# ----------------------------------------------------------------------------------------
# A.4 actor loss
# ----------------------------------------------------------------------------------------
def actor_loss(pa, pa_old, pc, mb: Dict[str, Tensor], eps_pi: Tensor, eps_kl: Tensor,
               hp: Hyper, sem: Semantics):
    """JAX ``actor_loss`` reppo.py:526-622 (forward-KL branch) / Torch actor ``update``
    :291-338.  ``eps_pi [mb, A]`` and ``eps_kl [16, mb, A]`` are injected standard normals;
    ``pc`` is the already-updated, frozen critic."""
    A = hp.action_dim
    mu, std = actor_forward(pa, mb["obs"], hp, sem)
    x = mu + std * eps_pi
    a = torch.tanh(x)
    if sem.clip_policy_sample:
        # torchrl/reppo.py:301: log_prob(actions.clip(...)) -> atanh, gradient through the clip
        xc = torch.atanh(torch.clamp(a, -sem.action_clip, sem.action_clip))
        logp = (_normal_logpdf(xc, mu, std) - _fldj(xc)).sum(-1)
    else:
        # distrax sample_and_log_prob: base log-prob at the pre-tanh sample minus fldj
        logp = (-0.5 * eps_pi ** 2 - torch.log(std) - HALF_LOG_2PI - _fldj(x)).sum(-1)
    q = critic_value(pc, mb["critic_obs"], a, hp, sem)

    if hp.reverse_kl and sem.name == "jax":
        # jaxrl/reppo.py:543-553: 16 reparameterised samples of the CURRENT policy (gradient flows through them and
        # through the clip), log pi at the pre-tanh sample (distrax sample_and_log_prob), log pi_old at the clipped action
        with torch.no_grad():
            mu_o, std_o = actor_forward(pa_old, mb["obs"], hp, sem)
        x_s = mu.unsqueeze(0) + std.unsqueeze(0) * eps_kl
        lp_pi = (-0.5 * eps_kl ** 2 - torch.log(std) - HALF_LOG_2PI - _fldj(x_s)).sum(-1).mean(0)
        u_s = torch.atanh(torch.clamp(torch.tanh(x_s), -sem.action_clip, sem.action_clip))
        lp_o = (_normal_logpdf(u_s, mu_o, std_o) - _fldj(u_s)).sum(-1).mean(0)
        kl = lp_pi - lp_o
        return _actor_loss_tail(pa, logp, q, a, kl, hp, sem)
    with torch.no_grad():
        mu_o, std_o = actor_forward(pa_old, mb["obs"], hp, sem)
        x_o = mu_o.unsqueeze(0) + std_o.unsqueeze(0) * eps_kl
        a_o = torch.clamp(torch.tanh(x_o), -sem.action_clip, sem.action_clip)
        u = torch.atanh(a_o)
        if sem.old_logp_at_clipped:
            lp_old = (_normal_logpdf(u, mu_o, std_o) - _fldj(u)).sum(-1).mean(0)
        else:
            lp_old = (-0.5 * eps_kl ** 2 - torch.log(std_o) - HALF_LOG_2PI - _fldj(x_o)).sum(-1).mean(0)
    lp_new = (_normal_logpdf(u, mu, std) - _fldj(u)).sum(-1).mean(0)
    kl = lp_old - lp_new
    return _actor_loss_tail(pa, logp, q, a, kl, hp, sem)


def _actor_loss_tail(pa, logp, q, a, kl, hp: Hyper, sem: Semantics):
    """Everything after the KL estimate (jaxrl/reppo.py:566-622; torchrl/reppo.py:313-338)."""
    A = hp.action_dim
    alpha = torch.exp(pa["log_temp"])
    beta = torch.exp(pa["log_lagrange"])
    rk = 1.0 if hp.reduce_kl else 0.0
    if sem.name == "torch":
        rk = 1.0  # the Torch path has no reduce_kl switch (torchrl/reppo.py:313-326)
    pathwise = alpha.detach() * logp - q
    if hp.actor_kl_clip_mode == "clipped":
        l_pi = torch.where(kl < hp.kl_bound, pathwise, kl * beta.detach() * rk)
    elif hp.actor_kl_clip_mode == "full":
        l_pi = pathwise + kl * beta.detach() * rk
    elif hp.actor_kl_clip_mode == "value":
        l_pi = pathwise
    else:
        raise ValueError(f"Unknown actor loss mode: {hp.actor_kl_clip_mode}")
    target_entropy = A * hp.ent_target_mult
    ent_loss = alpha * (target_entropy - logp).detach().mean()
    lag_loss = -beta * (kl - hp.kl_bound).detach().mean()
    loss = l_pi.mean()
    if hp.update_entropy_lagrangian or sem.name == "torch":
        loss = loss + ent_loss
    if hp.update_kl_lagrangian or sem.name == "torch":
        loss = loss + lag_loss
    metrics = {
        "loss": loss.detach(),
        "policy_loss": l_pi.mean().detach(),
        "kl": kl.mean().detach(),
        "entropy": (-logp).mean().detach(),
        "temperature": alpha.detach(),
        "lagrangian": beta.detach(),
        "entropy_loss": ent_loss.detach(),
        "lagrangian_loss": lag_loss.detach(),
        "q": q.mean().detach(),
        "abs_pred_action": a.abs().mean().detach(),
    }
    return loss, metrics


# ----------------------------------------------------------------------------------------
# A.6 optimizer
# ----------------------------------------------------------------------------------------
@dataclass
class AdamState:
    m: Dict[str, Tensor]
    v: Dict[str, Tensor]
    t: int = 0

    @staticmethod
    def zeros_like(params):
        return AdamState({k: torch.zeros_like(v) for k, v in params.items()},
                         {k: torch.zeros_like(v) for k, v in params.items()}, 0)


def clip_and_adam(params, grads, st: AdamState, lr, max_grad_norm, sem: Semantics,
                  b1=0.9, b2=0.999, eps=1e-8):
    """optax ``clip_by_global_norm`` -> ``adam`` (jaxrl/reppo.py:246-257) or torch
    ``clip_grad_norm_`` -> ``AdamW(weight_decay=0.01)`` (torchrl/reppo.py:270-273,527-534).
    Returns (new_params, new_state, grad_norm)."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)
    if max_grad_norm is None:
        scale = torch.ones_like(gn)
    elif sem.name == "jax":
        scale = torch.where(gn < max_grad_norm, torch.ones_like(gn), max_grad_norm / gn)
    else:
        scale = torch.clamp(max_grad_norm / (gn + sem.clip_eps), max=1.0)
    t = st.t + 1
    bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t
    new_p, new_m, new_v = {}, {}, {}
    for k, p in params.items():
        g = grads[k] * scale
        m = b1 * st.m[k] + (1 - b1) * g
        v = b2 * st.v[k] + (1 - b2) * g * g
        p2 = p * (1.0 - lr * sem.adamw_weight_decay) if sem.adamw_weight_decay else p
        new_p[k] = p2 - lr * (m / bc1) / (torch.sqrt(v / bc2) + eps)
        new_m[k], new_v[k] = m, v
    return new_p, AdamState(new_m, new_v, t), gn


def _value_and_grad(fn, params):
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    loss, metrics = fn(leaves)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(leaves, grads)}
    return loss.detach(), metrics, grads


# ----------------------------------------------------------------------------------------
# one minibatch step and the whole update
# ----------------------------------------------------------------------------------------
@dataclass
class TrainState:
    critic: Dict[str, Tensor]
    actor: Dict[str, Tensor]
    actor_old: Dict[str, Tensor]
    critic_opt: AdamState
    actor_opt: AdamState

    @staticmethod
    def create(critic, actor):
        return TrainState(critic, actor, {k: v.clone() for k, v in actor.items()},
                          AdamState.zeros_like(critic), AdamState.zeros_like(actor))


def critic_step(ts: TrainState, mb, hp, sem, lr=None):
    lr = hp.lr if lr is None else lr
    loss, metrics, grads = _value_and_grad(lambda p: critic_loss(p, mb, hp, sem), ts.critic)
    new_p, new_opt, gn = clip_and_adam(ts.critic, grads, ts.critic_opt, lr, hp.max_grad_norm, sem)
    metrics["grad_norm"] = gn
    return replace(ts, critic=new_p, critic_opt=new_opt), metrics, grads


def actor_step(ts: TrainState, mb, eps_pi, eps_kl, hp, sem, lr=None):
    lr = hp.lr if lr is None else lr
    loss, metrics, grads = _value_and_grad(
        lambda p: actor_loss(p, ts.actor_old, ts.critic, mb, eps_pi, eps_kl, hp, sem), ts.actor)
    new_p, new_opt, gn = clip_and_adam(ts.actor, grads, ts.actor_opt, lr, hp.max_grad_norm, sem)
    metrics["grad_norm"] = gn
    return replace(ts, actor=new_p, actor_opt=new_opt), metrics, grads


def compute_targets(batch: Dict[str, Tensor], hp: Hyper, sem: Semantics) -> Tensor:
    """``[T, N]`` targets.  Torch convention mutates ``batch['truncated'][-1]`` (:178)."""
    if sem.td_convention == "jax":
        return td_lambda_jax(batch["soft_reward"], batch["value"], batch["done"], batch["truncated"],
                             batch.get("importance_weight"), hp.gamma, hp.lmbda)
    return td_lambda_torch(batch["soft_reward"], batch["done"], batch["truncated"], batch["value"],
                           hp.gamma, hp.lmbda)


def flatten_batch(batch: Dict[str, Tensor], target: Tensor):
    """``[T, N, ...] -> [T*N, ...]``, row index ``t*N + n`` (jaxrl/reppo.py:452-455;
    torchrl/reppo.py:219)."""
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in batch.items() if v is not None}
    flat["target"] = target.reshape(-1)
    return flat


def take(flat, idx):
    return {k: v[idx] for k, v in flat.items()}


def learn_step(ts: TrainState, batch: Dict[str, Tensor], perms: Tensor, eps_pi: Tensor,
               eps_kl: Tensor, hp: Hyper, sem: Semantics, max_steps: Optional[int] = None,
               trace: bool = False):
    """The whole update (jaxrl/reppo.py:417-673; torchrl/reppo.py:614-632).

    batch: ``[T, N, ...]`` tensors obs, critic_obs, action, reward, soft_reward, value,
    next_emb, done, truncated (all float), importance_weight (optional).
    perms ``[epochs, S]`` int64, eps_pi ``[epochs, mb, A]``, eps_kl ``[epochs, 16, mb, A]`` --
    injected, because JAX threefry cannot be matched; the noise of an epoch is reused by
    every minibatch of that epoch (late-bound ``key``, :535,557).
    Returns (new_state, metrics of the last epoch averaged over minibatches, target_values,
    optional list of per-step metric dicts).
    """
    target = compute_targets(batch, hp, sem)
    flat = flatten_batch(batch, target)
    ts = replace(ts, actor_old={k: v.clone() for k, v in ts.actor.items()})  # :457-461
    mbs = hp.minibatch_size
    steps = 0
    traces = []
    last_epoch = []
    for e in range(hp.num_epochs):
        last_epoch = []
        for j in range(hp.num_mini_batches):
            if max_steps is not None and steps >= max_steps:
                break
            idx = perms[e, j * mbs:(j + 1) * mbs]
            mb = take(flat, idx)
            ts, cm, _ = critic_step(ts, mb, hp, sem)
            ts, am, _ = actor_step(ts, mb, eps_pi[e], eps_kl[e], hp, sem)
            m = {**{f"critic/{k}": v for k, v in cm.items()}, **{f"actor/{k}": v for k, v in am.items()}}
            last_epoch.append(m)
            if trace:
                traces.append(m)
            steps += 1
    metrics = {}
    if last_epoch:
        for k in last_epoch[0]:
            metrics[k] = torch.stack([torch.as_tensor(m[k]) for m in last_epoch]).mean(0)
    return ts, metrics, target, traces


# ----------------------------------------------------------------------------------------
# synthetic rollouts -- shared by tests and bench so every arm sees one input
# ----------------------------------------------------------------------------------------
def synthetic_rollout(hp: Hyper, seed: int = 0, terminate: bool = False, reward_scaling: float = 1.0,
                      iw_variant: bool = False, dtype=torch.float32, num_envs: Optional[int] = None,
                      env_offset: int = 0):
    """Seeded CPU generation of one ``[T, N, ...]`` rollout batch.  ``num_envs`` /
    ``env_offset`` select an env-column shard of the full batch (every rank regenerates the
    full fields and slices, so shards are consistent across GPU counts)."""
    T, N = hp.num_steps, hp.num_envs
    H, A = hp.critic_hidden_dim, hp.action_dim
    g = torch.Generator().manual_seed(seed)
    R = hp.vmax - hp.vmin
    b = {}
    b["obs"] = torch.randn(T, N, hp.obs_dim, generator=g)
    if hp.critic_obs_dim == hp.obs_dim:
        b["critic_obs"] = b["obs"].clone()
    else:
        b["critic_obs"] = torch.randn(T, N, hp.critic_obs_dim, generator=g)
    b["action"] = torch.clamp(torch.tanh(torch.randn(T, N, A, generator=g)), -0.999, 0.999)
    b["reward"] = torch.rand(T, N, generator=g) * reward_scaling
    b["value"] = hp.vmin + R * (0.05 + 0.55 * torch.rand(T, N, generator=g))
    trunc = (torch.rand(T, N, generator=g) < 1e-3).float()
    done = (torch.rand(T, N, generator=g) < 0.002).float() if terminate else torch.zeros(T, N)
    b["truncated"] = trunc
    b["done"] = torch.clamp(done + trunc, max=1.0)
    if iw_variant:
        b["importance_weight"] = torch.log(0.5 + 0.5 * torch.rand(T, N, generator=g))
    else:
        _ = torch.rand(T, N, generator=g)
        b["importance_weight"] = torch.zeros(T, N)
    ell = torch.randn(T, N, generator=g) - A / 2.0
    b["soft_reward"] = b["reward"] - hp.gamma * 0.01 * ell
    b["next_emb"] = torch.randn(T, N, H, generator=g)
    if num_envs is not None:
        b = {k: v[:, env_offset:env_offset + num_envs].contiguous() for k, v in b.items()}
    return {k: v.to(dtype) for k, v in b.items()}


def synthetic_noise(hp: Hyper, seed: int = 1, dtype=torch.float32, mb: Optional[int] = None):
    g = torch.Generator().manual_seed(seed)
    mb = hp.minibatch_size if mb is None else mb
    eps_pi = torch.randn(hp.num_epochs, mb, hp.action_dim, generator=g)
    eps_kl = torch.randn(hp.num_epochs, KL_SAMPLES, mb, hp.action_dim, generator=g)
    return eps_pi.to(dtype), eps_kl.to(dtype)


def synthetic_perms(hp: Hyper, seed: int = 2, size: Optional[int] = None):
    g = torch.Generator().manual_seed(seed)
    S = hp.batch_size if size is None else size
    return torch.stack([torch.randperm(S, generator=g) for _ in range(hp.num_epochs)])


def cast_state(ts: TrainState, dtype):
    c = lambda d: {k: v.to(dtype) for k, v in d.items()}
    return TrainState(c(ts.critic), c(ts.actor), c(ts.actor_old),
                      AdamState(c(ts.critic_opt.m), c(ts.critic_opt.v), ts.critic_opt.t),
                      AdamState(c(ts.actor_opt.m), c(ts.actor_opt.v), ts.actor_opt.t))
