"""Torch-facing façade: the reference's entry points for the update path, backed by the B200 engine.

Mirrors (same names, argument meaning, return keys, error behaviour):

* ``Actor`` / ``Critic``            -- src/networks/torch_models.py:209-335 (constructor arguments, forward
                                      tuple order, ``state_dict`` keys)
* ``TrainState``                    -- src/torchrl/reppo.py:46-58
* ``make_postprocess_fn``           -- src/torchrl/reppo.py:173-221 (TD(lambda) scan + flatten)
* ``make_critic_update_fn``         -- src/torchrl/reppo.py:224-285
* ``make_actor_update_fn``          -- src/torchrl/reppo.py:288-360
* ``reppo_update``                  -- the whole per-rollout update: src/jaxrl/reppo.py:417-673 (learn_step) /
                                      the loop body src/torchrl/reppo.py:614-632

The modules' parameters are views into the engine's flat arenas; forward and update run the
hand-written CUDA kernels through the C ABI.  ``TensorDict`` is not installed here, so batches are
``TensorBatch`` (a dict of tensors sharing leading batch dimensions) or any mapping of tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Any, Dict, Mapping, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .engine import EngineConfig, HyperParams, ReppoEngine


# -------------------------------------------------------------------------------------------------
# minimal TensorDict stand-in
# -------------------------------------------------------------------------------------------------
class TensorBatch(dict):
    """Mapping of tensors with common leading dims: supports the operations the reference loop uses
    (``data[indices]``, slicing, ``.contiguous()``, ``.float()``, ``.flatten(0, 1)``, ``.detach()``, ``.to``)."""

    def __init__(self, data: Mapping[str, torch.Tensor], batch_size=None, device=None):
        super().__init__(data)
        self.batch_size = tuple(batch_size) if batch_size is not None else None

    def _map(self, fn):
        return TensorBatch({k: fn(v) for k, v in self.items()})

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        return self._map(lambda v: v[key])

    def contiguous(self):
        return self._map(lambda v: v.contiguous())

    def float(self):
        return self._map(lambda v: v.float())

    def detach(self):
        return self._map(lambda v: v.detach())

    def flatten(self, a=0, b=1):
        return self._map(lambda v: v.flatten(a, b))

    def to(self, *args, **kw):
        return self._map(lambda v: v.to(*args, **kw))


# -------------------------------------------------------------------------------------------------
# modules
# -------------------------------------------------------------------------------------------------
def _attach(root: nn.Module, dotted: str, param: nn.Parameter):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], param)


class _ArenaModule(nn.Module):
    """nn.Module whose parameters are views into one flat fp32 tensor (the engine arena once bound)."""

    def _build(self, layout: Dict[str, tuple], device):
        n = sum(int(math.prod(s)) if s else 1 for s in layout.values())
        self._layout = {}
        off = 0
        flat = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=device)
        for k, shp in layout.items():
            cnt = int(math.prod(shp)) if shp else 1
            self._layout[k] = (off, tuple(shp))
            _attach(self, k, nn.Parameter(flat[off:off + cnt].view(shp), requires_grad=True))
            off += cnt
        self._flat = flat
        self._engine: Optional[ReppoEngine] = None

    def _rebind(self, flat: torch.Tensor, grads: Optional[torch.Tensor] = None):
        """Point every parameter at `flat` (same layout), keeping the current values."""
        flat[: self._flat.numel()].copy_(self._flat.to(flat.device))
        named = dict(self.named_parameters())
        for k, (off, shp) in self._layout.items():
            cnt = int(math.prod(shp)) if shp else 1
            named[k].data = flat[off:off + cnt].view(shp)
            if grads is not None:
                named[k].grad = grads[off:off + cnt].view(shp)
        self._flat = flat

    def flat_parameters(self) -> torch.Tensor:
        return self._flat


def _fcnn_layout(prefix, in_f, hidden, out_f, layers, input_activation, use_norm, out=None):
    out = {} if out is None else out
    shift = 1 if input_activation else 0
    for i in range(layers):
        fi = in_f if i == 0 else hidden
        fo = out_f if i == layers - 1 else hidden
        out[f"{prefix}.net.{shift + i}.0.weight"] = (fo, fi)
        out[f"{prefix}.net.{shift + i}.0.bias"] = (fo,)
        if use_norm and i < layers - 1:
            out[f"{prefix}.net.{shift + i}.1.weight"] = (fo,)
    return out


def _init_linear_(module: _ArenaModule, gen=None):
    """nn.Linear's default init (kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for weight and bias),
    RMSNorm scale 1 -- what the reference modules get from torch (torch_models.py:71-79)."""
    named = dict(module.named_parameters())
    for k, p in named.items():
        if k.endswith(".0.weight"):
            bound = 1.0 / math.sqrt(p.shape[1])
            with torch.no_grad():
                p.uniform_(-bound, bound, generator=gen)
                named[k[:-6] + "bias"].uniform_(-bound, bound, generator=gen)
        elif k.endswith(".1.weight"):
            with torch.no_grad():
                p.fill_(1.0)


class Critic(_ArenaModule):
    """Categorical HL-Gauss critic (src/networks/torch_models.py:209-285).  ``forward`` returns plain tensors computed by the
    library (no autograd graph; see ``Actor``): gradients come from ``make_critic_update_fn`` / ``reppo_update`` only."""

    def __init__(self, n_obs, n_act, num_atoms: int, vmin: float, vmax: float, hidden_dim=256, use_norm=True,
                 use_encoder_norm=False, encoder_layers=1, head_layers=1, pred_layers=1, device=None,
                 semantics: str = "torch"):
        super().__init__()
        if use_encoder_norm:
            raise NotImplementedError("use_encoder_norm=True is never set by the reference (src/torchrl/reppo.py:510-523)")
        if min(encoder_layers, head_layers, pred_layers) < 2:
            raise NotImplementedError("layers == 1 has divergent JAX/Torch semantics (SURVEY.md App. B)")
        self.n_obs, self.n_act, self.num_atoms, self.vmin, self.vmax = n_obs, n_act, num_atoms, float(vmin), float(vmax)
        self.hidden_dim, self.use_norm = hidden_dim, use_norm
        self.encoder_layers, self.head_layers, self.pred_layers = encoder_layers, head_layers, pred_layers
        self.semantics = semantics
        pred_out = hidden_dim + (1 if semantics == "jax" else 0)
        layout = {"zero_dist": (num_atoms,)}
        _fcnn_layout("feature_module", n_obs + n_act, hidden_dim, hidden_dim, encoder_layers, False, use_norm, layout)
        _fcnn_layout("critic_module", hidden_dim, hidden_dim, num_atoms, head_layers, True, use_norm, layout)
        _fcnn_layout("pred_module", hidden_dim, hidden_dim, pred_out, pred_layers, True, use_norm, layout)
        self._build(layout, device)
        _init_linear_(self)
        self.values = torch.linspace(vmin, vmax, num_atoms, device=device, dtype=torch.float32)
        with torch.no_grad():
            self.zero_dist.copy_(hl_gauss(torch.zeros(1, 1), vmin, vmax, num_atoms, torch_variant=(semantics == "torch")).reshape(-1))

    def forward(self, obs, action):
        eng = _engine_of(self)
        return eng.critic_forward(obs, action, params=self._flat)


class Actor(_ArenaModule):
    """Tanh-Gaussian policy with temperature / Lagrange multipliers (src/networks/torch_models.py:288-335).

    ``forward`` is computed by the library and returns plain tensors: there is NO autograd graph behind it (and the
    temperature / Lagrange multiplier come back detached), so a loss built from it cannot be back-propagated.  Gradients
    exist only through ``make_actor_update_fn`` / ``reppo_update``, which write them into the ``.grad`` views of these
    parameters."""

    def __init__(self, n_obs, n_act, ent_start: float, kl_start: float, hidden_dim=256, use_norm=True, layers=2,
                 min_std=0.1, device=None):
        super().__init__()
        if layers < 2:
            raise NotImplementedError("layers == 1 has divergent JAX/Torch semantics (SURVEY.md App. B)")
        self.n_obs, self.n_act, self.hidden_dim, self.use_norm, self.layers = n_obs, n_act, hidden_dim, use_norm, layers
        self.min_std = float(min_std)
        layout = {"log_temp": (), "log_lagrange": ()}
        _fcnn_layout("model", n_obs, hidden_dim, 2 * n_act, layers, False, use_norm, layout)
        self._build(layout, device)
        _init_linear_(self)
        with torch.no_grad():
            self.log_temp.fill_(math.log(ent_start))
            self.log_lagrange.fill_(math.log(kl_start))

    def forward(self, obs):
        eng = _engine_of(self)
        mean, std = eng.actor_forward(obs, min_std=self.min_std, params=self._flat)
        pi = torch.distributions.TransformedDistribution(
            torch.distributions.Normal(mean, std, validate_args=False), [torch.distributions.TanhTransform()])
        return pi, torch.tanh(mean), torch.exp(self.log_temp.detach()), torch.exp(self.log_lagrange.detach())


def hl_gauss(inp, vmin, vmax, num_atoms, torch_variant=True):
    """Host helper used only to INITIALISE ``zero_dist`` (torch_models.py:268-276 / jax_models.py:288-290);
    the training-time histogram is computed inside the fused CUDA loss kernel."""
    x = torch.clip(inp, vmin, vmax)
    bw = (vmax - vmin) / (num_atoms - 1)
    support = torch.linspace(vmin - bw / 2, vmax + bw / 2, num_atoms + 1)
    sigma = bw * 0.75
    e1, e2 = (1e-6, 1e-6) if torch_variant else (0.0, 0.0)
    cdf = torch.erf((support.unsqueeze(0) - x) / (math.sqrt(2.0) * sigma + e1))
    z = cdf[..., -1] - cdf[..., 0]
    return (cdf[..., 1:] - cdf[..., :-1]) / (z.unsqueeze(-1) + e2)


class ArenaOptimizer:
    """Stands where the reference keeps ``optim.AdamW`` (src/torchrl/reppo.py:527-534): the Adam moments live in
    the engine arenas and the step is the fused CUDA kernel; this object only carries ``lr``."""

    def __init__(self, lr: float):
        self.param_groups = [{"lr": float(lr)}]

    def zero_grad(self, set_to_none: bool = True):
        pass


@dataclass
class TrainState:
    """Field-compatible with src/torchrl/reppo.py:46-58."""
    device: Any = None
    obs: Any = None
    critic_obs: Any = None
    actor: Optional[Actor] = None
    old_actor: Optional[Actor] = None
    critic: Optional[Critic] = None
    normalizer: Any = None
    critic_normalizer: Any = None
    actor_optimizer: Any = None
    critic_optimizer: Any = None
    scaler: Any = None
    engine: Optional[ReppoEngine] = None

    def compile(self):  # the reference compiles modules with torch.compile; nothing to do here
        pass


def _engine_of(module: _ArenaModule) -> ReppoEngine:
    if module._engine is None:
        raise RuntimeError("module is not bound to a ReppoEngine: build it with make_train_state(...) or bind_engine(...)")
    return module._engine


def bind_engine(actor: Actor, critic: Critic, old_actor: Optional[Actor] = None, *, max_rows: int = 2048,
                semantics: Optional[str] = None, norm_type: str = "rms", gemm_mode: str = "fp32",
                device=None, max_batch_rows: int = 0) -> ReppoEngine:
    """Create the engine for (actor, critic) and move their parameters into its arenas."""
    semantics = semantics or critic.semantics
    ecfg = EngineConfig(
        obs_dim=actor.n_obs, action_dim=actor.n_act, critic_obs_dim=critic.n_obs, critic_hidden_dim=critic.hidden_dim,
        actor_hidden_dim=actor.hidden_dim, num_bins=critic.num_atoms, vmin=critic.vmin, vmax=critic.vmax,
        num_critic_encoder_layers=critic.encoder_layers, num_critic_head_layers=critic.head_layers,
        num_critic_pred_layers=critic.pred_layers, num_actor_layers=actor.layers, use_critic_norm=critic.use_norm,
        use_actor_norm=actor.use_norm, norm_type=norm_type, semantics=semantics, gemm_mode=gemm_mode, max_rows=max_rows,
        max_batch_rows=max_batch_rows)
    eng = ReppoEngine(ecfg, device=device)
    for mod, net in ((critic, eng.critic), (actor, eng.actor)):
        assert {k: v[1] for k, v in mod._layout.items()} == {k: tuple(v[1]) for k, v in net.layout.items()}, "layout mismatch"
        mod._rebind(net.params, net.grads)
        mod._engine = eng
    if old_actor is not None:
        old_actor._rebind(eng.actor_old)
        old_actor._engine = eng
    else:
        eng.sync_actor_old()

    def _regrad():   # dist.attach_multicast moved the gradient arenas into symmetric memory: re-point the .grad views
        critic._rebind(eng.critic.params, eng.critic.grads)
        actor._rebind(eng.actor.params, eng.actor.grads)
    eng.__dict__.setdefault("_grads_moved", []).append(_regrad)
    return eng


def make_train_state(cfg, n_obs: int, n_critic_obs: int, n_act: int, device=None, max_rows: Optional[int] = None) -> TrainState:
    """Networks + optimizers exactly as ``main`` builds them (src/torchrl/reppo.py:500-555), bound to one engine."""
    h = cfg.hyperparameters
    b200 = cfg.get("b200", {}) if hasattr(cfg, "get") else {}
    semantics = b200.get("semantics", "torch")
    device = torch.device(device if device is not None else f"cuda:{cfg.platform.get('device_rank', 0)}")
    actor = Actor(n_obs=n_obs, n_act=n_act, ent_start=h.ent_start, kl_start=h.kl_start, hidden_dim=h.actor_hidden_dim,
                  use_norm=h.use_actor_norm, layers=h.num_actor_layers, min_std=h.actor_min_std)
    critic = Critic(n_obs=n_critic_obs, n_act=n_act, num_atoms=h.num_bins, vmin=h.vmin, vmax=h.vmax,
                    hidden_dim=h.critic_hidden_dim, use_norm=h.use_critic_norm, use_encoder_norm=False,
                    encoder_layers=h.num_critic_encoder_layers, head_layers=h.num_critic_head_layers,
                    pred_layers=h.num_critic_pred_layers, semantics=semantics)
    old_actor = Actor(n_obs=n_obs, n_act=n_act, ent_start=h.ent_start, kl_start=h.kl_start, hidden_dim=h.actor_hidden_dim,
                      use_norm=h.use_actor_norm, layers=h.num_actor_layers, min_std=h.actor_min_std)
    old_actor.load_state_dict(actor.state_dict())
    mb = (int(h.num_steps) * int(h.num_envs)) // int(h.num_mini_batches)
    # platform.amp_dtype: bf16 -> tcgen05 on bf16 operands; tf32 -> tcgen05 on the fp32 operands (what
    # torch.set_float32_matmul_precision("high") asks of the reference's matmuls); anything else -> fp32 FFMA ("highest")
    amp = str(cfg.platform.get("amp_dtype", "f32"))
    gemm_mode = amp if amp in ("bf16", "tf32") else "fp32"
    # the workspace serves the minibatch steps AND the per-environment-step forwards of the collection loop (num_envs rows)
    eng = bind_engine(actor, critic, old_actor, max_rows=max_rows or max(mb, int(h.num_envs)), semantics=semantics,
                      norm_type=b200.get("norm_type", "rms"), gemm_mode=gemm_mode, device=device,
                      max_batch_rows=mb * int(h.num_mini_batches))
    return TrainState(device=device, actor=actor, old_actor=old_actor, critic=critic, normalizer=nn.Identity(),
                      critic_normalizer=nn.Identity(), actor_optimizer=ArenaOptimizer(h.lr),
                      critic_optimizer=ArenaOptimizer(h.lr), scaler=None, engine=eng)


def _hyper(cfg, train_state: TrainState, which: str = "critic") -> HyperParams:
    h = cfg.hyperparameters
    opt = train_state.critic_optimizer if which == "critic" else train_state.actor_optimizer
    lr = float(opt.param_groups[0]["lr"]) if opt is not None and hasattr(opt, "param_groups") else float(h.lr)
    return HyperParams(
        gamma=float(h.gamma), lmbda=float(h.lmbda), lr=lr, max_grad_norm=float(h.max_grad_norm), kl_bound=float(h.kl_bound),
        ent_target_mult=float(h.ent_target_mult), aux_loss_mult=float(h.aux_loss_mult),
        actor_min_std=float(train_state.actor.min_std), actor_kl_clip_mode=str(h.actor_kl_clip_mode),
        reverse_kl=bool(h.get("reverse_kl", False)), reduce_kl=bool(h.get("reduce_kl", True)), update_entropy_lagrangian=bool(h.get("update_entropy_lagrangian", True)),
        update_kl_lagrangian=bool(h.get("update_kl_lagrangian", True)),
        partial_reset=bool(cfg.env.get("partial_reset", False)), lr_anneal_steps=_anneal_steps(cfg))


def _anneal_steps(cfg) -> int:
    """cfg.anneal_lr: optax.linear_schedule(lr, 0, num_updates) with num_updates = (total_time_steps // (num_steps *
    num_envs)) * num_epochs * num_mini_batches (src/jaxrl/reppo.py:239-244); evaluated on the device at each network's
    optimizer step counter (optim.cu), so the captured graph of reppo_update is reused across updates."""
    h = cfg.hyperparameters
    if not bool(h.get("anneal_lr", False)):
        return 0
    iters = int(h.total_time_steps) // (int(h.num_steps) * int(h.num_envs))
    return max(1, iters * int(h.num_epochs) * int(h.num_mini_batches))


def _engine_for(cfg, train_state: TrainState) -> ReppoEngine:
    if train_state.engine is None:
        h = cfg.hyperparameters
        mb = (int(h.num_steps) * int(h.num_envs)) // int(h.num_mini_batches)
        train_state.engine = bind_engine(train_state.actor, train_state.critic, train_state.old_actor,
                                         max_rows=max(mb, int(h.num_envs)), device=train_state.device,
                                         max_batch_rows=mb * int(h.num_mini_batches))
    return train_state.engine


# -------------------------------------------------------------------------------------------------
# rollout side: observation normalisation, collection loop, checkpoints
# -------------------------------------------------------------------------------------------------
class EmpiricalNormalization(nn.Module):
    """Drop-in for src/torchrl/reppo_util.py:394-471 (same constructor, buffers ``_mean/_var/_std/count`` and
    ``state_dict`` keys); the Chan update and the normalisation run as two CUDA kernels (rollout.cu)."""

    def __init__(self, shape, device, eps=1e-2, until=None):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("reppo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.eps, self.until = eps, until
        self.device = torch.device(device)
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        self.register_buffer("_mean", torch.zeros(shape).unsqueeze(0).to(self.device))
        self.register_buffer("_var", torch.ones(shape).unsqueeze(0).to(self.device))
        self.register_buffer("_std", torch.ones(shape).unsqueeze(0).to(self.device))
        self.register_buffer("count", torch.tensor(0, dtype=torch.long).to(self.device))
        self._lib = L.load()

    @property
    def mean(self):
        return self._mean.squeeze(0).clone()

    @property
    def std(self):
        return self._std.squeeze(0).clone()

    def forward(self, x: torch.Tensor, center: bool = True) -> torch.Tensor:
        if x.shape[-1:] != self._mean.shape[-1:]:
            raise ValueError(f"Expected input of shape (*,{self._mean.shape[-1:]}), got {x.shape}")
        xf = x.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(xf)
        dim = xf.shape[-1]
        rows = xf.numel() // dim
        L.check(self._lib.reppo_obs_normalize(
            xf.data_ptr(), rows, dim, self._mean.data_ptr(), self._var.data_ptr(), self._std.data_ptr(), self.count.data_ptr(),
            float(self.eps), int(self.training), int(center), -1 if self.until is None else int(self.until), out.data_ptr(),
            torch.cuda.current_stream(self.device).cuda_stream), None, "reppo_obs_normalize")
        return out

    def update(self, x):
        self.training, was = True, self.training
        try:
            self.forward(x)
        finally:
            self.training = was

    def inverse(self, y):
        return y * (self._std + self.eps) + self._mean


def make_collect_fn(cfg, env, generator: Optional[torch.Generator] = None):
    """Collection loop of src/torchrl/reppo.py:89-171 over the engine's rollout kernels: per environment step one
    policy launch group (normalise -> actor -> sample + log-prob) and one bootstrap group (normalise -> actor -> next
    action + log-prob -> critic -> next value / embedding -> soft reward).  ``env`` needs ``step(actions) -> (next_obs,
    rewards, dones, truncations, infos)``, ``num_envs`` and ``asymmetric_obs`` like the reference's wrappers."""
    h = cfg.hyperparameters
    asymmetric_obs = getattr(env, "asymmetric_obs", False)

    def collect_fn(train_state: TrainState):
        eng = _engine_for(cfg, train_state)
        dev = eng.device
        transitions, info_list = [], []
        obs, critic_obs = train_state.obs, train_state.critic_obs
        A = eng.cfg.action_dim
        for _ in range(int(h.num_steps)):
            norm_obs = train_state.normalizer(obs)
            norm_critic_obs = train_state.critic_normalizer(critic_obs)
            eps = torch.randn(norm_obs.shape[0], A, device=dev, generator=generator)
            actions, log_probs, _ = eng.policy_act(norm_obs, eps, clip=0.999, min_std=float(train_state.actor.min_std))
            next_obs, rewards, dones, truncations, infos = env.step(actions)
            next_critic_obs = infos["observations"]["critic"] if asymmetric_obs else next_obs
            if cfg.env.get("has_final_obs", False) and cfg.env.get("partial_reset", False) and "final_observation" in infos:
                _next_obs = infos["final_observation"]
                _next_critic_obs = _next_obs
            else:
                _next_obs, _next_critic_obs = next_obs, next_critic_obs
            eps2 = torch.randn(norm_obs.shape[0], A, device=dev, generator=generator)
            b = eng.bootstrap(train_state.normalizer(_next_obs), train_state.critic_normalizer(_next_critic_obs), eps2, rewards,
                              float(h.gamma), clip=1.0 - 1e-6, min_std=float(train_state.actor.min_std))
            transitions.append({
                "observations": norm_obs, "critic_observations": norm_critic_obs, "actions": actions, "log_probs": log_probs,
                "rewards": b["soft_reward"].unsqueeze(-1), "raw_rewards": rewards.to(dev, torch.float32).reshape(-1, 1),
                "next_embeddings": b["next_emb"], "next_values": b["next_value"].unsqueeze(-1),
                "dones": dones.unsqueeze(-1).float(), "truncations": truncations.unsqueeze(-1).float()})
            info_list.append(infos)
            obs, critic_obs = next_obs, next_critic_obs
        train_state = replace(train_state, obs=obs, critic_obs=critic_obs)
        stacked = TensorBatch({k: torch.stack([t[k] for t in transitions], dim=0) for k in transitions[0]},
                              batch_size=(int(h.num_steps), transitions[0]["actions"].shape[0]))
        return train_state, stacked, info_list

    return collect_fn


def _cpu_state(sd):
    return {k: v.detach().to("cpu", non_blocking=True) for k, v in sd.items()}


def save_params(global_step, actor, qnet, qnet_target, obs_normalizer, critic_obs_normalizer, args, save_path):
    """Checkpoint with the reference's layout and keys (src/torchrl/reppo_util.py:723-753): the façade modules'
    ``state_dict()`` keys are the reference modules' keys, so the file loads into either implementation."""
    import os
    os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
    save_dict = {
        "actor_state_dict": _cpu_state(actor.state_dict()),
        "qnet_state_dict": _cpu_state(qnet.state_dict()),
        "qnet_target_state_dict": _cpu_state(qnet_target.state_dict()) if qnet_target is not None else None,
        "obs_normalizer_state": _cpu_state(obs_normalizer.state_dict()) if hasattr(obs_normalizer, "state_dict") else None,
        "critic_obs_normalizer_state": (_cpu_state(critic_obs_normalizer.state_dict())
                                        if hasattr(critic_obs_normalizer, "state_dict") else None),
        "args": dict(vars(args)) if hasattr(args, "__dict__") else dict(args),
        "global_step": global_step,
    }
    torch.save(save_dict, save_path, _use_new_zipfile_serialization=True)


def load_params(path, actor, qnet, obs_normalizer=None, critic_obs_normalizer=None, old_actor=None):
    """Inverse of ``save_params``: copies the tensors into the engine arenas through the façade modules."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    actor.load_state_dict(ck["actor_state_dict"])
    qnet.load_state_dict(ck["qnet_state_dict"])
    if old_actor is not None:
        old_actor.load_state_dict(ck["actor_state_dict"])
    for mod, key in ((obs_normalizer, "obs_normalizer_state"), (critic_obs_normalizer, "critic_obs_normalizer_state")):
        if mod is not None and ck.get(key) and hasattr(mod, "load_state_dict") and len(ck[key]) > 0:
            mod.load_state_dict(ck[key])
    return ck.get("global_step", 0)


# -------------------------------------------------------------------------------------------------
# the three factory closures
# -------------------------------------------------------------------------------------------------
def make_postprocess_fn(cfg, env=None):
    """TD(lambda) targets + flatten (src/torchrl/reppo.py:173-221).  Like the reference's ``compute_gve`` the Torch
    convention writes ``truncations[-1] = 1`` into the transition."""

    def postprocess(train_state: TrainState, transition: Mapping[str, torch.Tensor]):
        eng = _engine_for(cfg, train_state)
        h = cfg.hyperparameters
        T, N = int(h.num_steps), int(h.num_envs)
        dev = eng.device

        def tn(x):  # [T, N, 1] or [T, N] -> contiguous fp32 [T, N] on the device (a view when already so)
            x = x.to(dev)
            x = x.reshape(T, N)
            return x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()

        trunc_in = transition["truncations"]
        trunc = tn(trunc_in)
        conv = eng.cfg.semantics
        iw = transition.get("importance_weight") if hasattr(transition, "get") else None
        gve = eng.td_lambda(tn(transition["rewards"]), tn(transition["next_values"]), tn(transition["dones"]), trunc,
                            tn(iw) if iw is not None else None, float(h.gamma), float(h.lmbda), convention=conv)
        if conv == "torch" and trunc.data_ptr() != trunc_in.data_ptr():
            trunc_in[-1] = 1.0  # the kernel wrote into a converted copy: reproduce the in-place mutation (:178)
        data = TensorBatch({k: transition[k] for k in (
            "observations", "critic_observations", "actions", "rewards", "next_embeddings", "next_values", "dones",
            "truncations") if k in transition})
        data["truncations"] = trunc.reshape(T, N, 1)
        data["gve"] = gve.reshape(T, N, 1)
        if "raw_rewards" in transition:
            data["raw_rewards"] = transition["raw_rewards"]
        return data.to(dev).float().flatten(0, 1).detach()

    return postprocess


def _mini_batch(eng: ReppoEngine, data: Mapping[str, torch.Tensor]) -> L.Batch:
    flat = {
        "obs": data["observations"], "critic_obs": data["critic_observations"], "action": data["actions"],
        "done": data["dones"].reshape(-1), "truncated": data["truncations"].reshape(-1),
        "next_emb": data["next_embeddings"], "target": data["gve"].reshape(-1),
        "reward": data["raw_rewards"].reshape(-1) if "raw_rewards" in data else
        (data["rewards"].reshape(-1) if eng.cfg.semantics == "jax" else None),
    }
    return eng.make_batch(flat)


def make_critic_update_fn(cfg, train_state: TrainState):
    """``update(data) -> logs`` with the reference's keys (src/torchrl/reppo.py:224-285)."""
    eng = _engine_for(cfg, train_state)

    def update(data: Mapping[str, torch.Tensor]):
        rows = data["gve"].shape[0]
        m = eng.critic_step(_mini_batch(eng, data), rows, _hyper(cfg, train_state, "critic")).clone()
        i = L.METRIC_NAMES.index
        return {"critic_grad_norm": m[i("critic_grad_norm")], "qf_loss": m[i("critic_loss")], "qf_max": m[i("target_max")],
                "qf_min": m[i("target_min")], "qf_mean": m[i("target_mean")], "embedding_loss": m[i("aux_mean")]}

    return update


def make_actor_update_fn(cfg, train_state: TrainState, generator: Optional[torch.Generator] = None):
    """``update(data) -> logs`` (src/torchrl/reppo.py:288-360).  Fresh standard-normal noise per call from torch's
    device generator (the reference draws it inside ``rsample`` / ``sample((16,))``)."""
    eng = _engine_for(cfg, train_state)
    A, ks = eng.cfg.action_dim, eng.cfg.kl_samples

    def update(data: Mapping[str, torch.Tensor], eps_pi: Optional[torch.Tensor] = None, eps_kl: Optional[torch.Tensor] = None):
        rows = data["gve"].shape[0]
        if eps_pi is None:
            eps_pi = torch.randn(rows, A, device=eng.device, generator=generator)
        if eps_kl is None:
            eps_kl = torch.randn(ks, rows, A, device=eng.device, generator=generator)
        m = eng.actor_step(_mini_batch(eng, data), rows, eps_pi, eps_kl, _hyper(cfg, train_state, "actor")).clone()
        i = L.METRIC_NAMES.index
        return {"actor_grad_norm": m[i("actor_grad_norm")], "actor_loss": m[i("actor_loss")], "kl": m[i("kl")],
                "entropy": m[i("entropy")], "temperature": m[i("temperature")], "lagrangian": m[i("lagrangian")],
                "entropy_loss": m[i("entropy_loss")], "lagrangian_loss": m[i("lagrangian_loss")]}

    return update


# -------------------------------------------------------------------------------------------------
# the whole update in one call
# -------------------------------------------------------------------------------------------------
def _stage_dict(eng, index: int) -> dict:
    """Device staging buffers of an engine: sets 0 / 1 hold a rollout each, -1 the per-update scratch."""
    return eng.__dict__.setdefault("_stages", {}).setdefault(index, {})


def _rollout_fields(eng, transition, T: int, N: int):
    """(staging name, reference key, shape) of every field of a ``[T, N, ...]`` rollout that the update reads."""
    conv = eng.cfg.semantics
    fields = [("obs", "observations", (T, N, eng.cfg.obs_dim)),
              ("critic_obs", "critic_observations", (T, N, eng._shape.critic_obs_dim)),
              ("action", "actions", (T, N, eng.cfg.action_dim)), ("soft_reward", "rewards", (T, N)),
              ("next_emb", "next_embeddings", (T, N, eng.cfg.critic_hidden_dim)), ("value", "next_values", (T, N)),
              ("done", "dones", (T, N)), ("truncated", "truncations", (T, N))]
    if "raw_rewards" in transition:
        fields.append(("reward", "raw_rewards", (T, N)))
    if "importance_weight" in transition and conv == "jax":
        fields.append(("iw", "importance_weight", (T, N)))
    return fields


def _noise_shapes(eng, cfg):
    h = cfg.hyperparameters
    T, N, n_mb, E = int(h.num_steps), int(h.num_envs), int(h.num_mini_batches), int(h.num_epochs)
    S = T * N
    mb = S // n_mb
    A, ks = eng.cfg.action_dim, eng.cfg.kl_samples
    per_mb = eng.cfg.semantics == "torch"   # fresh draw per minibatch (Torch) / one draw per epoch (JAX)
    return (E, S), ((E, n_mb, mb, A) if per_mb else (E, mb, A)), ((E, n_mb, ks, mb, A) if per_mb else (E, ks, mb, A))


def _stage_buf(d: dict, name: str, shape, dtype, dev):
    t = d.get(name)
    if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
        t = d[name] = torch.empty(shape, dtype=dtype, device=dev)
    return t


def _draw_plan(stage: dict, eng, cfg, generator, dev):
    """Minibatch permutations and actor noise of one update, drawn on the current stream into `stage`."""
    ps, s1, s2 = _noise_shapes(eng, cfg)
    pbuf = _stage_buf(stage, "perms", ps, torch.int32, dev)
    for e in range(ps[0]):
        pbuf[e].copy_(torch.randperm(ps[1], device=dev, generator=generator))
    _stage_buf(stage, "eps_pi", s1, torch.float32, dev).normal_(generator=generator)
    _stage_buf(stage, "eps_kl", s2, torch.float32, dev).normal_(generator=generator)


def reppo_prefetch(train_state: TrainState, transition: Mapping[str, torch.Tensor], cfg, *,
                   generator: Optional[torch.Generator] = None, draw_plan: bool = True) -> None:
    """Start the host-to-device copy of the NEXT rollout while the current update is still running.

    The copy goes into the staging set that the running update does not read, on a dedicated copy stream, after that
    set's previous update has finished; ``reppo_update`` called with the same ``transition`` object then skips its own
    copies and waits for the copy's event instead.  With ``draw_plan`` the minibatch permutations and the actor noise of
    that update are drawn on the copy stream as well (pass the generator the update calls use).  (The reference keeps rollouts on the device, src/torchrl/reppo.py:
    89-171; this is the input pipeline for hosts whose simulator is not on the GPU.)"""
    eng = _engine_for(cfg, train_state)
    h = cfg.hyperparameters
    T, N = int(h.num_steps), int(h.num_envs)
    dev = eng.device
    si = 1 - eng.__dict__.get("_stage_set", 0)
    stage = _stage_dict(eng, si)
    cs = eng.__dict__.get("_copy_stream")
    if cs is None:
        cs = eng._copy_stream = torch.cuda.Stream(device=dev)
    free = eng.__dict__.get("_stage_free", {}).get(si)
    if free is not None:
        cs.wait_event(free)
    with torch.cuda.stream(cs):
        for name, key, shape in _rollout_fields(eng, transition, T, N):
            t = stage.get(name)
            if t is None or tuple(t.shape) != tuple(shape):
                t = stage[name] = torch.empty(shape, dtype=torch.float32, device=dev)
            t.copy_(transition[key].reshape(shape), non_blocking=True)
        if draw_plan:
            # the next update's permutations and noise are drawn here too (same generator, same order as the plain calls)
            _draw_plan(stage, eng, cfg, generator, dev)
        ev = torch.cuda.Event()
        ev.record(cs)
    eng._prefetched = {"src": transition, "set": si, "event": ev, "plan": bool(draw_plan)}


def reppo_update(train_state: TrainState, transition: Mapping[str, torch.Tensor], cfg, *,
                 perms: Optional[torch.Tensor] = None, eps_pi: Optional[torch.Tensor] = None,
                 eps_kl: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
                 use_graph: Optional[bool] = None, return_device_metrics: bool = False, world_size: int = 1,
                 prefetch_next: Optional[Mapping[str, torch.Tensor]] = None):
    """One full REPPO update over a rollout batch = ``learn_step`` (src/jaxrl/reppo.py:417-673) /
    lines 614-632 of src/torchrl/reppo.py: TD(lambda) scan, ``num_epochs`` x ``num_mini_batches`` critic+actor
    steps on shuffled minibatches, then ``old_actor <- actor``.

    ``transition``: ``[T, N, ...]`` tensors (host or device) with the reference's keys.  ``perms [E, S]``,
    ``eps_pi`` and ``eps_kl`` may be injected (parity tests); by default they are drawn on the device:
    JAX semantics reuses one noise draw per epoch, Torch semantics draws per minibatch.
    Returns the metrics of the last epoch averaged over its minibatches, as a dict of Python floats.

    ``prefetch_next``: the NEXT rollout (host tensors, ideally pinned).  Its host-to-device copy is issued on a copy
    stream into the idle staging set as soon as this update is enqueued, so it overlaps this update's kernels; the next
    ``reppo_update`` call that is given that same mapping finds it staged (see ``reppo_prefetch``).
    """
    eng = _engine_for(cfg, train_state)
    h = cfg.hyperparameters
    T, N, n_mb, E = int(h.num_steps), int(h.num_envs), int(h.num_mini_batches), int(h.num_epochs)
    S = T * N
    mb = S // n_mb
    dev = eng.device
    A, ks = eng.cfg.action_dim, eng.cfg.kl_samples
    conv = eng.cfg.semantics

    # Persistent device staging buffers: the rollout is copied INTO them (one async H2D per field from pinned host
    # memory), so every pointer the engine sees is stable across calls and the captured CUDA graph is replayed
    # instead of being re-captured.  There are TWO staging sets: ``reppo_prefetch`` (or ``prefetch_next=``) copies the
    # next rollout into the idle set on a copy stream while this update runs; the engine keeps one graph per set.
    pf = eng.__dict__.get("_prefetched")
    have_plan = False
    if pf is not None and pf["src"] is transition:
        si, staged = pf["set"], True
        have_plan = pf.get("plan", False) and perms is None and eps_pi is None and eps_kl is None
        torch.cuda.current_stream(dev).wait_event(pf["event"])
        eng._prefetched = None
    else:
        si, staged = eng.__dict__.get("_stage_set", 0), False
    eng._stage_set = si
    stage = _stage_dict(eng, si)
    shared = _stage_dict(eng, -1)   # per-update scratch that is not part of the rollout (same pointers for both sets)

    def buf(name, shape, dtype=torch.float32, where=None):
        d = stage if where is None else where
        t = d.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=dev)
            d[name] = t
        return t

    def put(name, src, shape):
        dst = buf(name, shape)
        if not staged:
            dst.copy_(src.reshape(shape), non_blocking=True)
        return dst

    fields = _rollout_fields(eng, transition, T, N)
    placed = {name: put(name, transition[key], shape) for name, key, shape in fields}
    obs, cobs, act, soft = placed["obs"], placed["critic_obs"], placed["action"], placed["soft_reward"]
    nemb, val, done, trunc = placed["next_emb"], placed["value"], placed["done"], placed["truncated"]
    raw = placed.get("reward", soft if conv == "jax" else None)
    iw = placed.get("iw")
    target = buf("target", (T, N))
    eng.td_lambda(soft, val, done, trunc, iw, float(h.gamma), float(h.lmbda), convention=conv, out=target)
    if conv == "torch":
        transition["truncations"][-1] = 1.0  # compute_gve writes this into the caller's batch (src/torchrl/reppo.py:178)
    flat = {"obs": obs.view(S, -1), "critic_obs": cobs.view(S, -1), "action": act.view(S, -1), "done": done.view(S),
            "truncated": trunc.view(S), "next_emb": nemb.view(S, -1), "target": target.view(S),
            "reward": raw.view(S) if raw is not None else None}
    batch = eng.make_batch(flat)
    if conv == "jax":
        eng.sync_actor_old()  # actor_target <- actor at the START of learn_step (src/jaxrl/reppo.py:457-464)
    per_mb = conv == "torch"
    # permutations and noise live in the staging set (one captured graph per set): already drawn by reppo_prefetch,
    # injected by the caller (parity tests), or drawn here
    pbuf = buf("perms", (E, S), torch.int32)
    e1 = buf("eps_pi", (E, n_mb, mb, A) if per_mb else (E, mb, A))
    e2 = buf("eps_kl", (E, n_mb, ks, mb, A) if per_mb else (E, ks, mb, A))
    if not have_plan:
        if perms is None:
            for e in range(E):
                pbuf[e].copy_(torch.randperm(S, device=dev, generator=generator))
        else:
            pbuf.copy_(perms.to(torch.int32), non_blocking=True)
        if eps_pi is None:
            e1.normal_(generator=generator)
        else:
            e1.copy_(eps_pi.reshape(e1.shape), non_blocking=True)
        if eps_kl is None:
            e2.normal_(generator=generator)
        else:
            e2.copy_(eps_kl.reshape(e2.shape), non_blocking=True)
    b200 = cfg.get("b200", {}) if hasattr(cfg, "get") else {}
    graph = bool(b200.get("use_graph", True)) if use_graph is None else use_graph
    hyp = _hyper(cfg, train_state)
    hyp.world_size = world_size  # data-parallel ranks sharing every minibatch (gradients all-reduced by the engine)
    m, sm = eng.update(batch, pbuf, e1, e2, hyp, n_mb, use_graph=graph, step_metrics=buf("step_metrics", (E * n_mb, L.NUM_METRICS), where=shared))
    eng.sync_actor_old()  # src/torchrl/reppo.py:631-632 (same state as the JAX ordering at the next call)
    # this staging set is free again once everything enqueued so far has run
    free = eng.__dict__.setdefault("_stage_free", {})
    ev = free.get(si)
    if ev is None:
        ev = free[si] = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    if prefetch_next is not None:
        reppo_prefetch(train_state, prefetch_next, cfg, generator=generator)
    if return_device_metrics:
        return m, sm
    if world_size > 1:
        # per-rank metrics are partial sums over this rank's rows (already divided by world * mb): one all-reduce per update
        from . import dist as _dp
        m = _dp.reduce_metrics(m)
    md = eng.metrics_dict(m)  # one device->host read per update
    logs = {  # Torch keys (src/torchrl/reppo.py:275-283,348-358)
        "critic_grad_norm": md["critic_grad_norm"], "qf_loss": md["critic_loss"], "qf_max": md["target_max"],
        "qf_min": md["target_min"], "qf_mean": md["target_mean"], "embedding_loss": md["aux_mean"],
        "actor_grad_norm": md["actor_grad_norm"], "actor_loss": md["actor_loss"], "kl": md["kl"], "entropy": md["entropy"],
        "temperature": md["temperature"], "lagrangian": md["lagrangian"], "entropy_loss": md["entropy_loss"],
        "lagrangian_loss": md["lagrangian_loss"],
        # JAX keys (src/jaxrl/reppo.py:514-524,609-622)
        "value_loss": md["value_loss"], "loss": md["actor_loss"], "q": md["q_mean"], "temp": md["temperature"],
        "abs_pred_action": md["abs_pred_action"], "reward_mean": md["reward_mean"], "target_values": md["target_mean"],
        "aux_loss": md["aux_mean"],
    }
    return logs
