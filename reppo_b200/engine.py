"""Host-side engine: PyTorch allocates the flat arenas / workspace, the C ABI does the work.

``ReppoEngine`` is the thin object the Torch-facing façade (``reppo_b200.torch_api``) and the
tests drive.  It owns, per network, one flat fp32 arena each for parameters, gradients and the
two Adam moments (layout = the reference ``state_dict`` order), a device step
counter, a copy of the behaviour-policy parameters (``actor_old`` = ``actor_target``,
src/jaxrl/reppo.py:457-464) and the kernel workspace.  All compute goes through
``libreppo_b200.so``; nothing here has a CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Mapping, Optional

import torch

from . import _lib as L


@dataclass
class EngineConfig:
    obs_dim: int
    action_dim: int
    critic_obs_dim: Optional[int] = None
    critic_hidden_dim: int = 512
    actor_hidden_dim: int = 512
    num_bins: int = 151
    vmin: float = 0.0
    vmax: float = 150.0
    num_critic_encoder_layers: int = 2
    num_critic_head_layers: int = 2
    num_critic_pred_layers: int = 2
    num_actor_layers: int = 3
    use_critic_norm: bool = True
    use_actor_norm: bool = True
    norm_type: str = "rms"          # "rms" (jax_models.py:46) | "layer" (reppo_mj_playground.py:65-72)
    semantics: str = "jax"          # "jax" parity target | "torch" API semantics
    gemm_mode: str = "fp32"         # "fp32" (FFMA) | "bf16" (tcgen05, bf16 operands) | "tf32" (tcgen05 on the fp32 operands)
    max_rows: int = 2048
    kl_samples: int = 16
    max_batch_rows: int = 0         # rows per epoch of `update` (n_mb * mb): enables the per-epoch pre-gather (0 = per-step gathers)

    def shape(self) -> L.Shape:
        cod = self.critic_obs_dim if self.critic_obs_dim is not None else self.obs_dim
        norm = L.NORMS[self.norm_type]
        return L.Shape(
            obs_dim=self.obs_dim, critic_obs_dim=cod, action_dim=self.action_dim,
            critic_hidden=self.critic_hidden_dim, actor_hidden=self.actor_hidden_dim, num_bins=self.num_bins,
            enc_layers=self.num_critic_encoder_layers, head_layers=self.num_critic_head_layers,
            pred_layers=self.num_critic_pred_layers, actor_layers=self.num_actor_layers,
            critic_norm=norm if self.use_critic_norm else L.NORM_NONE,
            actor_norm=norm if self.use_actor_norm else L.NORM_NONE,
            semantics=L.SEM_JAX if self.semantics == "jax" else L.SEM_TORCH,
            gemm_mode=L.GEMM_MODES[self.gemm_mode],
            max_rows=self.max_rows, kl_samples=self.kl_samples, vmin=self.vmin, vmax=self.vmax,
            max_batch_rows=self.max_batch_rows)


@dataclass
class HyperParams:
    """The slice of config/reppo.yaml:16-70 the update consumes."""
    gamma: float = 0.99
    lmbda: float = 0.95
    lr: float = 3e-4
    max_grad_norm: float = 0.5
    kl_bound: float = 0.1
    ent_target_mult: float = 0.5
    aux_loss_mult: float = 1.0
    actor_min_std: float = 0.0
    actor_kl_clip_mode: str = "clipped"
    reduce_kl: bool = True
    update_entropy_lagrangian: bool = True
    update_kl_lagrangian: bool = True
    partial_reset: bool = False
    world_size: int = 1
    reverse_kl: bool = False        # cfg.reverse_kl (src/jaxrl/reppo.py:543-553): KL(pi || pi_old) from samples of the current policy
    lr_anneal_steps: int = 0        # cfg.anneal_lr: optimizer steps over which lr decays linearly to 0 (src/jaxrl/reppo.py:239-244)

    def c_struct(self) -> L.Hyper:
        if self.actor_kl_clip_mode not in L.KL_MODES:
            raise ValueError(f"Unknown actor loss mode: {self.actor_kl_clip_mode}")  # src/jaxrl/reppo.py:586-588
        return L.Hyper(
            gamma=self.gamma, lmbda=self.lmbda, lr=self.lr,
            max_grad_norm=self.max_grad_norm if self.max_grad_norm is not None else 0.0,
            kl_bound=self.kl_bound, ent_target_mult=self.ent_target_mult, aux_loss_mult=self.aux_loss_mult,
            actor_min_std=self.actor_min_std, kl_clip_mode=L.KL_MODES[self.actor_kl_clip_mode],
            reduce_kl=int(self.reduce_kl), update_entropy_lagrangian=int(self.update_entropy_lagrangian),
            update_kl_lagrangian=int(self.update_kl_lagrangian), partial_reset=int(self.partial_reset),
            world_size=self.world_size, lr_anneal_steps=int(self.lr_anneal_steps), reverse_kl=int(self.reverse_kl))


class _Net:
    """One network's arenas + named views."""

    def __init__(self, lib, ctx, net: int, device):
        self.count = int(lib.reppo_param_count(ctx, net))
        n_alloc = (self.count + 3) // 4 * 4
        z = lambda: torch.zeros(n_alloc, dtype=torch.float32, device=device)
        self.params, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z()
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.layout: Dict[str, tuple] = {}
        name = C.create_string_buffer(128)
        off, rows, cols = C.c_int64(), C.c_int32(), C.c_int32()
        for i in range(lib.reppo_param_num_tensors(ctx, net)):
            L.check(lib.reppo_param_tensor(ctx, net, i, name, 128, C.byref(off), C.byref(rows), C.byref(cols)), ctx,
                    "reppo_param_tensor")
            shp = () if rows.value == 0 else ((rows.value,) if cols.value == 0 else (rows.value, cols.value))
            self.layout[name.value.decode()] = (off.value, shp)

    def view(self, arena: torch.Tensor, key: str) -> torch.Tensor:
        off, shp = self.layout[key]
        n = 1
        for s in shp:
            n *= s
        return arena[off:off + n].view(shp)

    def as_dict(self, arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: self.view(arena, k) for k in self.layout}

    def c_state(self) -> L.NetState:
        return L.NetState(self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                          self.adam_v.data_ptr(), self.step.data_ptr())


class ReppoEngine:
    def __init__(self, cfg: EngineConfig, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("reppo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = L.load()
        self._shape = cfg.shape()
        ctx = C.c_void_p()
        L.check(self.lib.reppo_ctx_create(C.byref(self._shape), C.byref(ctx)), None, "reppo_ctx_create")
        self.ctx = ctx
        with torch.cuda.device(self.device):
            self.critic = _Net(self.lib, ctx, L.NET_CRITIC, self.device)
            self.actor = _Net(self.lib, ctx, L.NET_ACTOR, self.device)
            self.actor_old = torch.zeros_like(self.actor.params)
            nbytes = int(self.lib.reppo_workspace_bytes(ctx))
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            L.check(self.lib.reppo_set_workspace(ctx, base, nbytes), ctx, "reppo_set_workspace")
            # bin edges / support exactly as the reference builds them (torch.linspace, fp32)
            bw = (cfg.vmax - cfg.vmin) / (cfg.num_bins - 1)
            self._edges = torch.linspace(cfg.vmin - bw / 2, cfg.vmax + bw / 2, cfg.num_bins + 1, dtype=torch.float32)
            self._support = torch.linspace(cfg.vmin, cfg.vmax, cfg.num_bins, dtype=torch.float32)
            L.check(self.lib.reppo_set_bins(ctx, self._edges.data_ptr(), self._support.data_ptr()), ctx, "reppo_set_bins")
            self.metrics = torch.zeros(L.NUM_METRICS, dtype=torch.float32, device=self.device)
        self.pred_dim = cfg.critic_hidden_dim + (1 if cfg.semantics == "jax" else 0)
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.reppo_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- plumbing ---------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _f32(self, t: torch.Tensor) -> torch.Tensor:
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.float32).contiguous()
        return t

    def c_state(self) -> L.State:
        return L.State(self.critic.c_state(), self.actor.c_state(), self.actor_old.data_ptr())

    def load_params(self, critic: Optional[Mapping[str, torch.Tensor]] = None,
                    actor: Optional[Mapping[str, torch.Tensor]] = None, sync_old: bool = True):
        for net, src in ((self.critic, critic), (self.actor, actor)):
            if src is None:
                continue
            missing = set(net.layout) - set(src)
            extra = set(src) - set(net.layout)
            if missing or extra:
                raise KeyError(f"parameter keys mismatch: missing={sorted(missing)} unexpected={sorted(extra)}")
            for k in net.layout:
                net.view(net.params, k).copy_(src[k].to(self.device, torch.float32))
        if actor is not None and sync_old:
            self.sync_actor_old()

    def sync_actor_old(self):
        """actor_target <- actor (src/jaxrl/reppo.py:457-464; src/torchrl/reppo.py:631-632)."""
        self.actor_old.copy_(self.actor.params)

    def reset_optimizer(self):
        for net in (self.critic, self.actor):
            net.adam_m.zero_(); net.adam_v.zero_(); net.step.zero_()

    def make_batch(self, flat: Mapping[str, torch.Tensor]) -> L.Batch:
        """flat: obs, critic_obs, action, reward, done, truncated, next_emb, target -- [S, ...] tensors."""
        keep = {k: self._f32(flat[k]) for k in ("obs", "critic_obs", "action", "done", "truncated", "next_emb", "target")}
        keep["reward"] = self._f32(flat["reward"]) if flat.get("reward") is not None else None
        self._keep = [keep]
        S = keep["target"].numel()
        ptr = lambda t: t.data_ptr() if t is not None else None
        return L.Batch(ptr(keep["obs"]), ptr(keep["critic_obs"]), ptr(keep["action"]), ptr(keep["reward"]),
                       ptr(keep["done"]), ptr(keep["truncated"]), ptr(keep["next_emb"]), ptr(keep["target"]), S)

    # ---- K1 ---------------------------------------------------------------------------------
    def td_lambda(self, soft_reward, value, done, truncated, importance_weight=None, gamma=0.99, lmbda=0.95,
                  convention: Optional[str] = None, mutate_truncated: bool = True, out=None) -> torch.Tensor:
        """[T, N] fp32 device tensors -> TD(lambda) targets [T, N]."""
        conv = convention or self.cfg.semantics
        T, N = soft_reward.shape
        for t in (soft_reward, value, done, truncated):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (T, N)
        out = torch.empty_like(soft_reward) if out is None else out
        iw = importance_weight if (importance_weight is not None and conv == "jax") else None
        L.check(self.lib.reppo_td_lambda(
            soft_reward.data_ptr(), value.data_ptr(), done.data_ptr(), truncated.data_ptr(),
            iw.data_ptr() if iw is not None else None, out.data_ptr(), T, N, float(gamma), float(lmbda),
            L.SEM_JAX if conv == "jax" else L.SEM_TORCH, int(mutate_truncated), self._stream()), None, "reppo_td_lambda")
        return out

    # ---- forwards ---------------------------------------------------------------------------
    def critic_forward(self, critic_obs, action, params: Optional[torch.Tensor] = None):
        co, ac = self._f32(critic_obs), self._f32(action)
        rows = co.shape[0]
        dev = self.device
        value = torch.empty(rows, device=dev); logits = torch.empty(rows, self.cfg.num_bins, device=dev)
        pred = torch.empty(rows, self.pred_dim, device=dev); feat = torch.empty(rows, self.cfg.critic_hidden_dim, device=dev)
        p = self.critic.params if params is None else params
        L.check(self.lib.reppo_critic_forward(self.ctx, p.data_ptr(), co.data_ptr(), ac.data_ptr(), rows, value.data_ptr(),
                                              logits.data_ptr(), pred.data_ptr(), feat.data_ptr(), self._stream()),
                self.ctx, "reppo_critic_forward")
        return value, logits, pred, feat

    def actor_forward(self, obs, min_std: float = 0.0, params: Optional[torch.Tensor] = None):
        ob = self._f32(obs)
        rows = ob.shape[0]
        mean = torch.empty(rows, self.cfg.action_dim, device=self.device); std = torch.empty_like(mean)
        p = self.actor.params if params is None else params
        L.check(self.lib.reppo_actor_forward(self.ctx, p.data_ptr(), ob.data_ptr(), rows, float(min_std), mean.data_ptr(),
                                             std.data_ptr(), self._stream()), self.ctx, "reppo_actor_forward")
        return mean, std

    def linear(self, op: int, a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One Linear layer's GEMM in isolation (op 0 forward y = x w^T + b, 1 dgrad, 2 wgrad); see reppo_linear."""
        if op == 0:
            rows, in_f = a.shape; out_f = b.shape[0]; shp = (rows, out_f)
        elif op == 1:
            rows, out_f = a.shape; in_f = b.shape[1]; shp = (rows, in_f)
        else:
            rows, out_f = a.shape; in_f = b.shape[1]; shp = (out_f, in_f)
        if out is None:
            out = torch.zeros(shp, device=self.device, dtype=torch.float32)
        elif op == 2:
            out.zero_()
        L.check(self.lib.reppo_linear(self.ctx, op, rows, in_f, out_f, a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                      bias.data_ptr() if bias is not None else None, self._stream()), self.ctx, "reppo_linear")
        return out

    def kernel_probe(self, which: int, in0, in1, in2, in3, out0, scratch, hyper: "HyperParams") -> None:
        """One loss / sampling kernel in isolation at any row count (reppo_kernel_probe): 0 = HL-Gauss cross-entropy,
        1 = next-embedding regression, 2 = tanh-Gaussian sample + KL.  All tensors are caller-allocated fp32 CUDA."""
        import ctypes as C
        rows = int(in0.shape[0])
        hs = hyper.c_struct()
        L.check(self.lib.reppo_kernel_probe(self.ctx, which, rows, in0.data_ptr(), in1.data_ptr(), in2.data_ptr(), in3.data_ptr(),
                                            out0.data_ptr(), scratch.data_ptr(), C.byref(hs), self._stream()), self.ctx,
                "reppo_kernel_probe")

    # ---- rollout side ------------------------------------------------------------
    def policy_act(self, obs, eps, scale: Optional[torch.Tensor] = None, clip: float = 0.999, store_clipped: bool = False,
                   min_std: float = 0.0, params: Optional[torch.Tensor] = None, want_scaled: bool = False):
        """Policy step of the collection loop (src/torchrl/reppo.py:104-109,146; src/jaxrl/reppo.py:347-364):
        action = tanh(mu + std * scale * eps), log-prob of the action clipped to +-clip (clip <= 0: at the pre-tanh
        sample).  Returns (action [rows, A], log_prob [rows], log_prob under the scaled policy or None)."""
        ob, e = self._f32(obs), self._f32(eps)
        rows = ob.shape[0]
        action = torch.empty(rows, self.cfg.action_dim, device=self.device)
        logp = torch.empty(rows, device=self.device)
        logps = torch.empty(rows, device=self.device) if (want_scaled or scale is not None) else None
        sc = self._f32(scale).reshape(-1) if scale is not None else None
        p = self.actor.params if params is None else params
        L.check(self.lib.reppo_policy_act(self.ctx, p.data_ptr(), ob.data_ptr(), e.data_ptr(), sc.data_ptr() if sc is not None else None,
                                          rows, float(min_std), float(clip), int(store_clipped), action.data_ptr(), logp.data_ptr(),
                                          logps.data_ptr() if logps is not None else None, self._stream()),
                self.ctx, "reppo_policy_act")
        return action, logp, logps

    def bootstrap(self, next_obs, next_critic_obs, eps, reward, gamma: float, clip: float = 1.0 - 1e-6, min_std: float = 0.0):
        """Bootstrap half of the collection step (src/torchrl/reppo.py:117-138; src/jaxrl/reppo.py:367-374)."""
        ob, co, e, r = self._f32(next_obs), self._f32(next_critic_obs), self._f32(eps), self._f32(reward).reshape(-1)
        rows = ob.shape[0]
        dev = self.device
        out = {"next_value": torch.empty(rows, device=dev), "next_emb": torch.empty(rows, self.cfg.critic_hidden_dim, device=dev),
               "soft_reward": torch.empty(rows, device=dev), "next_action": torch.empty(rows, self.cfg.action_dim, device=dev),
               "next_log_prob": torch.empty(rows, device=dev)}
        L.check(self.lib.reppo_bootstrap(self.ctx, self.actor.params.data_ptr(), self.critic.params.data_ptr(), ob.data_ptr(),
                                         co.data_ptr(), e.data_ptr(), r.data_ptr(), rows, float(gamma), float(min_std), float(clip),
                                         out["next_value"].data_ptr(), out["next_emb"].data_ptr(), out["soft_reward"].data_ptr(),
                                         out["next_action"].data_ptr(), out["next_log_prob"].data_ptr(), self._stream()),
                self.ctx, "reppo_bootstrap")
        return out

    # ---- steps ------------------------------------------------------------------------------
    def _idx(self, idx):
        if isinstance(idx, int):  # identity gather over the first `idx` rows of the batch
            return None, idx
        t = idx.to(device=self.device, dtype=torch.int32).contiguous()
        return t, t.numel()

    def critic_grad(self, batch: L.Batch, idx, hp: HyperParams, step: bool = False, metrics=None):
        idx, rows = self._idx(idx)
        m = self.metrics if metrics is None else metrics
        st, h = self.c_state(), hp.c_struct()
        fn = self.lib.reppo_critic_step if step else self.lib.reppo_critic_grad
        L.check(fn(self.ctx, C.byref(st), C.byref(batch), idx.data_ptr() if idx is not None else None, rows, C.byref(h), m.data_ptr(),
                   self._stream()), self.ctx, "reppo_critic_step" if step else "reppo_critic_grad")
        return m

    def actor_grad(self, batch: L.Batch, idx, eps_pi, eps_kl, hp: HyperParams, step: bool = False, metrics=None):
        idx, rows = self._idx(idx)
        e1, e2 = self._f32(eps_pi), self._f32(eps_kl)
        m = self.metrics if metrics is None else metrics
        st, h = self.c_state(), hp.c_struct()
        fn = self.lib.reppo_actor_step if step else self.lib.reppo_actor_grad
        L.check(fn(self.ctx, C.byref(st), C.byref(batch), idx.data_ptr() if idx is not None else None, rows, e1.data_ptr(), e2.data_ptr(),
                   C.byref(h), m.data_ptr(), self._stream()), self.ctx, "reppo_actor_step" if step else "reppo_actor_grad")
        return m

    def critic_step(self, batch, idx, hp, metrics=None):
        return self.critic_grad(batch, idx, hp, step=True, metrics=metrics)

    def actor_step(self, batch, idx, eps_pi, eps_kl, hp, metrics=None):
        return self.actor_grad(batch, idx, eps_pi, eps_kl, hp, step=True, metrics=metrics)

    def update(self, batch: L.Batch, perms: torch.Tensor, eps_pi: torch.Tensor, eps_kl: torch.Tensor, hp: HyperParams,
               num_mini_batches: int, use_graph: bool = False, step_metrics: Optional[torch.Tensor] = None,
               metrics: Optional[torch.Tensor] = None):
        """All epochs x minibatches (src/jaxrl/reppo.py:643-673).  perms [E, n_mb*mb] int32, eps_pi [E, mb, A],
        eps_kl [E, 16, mb, A] device tensors (kept alive by the caller while the stream work runs)."""
        E = perms.shape[0]
        mb = perms.shape[1] // num_mini_batches
        A, ks = self.cfg.action_dim, self.cfg.kl_samples
        assert perms.dtype == torch.int32 and perms.is_cuda and perms.is_contiguous()
        assert eps_pi.is_cuda and eps_kl.is_cuda and eps_pi.is_contiguous() and eps_kl.is_contiguous()
        per_mb = eps_pi.dim() == 4  # [E, n_mb, mb, A]: fresh noise for every step (Torch semantics)
        if per_mb:
            assert tuple(eps_pi.shape) == (E, num_mini_batches, mb, A), eps_pi.shape
            assert tuple(eps_kl.shape) == (E, num_mini_batches, ks, mb, A), eps_kl.shape
        else:
            assert tuple(eps_pi.shape) == (E, mb, A), eps_pi.shape
            assert tuple(eps_kl.shape) == (E, ks, mb, A), eps_kl.shape
        if step_metrics is None:
            step_metrics = torch.empty(E * num_mini_batches, L.NUM_METRICS, device=self.device)
        m = self.metrics if metrics is None else metrics
        S = int(batch.num_rows)
        if getattr(self, "_old_out", None) is None or self._old_out.shape[0] != S:
            self._old_out = torch.empty(S, 2 * A, device=self.device, dtype=torch.float32)
        plan = L.Plan(E, num_mini_batches, mb, perms.data_ptr(), eps_pi.data_ptr(), eps_kl.data_ptr(),
                      step_metrics.data_ptr(), m.data_ptr(), int(use_graph), int(per_mb), self._old_out.data_ptr())
        st, h = self.c_state(), hp.c_struct()
        L.check(self.lib.reppo_update(self.ctx, C.byref(st), C.byref(batch), C.byref(plan), C.byref(h), self._stream()),
                self.ctx, "reppo_update")
        return m, step_metrics

    def launch_count(self) -> int:
        return int(self.lib.reppo_launch_count(self.ctx))

    @staticmethod
    def metrics_dict(m: torch.Tensor) -> Dict[str, float]:
        v = m.detach().float().cpu().tolist()
        return {k: v[i] for i, k in enumerate(L.METRIC_NAMES)}
