"""Oracle B -- the reference's OWN Torch closures (TEST INFRASTRUCTURE ONLY), imported from /root/reference in the
build container and, where that does not exist (the GPU box), from the byte-compiled copy ``oracle/build_ref.py``
leaves under ``oracle/_ref``.

``src/torchrl/reppo.py`` imports tensordict / hydra / omegaconf / torchinfo, none of which is
installed here; ``oracle/_stubs`` supplies minimal stand-ins so that the reference's
``make_postprocess_fn`` (:173-221), ``make_critic_update_fn`` (:224-285),
``make_actor_update_fn`` (:288-360), ``Actor`` / ``Critic`` (torch_models.py:209-335) and
``hl_gauss`` (reppo_util.py:756-777) run unmodified on the CPU.  ``oracle/gen_golden.py`` uses
this to produce the committed fixtures under ``tests/golden/``.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "_stubs")


def reference_root():
    """/root/reference (sources, build container) or oracle/_ref (sourceless .pyc, travels to the GPU box); None if neither."""
    if os.path.isdir(os.path.join("/root/reference", "src", "torchrl")):
        return "/root/reference"
    ref = os.path.join(_HERE, "_ref")
    if os.path.exists(os.path.join(ref, "src", "torchrl", "reppo.pyc.bin")):
        return ref
    return None


class _CompiledRefImporter:
    """Imports ``src.*`` from the byte-compiled files oracle/build_ref.py wrote (``<module>.pyc.bin`` = a .pyc under a suffix
    the gpurun snapshot does not skip)."""

    def __init__(self, root):
        self.root = root

    def _path(self, fullname):
        base = os.path.join(self.root, *fullname.split("."))
        if os.path.exists(os.path.join(base, "__init__.pyc.bin")):
            return os.path.join(base, "__init__.pyc.bin"), True
        if os.path.exists(base + ".pyc.bin"):
            return base + ".pyc.bin", False
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        import importlib.util
        if fullname != "src" and not fullname.startswith("src."):
            return None
        file, is_pkg = self._path(fullname)
        if file is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, self, origin=file, is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [os.path.dirname(file)]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        with open(module.__spec__.origin, "rb") as f:
            code = marshal.loads(f.read()[16:])   # skip the .pyc header (magic, flags, hash)
        module.__file__ = module.__spec__.origin
        exec(code, module.__dict__)


REFERENCE_ROOT = reference_root()


def available() -> bool:
    return reference_root() is not None


def load_reference():
    """Import the reference Torch modules with the stubs on the path."""
    root = reference_root()
    if root is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref (python -m oracle.build_ref) is present -- Oracle B cannot run")
    if root == "/root/reference":
        paths = (_STUBS, root)
    else:
        paths = (_STUBS,)
        if not any(isinstance(f, _CompiledRefImporter) for f in sys.meta_path):
            sys.meta_path.insert(0, _CompiledRefImporter(root))
    for p in paths:
        if p not in sys.path:
            sys.path.insert(0, p)
    # src/torchrl/envs.py pulls simulator wrappers lazily inside make_envs; importing is safe
    omp = os.environ.get("OMP_NUM_THREADS")
    import src.torchrl.reppo as ref_reppo  # noqa: E402  (sets OMP_NUM_THREADS=1 at import, :37)
    import src.networks.torch_models as ref_models  # noqa: E402
    import src.torchrl.reppo_util as ref_util  # noqa: E402
    if omp is None:
        os.environ.pop("OMP_NUM_THREADS", None)
    else:
        os.environ["OMP_NUM_THREADS"] = omp
    torch.set_float32_matmul_precision("highest")  # undo :35 for CPU determinism (no effect on CPU math)
    return types.SimpleNamespace(reppo=ref_reppo, models=ref_models, util=ref_util)


def make_cfg(hp, partial_reset: bool = False):
    """Attr-dict with the fields the closures read (cfg.hyperparameters.*, cfg.env.get, cfg.platform.*)."""
    from omegaconf import DictConfig

    h = DictConfig(
        num_steps=hp.num_steps, num_envs=hp.num_envs, num_mini_batches=hp.num_mini_batches,
        num_epochs=hp.num_epochs, gamma=hp.gamma, lmbda=hp.lmbda, lr=hp.lr,
        max_grad_norm=hp.max_grad_norm, vmin=hp.vmin, vmax=hp.vmax, num_bins=hp.num_bins,
        kl_bound=hp.kl_bound, actor_kl_clip_mode=hp.actor_kl_clip_mode,
        ent_target_mult=hp.ent_target_mult, aux_loss_mult=hp.aux_loss_mult,
    )
    env = DictConfig(partial_reset=partial_reset)
    platform = DictConfig(amp_enabled=False, cuda=False, amp_dtype="f32", amp_device="cpu")
    return DictConfig(hyperparameters=h, env=env, platform=platform)


def build_train_state(ref, hp, critic_params, actor_params):
    """Reference ``Actor``/``Critic``/``AdamW`` exactly as ``main`` builds them (:500-555),
    with injected weights."""
    dev = torch.device("cpu")
    actor = ref.models.Actor(
        n_obs=hp.obs_dim, n_act=hp.action_dim, ent_start=hp.ent_start, kl_start=hp.kl_start,
        hidden_dim=hp.actor_hidden_dim, use_norm=hp.use_actor_norm, layers=hp.num_actor_layers,
        min_std=hp.actor_min_std, device=dev)
    critic = ref.models.Critic(
        n_obs=hp.critic_obs_dim, n_act=hp.action_dim, num_atoms=hp.num_bins, vmin=hp.vmin, vmax=hp.vmax,
        hidden_dim=hp.critic_hidden_dim, use_norm=hp.use_critic_norm, use_encoder_norm=False,
        encoder_layers=hp.num_critic_encoder_layers, head_layers=hp.num_critic_head_layers,
        pred_layers=hp.num_critic_pred_layers, device=dev)
    actor.load_state_dict({k: v.clone() for k, v in actor_params.items()})
    critic.load_state_dict({k: v.clone() for k, v in critic_params.items()})
    import copy
    old_actor = copy.deepcopy(actor)
    q_opt = torch.optim.AdamW(list(critic.parameters()), lr=torch.tensor(hp.lr))
    a_opt = torch.optim.AdamW(list(actor.parameters()), lr=torch.tensor(hp.lr))
    scaler = torch.amp.GradScaler(enabled=False)
    return ref.reppo.TrainState(
        device=dev, obs=None, critic_obs=None, actor=actor, old_actor=old_actor, critic=critic,
        normalizer=torch.nn.Identity(), critic_normalizer=torch.nn.Identity(),
        actor_optimizer=a_opt, critic_optimizer=q_opt, scaler=scaler)


@contextlib.contextmanager
def injected_noise(eps_pi, eps_kl):
    """Replay fixed standard-normal draws through ``Normal.rsample`` (policy sample,
    torchrl/reppo.py:300) and ``Normal.sample((16,))`` (KL samples, :308)."""
    N = torch.distributions.Normal
    orig_r, orig_s = N.rsample, N.sample

    def rsample(self, sample_shape=torch.Size()):
        assert tuple(sample_shape) == (), sample_shape
        return self.loc + self.scale * eps_pi.to(self.loc.dtype)

    def sample(self, sample_shape=torch.Size()):
        assert tuple(sample_shape) == (eps_kl.shape[0],), sample_shape
        with torch.no_grad():
            return self.loc.unsqueeze(0) + self.scale.unsqueeze(0) * eps_kl.to(self.loc.dtype)

    N.rsample, N.sample = rsample, sample
    try:
        yield
    finally:
        N.rsample, N.sample = orig_r, orig_s


def to_ref_transition(batch):
    """``[T, N]`` oracle batch -> the reference's per-step TensorDict layout (:143-158)."""
    from tensordict import TensorDict

    T, N = batch["soft_reward"].shape
    return TensorDict({
        "observations": batch["obs"].clone(),
        "critic_observations": batch["critic_obs"].clone(),
        "actions": batch["action"].clone(),
        "rewards": batch["soft_reward"].clone().unsqueeze(-1),
        "next_embeddings": batch["next_emb"].clone(),
        "next_values": batch["value"].clone().unsqueeze(-1),
        "dones": batch["done"].clone().unsqueeze(-1),
        "truncations": batch["truncated"].clone().unsqueeze(-1),
    }, batch_size=(T, N))


def run_update(ref, ts, hp, batch, perms, eps_pi, eps_kl, max_steps=None, trace=True):
    """Drive the reference closures exactly like the trainer loop (:613-632), but with the
    permutation and noise injected.  Returns (flat data with 'gve', list of per-step logs)."""
    cfg = make_cfg(hp, partial_reset=hp.partial_reset)
    post = ref.reppo.make_postprocess_fn(cfg, None)
    upd_c = ref.reppo.make_critic_update_fn(cfg, ts)
    upd_a = ref.reppo.make_actor_update_fn(cfg, ts)
    data = post(ts, to_ref_transition(batch))
    mbs = hp.minibatch_size
    logs, steps = [], 0
    for e in range(hp.num_epochs):
        # the reference re-shuffles the already shuffled data (:621); composing the injected
        # permutations reproduces "minibatch j of epoch e = rows perms[e][j*mb:(j+1)*mb]"
        for j in range(hp.num_mini_batches):
            if max_steps is not None and steps >= max_steps:
                break
            idx = perms[e, j * mbs:(j + 1) * mbs]
            mini = data[idx]
            cl = upd_c(mini)
            with injected_noise(eps_pi[e], eps_kl[e]):
                al = upd_a(mini)
            logs.append({**{k: v.clone() for k, v in cl.items()}, **{k: v.clone() for k, v in al.items()}})
            steps += 1
    for p, tp in zip(ts.actor.parameters(), ts.old_actor.parameters()):  # :631-632
        tp.data.copy_(p.data)
    return data, logs
