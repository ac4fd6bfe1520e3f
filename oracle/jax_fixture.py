"""Fixture format of a REAL JAX ``learn_step`` run (src/jaxrl/reppo.py:417-673) and its loader.  TEST INFRASTRUCTURE.

``oracle/gen_golden_jax.py`` (run on a machine that has the reference's JAX stack; there is none in the build image)
writes ``tests/golden/jax_<case>.npz``; ``tests/test_jax_fixture_hook.py`` loads every such file it finds and holds the
oracle (CPU) and the CUDA path (GPU) against it.  The file is flat numpy, self-describing, and carries the network
parameters under their flax-nnx state paths, so nothing here depends on JAX:

    hyper                    JSON of the ReppoConfig fields the update consumes + obs/action dims
    batch/<field>            [T, N, ...] Transition fields (obs, critic_obs, action, reward, soft_reward, next_emb, value,
                             done, truncated, importance_weight)
    perms [E, S], eps_pi [E, mb, A], eps_kl [E, 16, mb, A]
                             the permutations / standard-normal draws learn_step made from its key (re-derived by the
                             dumper with the same split sequence and the same distrax sampling call)
    in/critic/<path>, in/actor/<path>      parameters before the update, nnx paths joined with '/'
    out/critic/<path>, out/actor/<path>    parameters after the update
    out/metrics/<name>                     last-epoch means; out/target_values [T, N] optional (oracle-written files only)

nnx path -> oracle key (jax_models.py:57-125: FCNN = input_layer, main_layers[j], output_layer, each an nnx.Sequential of
Linear [, RMSNorm], so ``.../layers/0/kernel|bias`` and ``.../layers/1/scale``; a Linear kernel is [in, out] -> transposed):
"""
from __future__ import annotations

import json
import re
from typing import Dict

import numpy as np
import torch

from . import reppo_oracle as O

_BLOCK = re.compile(r"^(?P<mod>[a-z_]+)/(?P<blk>input_layer|output_layer|main_layers/(?P<j>\d+))/layers/(?P<slot>\d+)/(?P<leaf>kernel|bias|scale)$")
_MODULES = {"feature_module": ("feature_module", "num_critic_encoder_layers", False),
            "critic_module": ("critic_module", "num_critic_head_layers", True),
            "pred_module": ("pred_module", "num_critic_pred_layers", True),
            "actor_module": ("model", "num_actor_layers", False)}
_IGNORED = re.compile(r"^[a-z_]+/norm/scale$")   # FCNN.norm: constructed, never applied (jax_models.py:100,115)


def oracle_key(path: str, hp: O.Hyper):
    """nnx state path -> (oracle parameter name, needs_transpose); None for parameters the update never touches."""
    path = path.replace(".", "/").removesuffix("/value")
    if _IGNORED.match(path):
        return None
    if path == "zero_dist":
        return "zero_dist", False
    if path == "temperature_log_param":
        return "log_temp", False
    if path == "lagrangian_log_param":
        return "log_lagrange", False
    m = _BLOCK.match(path)
    if m is None or m["mod"] not in _MODULES:
        raise KeyError(f"unrecognised nnx parameter path '{path}': extend oracle/jax_fixture.py")
    prefix, layers_attr, inp_act = _MODULES[m["mod"]]
    layers = getattr(hp, layers_attr)
    i = 0 if m["blk"] == "input_layer" else (layers - 1 if m["blk"] == "output_layer" else 1 + int(m["j"]))
    blk = O._fcnn_keys(prefix, layers, inp_act)[i]
    if m["leaf"] == "scale":
        return f"{blk}.1.weight", False
    return (f"{blk}.0.weight", True) if m["leaf"] == "kernel" else (f"{blk}.0.bias", False)


def nnx_path(key: str, hp: O.Hyper) -> str:
    """Inverse of ``oracle_key`` (used to write a fixture from oracle-side tensors in the hook's self-test)."""
    if key in ("zero_dist",):
        return key
    if key == "log_temp":
        return "temperature_log_param"
    if key == "log_lagrange":
        return "lagrangian_log_param"
    prefix, _, idx, slot, leaf = re.match(r"^([a-z_]+)\.(net)\.(\d+)\.(\d)\.(weight|bias)$", key).groups()
    mod = {v[0]: k for k, v in _MODULES.items()}[prefix]
    _, layers_attr, inp_act = _MODULES[mod]
    layers = getattr(hp, layers_attr)
    i = int(idx) - (1 if inp_act else 0)
    blk = "input_layer" if i == 0 else ("output_layer" if i == layers - 1 else f"main_layers/{i - 1}")
    leaf = "scale" if slot == "1" else ("kernel" if leaf == "weight" else "bias")
    return f"{mod}/{blk}/layers/{slot}/{leaf}"


def _params(z, group: str, hp: O.Hyper, want: Dict[str, tuple]) -> Dict[str, torch.Tensor]:
    out = {}
    for name in z.files:
        if not name.startswith(group + "/"):
            continue
        mapped = oracle_key(name[len(group) + 1:], hp)
        if mapped is None:
            continue
        key, tr = mapped
        t = torch.from_numpy(np.asarray(z[name], dtype=np.float32))
        t = t.t().contiguous() if tr else t
        out[key] = t.reshape(want[key]) if key in want else t
    missing = sorted(set(want) - set(out))
    extra = sorted(set(out) - set(want))
    if missing or extra:
        raise KeyError(f"fixture group '{group}' does not match the oracle layout: missing {missing}, unexpected {extra}")
    return out


def load(path: str):
    """-> (hp, dict(batch, perms, eps_pi, eps_kl, critic, actor, out_critic, out_actor, target_values, metrics))."""
    z = np.load(path, allow_pickle=False)
    hp = O.Hyper(**json.loads(str(z["hyper"])))
    sem = O.JAX
    cshape, ashape = O.critic_param_shapes(hp, sem), O.actor_param_shapes(hp)
    t = lambda a: torch.from_numpy(np.asarray(a))
    return hp, {
        "batch": {n[6:]: t(z[n]).float() for n in z.files if n.startswith("batch/")},
        "perms": t(z["perms"]).long(), "eps_pi": t(z["eps_pi"]).float(), "eps_kl": t(z["eps_kl"]).float(),
        "critic": _params(z, "in/critic", hp, cshape), "actor": _params(z, "in/actor", hp, ashape),
        "out_critic": _params(z, "out/critic", hp, cshape), "out_actor": _params(z, "out/actor", hp, ashape),
        "target_values": t(z["out/target_values"]).float() if "out/target_values" in z.files else None,
        "metrics": {n[12:]: float(z[n]) for n in z.files if n.startswith("out/metrics/")},
    }


def dump_from_oracle(path: str, hp: O.Hyper, seed: int = 0):
    """Write a file in the fixture format from an ORACLE run (JAX semantics).  It pins nothing -- it exists so that the
    loader, the name mapping and the comparison code of the hook are exercised while no JAX machine has produced a real
    fixture."""
    sem = O.JAX
    cp, ap = O.init_params(hp, sem, seed=seed)
    batch = O.synthetic_rollout(hp, seed=seed, terminate=True, iw_variant=True)
    perms, (eps_pi, eps_kl) = O.synthetic_perms(hp), O.synthetic_noise(hp)
    ts, metrics, target, _ = O.learn_step(O.TrainState.create(cp, ap), {k: v.clone() for k, v in batch.items()}, perms, eps_pi,
                                          eps_kl, hp, sem)
    arrays = {"hyper": np.array(json.dumps({k: getattr(hp, k) for k in hp.__dataclass_fields__}))}
    arrays.update({f"batch/{k}": v.numpy() for k, v in batch.items()})
    arrays.update(perms=perms.numpy(), eps_pi=eps_pi.numpy(), eps_kl=eps_kl.numpy())

    def put(group, params):
        for k, v in params.items():
            p = nnx_path(k, hp)
            a = v.detach().numpy()
            arrays[f"{group}/{p}"] = a.T.copy() if p.endswith("kernel") else (a.reshape(1) if a.ndim == 0 else a)
    put("in/critic", cp), put("in/actor", ap), put("out/critic", ts.critic), put("out/actor", ts.actor)
    arrays["out/target_values"] = target.numpy()
    for k, v in metrics.items():
        arrays[f"out/metrics/{k}"] = np.asarray(float(v))
    np.savez(path, **arrays)
