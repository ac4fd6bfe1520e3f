"""Hydra-style config composer for the REPPO config surface (hydra / omegaconf are not installed
in this image).  Reproduces what the reference entry points rely on:

* the defaults list of ``config/reppo.yaml:1-5`` (``env``, ``platform``, ``experiment_overrides``, ``_self_``),
* CLI overrides ``group=name`` (config-group choice) and ``a.b.c=value`` (value override),
* ``${a.b}`` interpolation (``vmin: ${env.vmin}``, ``reppo.yaml:40-41``),
* the merge of ``experiment_overrides.hyperparameters`` into ``hyperparameters``
  (``src/jaxrl/reppo.py:932``; the Torch entry point forgets it, ``src/torchrl/reppo.py:442-443`` --
  pass ``merge_overrides=False`` to mimic that),
* attribute access ``cfg.hyperparameters.lr`` and ``cfg.env.get("partial_reset", False)``.
"""
from __future__ import annotations

import os
import re
from typing import Any, Iterable, Optional

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
GROUPS = ("env", "platform", "experiment_overrides")


class DictConfig(dict):
    """dict with attribute access (enough of omegaconf.DictConfig for the update path)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(x):
        if isinstance(x, dict):
            return DictConfig({k: DictConfig.wrap(v) for k, v in x.items()})
        if isinstance(x, list):
            return [DictConfig.wrap(v) for v in x]
        return x


class _Loader(yaml.SafeLoader):
    pass


# YAML 1.1 does not read "3e-4" as a float; hydra (YAML 1.2-ish) does
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^[-+]?(?:\d[\d_]*\.\d*(?:[eE][-+]?\d+)?|\d[\d_]*[eE][-+]?\d+|\.\d+(?:[eE][-+]?\d+)?|\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$"),
    list("-+0123456789."))


def _load_yaml(path: str) -> dict:
    with open(path) as f:
        return yaml.load(f, Loader=_Loader) or {}


def _parse_value(s: str) -> Any:
    return yaml.load(s, Loader=_Loader)


def _deep_merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _set_path(d: dict, path: str, value: Any):
    keys = path.split(".")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def _get_path(d: dict, path: str):
    for k in path.split("."):
        d = d[k]
    return d


_INTERP = re.compile(r"\$\{([^}]+)\}")


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node.strip())
        if m:
            return _resolve(_get_path(root, m.group(1)), root)
        return _INTERP.sub(lambda mm: str(_resolve(_get_path(root, mm.group(1)), root)), node)
    return node


def compose(overrides: Iterable[str] = (), config_name: str = "reppo", config_dir: Optional[str] = None,
            merge_overrides: bool = True) -> DictConfig:
    config_dir = config_dir or CONFIG_DIR
    primary = _load_yaml(os.path.join(config_dir, f"{config_name}.yaml"))
    defaults = primary.pop("defaults", [])
    choices = {}
    order = []
    for item in defaults:
        if item == "_self_":
            order.append("_self_")
        else:
            (g, name), = item.items()
            choices[g] = name
            order.append(g)
    if "_self_" not in order:
        order.append("_self_")
    value_overrides = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        key = key.lstrip("+")
        if key in choices or key in GROUPS:
            choices[key] = val
            if key not in order:
                order.insert(0, key)
        else:
            value_overrides.append((key, _parse_value(val)))
    cfg: dict = {}
    for g in order:
        if g == "_self_":
            _deep_merge(cfg, primary)
        else:
            path = os.path.join(config_dir, g, f"{choices[g]}.yaml")
            if not os.path.exists(path):
                raise FileNotFoundError(f"config group '{g}' has no option '{choices[g]}' ({path})")
            _deep_merge(cfg, {g: _load_yaml(path)})
    for key, val in value_overrides:
        _set_path(cfg, key, val)
    if merge_overrides:
        _deep_merge(cfg.setdefault("hyperparameters", {}), cfg.get("experiment_overrides", {}).get("hyperparameters", {}) or {})
        # explicit CLI value overrides win over the experiment override file
        for key, val in value_overrides:
            if key.startswith("hyperparameters."):
                _set_path(cfg, key, val)
    cfg = _resolve(cfg, cfg)
    return DictConfig.wrap(cfg)


def validate_update_config(cfg: DictConfig):
    """Reject the config switches whose code paths this implementation does not carry
    (SURVEY.md A.8): they are dead or ablation-only paths in the reference."""
    h = cfg.hyperparameters
    if not h.get("hl_gauss", True):
        raise NotImplementedError("hl_gauss=false (scalar MSE critic, src/networks/jax_models.py:127-208) is not implemented")
    for k in ("use_simplical_embedding", "use_critic_skip", "use_actor_skip"):
        if h.get(k, False):
            raise NotImplementedError(f"{k}=true is a dead path in the reference and is not implemented")
    if h.get("reverse_kl", False) and cfg.get("b200", {}).get("semantics", "torch") != "jax":
        # only the JAX update has the switch (src/jaxrl/reppo.py:543-553); src/torchrl/reppo.py never reads it
        raise NotImplementedError("reverse_kl=true exists in the JAX update only: set b200.semantics=jax")
    if h.actor_kl_clip_mode not in ("clipped", "full", "value"):
        raise ValueError(f"Unknown actor loss mode: {h.actor_kl_clip_mode}")
    for k in ("num_critic_encoder_layers", "num_critic_head_layers", "num_critic_pred_layers", "num_actor_layers"):
        if int(h[k]) < 2:
            raise NotImplementedError(f"{k}=1 has divergent JAX/Torch semantics in the reference (SURVEY.md App. B)")
    if (int(h.num_steps) * int(h.num_envs)) % int(h.num_mini_batches):
        raise ValueError("num_steps * num_envs must be divisible by num_mini_batches")
