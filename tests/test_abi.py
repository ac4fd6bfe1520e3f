"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/reppo_b200.h declares, and lays the parameter arenas out in the reference's state_dict
order (no compute calls -- those need a GPU)."""
import ctypes as C
import os
import re

import pytest

from oracle import reppo_oracle as O
from reppo_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "reppo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reppo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/reppo_b200.h but not exported"
        assert n in L.SYMBOLS, f"{n} has no ctypes signature in reppo_b200/_lib.py"
    assert set(L.SYMBOLS) == set(names)
    assert lib.reppo_version() >= 1


def _ctx(hp: O.Hyper, sem):
    from reppo_b200.engine import EngineConfig
    cfg = EngineConfig(obs_dim=hp.obs_dim, action_dim=hp.action_dim, critic_obs_dim=hp.critic_obs_dim,
                       critic_hidden_dim=hp.critic_hidden_dim, actor_hidden_dim=hp.actor_hidden_dim,
                       num_bins=hp.num_bins, vmin=hp.vmin, vmax=hp.vmax, norm_type=hp.norm_type, semantics=sem.name,
                       use_critic_norm=hp.use_critic_norm, use_actor_norm=hp.use_actor_norm,
                       num_actor_layers=hp.num_actor_layers, num_critic_encoder_layers=hp.num_critic_encoder_layers)
    lib = L.load()
    shape = cfg.shape()
    ctx = C.c_void_p()
    L.check(lib.reppo_ctx_create(C.byref(shape), C.byref(ctx)), None, "create")
    return lib, ctx


def _layout(lib, ctx, net):
    out = []
    name = C.create_string_buffer(128)
    off, rows, cols = C.c_int64(), C.c_int32(), C.c_int32()
    for i in range(lib.reppo_param_num_tensors(ctx, net)):
        assert lib.reppo_param_tensor(ctx, net, i, name, 128, C.byref(off), C.byref(rows), C.byref(cols)) == 0
        shp = () if rows.value == 0 else ((rows.value,) if cols.value == 0 else (rows.value, cols.value))
        out.append((name.value.decode(), off.value, shp))
    return out


@pytest.mark.parametrize("sem", [O.JAX, O.TORCH])
@pytest.mark.parametrize("norm", ["rms", "layer"])
def test_param_layout_is_reference_state_dict_order(sem, norm):
    hp = O.Hyper(norm_type=norm)  # CheetahRun-shaped defaults (C2)
    lib, ctx = _ctx(hp, sem)
    try:
        for net, shapes in ((L.NET_CRITIC, O.critic_param_shapes(hp, sem)), (L.NET_ACTOR, O.actor_param_shapes(hp))):
            lay = _layout(lib, ctx, net)
            assert [n for n, _, _ in lay] == list(shapes)
            off = 0
            for (n, o, shp), want in zip(lay, shapes.values()):
                assert tuple(shp) == tuple(want), n
                assert o == off
                k = 1
                for s in shp:
                    k *= s
                off += k
            assert lib.reppo_param_count(ctx, net) == off
        if sem.name == "torch" and norm == "rms":
            assert lib.reppo_param_count(ctx, L.NET_CRITIC) == 1_142_062
            assert lib.reppo_param_count(ctx, L.NET_ACTOR) == 279_054
        assert lib.reppo_workspace_bytes(ctx) > 0
    finally:
        lib.reppo_ctx_destroy(ctx)


def test_unsupported_shapes_fail_loudly():
    lib = L.load()
    from reppo_b200.engine import EngineConfig
    for kw, frag in ((dict(num_actor_layers=1), b"layers == 1"), (dict(num_bins=300), b"num_bins"),
                     (dict(critic_hidden_dim=4096), b"hidden_dim")):
        shape = EngineConfig(obs_dim=3, action_dim=2, **kw).shape()
        ctx = C.c_void_p()
        assert lib.reppo_ctx_create(C.byref(shape), C.byref(ctx)) < 0
        assert frag in lib.reppo_last_error(None)


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from reppo_b200.engine import EngineConfig, ReppoEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ReppoEngine(EngineConfig(obs_dim=3, action_dim=2))
