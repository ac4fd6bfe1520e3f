"""ctypes binding of ``libreppo_b200.so`` (the C ABI declared in ``include/reppo_b200.h``).

There is no fallback: if the shared library is missing it is built with nvcc, and if that is
impossible the import fails loudly.  Every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libreppo_b200.so")

# ---- enums (mirror include/reppo_b200.h) -------------------------------------------------------
SEM_JAX, SEM_TORCH = 0, 1
NORM_NONE, NORM_RMS, NORM_LAYER = 0, 1, 2
KL_CLIPPED, KL_FULL, KL_VALUE = 0, 1, 2
GEMM_FP32, GEMM_BF16, GEMM_TF32 = 0, 1, 2
GEMM_MODES = {"fp32": GEMM_FP32, "bf16": GEMM_BF16, "tf32": GEMM_TF32}
NET_CRITIC, NET_ACTOR = 0, 1
NUM_METRICS = 24
METRIC_NAMES = [
    "critic_loss", "ce_mean", "aux_mean", "value_loss", "q_mean", "target_mean", "target_max", "target_min",
    "critic_grad_norm", "actor_loss", "policy_loss", "kl", "entropy", "temperature", "lagrangian", "entropy_loss",
    "lagrangian_loss", "actor_q_mean", "abs_pred_action", "actor_grad_norm", "reward_mean", "abs_batch_action",
]
KL_MODES = {"clipped": KL_CLIPPED, "full": KL_FULL, "value": KL_VALUE}
NORMS = {"none": NORM_NONE, "rms": NORM_RMS, "layer": NORM_LAYER}

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "obs_dim", "critic_obs_dim", "action_dim", "critic_hidden", "actor_hidden", "num_bins",
        "enc_layers", "head_layers", "pred_layers", "actor_layers", "critic_norm", "actor_norm",
        "semantics", "gemm_mode", "max_rows", "kl_samples")] + [("vmin", C.c_float), ("vmax", C.c_float),
                                                                      ("max_batch_rows", C.c_int32)]


class Hyper(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "gamma", "lmbda", "lr", "max_grad_norm", "kl_bound", "ent_target_mult", "aux_loss_mult", "actor_min_std")] + [
        (n, C.c_int32) for n in ("kl_clip_mode", "reduce_kl", "update_entropy_lagrangian", "update_kl_lagrangian",
                                 "partial_reset", "world_size", "lr_anneal_steps", "reverse_kl")]


class NetState(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
                ("step", C.c_void_p)]


class State(C.Structure):
    _fields_ = [("critic", NetState), ("actor", NetState), ("actor_old", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "critic_obs", "action", "reward", "done", "truncated", "next_emb",
                                          "target")] + [("num_rows", C.c_int64)]


class Plan(C.Structure):
    _fields_ = [("num_epochs", C.c_int32), ("num_mini_batches", C.c_int32), ("minibatch", C.c_int32),
                ("perms", C.c_void_p), ("eps_pi", C.c_void_p), ("eps_kl", C.c_void_p),
                ("step_metrics", C.c_void_p), ("metrics", C.c_void_p), ("use_graph", C.c_int32),
                ("noise_per_minibatch", C.c_int32), ("old_policy_out", C.c_void_p)]


class Multicast(C.Structure):
    _fields_ = [("local_base", C.c_void_p), ("mc_base", C.c_void_p), ("peer_base", C.c_void_p * 16),
                ("critic_grads_off", C.c_int64), ("actor_grads_off", C.c_int64), ("flags_off", C.c_int64),
                ("flags_bytes", C.c_int64), ("rank", C.c_int32), ("world_size", C.c_int32)]


MC_FLAG_BYTES = (3 * 256 * 16 + 16) * 4   # kMcChains * kMcFlagBlocks * kMcMaxRanks flag words + epoch words (csrc/kernels.h)

# every symbol include/reppo_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
SYMBOLS = {
    "reppo_version": (C.c_int, []),
    "reppo_ctx_create": (C.c_int, [C.POINTER(Shape), C.POINTER(_vp)]),
    "reppo_ctx_destroy": (None, [_vp]),
    "reppo_last_error": (C.c_char_p, [_vp]),
    "reppo_param_count": (_i64, [_vp, C.c_int]),
    "reppo_param_num_tensors": (_i32, [_vp, C.c_int]),
    "reppo_param_tensor": (C.c_int, [_vp, C.c_int, _i32, C.c_char_p, _i32, C.POINTER(_i64), i32p, i32p]),
    "reppo_workspace_bytes": (C.c_size_t, [_vp]),
    "reppo_set_workspace": (C.c_int, [_vp, _vp, C.c_size_t]),
    "reppo_set_bins": (C.c_int, [_vp, _vp, _vp]),
    "reppo_td_lambda": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _d, _d, _i32, _i32, _vp]),
    "reppo_critic_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "reppo_actor_forward": (C.c_int, [_vp, _vp, _vp, _i32, _f, _vp, _vp, _vp]),
    "reppo_obs_normalize": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _f, _i32, _i32, _i64, _vp, _vp]),
    "reppo_policy_act": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _i32, _vp, _vp, _vp, _vp]),
    "reppo_bootstrap": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "reppo_linear": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "reppo_kernel_probe": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "reppo_critic_step": (C.c_int, [_vp, C.POINTER(State), C.POINTER(Batch), _vp, _i32, C.POINTER(Hyper), _vp, _vp]),
    "reppo_actor_step": (C.c_int, [_vp, C.POINTER(State), C.POINTER(Batch), _vp, _i32, _vp, _vp, C.POINTER(Hyper), _vp, _vp]),
    "reppo_critic_grad": (C.c_int, [_vp, C.POINTER(State), C.POINTER(Batch), _vp, _i32, C.POINTER(Hyper), _vp, _vp]),
    "reppo_actor_grad": (C.c_int, [_vp, C.POINTER(State), C.POINTER(Batch), _vp, _i32, _vp, _vp, C.POINTER(Hyper), _vp, _vp]),
    "reppo_clip_adam": (C.c_int, [_vp, C.POINTER(NetState), _i64, C.POINTER(Hyper), _vp, _vp]),
    "reppo_update": (C.c_int, [_vp, C.POINTER(State), C.POINTER(Batch), C.POINTER(Plan), C.POINTER(Hyper), _vp]),
    "reppo_launch_count": (_i64, [_vp]),
    "reppo_nccl_unique_id": (C.c_int, [_vp, _vp]),
    "reppo_nccl_init": (C.c_int, [_vp, _vp, _i32, _i32]),
    "reppo_nccl_finalize": (C.c_int, [_vp]),
    "reppo_dp_attach_multicast": (C.c_int, [_vp, C.POINTER(Multicast)]),
    "reppo_dp_status": (C.c_int, [_vp, i32p]),
    "reppo_dp_stamps": (C.c_int, [_vp, _vp, _i64]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the library (building it in-tree first when absent) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m reppo_b200.build` (no CPU fallback exists)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ReppoError(RuntimeError):
    pass


def check(code: int, ctx=None, what: str = ""):
    if code != 0:
        msg = load().reppo_last_error(ctx)
        raise ReppoError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")
