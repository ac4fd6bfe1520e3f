"""Helpers to load the reference-generated fixtures (tests/golden/*.pt, made by oracle/gen_golden.py)."""
import os

import torch

from oracle import reppo_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _digest_ok(t, d, tol=0.0):
    # the sums are re-computed on another CPU (different SIMD width -> different summation order):
    # compare them to 1e-10 relative, the 64-entry head exactly
    f = t.detach().double().reshape(-1)
    tol = max(tol, 1e-10 * max(1.0, d["abs_sum"].item()))
    return (tuple(t.shape) == tuple(d["shape"]) and torch.equal(t.reshape(-1)[:64], d["head"])
            and abs(f.sum().item() - d["sum"].item()) <= tol and abs(f.abs().sum().item() - d["abs_sum"].item()) <= tol)


def load_case(name):
    """Returns (hp, golden dict with all inputs materialised)."""
    g = torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)
    hp = O.Hyper(**g["hyper"])
    inp = g["inputs"]
    if inp["critic"] is None:
        cp, ap = O.init_params(hp, O.TORCH, seed=0)
        for k, v in cp.items():
            assert _digest_ok(v, inp["param_digest"]["critic"][k]), f"regenerated critic param {k} differs"
        for k, v in ap.items():
            assert _digest_ok(v, inp["param_digest"]["actor"][k]), f"regenerated actor param {k} differs"
        inp["critic"], inp["actor"] = cp, ap
    if "next_emb" not in inp["batch"]:
        full = O.synthetic_rollout(hp, seed=0, terminate=g["terminate"])
        assert _digest_ok(full["next_emb"], inp["next_emb_digest"]), "regenerated next_emb differs"
        for k, v in inp["batch"].items():
            assert torch.equal(full[k], v), f"regenerated batch field {k} differs"
        inp["batch"]["next_emb"] = full["next_emb"]
    return hp, g
