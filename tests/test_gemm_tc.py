"""The tcgen05 / TMEM / TMA GEMM (reppo_b200/csrc/gemm_bf16_tc.cu) against a plain PyTorch fp32
reference of the same op on the same bf16-rounded operands (so the only difference is the fp32
accumulation order inside the tensor core), through the C ABI entry ``reppo_linear``.
Tolerance: |err| <= 2e-3 * sqrt(K) * max|a| * max|b| * 2^-8 ... in practice we assert
rtol 1e-4 / atol 1e-3 * sqrt(K/512) on O(1) data, far below bf16 operand rounding (2^-9)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(max_rows, H=512, mode="bf16", D=17, A=6, B=151):
    from reppo_b200.engine import EngineConfig, ReppoEngine
    return ReppoEngine(EngineConfig(obs_dim=D, action_dim=A, critic_hidden_dim=H, actor_hidden_dim=H, num_bins=B,
                                    semantics="jax", gemm_mode=mode, max_rows=max_rows))


def _bf(x):
    return x.bfloat16().float()


SHAPES = [(2048, 512, 512), (128, 64, 64), (300, 23, 151), (2048, 512, 513), (1000, 512, 12), (77, 17, 512), (2048, 151, 512),
          (256, 512, 1)]


@pytest.mark.parametrize("rows,in_f,out_f", SHAPES)
def test_forward_dgrad_wgrad_vs_torch(rows, in_f, out_f):
    eng = _engine(max_rows=2048)
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + in_f)
    x = torch.randn(rows, in_f, device="cuda", generator=g)
    w = torch.randn(out_f, in_f, device="cuda", generator=g) / math.sqrt(in_f)
    b = torch.randn(out_f, device="cuda", generator=g)
    dy = torch.randn(rows, out_f, device="cuda", generator=g)
    y = eng.linear(0, x, w, bias=b)
    torch.testing.assert_close(y, _bf(x) @ _bf(w).T + b, rtol=1e-4, atol=2e-4 * math.sqrt(in_f))
    dx = eng.linear(1, dy, w)
    torch.testing.assert_close(dx, _bf(dy) @ _bf(w), rtol=1e-4, atol=2e-4 * math.sqrt(out_f))
    dw = eng.linear(2, dy, x)
    torch.testing.assert_close(dw, _bf(dy).T @ _bf(x), rtol=1e-4, atol=2e-4 * math.sqrt(rows))


# Shapes that fill the machine with 256-row pair tiles run the persistent cta_group::2 kernel (gemm_bf16_tc2.cu):
# full tiles, ragged rows / columns (scalar epilogue, TMA out-of-bounds fill), split-K wgrad, 128- and 256-wide tiles.
LARGE_SHAPES = [(16384, 1024, 1024), (4100, 1024, 1025), (8192, 512, 768), (5000, 1024, 384), (16384, 256, 1024)]


@pytest.mark.parametrize("rows,in_f,out_f", LARGE_SHAPES)
def test_large_shapes_pair_kernel_vs_torch(rows, in_f, out_f):
    eng = _engine(max_rows=16384, H=1024)
    g = torch.Generator(device="cuda").manual_seed(rows + in_f)
    x = torch.randn(rows, in_f, device="cuda", generator=g)
    w = torch.randn(out_f, in_f, device="cuda", generator=g) / math.sqrt(in_f)
    b = torch.randn(out_f, device="cuda", generator=g)
    dy = torch.randn(rows, out_f, device="cuda", generator=g)
    y = eng.linear(0, x, w, bias=b)
    torch.testing.assert_close(y, _bf(x) @ _bf(w).T + b, rtol=1e-4, atol=2e-4 * math.sqrt(in_f))
    assert torch.equal(y, eng.linear(0, x, w, bias=b))          # deterministic (no split-K in the forward)
    dx = eng.linear(1, dy, w)
    torch.testing.assert_close(dx, _bf(dy) @ _bf(w), rtol=1e-4, atol=2e-4 * math.sqrt(out_f))
    dw = eng.linear(2, dy, x)
    torch.testing.assert_close(dw, _bf(dy).T @ _bf(x), rtol=1e-4, atol=2e-4 * math.sqrt(rows))


def test_fp32_mode_linear_vs_torch():
    eng = _engine(max_rows=512, mode="fp32")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(300, 23, device="cuda", generator=g)
    w = torch.randn(151, 23, device="cuda", generator=g)
    b = torch.randn(151, device="cuda", generator=g)
    dy = torch.randn(300, 151, device="cuda", generator=g)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.testing.assert_close(eng.linear(0, x, w, bias=b), x @ w.T + b, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(eng.linear(1, dy, w), dy @ w, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(eng.linear(2, dy, x), dy.T @ x, rtol=1e-5, atol=1e-4)


# tf32 mode: tcgen05.mma.kind::tf32 reads the fp32 operands themselves (10-bit mantissa products, fp32 accumulation).  Shapes whose
# operands TMA can address (leading dimensions multiples of 4 floats) are read in place in all three arrangements, K-major and
# MN-major; ragged ones are re-laid into row-padded scratch first -- both must agree with torch fp32 within tf32
# operand precision (2^-11 relative per operand), and the tensor-core ones must be clearly closer than bf16 operands would be.
TF32_SHAPES = [(2048, 512, 512), (128, 64, 64), (300, 24, 152), (2048, 512, 516), (1000, 512, 12), (77, 16, 512), (2048, 152, 512),
               (300, 23, 151), (256, 512, 1), (640, 1024, 256)]


@pytest.mark.parametrize("rows,in_f,out_f", TF32_SHAPES)
def test_tf32_mode_forward_dgrad_wgrad_vs_torch(rows, in_f, out_f):
    eng = _engine(max_rows=2048, mode="tf32", H=1024)
    g = torch.Generator(device="cuda").manual_seed(rows * 5 + out_f)
    x = torch.randn(rows, in_f, device="cuda", generator=g)
    w = torch.randn(out_f, in_f, device="cuda", generator=g) / math.sqrt(in_f)
    b = torch.randn(out_f, device="cuda", generator=g)
    dy = torch.randn(rows, out_f, device="cuda", generator=g)
    torch.backends.cuda.matmul.allow_tf32 = False
    cases = [(eng.linear(0, x, w, bias=b), x @ w.T + b, _bf(x) @ _bf(w).T + b, in_f),
             (eng.linear(1, dy, w), dy @ w, _bf(dy) @ _bf(w), out_f),
             (eng.linear(2, dy, x), dy.T @ x, _bf(dy).T @ _bf(x), rows)]
    for got, ref, ref_bf16, k in cases:
        scale = ref.abs().max().item()
        torch.testing.assert_close(got, ref, rtol=2e-3, atol=1.5e-3 * scale)
        # ragged operands (leading dimension not a multiple of four floats) go through row-padded copies: every shape within the
        # context's limits runs on the tensor cores
        if k >= 64 and ref.numel() >= 4096:
            # not bf16 in disguise: the error against fp32 is several times below that of bf16-rounded operands
            assert (got - ref).abs().mean().item() < 0.35 * (ref_bf16 - ref).abs().mean().item()


def test_forward_is_deterministic_and_row_independent():
    """Properties at the full C2 minibatch size: repeated launches are bit-identical, and any
    row subset of the input gives the same rows of the output (no cross-row leakage through the
    128-row tiles / TMA out-of-bounds fill)."""
    eng = _engine(max_rows=2048)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2048, 512, device="cuda", generator=g)
    w = torch.randn(512, 512, device="cuda", generator=g) / 22.6
    y1 = eng.linear(0, x, w).clone()
    y2 = eng.linear(0, x, w).clone()
    assert torch.equal(y1, y2)
    y3 = eng.linear(0, x[:1000].contiguous(), w)
    assert torch.equal(y3, y1[:1000])


@pytest.mark.parametrize("H,norm", [(512, "rms"), (256, "layer"), (1024, "rms"), (128, "rms")])
def test_fused_norm_epilogue_matches_unfused(H, norm, monkeypatch):
    """Linear + bias + norm + swish fused into the tcgen05 epilogue (cluster-wide row statistics over DSMEM) against
    the unfused GEMM + row kernel on the same bf16 operands, and both against the fp32 oracle forward."""
    from oracle import reppo_oracle as O
    from reppo_b200.engine import EngineConfig, ReppoEngine
    hp = O.Hyper(critic_hidden_dim=H, actor_hidden_dim=H, norm_type=norm, obs_dim=17, critic_obs_dim=17, action_dim=6,
                 num_bins=151, vmin=0.0, vmax=150.0)
    cp, ap = O.init_params(hp, O.JAX, seed=3)
    g = torch.Generator().manual_seed(0)
    for d in (cp, ap):
        for k in d:
            if k.endswith(".1.weight") or k.endswith(".1.bias") or k.endswith(".0.bias"):
                d[k] = d[k] + 0.1 * torch.randn(d[k].shape, generator=g)
    rows = 300
    co, ob = torch.randn(rows, 17, generator=g), torch.randn(rows, 17, generator=g)
    ac = torch.tanh(torch.randn(rows, 6, generator=g))
    outs = []
    for fuse in (True, False):
        if fuse:
            monkeypatch.delenv("REPPO_NO_FUSE_NORM", raising=False)
        else:
            monkeypatch.setenv("REPPO_NO_FUSE_NORM", "1")
        eng = ReppoEngine(EngineConfig(obs_dim=17, action_dim=6, critic_hidden_dim=H, actor_hidden_dim=H, num_bins=151,
                                       norm_type=norm, semantics="jax", gemm_mode="bf16", max_rows=512))
        eng.load_params(cp, ap)
        v, lg, pr, ft = eng.critic_forward(co, ac)
        mu, sd = eng.actor_forward(ob)
        outs.append([t.clone() for t in (v, lg, pr, ft, mu, sd)])
    for a, b in zip(*outs):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-3 * max(1.0, b.abs().max().item()))
    v0, l0, p0, f0 = O.critic_forward(cp, co, ac, hp, O.JAX)
    m0, s0 = O.actor_forward(ap, ob, hp, O.JAX)
    for got, want in zip(outs[0], (v0, l0, p0, f0, m0, s0)):
        scale = want.abs().max().item()
        assert (got.cpu() - want).abs().max().item() <= 3e-2 * max(scale, 1.0)
