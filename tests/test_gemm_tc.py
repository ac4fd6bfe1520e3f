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
