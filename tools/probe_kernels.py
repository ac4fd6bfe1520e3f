"""Run the loss / sampling kernels alone through reppo_kernel_probe (for ncu):  python tools/probe_kernels.py [rows] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from reppo_b200.engine import EngineConfig, HyperParams, ReppoEngine  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
eng = ReppoEngine(EngineConfig(obs_dim=17, action_dim=6, semantics="jax", gemm_mode="bf16", max_rows=2048))
hyp = HyperParams()
B, H, A = 151, 512, 6
x = torch.randn(rows, B, device=dev); o = torch.empty(rows, B, device=dev)
tg = 150.0 * torch.rand(rows, device=dev)
ce = (x, tg, torch.zeros(rows, device=dev), torch.zeros(B, device=dev), o, torch.zeros(24 + 2 * B, device=dev))
pa = 0.3 * torch.randn(rows, 2 * A, device=dev); po = 0.3 * torch.randn(rows, 2 * A, device=dev)
sa = (pa, po, torch.randn(rows, A, device=dev), torch.randn(16, rows, A, device=dev), torch.empty(rows, A, device=dev),
      torch.zeros(rows * (2 + 2 * A), device=dev))
for _ in range(reps):
    eng.kernel_probe(0, *ce[:4], ce[4], ce[5], hyp)
    eng.kernel_probe(2, *sa[:4], sa[4], sa[5], hyp)
torch.cuda.synchronize()
print("ok")
