"""Isolated timing + correctness of the Linear GEMMs at large shapes (the C5 sweep) through reppo_linear.

    python tools/gemm_bench.py [--rows 16384] [--dims 1024,2048] [--check]

Each (op, rows, in, out): checked against torch on the same bf16-rounded operands, then 20 back-to-back launches
replayed as a CUDA graph between CUDA events.  Compare REPPO_TC2=0 (one-tile-per-CTA kernel) with the default
(persistent CTA-pair kernel)."""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(rows, H, reps=20, check=True, peaks=None):
    """fwd / dgrad / wgrad of one H x H Linear at `rows` rows: microseconds, TFLOP/s, fraction of the measured bf16 burst peak."""
    from reppo_b200.engine import EngineConfig, ReppoEngine
    if peaks is None:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    dev = torch.device("cuda:0")
    eng = ReppoEngine(EngineConfig(obs_dim=17, action_dim=6, critic_hidden_dim=H, actor_hidden_dim=H, num_bins=151,
                                   semantics="jax", gemm_mode="bf16", max_rows=rows))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(rows, H, device=dev, generator=g)
    w = torch.randn(H, H, device=dev, generator=g) / math.sqrt(H)
    dy = torch.randn(rows, H, device=dev, generator=g)
    b = torch.randn(H, device=dev, generator=g)
    res = []
    for op, (a_, b_), name in ((0, (x, w), "fwd"), (1, (dy, w), "dgrad"), (2, (dy, x), "wgrad")):
        out = eng.linear(op, a_, b_, bias=b if op == 0 else None)
        torch.cuda.synchronize()
        err = None
        if check:
            bf = lambda t: t.bfloat16().float()
            ref = bf(a_) @ bf(b_).T + b if op == 0 else (bf(a_) @ bf(b_) if op == 1 else bf(a_).T @ bf(b_))
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            del ref
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                for _ in range(reps):
                    eng.lib.reppo_linear(eng.ctx, op | 0x100, rows, H, H, a_.data_ptr(), b_.data_ptr(), out.data_ptr(), None,
                                         eng._stream())
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(2):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (3 * reps)
        tf = 2.0 * rows * H * H / (ms * 1e-3) / 1e12
        res.append({"op": name, "rows": rows, "H": H, "us": round(ms * 1e3, 2), "tflops": round(tf, 1),
                    "frac_burst": round(tf / peaks["bf16_tflops"], 3), "rel_err": err})
        del gr
    del eng
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="16384")
    ap.add_argument("--dims", default="1024")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    for rows in [int(r) for r in args.rows.split(",")]:
        for H in [int(d) for d in args.dims.split(",")]:
            for r in measure(rows, H, args.reps, not args.no_check):
                r.update({"tc2": os.environ.get("REPPO_TC2", "1"), "bn": os.environ.get("REPPO_TC2_BN", "auto")})
                print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
