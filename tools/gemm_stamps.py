"""Where does an isolated C2 Linear GEMM spend its time?  Per-CTA SM-clock stamps from the instrumented side build.

    python -m reppo_b200.build --variant=stamps -DREPPO_TC_STAMPS      (build host)
    python tools/gemm_stamps.py [--rows 2048] [--dim 512] [--chain 20]  (GPU box)

A chain of `chain` identical, stream-ordered launches (programmatic dependent launch on, replayed as a CUDA graph) of the
forward / dgrad / wgrad GEMM; every CTA records clock64 at: kernel entry, before / after griddepcontrol.wait, first / last
operand stage landed, accumulator complete, TMEM drained to shared memory, global stores done, CTA exit.  Printed: the
median (over CTAs and launches) length of every phase in SM cycles, and per SM the gap between the exit of launch L and the
release of launch L+1's dependency wait (the smallest gap over SMs = what one link of a dependent kernel chain costs).
"""
import argparse
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("REPPO_LIB_PATH", os.path.join(ROOT, "reppo_b200", "lib_stamps.so"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2048)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--chain", type=int, default=20)
    args = ap.parse_args()
    from reppo_b200.engine import EngineConfig, ReppoEngine
    rows, H, n = args.rows, args.dim, args.chain
    eng = ReppoEngine(EngineConfig(obs_dim=17, action_dim=6, critic_hidden_dim=H, actor_hidden_dim=H, num_bins=151,
                                   semantics="jax", gemm_mode="bf16", max_rows=rows))
    lib = eng.lib
    lib.reppo_debug_tc_stamps.restype = C.c_int
    lib.reppo_debug_tc_stamps.argtypes = [C.c_void_p, C.c_uint]
    lib.reppo_debug_tc_stamp_count.restype = C.c_uint
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(rows, H, device=dev, generator=g)
    w = torch.randn(H, H, device=dev, generator=g) / math.sqrt(H)
    dy = torch.randn(rows, H, device=dev, generator=g)
    b = torch.randn(H, device=dev, generator=g)
    cap = 64 * 1024
    stamps = torch.zeros(cap, 12, dtype=torch.int64, device=dev)
    names = ["entry->wait", "in wait", "wait->first stage", "first->last stage", "last stage->acc ready", "tmem->smem",
             "global stores", "stores->exit"]
    pairs = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 6), (6, 7), (7, 8), (8, 9)]
    for op, (a_, b_), name in ((0, (x, w), "fwd"), (1, (dy, w), "dgrad"), (2, (dy, x), "wgrad")):
        out = eng.linear(op, a_, b_, bias=b if op == 0 else None)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                for i in range(n):
                    # 0x100: operands already staged; 0x200 (after the first launch): keep the stream's chain cursor, so that
                    # every launch is linked to its predecessor (chain.cuh) unless REPPO_NO_CHAIN=1
                    lib.reppo_linear(eng.ctx, op | 0x100 | (0x200 if i else 0), rows, H, H, a_.data_ptr(), b_.data_ptr(),
                                     out.data_ptr(), None, eng._stream())
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(2):
            gr.replay()
        torch.cuda.synchronize()
        assert lib.reppo_debug_tc_stamps(stamps.data_ptr(), cap) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        cnt = lib.reppo_debug_tc_stamp_count()
        us = e0.elapsed_time(e1) * 1e3 / n
        ctas = cnt // n
        s = stamps[:cnt].cpu().view(n, ctas, 12)
        print(f"== {name} {rows}x{H}x{H}: {us:.2f} us per launch ({us * 1.965e3:.0f} clk at 1965 MHz), {ctas} CTAs per launch")
        mid = s[2:]  # skip the first launches (graph start)
        for (i, j), nm in zip(pairs, names):
            d = (mid[:, :, j] - mid[:, :, i]).flatten().float()
            print(f"   {nm:24s} median {d.median().item():7.0f}  p10 {d.quantile(0.1).item():7.0f}  p90 {d.quantile(0.9).item():7.0f} clk")
        tot = (mid[:, :, 9] - mid[:, :, 2]).flatten().float()
        print(f"   {'wait release -> exit':24s} median {tot.median().item():7.0f}  max {tot.max().item():7.0f} clk")
        # per SM: exit of launch L -> wait release of launch L+1 (same SM clock)
        smid = (s[:, :, 11] & 0xFFFF)
        gaps, spans = [], []
        for L in range(2, n - 1):
            last_exit = {}
            for c in range(ctas):
                sm = int(smid[L, c])
                last_exit[sm] = max(last_exit.get(sm, -1), int(s[L, c, 9]))
            g_ = [int(s[L + 1, c, 2]) - last_exit[int(smid[L + 1, c])] for c in range(ctas) if int(smid[L + 1, c]) in last_exit]
            if g_:
                gaps.append(min(g_))
            # launch period seen by one SM: wait release of L -> wait release of L+1
            p_ = {}
            for c in range(ctas):
                p_[int(smid[L, c])] = int(s[L, c, 2])
            q_ = [int(s[L + 1, c, 2]) - p_[int(smid[L + 1, c])] for c in range(ctas) if int(smid[L + 1, c]) in p_]
            if q_:
                spans.append(sorted(q_)[len(q_) // 2])
        if gaps:
            gaps.sort(); spans.sort()
            print(f"   last CTA exit -> next launch's wait released (min over SMs): median {gaps[len(gaps) // 2]} clk")
            print(f"   wait release -> next wait release on the same SM: median {spans[len(spans) // 2]} clk")
        del gr


if __name__ == "__main__":
    main()
