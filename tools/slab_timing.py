"""Per-step clock stamps of the slab programs at a bench shape (profiling aid; run on the GPU box).

    REPPO_SLAB_TIMING=1 python tools/slab_timing.py [C2]
"""
import ctypes as C
import os
import sys

os.environ.setdefault("REPPO_SLAB_TIMING", "1")
os.environ.setdefault("REPPO_SLAB", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import reppo_oracle as O  # noqa: E402
from reppo_b200 import torch_api as T  # noqa: E402
from reppo_b200.engine import HyperParams  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    cfg, D, A, term, rs = bench.make_cfg(name, "jax")
    cfg.platform.amp_dtype = "bf16"
    hp = bench.oracle_hyper(cfg, D, A)
    dev = torch.device("cuda:0")
    ts = T.make_train_state(cfg, D, D, A, device=dev)
    eng = ts.engine
    cp, ap = O.init_params(hp, O.JAX, seed=0)
    eng.load_params(cp, ap)
    batch = O.synthetic_rollout(hp, seed=0, terminate=term, reward_scaling=rs)
    devb = {k: v.to(dev) for k, v in batch.items()}
    flat = {k: v.reshape(-1, *v.shape[2:]) for k, v in devb.items()}
    flat["target"] = flat["value"].clone()
    cb = eng.make_batch(flat)
    S, mb = hp.batch_size, hp.minibatch_size
    g = torch.Generator(device=dev).manual_seed(0)
    idx = torch.randperm(S, device=dev, generator=g)[:mb].to(torch.int32)
    eps_pi = torch.randn(mb, A, device=dev, generator=g)
    eps_kl = torch.randn(16, mb, A, device=dev, generator=g)
    hyp = HyperParams()
    ptr, nbytes = C.c_void_p(), C.c_size_t()
    assert eng.lib.reppo_debug_buffer(eng.ctx, b"slab_timing", C.byref(ptr), C.byref(nbytes)) == 0
    base = eng.workspace.data_ptr()
    off = ptr.value - base
    view = eng.workspace[off:off + nbytes.value].view(torch.int64).view(3, 448)
    for rep in range(3):
        view.zero_()
        eng.critic_grad(cb, idx, hyp)
        eng.actor_grad(cb, idx, eps_pi, eps_kl, hyp)
        torch.cuda.synchronize()
    t = view.cpu()
    names = ["critic", "policy prefix", "policy main"]
    for pi in range(3):
        row = [int(x) for x in t[pi][:64] if int(x) != 0]
        if len(row) < 2:
            continue
        d = [(row[i + 1] - row[i]) for i in range(len(row) - 1)]
        print(f"{names[pi]}: {len(d)} steps, total {sum(d)} cycles ({sum(d) / 1.965e3:.1f} us at 1965 MHz)")
        print("   per step (cycles): " + " ".join(str(x) for x in d))
        print("   step: [tma issued, first full, mma issued, tmem full, part1 end, stats sync, part2 end, p2 tmem/dsmem read, p2 staged, p2 loop end] relative to step start")
        for si in range(len(d)):
            st = [int(x) for x in t[pi][64 + si * 12: 64 + si * 12 + 12]]
            rel = [(x - row[si]) if x else 0 for x in st[1:]]
            print(f"   {si:2d}: {rel}")


if __name__ == "__main__":
    main()
