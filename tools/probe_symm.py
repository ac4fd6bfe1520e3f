"""Probe: does torch's symmetric memory (CUDA VMM + NVSwitch multicast) rendezvous work on this box?  torchrun, 2+ ranks."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
t.fill_(rank + 1)
hdl.barrier()
print(f"rank {rank}: world {hdl.world_size} multicast_ptr {hdl.multicast_ptr:#x} buffers {[hex(p) for p in hdl.buffer_ptrs]} "
      f"signal_pads {[hex(p) for p in hdl.signal_pad_ptrs]} pad_size {getattr(hdl, 'signal_pad_size', None)}", flush=True)
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(f"rank {rank}: peer value {peer[0].item()}", flush=True)
# torch's own multimem all-reduce as a reference point for latency
x = symm_mem.empty(1142575 + 1, dtype=torch.float32, device=dev)[:1142572]
symm_mem.rendezvous(x, dist.group.WORLD.group_name)
for name, fn in (("multimem_all_reduce_", lambda: torch.ops.symm_mem.multimem_all_reduce_(x, "sum", dist.group.WORLD.group_name)),
                 ("two_shot_all_reduce_", lambda: torch.ops.symm_mem.two_shot_all_reduce_(x, "sum", dist.group.WORLD.group_name)),
                 ("one_shot_all_reduce", lambda: torch.ops.symm_mem.one_shot_all_reduce(x, "sum", dist.group.WORLD.group_name)),
                 ("nccl all_reduce", lambda: dist.all_reduce(x))):
    try:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            fn()
        b.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"{name}: {a.elapsed_time(b) / 200 * 1000:.1f} us per 4.57 MB all-reduce", flush=True)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
dist.destroy_process_group()
