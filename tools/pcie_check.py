"""Concurrent pinned D2H bandwidth per rank (run under torchrun): is the end-to-end N>1 number bound by the host?"""
import os, sys, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 454_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for mode in ("alone", "together"):
    for r in range(dist.get_world_size() if mode == "alone" else 1):
        dist.barrier(); torch.cuda.synchronize()
        if mode == "together" or r == dist.get_rank():
            for _ in range(2): h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): h.copy_(d, non_blocking=True)
            b.record(); torch.cuda.synchronize()
            print(f"rank {dist.get_rank()} {mode}: D2H {5 * n / a.elapsed_time(b) / 1e6:.1f} GB/s", flush=True)
        dist.barrier()
dist.destroy_process_group()
