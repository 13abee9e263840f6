#!/usr/bin/env python
"""Host -> device copy bandwidth of every rank when all ranks copy at once (the input prefetch of the e2e path moves
~60 MB of crops per step and rank): torchrun --nproc-per-node N tools/h2d_bw.py"""
import os, time, torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
n = 60 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out = [None] * world
if world > 1:
    dist.all_gather_object(out, (rank, ms))
else:
    out = [(rank, ms)]
if rank == 0:
    print("60 MB H2D, all ranks at once: " + ", ".join(f"rank {r}: {m:.2f} ms = {n / m / 1e6:.1f} GB/s" for r, m in sorted(out)))
if world > 1:
    dist.destroy_process_group()
