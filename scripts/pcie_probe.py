#!/usr/bin/env python
"""Raw host<->device copy bandwidth of the box with N ranks copying at once (torchrun --nproc-per-node N): the ceiling of
bench.py's end-to-end legs, which move 2 GiB per stream in each direction through pinned host memory.
Prints per-rank and aggregate GB/s for D2H alone, H2D alone and both directions together."""
import os
import time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GIB = 1 << 30
dev = torch.empty(2 * GIB, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(2 * GIB, dtype=torch.uint8, device="cuda")
host = torch.empty(2 * GIB, dtype=torch.uint8, pin_memory=True)
host2 = torch.empty(2 * GIB, dtype=torch.uint8, pin_memory=True)
host.zero_(); host2.zero_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 6


def run(d2h, h2d):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(REPS):
        if d2h:
            with torch.cuda.stream(s1):
                host.copy_(dev, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2):
                dev2.copy_(host2, non_blocking=True)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    t = torch.tensor([sec], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name, a, b in (("D2H", True, False), ("H2D", False, True), ("both", True, True)):
    run(a, b)
    sec = run(a, b)
    nbytes = REPS * 2 * GIB * (int(a) + int(b))
    if rank == 0:
        print(f"{name:5s} {world} ranks: {nbytes / sec / 1e9:7.1f} GB/s per rank, {world * nbytes / sec / 1e9:7.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
