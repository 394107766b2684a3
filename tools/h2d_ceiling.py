#!/usr/bin/env python
"""Host->device ceiling of the box with N ranks copying at once (the limiter of bench.py's e2e leg at 8 GPUs).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py [MB]
Every rank copies a pinned buffer of MB megabytes to its GPU back to back for ~2 s (two streams, like the library's
double-buffered staging); rank 0 prints per-rank and aggregate GB/s, alone (ranks one after the other) and together."""
import os
import sys
import time

import torch
import torch.distributed as dist

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rank, ws, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=dev)
h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = [torch.empty(mb << 20, dtype=torch.uint8, device=dev) for _ in range(2)]
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]


def run(seconds):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < seconds:
        for s in range(2):
            with torch.cuda.stream(streams[s]):
                d[s].copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        reps += 2
    return reps * mb / 1024.0 / (time.perf_counter() - t0)


run(0.3)
alone = torch.zeros(ws, device=dev)
for r in range(ws):
    if ws > 1:
        dist.barrier()
    if r == rank:
        alone[r] = run(1.0)
if ws > 1:
    dist.all_reduce(alone)
    dist.barrier()
together = torch.zeros(ws, device=dev)
together[rank] = run(2.0)
if ws > 1:
    dist.all_reduce(together)
if rank == 0:
    print("H2D pinned, %d MB buffers, %d ranks" % (mb, ws))
    print("  alone    GB/s per rank:", " ".join("%.1f" % v for v in alone.tolist()))
    print("  together GB/s per rank:", " ".join("%.1f" % v for v in together.tolist()), " aggregate %.1f" % float(together.sum()))
if ws > 1:
    dist.destroy_process_group()
