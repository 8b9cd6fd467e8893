#!/usr/bin/env python
"""Aggregate host->device copy bandwidth of the box when every rank pushes a frame's worth of TRS at once
(torchrun, one rank per GPU): the ceiling of the end-to-end 'all entities dirty' figure at N GPUs.
Usage: python -m torch.distributed.run --nproc-per-node N tools/h2d_ceiling.py [bytes_per_rank]"""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 16_777_216 * 40
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
back = torch.empty(nbytes // 8, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms_h2d = timed(lambda: dev.copy_(host, non_blocking=True))
ms_d2h = timed(lambda: back.copy_(dev[: nbytes // 8], non_blocking=True))
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_rank": nbytes, "h2d_ms_max_over_ranks": ms_h2d,
                      "h2d_GBps_per_rank": nbytes / ms_h2d / 1e6, "h2d_GBps_aggregate": world * nbytes / ms_h2d / 1e6,
                      "d2h_GBps_per_rank": (nbytes // 8) / ms_d2h / 1e6,
                      "numa": open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else None,
                      "cpus": len(os.sched_getaffinity(0))}))
if world > 1:
    dist.destroy_process_group()
