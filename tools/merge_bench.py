"""Timing of the f3 global draw-list merge (astre_gpu_merge_lists).

  python tools/merge_bench.py --lists 8 --entities 16777216        one GPU holds G lists (kernel cost, all local)
  torchrun --nproc-per-node N tools/merge_bench.py --entities ...   one list per GPU: p2p (NVLink reads inside the
                                                                    merge kernel) against nccl (all-gather + merge)
Checks on the device that the result is ordered (class, source, slot); parity with the oracle is tests/test_merge.py."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astre_b200 import GpuContext, multi, scenes  # noqa: E402


def check_order(keys, src):
    cls = keys >> 26
    rank_slot = (src.to(torch.int64) << 26) | (keys & ((1 << 26) - 1))
    ok = (cls[1:] > cls[:-1]) | ((cls[1:] == cls[:-1]) & (rank_slot[1:] > rank_slot[:-1]))
    return bool(ok.all().item())


def merged_tensors(ctx, n, device):
    kp, sp = ctx.device_merged_ptrs()
    return multi.device_view(kp, n, "<i8", device), multi.device_view(sp, n, "|u1", device)


def local(args):
    dev = torch.device("cuda", 0)
    n, g = args.entities, args.lists
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lists, counts = [], []
    with GpuContext(n, max_views=5) as ctx:
        ctx.set_stream(stream.cuda_stream)
        for r in range(g):
            scenes.load_into(ctx, scenes.config4(n * g, r * n, (r + 1) * n))
            ctx.frame()
            ctx.publish_keys(0)
            total = ctx.read_total()
            lists.append(multi.device_view(ctx.device_published_keys_ptr(0), total, "<i8", dev).clone())
            counts.append(total)
        ptrs = [t.data_ptr() for t in lists]
        ms = []
        for i in range(args.warmup + args.steps):
            ctx.merge_lists(ptrs, counts)
            if i >= args.warmup:
                ms.append(ctx.last_merge_ms())
        total = sum(counts)
        keys, src = merged_tensors(ctx, total, dev)
        ok = check_order(keys, src)
        t = float(np.mean(ms))
        print(json.dumps({"mode": "local", "lists": g, "keys": total, "merge_ms": t, "min_ms": float(np.min(ms)),
                          "GBps": total * 17 / t / 1e6, "ordered": ok}))


def distributed(args):
    import torch.distributed as dist
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n = args.entities
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    max_keys = n * 2
    with GpuContext(n, max_views=5, max_keys=max_keys, device=lr) as ctx:
        ctx.set_stream(stream.cuda_stream)
        scenes.load_into(ctx, scenes.config4(n * world, rank * n, (rank + 1) * n))
        frame = [0]
        out = {}

        def timed(fn):
            for _ in range(args.warmup):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                fn()
            e1.record(stream)
            dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        counts_dev = multi.device_view(ctx.device_counts_ptr(), 5, "<i4", dev)

        def frame_only():
            ctx.frame()
            multi.exchange_counts(counts_dev).cpu()
        out["frame+counts_ms"] = timed(frame_only)
        for mode in ("p2p", "nccl"):
            g = multi.GlobalDrawList(ctx, 5, max_keys, mode=mode)
            for sliced in (False, True):
                def step():
                    ctx.frame()
                    g.build(frame[0], sliced=sliced)
                    frame[0] += 1
                name = f"{mode}{'_sliced' if sliced else ''}"
                out[name + "_ms"] = timed(step)
                out[name + "_merge_kernels_ms"] = ctx.last_merge_ms()
                all_counts, first, count = g.build(frame[0], sliced=sliced); frame[0] += 1
                keys, src = merged_tensors(ctx, count, dev)
                ok = torch.tensor([1 if check_order(keys, src) else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                out[name + "_ordered"] = bool(ok.item())
                out["keys_total"] = int(all_counts.sum())
        # peer to peer with the key counts left on the device: no host round trip per frame
        g = multi.GlobalDrawList(ctx, 5, max_keys, mode="p2p")
        for sliced in (False, True):
            def step():
                ctx.frame()
                g.build_enqueue(frame[0], sliced=sliced)
                frame[0] += 1
            name = f"p2p_dev{'_sliced' if sliced else ''}"
            out[name + "_ms"] = timed(step)
            out[name + "_merge_kernels_ms"] = ctx.last_merge_ms()
            count = ctx.merged_count()
            keys, src = merged_tensors(ctx, count, dev)
            ok = torch.tensor([1 if check_order(keys, src) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            out[name + "_ordered"] = bool(ok.item())

        g2 = multi.GlobalDrawList(ctx, 5, max_keys, mode="p2p")      # no NCCL: flags + counts in the publish buffers
        for sliced in (False, True):
            def step():
                ctx.frame()
                g2.build_flags(sliced=sliced)
            name = f"p2p_flags{'_sliced' if sliced else ''}"
            out[name + "_ms"] = timed(step)
            out[name + "_merge_kernels_ms"] = ctx.last_merge_ms()
            count = ctx.merged_count()
            keys, src = merged_tensors(ctx, count, dev)
            ok = torch.tensor([1 if check_order(keys, src) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            out[name + "_ordered"] = bool(ok.item())

        def frame_no_sync():
            ctx.frame()
            multi.exchange_counts(counts_dev)
        out["frame+counts_nosync_ms"] = timed(frame_no_sync)
        dist.barrier()
    if rank == 0:
        out.update({"mode": "distributed", "gpus": world, "entities_per_gpu": n})
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lists", type=int, default=8)
    ap.add_argument("--entities", type=int, default=4_194_304)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    distributed(a) if "WORLD_SIZE" in os.environ else local(a)
