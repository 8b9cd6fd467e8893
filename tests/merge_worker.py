"""Worker of test_multi_gpu_merge (one process per GPU under torchrun): builds the global
key-sorted draw list (f3) in both exchange modes and checks it against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib  # noqa: E402
from astre_b200 import GpuContext, multi, scenes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    oracle = oracle_lib.load()
    scene = scenes.config3(n)
    lists = [oracle.frame(scene.shard(r, world))[1] for r in range(world)]
    o_keys, o_src = oracle.merge_lists(lists, 26)

    shard = scene.shard(rank, world)
    max_keys = scene.shard(0, world).n * scene.views_per_world      # the same on every rank (publish buffer layout)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    with GpuContext(shard.n, max_views=scene.views_per_world, max_keys=max_keys, device=local) as ctx:
        ctx.set_stream(stream.cuda_stream)
        scenes.load_into(ctx, shard)
        frame = 0
        for mode in ("p2p", "nccl"):
            g = multi.GlobalDrawList(ctx, scene.views_per_world, max_keys, mode=mode)
            for sliced in (False, True, False):
                ctx.frame()
                all_counts, first, count = g.build(frame, sliced=sliced)
                frame += 1
                assert [int(c) for c in all_counts.sum(axis=1)] == [len(a) for a in lists], (mode, all_counts)
                keys, src = ctx.read_merged(0, count)
                assert np.array_equal(keys, o_keys[first:first + count]), (mode, sliced, "keys differ")
                assert np.array_equal(src, o_src[first:first + count]), (mode, sliced, "sources differ")
                if sliced:
                    assert count < len(o_keys)
            if mode == "p2p":                     # no NCCL at all: flags and counts polled over NVLink
                for sliced in (False, True, True, False):
                    ctx.frame()
                    g.build_flags(sliced=sliced)
                    per = -(-len(o_keys) // world)
                    first = min(rank * per, len(o_keys)) if sliced else 0
                    count = min(per, len(o_keys) - first) if sliced else len(o_keys)
                    assert ctx.merged_count() == count, (ctx.merged_count(), count)
                    keys, src = ctx.read_merged(0, count)
                    assert np.array_equal(keys, o_keys[first:first + count]), ("flags", sliced, "keys differ")
                    assert np.array_equal(src, o_src[first:first + count]), ("flags", sliced, "sources differ")
            if mode == "p2p":                     # same flow, key counts never leave the device
                for sliced in (True, False):
                    ctx.frame()
                    g.build_enqueue(frame, sliced=sliced)
                    frame += 1
                    per = -(-len(o_keys) // world)
                    first = min(rank * per, len(o_keys)) if sliced else 0
                    count = min(per, len(o_keys) - first) if sliced else len(o_keys)
                    assert ctx.merged_count() == count, (ctx.merged_count(), count)
                    keys, src = ctx.read_merged(0, count)
                    assert np.array_equal(keys, o_keys[first:first + count]), ("dev", sliced, "keys differ")
                    assert np.array_equal(src, o_src[first:first + count]), ("dev", sliced, "sources differ")
        dist.barrier()
    if rank == 0:
        print(f"MERGE_WORKER_OK world={world} keys={len(o_keys)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
