"""Worker of test_multi_gpu_count_boards (one process per GPU under torchrun): the per-frame count exchange of
SURVEY.md 8e through count boards -- one kernel stores the counts into every rank's board over NVLink (CUDA IPC
mappings), waits for everybody's and derives the offsets -- checked against the CPU oracle's counts of every shard,
and against the NCCL all-gather exchange on the same frames."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib  # noqa: E402
from astre_b200 import GpuContext, gpu, multi, scenes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    oracle = oracle_lib.load()
    for scene, rank_major in ((scenes.config3(n), False), (scenes.config5(n_worlds=8 * world, entities_per_world=2048), True)):
        shards = [scene.shard(r, world) for r in range(world)]
        want = np.stack([oracle.frame(s)[2] for s in shards])              # [world, n_counts]
        mine = shards[rank]
        n_counts = mine.n_worlds * mine.views_per_world
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        with GpuContext(mine.n, max_views=mine.views_per_world, device=local) as ctx:
            ctx.set_stream(stream.cuda_stream)
            scenes.load_into(ctx, mine)
            ex = multi.PeerCountExchange(ctx, n_counts, rank_major=rank_major)
            for k in range(23):                  # more frames than board slots; ranks are free to drift
                if k == 9:                       # every rank's scene changes at the same frame
                    for s in shards:
                        s.pos = (s.pos * np.float32(0.5)).astype(np.float32)
                    ctx.update_trs_range(0, mine.n, mine.pos, mine.rot, mine.scale)
                    want = np.stack([oracle.frame(s)[2] for s in shards])
                if k == 15 and rank == world - 1:
                    torch.cuda._sleep(200_000_000)          # ~0.1 s: one rank falls behind, the others fill their 4 slots
                ex.before_frame()
                ctx.frame(gpu.FRAME_ALL | (gpu.FRAME_GRAPH if k % 4 else 0))
                ex.after_frame()
                if k in (0, 8, 9, 14, 22):
                    table, off, tot = ex.last()
                    assert np.array_equal(table, want), (rank, k, table.tolist(), want.tolist())
                    h_off, h_tot = gpu.global_offsets(table, rank, rank_major)
                    assert np.array_equal(off, h_off) and np.array_equal(tot, h_tot), (rank, k)
            # the NCCL exchange on the same scene gives the same table and offsets
            ex.finish()
            torch.cuda.synchronize()
            nc = multi.CountExchange(ctx, n_counts, rank_major=rank_major)
            for _ in range(3):
                nc.before_frame()
                ctx.frame(gpu.FRAME_ALL | gpu.FRAME_GRAPH)
                nc.after_frame()
            t2, o2, s2 = nc.last()
            assert np.array_equal(t2, want) and np.array_equal(o2, off) and np.array_equal(s2, tot)
            nc.finish()
            torch.cuda.synchronize()
            ctx.set_counts_mirror(ex.slot[0].data_ptr(), ex.slot[1].data_ptr())      # back to the board exchange's slots
            for _ in range(2):
                ex.before_frame()
                ctx.frame(gpu.FRAME_ALL | gpu.FRAME_GRAPH)
                ex.after_frame()
            assert np.array_equal(ex.last()[0], want)
            ex.close()
            assert ctx.count_board_status() == (25, 0), ctx.count_board_status()
            ctx.set_stream(None)
        dist.barrier()
    if rank == 0:
        print(f"BOARD_WORKER_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
