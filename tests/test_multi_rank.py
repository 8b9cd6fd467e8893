"""N>1 host-side path on CPU: two gloo ranks shard a scene, exchange visible counts and place their
key segments in the global draw list.  The per-rank frame is played by the oracle here (there is no
GPU in this container); the exchange and the offset arithmetic are the product's (astre_b200.multi)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib
from astre_b200 import multi, scenes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, kind, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        oracle = oracle_lib.load()
        if kind == "flat":
            scene = scenes.config3(40_000)
            lo, hi = multi.shard_range(scene.n, rank, world_size)
            shard = scene.shard(rank, world_size)
            assert shard.n == hi - lo
        else:
            scene = scenes.config5(n_worlds=8, entities_per_world=1024)
            wlo, whi = multi.shard_worlds(scene.n_worlds, rank, world_size)
            shard = scene.shard(rank, world_size)
            assert shard.n_worlds == whi - wlo
        _, keys, counts = oracle.frame(shard)                 # stand-in for the rank's GPU frame
        all_counts = multi.exchange_counts(torch.from_numpy(counts.astype(np.int32)))
        offsets, totals = multi.segment_offsets(all_counts, rank, rank_major=(kind != "flat"))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), keys=keys, counts=counts, offsets=offsets, totals=totals,
                 all_counts=all_counts.numpy())
    finally:
        dist.destroy_process_group()


def _run(kind, tmp_path, world_size=2):
    port = _free_port()
    mp.spawn(_worker, args=(world_size, port, kind, str(tmp_path)), nprocs=world_size, join=True)
    return [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world_size)]


def test_flat_shards_assemble_the_global_list(tmp_path, oracle):
    res = _run("flat", tmp_path)
    scene = scenes.config3(40_000)
    F = scene.views_per_world
    total = int(sum(r["counts"].sum() for r in res))
    assert np.array_equal(res[0]["totals"], res[1]["totals"]) and int(res[0]["totals"].sum()) == total
    # place every rank's per-view segments at its offsets: the list must come out dense and ordered (view, rank, key)
    glob = np.full(total, np.uint64(0xFFFFFFFFFFFFFFFF))
    owner = np.full(total, -1)
    for rank, r in enumerate(res):
        start = 0
        for v in range(F):
            c = int(r["counts"][v])
            o = int(r["offsets"][v])
            assert np.all(owner[o:o + c] == -1)
            glob[o:o + c] = r["keys"][start:start + c]
            owner[o:o + c] = rank
            start += c
    assert np.all(owner >= 0)
    expect_views = np.concatenate([np.full(int(res[0]["totals"][v]), v) for v in range(F)])
    assert np.array_equal((glob >> np.uint64(59)).astype(np.int64), expect_views)
    # same visible set as the unsharded scene: a shard's key carries its local slot, so add the range start back
    _, ref_keys, ref_counts = oracle.frame(scene)
    assert np.array_equal(res[0]["totals"].astype(np.uint32), ref_counts)
    slot_mask = np.uint64((1 << 26) - 1)
    rebuilt = []
    for rank, r in enumerate(res):
        lo, _ = multi.shard_range(scene.n, rank, 2)
        rebuilt.append((r["keys"] & ~slot_mask) | ((r["keys"] & slot_mask) + np.uint64(lo)))
    assert np.array_equal(np.sort(np.concatenate(rebuilt)), ref_keys)
    # inside a view, rank 0's segment precedes rank 1's and each segment is sorted
    for v in range(F):
        o0, o1 = int(res[0]["offsets"][v]), int(res[1]["offsets"][v])
        assert o1 == o0 + int(res[0]["counts"][v])
        for r, o in ((res[0], o0), (res[1], o1)):
            seg = glob[o:o + int(r["counts"][v])]
            assert np.all(seg[1:] > seg[:-1])


def test_world_shards(tmp_path, oracle):
    res = _run("worlds", tmp_path)
    scene = scenes.config5(n_worlds=8, entities_per_world=1024)
    _, ref_keys, ref_counts = oracle.frame(scene)
    got = np.concatenate([r["counts"] for r in res])
    assert np.array_equal(got, ref_counts)                      # worlds [0,4) on rank 0, [4,8) on rank 1
    # whole-world shards: the global list is the concatenation of the ranks' lists (rank-major = world-major)
    n0 = int(res[0]["counts"].sum())
    assert res[0]["offsets"].tolist() == np.concatenate([[0], np.cumsum(res[0]["counts"])[:-1]]).tolist()
    assert res[1]["offsets"].tolist() == (n0 + np.concatenate([[0], np.cumsum(res[1]["counts"])[:-1]])).tolist()
    assert np.array_equal(res[1]["totals"], res[1]["counts"])                 # segment lengths
    # a shard numbers its worlds from 0: shift the world field back and compare with the unsharded list
    wb_shard, wb_full = 2, 3
    rebuilt = []
    for rank, r in enumerate(res):
        k = r["keys"]
        world = (k >> np.uint64(64 - wb_shard)) + np.uint64(rank * 4)
        rest = k & np.uint64((1 << (64 - wb_shard)) - 1)
        # fields below the world keep their order; re-pack for the 3-bit world field of the full scene
        local = rest & np.uint64((1 << (26 - wb_shard)) - 1)
        upper = rest >> np.uint64(26 - wb_shard)
        rebuilt.append((((world << np.uint64(38)) | upper) << np.uint64(26 - wb_full)) | local)
    assert np.array_equal(np.concatenate(rebuilt), ref_keys)


def test_shard_helpers():
    assert multi.shard_range(10, 0, 4) == (0, 4) and multi.shard_range(10, 3, 4) == (10, 10)
    assert multi.shard_worlds(4096, 7, 8) == (3584, 4096)
    a = torch.tensor([[3, 1], [2, 5], [0, 7]], dtype=torch.int32)
    off, tot = multi.segment_offsets(a, 2)
    assert tot.tolist() == [5, 13] and off.tolist() == [5, 5 + 6]
    off, seg = multi.segment_offsets(a, 2, rank_major=True)
    assert off.tolist() == [11, 11] and seg.tolist() == [0, 7]
    off, seg = multi.segment_offsets(a, 1, rank_major=True)
    assert off.tolist() == [4, 6] and seg.tolist() == [2, 5]


def test_shard_subtrees_partition_properties():
    """Every entity lands on exactly one rank, together with its whole ancestor chain; loads are balanced."""
    from astre_b200 import scenes, multi
    s = scenes.config2(40_000, levels=6)
    parts = multi.shard_subtrees(s.parent, 3)
    seen = np.concatenate([m for m, _ in parts])
    assert np.array_equal(np.sort(seen), np.arange(s.n))
    for mine, lp in parts:
        glob_parent = s.parent[mine]
        has = glob_parent != 0xFFFFFFFF
        assert np.array_equal(mine[lp[has]], glob_parent[has])           # local parent is the same entity
        assert (lp[~has] == 0xFFFFFFFF).all()
    sizes = [len(m) for m, _ in parts]
    biggest_subtree = 2 ** 6
    assert max(sizes) - min(sizes) <= 2 * biggest_subtree
    with pytest.raises(ValueError):
        multi.shard_subtrees(np.array([1, 0], np.uint32), 2)             # cycle


def test_subtree_shards_reproduce_the_unsharded_world_matrices(oracle):
    from astre_b200 import scenes
    s = scenes.config2(30_000, levels=8)
    full_world, full_keys, full_counts = oracle.frame(s)
    total = 0
    for r in range(4):
        sh = s.shard(r, 4)
        w, k, c = oracle.frame(sh)
        assert np.array_equal(oracle_lib.bits(w), oracle_lib.bits(full_world[sh.global_index]))
        total += int(c.sum())
    assert total == int(full_counts.sum())
