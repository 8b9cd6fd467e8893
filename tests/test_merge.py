"""f3: the globally key-sorted draw list across shards.

CPU part: the oracle's merge against an independent numpy statement of the order
(key >> slot_bits, shard, key & slot_mask).  GPU part (one GPU holds every shard, in separate
contexts): astre_gpu_merge_lists against the oracle, and against the property that the merged list
of G range shards, with (shard, local slot) mapped back to global slots, IS the sorted list of the
unsharded scene."""
import numpy as np
import pytest

import oracle_lib
from astre_b200 import scenes

SLOT_BITS = 26


def np_merge(lists):
    keys = np.concatenate(lists) if lists else np.empty(0, np.uint64)
    src = np.concatenate([np.full(len(a), r, np.uint8) for r, a in enumerate(lists)])
    cls = keys >> np.uint64(SLOT_BITS)
    local = keys & np.uint64((1 << SLOT_BITS) - 1)
    order = np.lexsort((local, src, cls))
    return keys[order], src[order]


def random_lists(rng, g, n_max, n_classes):
    """Sorted lists of unique keys with many class collisions across (and inside) lists."""
    out = []
    for _ in range(g):
        n = int(rng.integers(0, n_max + 1))
        cls = rng.integers(0, n_classes, n).astype(np.uint64) * np.uint64(977)
        slot = rng.choice(1 << 20, n, replace=False).astype(np.uint64)
        out.append(np.sort((cls << np.uint64(SLOT_BITS)) | slot))
    return out


@pytest.mark.parametrize("g,n_max,n_classes", [(1, 100, 5), (2, 1000, 7), (3, 5000, 1), (8, 3000, 50), (8, 20, 3), (4, 0, 1)])
def test_oracle_merge_matches_numpy(oracle, g, n_max, n_classes):
    rng = np.random.default_rng(g * 1000 + n_max)
    lists = random_lists(rng, g, n_max, n_classes)
    keys, src = oracle.merge_lists(lists, SLOT_BITS)
    k2, s2 = np_merge(lists)
    assert np.array_equal(keys, k2) and np.array_equal(src, s2)


def test_oracle_merge_of_range_shards_is_the_unsharded_list(oracle):
    scene = scenes.config3(60_000)
    _, full_keys, _ = oracle.frame(scene)
    g = 4
    lists, bases = [], []
    per = (-(-scene.n // g) + 3) & ~3
    for r in range(g):
        _, k, _ = oracle.frame(scene.shard(r, g))
        lists.append(k)
        bases.append(min(r * per, scene.n))
    keys, src = oracle.merge_lists(lists, SLOT_BITS)
    glob = keys + np.array(bases, np.uint64)[src]
    assert np.array_equal(glob, full_keys)


# ---- GPU ---------------------------------------------------------------------

def _gpu_shard_lists(scene, g):
    from astre_b200 import GpuContext
    ctxs, ptrs, counts, host = [], [], [], []
    per = max(scene.shard(0, g).n, 1)
    for r in range(g):
        sh = scene.shard(r, g)
        ctx = GpuContext(per, max_views=scene.views_per_world, max_keys=per * scene.views_per_world)   # same max_keys on every "rank"
        scenes.load_into(ctx, sh)
        ctx.frame()
        ctx.publish_keys(r & 1)
        host.append(ctx.read_all_keys())
        counts.append(len(host[-1]))
        ptrs.append(ctx.device_published_keys_ptr(r & 1))
        ctxs.append(ctx)
    return ctxs, ptrs, counts, host


@pytest.mark.gpu
@pytest.mark.parametrize("n,g", [(200_000, 2), (1_000_000, 8), (50_000, 3), (4_000, 8), (300, 5)])
def test_gpu_merge_of_range_shards(oracle, n, g):
    scene = scenes.config3(n)
    ctxs, ptrs, counts, host = _gpu_shard_lists(scene, g)
    try:
        o_keys, o_src = oracle.merge_lists(host, SLOT_BITS)
        ctxs[0].merge_lists(ptrs, counts)
        keys, src = ctxs[0].read_merged(0, len(o_keys))
        assert np.array_equal(keys, o_keys) and np.array_equal(src, o_src)
        # property: with (shard, local slot) -> global slot this is the unsharded frame's sorted list
        per = (-(-scene.n // g) + 3) & ~3
        bases = np.array([min(r * per, scene.n) for r in range(g)], np.uint64)
        _, full_keys, _ = oracle.frame(scene)
        assert np.array_equal(keys + bases[src], full_keys)
        # a rank's slice only
        total = len(o_keys)
        lo, hi = total // 3, total // 3 + total // 2
        ctxs[0].merge_lists(ptrs, counts, out_first=lo, out_count=hi - lo)
        k2, s2 = ctxs[0].read_merged(0, hi - lo)
        assert np.array_equal(k2, o_keys[lo:hi]) and np.array_equal(s2, o_src[lo:hi])
        # the peer-to-peer flow (here every "peer" buffer is on the same GPU): each rank ranks the
        # samples in its own list, then any rank builds the list from the published cut tables
        for r in range(g):
            ctxs[r].merge_prepare(r, ptrs, counts)
            ctxs[r].sync()
        last = ctxs[g - 1]
        last.merge_published(ptrs, counts)
        k3, s3 = last.read_merged(0, total)
        assert np.array_equal(k3, o_keys) and np.array_equal(s3, o_src)
        last.merge_published(ptrs, counts, out_first=lo, out_count=hi - lo)
        k4, s4 = last.read_merged(0, hi - lo)
        assert np.array_equal(k4, o_keys[lo:hi]) and np.array_equal(s4, o_src[lo:hi])
        # the same flow with the counts left on the device (the all-gathered [G, F] table of per-view counts)
        import torch
        table = torch.from_numpy(np.stack([c.read_counts() for c in ctxs]).astype(np.int32)).cuda()
        for r in range(g):
            ctxs[r].merge_prepare_dev(r, ptrs, table.data_ptr(), scene.views_per_world)
            ctxs[r].sync()
        last.merge_published_dev(ptrs)
        assert last.merged_count() == total
        k5, s5 = last.read_merged(0, total)
        assert np.array_equal(k5, o_keys) and np.array_equal(s5, o_src)
        # no NCCL, no host: flags and counts in the publish buffers, polled by one-warp kernels.  Every "rank"
        # enqueues its whole chain; the chains wait for one another on the device (each context has its own stream)
        p0 = [c.device_published_keys_ptr(0) for c in ctxs]
        p1 = [c.device_published_keys_ptr(1) for c in ctxs]
        for r in range(g):      # size every context's buffers first: in ONE process a reallocation (cudaFree waits for the
            ctxs[r].merge_prepare_dev(r, ptrs, table.data_ptr(), scene.views_per_world)     # device) would stall the other
            ctxs[r].merge_published_dev(ptrs)                                               # "ranks" chains mid-wait
            ctxs[r].sync()
        for epoch in (1, 2, 3):
            for r in range(g):
                ctxs[r].merge_p2p(r, p0, p1, epoch)
            for r in (0, g - 1):
                assert ctxs[r].merged_count() == total
                k7, s7 = ctxs[r].read_merged(0, total)
                assert np.array_equal(k7, o_keys) and np.array_equal(s7, o_src), f"flags flow, epoch {epoch}, rank {r}"
        per = -(-total // g)
        for r in (0, g - 1):
            ctxs[r].merge_prepare_dev(r, ptrs, table.data_ptr(), scene.views_per_world)
            ctxs[r].merge_published_dev(ptrs, slice_rank=r, slice_world=g)
            first = min(r * per, total)
            cnt = min(per, total - first)
            assert ctxs[r].merged_count() == cnt
            k6, s6 = ctxs[r].read_merged(0, cnt)
            assert np.array_equal(k6, o_keys[first:first + cnt]) and np.array_equal(s6, o_src[first:first + cnt])
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.gpu
def test_gpu_merge_dense_ties_and_empty_lists(oracle):
    """All entities visible in every view (long equal-class runs across shards) and shards with no keys."""
    from astre_b200 import GpuContext, gpu
    s = scenes.config1(120_000)
    s.pos = (s.pos * np.float32(0.02)).astype(np.float32)
    s.pos[:, 2] -= np.float32(40.0)
    s.mesh[:] = 1
    s.shader[:] = 1
    s.flags[:] = gpu.make_flags(np.ones(s.n, bool), np.full(s.n, 5, np.uint8))
    s.flags[: s.n // 4] = 0            # shard 0 of 4 emits nothing
    ctxs, ptrs, counts, host = _gpu_shard_lists(s, 4)
    try:
        assert counts[0] == 0 and min(counts[1:]) > 0
        o_keys, o_src = oracle.merge_lists(host, SLOT_BITS)
        ctxs[1].merge_lists(ptrs, counts)
        keys, src = ctxs[1].read_merged(0, len(o_keys))
        assert np.array_equal(keys, o_keys) and np.array_equal(src, o_src)
        # all lists empty
        ctxs[1].merge_lists(ptrs, [0, 0, 0, 0])
        k0, _ = ctxs[1].read_merged(0, 0)
        assert len(k0) == 0
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.gpu
def test_gpu_merge_rejects_worlds_and_bad_arguments():
    from astre_b200 import AstreError, GpuContext
    s = scenes.config5(8, 256)
    with GpuContext(s.n, max_views=1) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        ctx.publish_keys(0)
        with pytest.raises(AstreError):
            ctx.merge_lists([ctx.device_published_keys_ptr(0)], [1])
    with GpuContext(1024, max_views=1) as ctx:
        with pytest.raises(AstreError):
            ctx.merge_lists([None] * 9, [0] * 9)
        with pytest.raises(AstreError):
            ctx.merge_lists([None], [5])


@pytest.mark.gpu
def test_multi_gpu_merge():
    """Two processes, two GPUs: peers' lists read over NVLink (CUDA IPC) and the NCCL all-gather mode."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(here, "merge_worker.py"), "400000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MERGE_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_multi_gpu_count_boards():
    """One process per GPU (all of the box's GPUs, at least two): the count exchange as stores into peer memory."""
    import os
    import subprocess
    import sys
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n_gpus, 8)), "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(here, "board_worker.py"), "400000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "BOARD_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def _merge_kat():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "merge_kat_hand_derived.json")) as f:
        d = json.load(f)
    lists = [np.array([int(x, 16) for x in l], np.uint64) for l in d["lists"]]
    return lists, np.array([int(x, 16) for x in d["keys"]], np.uint64), np.array(d["sources"], np.uint8)


def test_oracle_merge_hand_derived_vector(oracle):
    lists, keys, src = _merge_kat()
    k, s = oracle.merge_lists(lists, SLOT_BITS)
    assert np.array_equal(k, keys) and np.array_equal(s, src)


@pytest.mark.gpu
def test_gpu_merge_hand_derived_vector():
    """The CUDA merge against the hand-derived vector (no oracle involved)."""
    import torch
    from astre_b200 import GpuContext
    lists, keys, src = _merge_kat()
    dev = [torch.from_numpy(l.view(np.int64)).cuda() for l in lists]
    with GpuContext(1024, max_views=1) as ctx:
        ctx.merge_lists([t.data_ptr() if t.numel() else None for t in dev], [len(l) for l in lists])
        k, s = ctx.read_merged(0, len(keys))
    assert np.array_equal(k, keys) and np.array_equal(s, src)


@pytest.mark.gpu
def test_gpu_merge_fuzz_random_lists(oracle):
    """Synthetic lists straight into the merge: 1..8 lists, empty to 60k keys, from one class for everything
    (ties across all lists over thousands of keys) to all-distinct classes, skewed sizes."""
    import torch
    from astre_b200 import GpuContext
    rng = np.random.default_rng(99)
    with GpuContext(1024, max_views=1) as ctx:
        for i in range(60):
            g = int(rng.integers(1, 9))
            n_classes = int(rng.choice([1, 2, 17, 1000, 1 << 20]))
            lists = []
            for _ in range(g):
                n = int(rng.choice([0, 1, 127, 128, 129, 5000, 60000])) if rng.random() < 0.7 else int(rng.integers(0, 60001))
                cls = rng.integers(0, n_classes, n).astype(np.uint64)
                slot = rng.choice(1 << 22, n, replace=False).astype(np.uint64)
                lists.append(np.sort((cls << np.uint64(SLOT_BITS)) | slot))
            o_keys, o_src = oracle.merge_lists(lists, SLOT_BITS)
            dev = [torch.from_numpy(l.view(np.int64)).cuda() for l in lists]
            ctx.merge_lists([t.data_ptr() if t.numel() else None for t in dev], [len(l) for l in lists])
            k, s = ctx.read_merged(0, len(o_keys))
            assert np.array_equal(k, o_keys) and np.array_equal(s, o_src), f"case {i}: g={g} classes={n_classes} sizes={[len(l) for l in lists]}"
            if len(o_keys) > 10:
                lo, hi = len(o_keys) // 4, len(o_keys) // 4 + len(o_keys) // 3
                ctx.merge_lists([t.data_ptr() if t.numel() else None for t in dev], [len(l) for l in lists], out_first=lo, out_count=hi - lo)
                k, s = ctx.read_merged(0, hi - lo)
                assert np.array_equal(k, o_keys[lo:hi]) and np.array_equal(s, o_src[lo:hi]), f"case {i} (window)"
