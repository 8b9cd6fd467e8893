"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Bit-exact:
world matrices (0 ULP, NaN payloads aside), per-view counts and sorted draw keys."""
import numpy as np
import pytest

import np_reference
import oracle_lib
from astre_b200 import AstreError, GpuContext, gpu, scenes

pytestmark = pytest.mark.gpu


def run_scene(scene, flags=gpu.FRAME_ALL, max_keys=0):
    with GpuContext(max(scene.n, 1), max_views=max(scene.views_per_world, 1), max_keys=max_keys) as ctx:
        scenes.load_into(ctx, scene)
        ctx.frame(flags)
        world = ctx.read_world(0, scene.n)
        counts = ctx.read_counts()
        keys = ctx.read_all_keys()
        per_view = [[ctx.read_keys(v, w) for v in range(scene.views_per_world)] for w in range(scene.n_worlds)] \
            if scene.n_worlds * scene.views_per_world <= 64 else None
        return world, counts, keys, per_view


def check_scene(oracle, scene):
    o_world, o_keys, o_counts = oracle.frame(scene)
    world, counts, keys, per_view = run_scene(scene)
    assert np.array_equal(oracle_lib.bits(world), oracle_lib.bits(o_world)), \
        f"world matrices differ: max ulp {oracle_lib.ulp_diff(world, o_world).max()}"
    assert np.array_equal(counts, o_counts)
    assert np.array_equal(keys, o_keys)
    if per_view is not None:
        flat = np.concatenate([k for w in per_view for k in w]) if len(o_keys) else np.empty(0, np.uint64)
        assert np.array_equal(flat, o_keys)
    return len(o_keys)


@pytest.fixture(params=["thread-per-entity", "tiled"])
def small_flat_kernel(request, monkeypatch):
    """Flat frames below 262,144 entities take k_transform_cull_indexed; ASTRE_SMALL_FLAT=0 sends them through the tiled
    TMA kernel (k_frame_fused) like big frames.  Small scenes are run through BOTH."""
    if request.param == "tiled":
        monkeypatch.setenv("ASTRE_SMALL_FLAT", "0")
    else:
        monkeypatch.delenv("ASTRE_SMALL_FLAT", raising=False)
    return request.param


def test_config1_flat_10k(oracle, small_flat_kernel):
    assert check_scene(oracle, scenes.config1()) > 0


@pytest.mark.parametrize("n", [1, 3, 127, 128, 129, 1023, 1024, 1025, 4099])
def test_ragged_sizes(oracle, n, small_flat_kernel):
    check_scene(oracle, scenes.config1(n))


def test_config2_hierarchy_small(oracle):
    check_scene(oracle, scenes.config2(n=65_536 + 13))


def test_config2_hierarchy_1m(oracle):
    assert check_scene(oracle, scenes.config2()) > 0


def _bfs_forest(rng, n_roots, n_levels, n_total):
    """Random forest in breadth-first slot order: levels are contiguous, parents non-decreasing along a level, with
    childless nodes and uneven branching (the shape the cone kernel takes)."""
    sizes = [n_roots]
    rest = n_total - n_roots
    for l in range(1, n_levels):
        take = rest if l == n_levels - 1 else int(rest * rng.uniform(0.2, 0.6))
        sizes.append(max(take, 1))
        rest -= sizes[-1]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    parent = np.full(starts[-1], gpu.NO_PARENT, np.uint32)
    for l in range(1, n_levels):
        up = np.sort(rng.integers(0, sizes[l - 1], sizes[l]))          # some parents get many children, some none
        parent[starts[l]:starts[l + 1]] = (starts[l - 1] + up).astype(np.uint32)
    return parent


@pytest.mark.parametrize("cones", ["on", "off"])
def test_hierarchy_breadth_first_forests(oracle, cones, monkeypatch):
    """Breadth-first forests take k_hier_cones (a CTA per run of root subtrees); ASTRE_NO_CONES=1 takes the per-level
    launches.  Both must equal the oracle."""
    if cones == "off":
        monkeypatch.setenv("ASTRE_NO_CONES", "1")
    else:
        monkeypatch.delenv("ASTRE_NO_CONES", raising=False)
    rng = np.random.default_rng(41)
    for n_roots, n_levels, n_total in ((1, 2, 40), (3, 5, 700), (500, 4, 60_000), (4000, 9, 300_000), (100_000, 3, 250_000)):
        parent = _bfs_forest(rng, n_roots, n_levels, n_total)
        s = scenes.config2(len(parent), levels=2)
        s.parent = parent
        s.pos = (s.pos * np.float32(0.3)).astype(np.float32)
        check_scene(oracle, s)
        with GpuContext(s.n, max_views=1) as ctx:                        # transform only, and frames repeat
            scenes.load_into(ctx, s)
            ctx.frame(gpu.FRAME_TRANSFORM)
            ctx.frame(gpu.FRAME_TRANSFORM)
            o_world, _ = oracle.transform_hierarchy(s.pos, s.rot, s.scale, s.parent)
            assert np.array_equal(oracle_lib.bits(ctx.read_world(0, s.n)), oracle_lib.bits(o_world))


def test_hierarchy_arbitrary_slot_order(oracle):
    """Levels that are not contiguous slot ranges take the indexed kernel."""
    s = scenes.config2(n=20_000)
    rng = np.random.default_rng(7)
    perm = rng.permutation(s.n).astype(np.uint32)      # new slot of old entity i = perm[i]
    inv = np.empty_like(perm)
    inv[perm] = np.arange(s.n, dtype=np.uint32)
    for name in ("pos", "rot", "scale", "mesh", "shader", "flags"):
        setattr(s, name, np.ascontiguousarray(getattr(s, name)[inv]))
    old_parent = s.parent[inv]
    s.parent = np.where(old_parent == gpu.NO_PARENT, gpu.NO_PARENT, perm[np.minimum(old_parent, s.n - 1)]).astype(np.uint32)
    check_scene(oracle, s)


def test_hierarchy_subtree_shards(oracle):
    """8e: a hierarchy sharded by root subtree -- every shard reproduces the unsharded world matrices of
    its entities, and the shards' visible counts add up."""
    s = scenes.config2(200_000, levels=8)
    full_world, _, full_counts = oracle.frame(s)
    total = 0
    for r in range(3):
        sh = s.shard(r, 3)
        check_scene(oracle, sh)
        world, counts, _, _ = run_scene(sh)
        assert np.array_equal(oracle_lib.bits(world), oracle_lib.bits(full_world[sh.global_index]))
        total += int(counts.sum())
    assert total == int(full_counts.sum())


def test_config3_five_views_1m(oracle):
    assert check_scene(oracle, scenes.config3(1 << 20)) > 0


def test_config3_full_16m(oracle):
    """BASELINE.json's headline configuration at full size."""
    s = scenes.config3()
    n_keys = check_scene(oracle, s)
    assert n_keys > 100_000


def test_config4_full_64m(oracle):
    """BASELINE.json configs[3] at full size on one GPU: 67,108,864 entities, 5 views (26-bit slot field full)."""
    s = scenes.config4()
    assert s.n == 67_108_864
    assert check_scene(oracle, s) > 2_000_000


def test_config5_full_4096_worlds(oracle):
    """BASELINE.json configs[4] at full size: 4096 worlds x 16,384 entities (12 world bits + 14 local-slot bits)."""
    s = scenes.config5()
    assert s.n_worlds == 4096 and s.n == 4096 * 16_384
    o_world, o_keys, o_counts = oracle.frame(s)
    with GpuContext(s.n, max_views=1) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        assert np.array_equal(ctx.read_counts(), o_counts)
        keys = ctx.read_all_keys()
        assert np.array_equal(keys, o_keys)
        assert int(keys[-1] >> np.uint64(52)) == 4095                   # the world field reaches its last value
        assert np.array_equal(ctx.read_keys(0, 4095), o_keys[len(o_keys) - int(o_counts[-1]):])
        runs = ctx.read_runs()                                          # f2 with 12 world bits (classes wider than 32 bits)
        world = ctx.read_world(0, s.n)
    assert np.array_equal(oracle_lib.bits(world), oracle_lib.bits(o_world))
    cls = o_keys >> np.uint64(40 - 12)
    heads = np.flatnonzero(np.concatenate([[True], cls[1:] != cls[:-1]]))
    assert len(runs) == len(heads)
    assert np.array_equal(runs[:, 0], heads)
    assert np.array_equal((runs[:, 2] >> 5).astype(np.uint64), o_keys[heads] >> np.uint64(52))
    assert np.array_equal(runs[:, 3].astype(np.uint64), (o_keys[heads] >> np.uint64(28)) & np.uint64((1 << 19) - 1))


def test_config5_worlds(oracle):
    s = scenes.config5(n_worlds=24, entities_per_world=4096)
    assert check_scene(oracle, s) > 0


def test_all_visible_multi_tile_sort(oracle, small_flat_kernel):
    """Every entity inside every view: exercises compaction at full rate and a
    multi-tile sort with all eight digit passes live in the low bits."""
    s = scenes.config1(300_000)
    s.pos = (s.pos * np.float32(0.02)).astype(np.float32)
    s.pos[:, 2] -= np.float32(40.0)
    s.flags[:] = gpu.make_flags(np.ones(s.n, bool), np.full(s.n, 5, np.uint8))
    s.clip = np.repeat(s.clip, 3, axis=1)
    s.masks = np.array([1, 4, 1], np.uint8)
    n_keys = check_scene(oracle, s)
    assert n_keys > 2 * s.n


def test_edge_values_bit_exact(oracle, small_flat_kernel):
    """-0, denormals, Inf, NaN, negative and zero scale, un-normalised and zero quaternions."""
    s = scenes.config1(4096)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 1e-42, -1e-42, 3e38, -3e38, np.inf, -np.inf, np.nan,
                         0.70710678, -0.70710678, 0.5, -0.5, 2.0], np.float32)
    rng = np.random.default_rng(11)
    for arr in (s.pos, s.rot, s.scale):
        pick = rng.random(arr.shape) < 0.35
        arr[pick] = specials[rng.integers(0, len(specials), pick.sum())]
    # axis-aligned rotations with negative components and mirrored scales: the -0 cases
    s.rot[:64] = 0
    s.rot[:64, 1] = np.float32(-0.70710678)
    s.rot[:64, 3] = np.float32(0.70710678)
    s.scale[:64:2, 0] = np.float32(-1.0)
    check_scene(oracle, s)


def test_invisible_and_phase_masks(oracle, small_flat_kernel):
    s = scenes.config3(50_000)
    s.pos = (s.pos * np.float32(0.01)).astype(np.float32)
    s.pos[:, 2] -= np.float32(30.0)
    vis = (np.arange(s.n) % 3) != 0
    phases = (np.arange(s.n) % 16).astype(np.uint8)
    s.flags = gpu.make_flags(vis, phases)
    s.mesh[::7] = 2000        # beyond the bounds table: never drawn
    check_scene(oracle, s)


def test_no_views_transform_only(oracle):
    s = scenes.config1(5000)
    with GpuContext(s.n, max_views=4) as ctx:
        ctx.upload_entities(0, s.pos, s.rot, s.scale, None, s.mesh, s.shader, s.flags)
        ctx.frame(gpu.FRAME_TRANSFORM)
        world = ctx.read_world(0, s.n)
        assert ctx.read_total() == 0
    assert np.array_equal(oracle_lib.bits(world), oracle_lib.bits(oracle.transform_flat(s.pos, s.rot, s.scale)))


def test_defaults_when_components_absent(oracle):
    """has_position/has_rotation/has_scale == false -> 0 / identity / 1 (transform_system.cpp:24-49)."""
    n = 100
    with GpuContext(n) as ctx:
        ctx.upload_entities(0, count=n)
        ctx.frame(gpu.FRAME_TRANSFORM)
        world = ctx.read_world(0, n)
    assert np.array_equal(world, np.tile(np.eye(4, dtype=np.float32).reshape(-1), (n, 1)))


def test_basis_vectors(oracle):
    s = scenes.config1(10_000)
    with GpuContext(s.n) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame(gpu.FRAME_ALL | gpu.FRAME_BASIS)
        f, u, r = ctx.read_basis(0, s.n)
    _, of, ou, orr = oracle.transform_flat(s.pos, s.rot, s.scale, basis=True)
    for a, b in ((f, of), (u, ou), (r, orr)):
        assert np.array_equal(oracle_lib.bits(a), oracle_lib.bits(b))


def test_dirty_updates(oracle, small_flat_kernel):
    s = scenes.config1(20_000)
    with GpuContext(s.n) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        rng = np.random.default_rng(3)
        slots = rng.choice(s.n, 777, replace=False).astype(np.uint32)
        s.pos[slots] = (rng.random((777, 3), np.float32) * 20 - 10).astype(np.float32)
        s.rot[slots[:100]] = np.array([0, 0, 0, 1], np.float32)
        ctx.update_trs(slots, pos=s.pos[slots])
        ctx.update_trs(slots[:100], rot=s.rot[slots[:100]])
        lo, cnt = 5000, 3000
        s.scale[lo:lo + cnt] *= np.float32(1.5)
        ctx.update_trs_range(lo, cnt, scale=s.scale[lo:lo + cnt])
        fl_slots = np.arange(0, s.n, 5, dtype=np.uint32)
        s.flags[fl_slots] = gpu.make_flags(np.zeros(len(fl_slots), bool), np.full(len(fl_slots), 5, np.uint8))
        ctx.update_flags(fl_slots, s.flags[fl_slots])
        ctx.frame()
        world, counts, keys = ctx.read_world(0, s.n), ctx.read_counts(), ctx.read_all_keys()
    o_world, o_keys, o_counts = oracle.frame(s)
    assert np.array_equal(oracle_lib.bits(world), oracle_lib.bits(o_world))
    assert np.array_equal(counts, o_counts) and np.array_equal(keys, o_keys)


def test_visible_world_gather(oracle):
    s = scenes.config1()
    with GpuContext(s.n) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        keys = ctx.read_all_keys()
        vis = ctx.read_visible_world(0, len(keys))
        world = ctx.read_world(0, s.n)
    slots = (keys & np.uint64((1 << 26) - 1)).astype(np.int64)
    assert np.array_equal(oracle_lib.bits(vis), oracle_lib.bits(world[slots]))


def test_buffer_guards_are_armed(gpu_lib):
    """The canaries around the library's device buffers (ASTRE_GUARD=1, set in conftest.py) see a deliberate overrun,
    and no test so far has damaged one."""
    assert gpu_lib.astre_gpu_debug_guard_selftest() == 4
    with GpuContext(5000, max_views=2) as ctx:
        s = scenes.config1(5000)
        scenes.load_into(ctx, s)
        ctx.frame()
        ctx.read_all_keys()
        assert gpu_lib.astre_gpu_debug_guard_check() == 0


def test_capacity_overflow_is_reported():
    s = scenes.config1(50_000)
    s.pos = (s.pos * np.float32(0.01)).astype(np.float32)
    s.pos[:, 2] -= np.float32(30.0)
    with GpuContext(s.n, max_views=1, max_keys=1000) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        with pytest.raises(AstreError) as e:
            ctx.read_counts()
        assert e.value.code == 3


def test_bad_hierarchy_is_rejected():
    n = 16
    parent = np.full(n, gpu.NO_PARENT, np.uint32)
    parent[3], parent[4] = 4, 3      # cycle
    with GpuContext(n) as ctx:
        ctx.upload_entities(0, parent=parent, count=n)
        with pytest.raises(AstreError) as e:
            ctx.frame()
        assert e.value.code == 4


def test_repeated_frames_are_identical(oracle):
    s = scenes.config3(200_000)
    with GpuContext(s.n, max_views=5) as ctx:
        scenes.load_into(ctx, s)
        ref = None
        for _ in range(5):
            ctx.frame()
            k = ctx.read_all_keys().copy()
            if ref is None:
                ref = k
            assert np.array_equal(k, ref)
    assert np.array_equal(ref, oracle.frame(s)[1])


def test_draw_runs(oracle):
    """f2: the sorted list cut into (view, shader, mesh) runs, against numpy on the oracle's keys."""
    s = scenes.config3(300_000)
    with GpuContext(s.n, max_views=5) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        keys = ctx.read_all_keys()
        runs = ctx.read_runs()
    _, o_keys, _ = oracle.frame(s)
    assert np.array_equal(keys, o_keys)
    cls = (o_keys >> np.uint64(40)).astype(np.uint64)
    heads = np.flatnonzero(np.concatenate([[True], cls[1:] != cls[:-1]]))
    assert np.array_equal(runs[:, 0], heads)
    assert np.array_equal(runs[:, 1], np.diff(np.concatenate([heads, [len(o_keys)]])))
    assert np.array_equal(runs[:, 2], (cls[heads] >> np.uint64(19)).astype(np.uint32))
    assert np.array_equal(runs[:, 3], (cls[heads] & np.uint64((1 << 19) - 1)).astype(np.uint32))
    assert runs[:, 1].sum() == len(o_keys) and len(runs) <= 5 * 4 * 8


def test_indirect_draw_commands(oracle):
    """f2, on the device: one DrawElementsIndirectCommand per run (mesh index range, instances = keys of the run,
    base instance = the run's first list position), against numpy on the oracle's keys; the command buffer stays on
    the GPU (pointer returned), a mesh id beyond the table gives an empty command, and the next frame rebuilds it."""
    import torch
    from astre_b200 import multi
    for s in (scenes.config3(300_000), scenes.config5(n_worlds=300, entities_per_world=256)):
        info = np.array([[36, 0, 0], [6, 36, 24], [96, 42, 28], [192, 138, 61], [60, 330, 94], [240, 390, 106]], np.int64)  # meshes 6, 7: no entry
        with GpuContext(s.n, max_views=s.views_per_world) as ctx:
            scenes.load_into(ctx, s)
            ctx.set_mesh_draw_info(info)
            for frame in range(2):
                if frame == 1:
                    s.pos = (s.pos * np.float32(0.5)).astype(np.float32)
                    ctx.update_trs_range(0, s.n, s.pos, s.rot, s.scale)
                ctx.frame(gpu.FRAME_ALL | gpu.FRAME_GATHER)
                n, dev = ctx.build_indirect()
                cmds = ctx.read_indirect()
                runs = ctx.read_runs()
                _, o_keys, _ = oracle.frame(s)
                world_bits = int(np.ceil(np.log2(s.n_worlds))) if s.n_worlds > 1 else 0
                want = np_reference.indirect_commands(o_keys, info, world_bits)        # pinned by tests/test_oracle.py
                assert n == len(want) and dev and np.array_equal(cmds, want)
                assert (want[:, 0] == 0).any() and (want[:, 0] != 0).any()
                assert np.array_equal(runs[:, 0], want[:, 4]) and np.array_equal(runs[:, 1], want[:, 1])      # command i belongs to run i
                on_device = multi.device_view(dev, n * 5, "<i4", torch.device("cuda", 0)).cpu().numpy().reshape(-1, 5)
                assert np.array_equal(on_device, want.astype(np.int32))


def test_interpolate(oracle):
    """f1: interpolateFrame's per-proxy matrix (render.cpp:26-39).  mix / mix and the T*R*S chain are bit-exact; the
    only inexact step is glm::slerp's acos / sin, which the reference takes from the platform's libm: the kernel
    evaluates them in double and rounds once, so each is within 1 ULP of any libm's float result.  Asserted here:
    translation column and last row bit-exact, non-finite elements in the same places, and every element of the
    rotation block within 8 ULP OF ITS COLUMN'S SCALE of the oracle (glibc).  Measured on B200: 6.83 -- a 1-ULP
    difference in a sine moves a quaternion component by ~1 ULP, and toMat4's sums of products of components
    (2(xy + wz) ...) amplify that a few times.  (The transform path proper is 0 ULP; the north star's 4-ULP
    allowance concerns it.)"""
    a, b = scenes.config1(20_000), scenes.config1(20_000)
    rng = np.random.default_rng(9)
    b.pos = (a.pos + rng.normal(size=a.pos.shape).astype(np.float32)).astype(np.float32)
    b.scale = (a.scale * np.float32(1.25)).astype(np.float32)
    q = rng.normal(size=a.rot.shape).astype(np.float32)
    b.rot = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    b.rot[::3] = a.rot[::3]                 # equal rotations -> lerp branch when |q| == 1
    unit = np.abs(np.linalg.norm(a.rot.astype(np.float64), axis=1) - 1.0) < 1e-3      # the stream also has raw / zero quaternions
    worst = 0.0
    with GpuContext(a.n) as ctx:
        scenes.load_into(ctx, a)
        ctx.frame(gpu.FRAME_TRANSFORM)
        ctx.snapshot_trs()
        ctx.update_trs_range(0, a.n, b.pos, b.rot, b.scale)
        for alpha in (0.0, 0.3, 1.0, 2.5):
            ctx.interpolate(alpha)
            got = ctx.read_world(0, a.n)
            exp = oracle.interpolate(alpha, (a.pos, a.rot, a.scale), (b.pos, b.rot, b.scale))
            assert np.array_equal(np.isfinite(got), np.isfinite(exp)), alpha
            # translation column and the last row never touch the trig path: exact
            assert np.array_equal(oracle_lib.bits(got[:, 12:16]), oracle_lib.bits(exp[:, 12:16]))
            assert np.array_equal(oracle_lib.bits(got[:, 3::4]), oracle_lib.bits(exp[:, 3::4]))
            t = np.float32(min(max(alpha, 0.0), 1.0))
            scale = (a.scale * (np.float32(1) - t) + b.scale * t).astype(np.float64)          # column c is scaled by s_c
            for c in range(3):
                g, e = got[:, 4 * c:4 * c + 3].astype(np.float64), exp[:, 4 * c:4 * c + 3].astype(np.float64)
                ok = np.isfinite(e).all(axis=1) & unit
                ulp = np.abs(scale[ok, c]) * 2.0 ** -23
                err = np.abs(g[ok] - e[ok]).max(axis=1) / ulp
                worst = max(worst, float(err.max()))
        assert worst <= 8.0, f"rotation block: {worst:.2f} ULP of the column scale"
        ctx.frame(gpu.FRAME_TRANSFORM)       # the plain transform of the current TRS is unaffected
        assert np.array_equal(oracle_lib.bits(ctx.read_world(0, a.n)), oracle_lib.bits(oracle.transform_flat(b.pos, b.rot, b.scale)))
    print(f"f1 rotation block: max {worst:.3f} ULP of the column scale")


@pytest.mark.parametrize("graph", [False, True])
def test_stress_moving_camera_and_dirty_sets(oracle, graph, small_flat_kernel):
    """Many frames on one context: the camera turns, entities move, flags flip -- every frame equals the oracle
    (shakes out stale state between frames: look-back epochs, bucket counts, counters, key buffers).  With the frame
    replayed as a CUDA graph the view records of every frame are patched into the captured launch."""
    s = scenes.config3(120_000)
    s.pos = (s.pos * np.float32(0.1)).astype(np.float32)
    rng = np.random.default_rng(21)
    with GpuContext(s.n, max_views=5) as ctx:
        scenes.load_into(ctx, s)
        for it in range(24):
            clips = [scenes.camera_clip(yaw_deg=15.0 * it)] + scenes.cascade_clips()
            s.clip = np.stack(clips)[None].astype(np.float32)
            ctx.set_views(s.clip[0], s.masks)
            slots = rng.choice(s.n, 3000, replace=False).astype(np.uint32)
            s.pos[slots] += rng.normal(scale=5.0, size=(3000, 3)).astype(np.float32)
            ctx.update_trs(slots, pos=s.pos[slots])
            fl = rng.choice(s.n, 500, replace=False).astype(np.uint32)
            s.flags[fl] ^= 1
            ctx.update_flags(fl, s.flags[fl])
            ctx.frame(gpu.FRAME_ALL | (gpu.FRAME_GRAPH if graph else 0))
            keys, counts = ctx.read_all_keys(), ctx.read_counts()
            if it % 4 == 3 or it < 2:
                _, o_keys, o_counts = oracle.frame(s)
                assert np.array_equal(counts, o_counts), it
                assert np.array_equal(keys, o_keys), it
            else:
                assert np.all(keys[1:] > keys[:-1]) and counts.sum() == len(keys)


def _stacked_scene(n=100_000):
    """Entities stacked at a few positions: keys that agree on view, shader, mesh AND depth -- the bucket sort's
    worst case (one bucket of ~37k keys), runs of every length, plus a spread-out tail."""
    s = scenes.config1(n)
    rng = np.random.default_rng(17)
    spots = np.array([[0, 5, -20], [3, 5, -40], [-6, 4, -80], [10, 8, -200]], np.float32)
    group = np.zeros(n, np.int64)
    group[40_000:70_000] = 1
    group[70_000:70_040] = 2            # 40 keys per (shader, mesh) class
    group[70_040:] = 3
    s.pos = spots[group].copy()
    s.pos[90_000:] += rng.normal(scale=30.0, size=(n - 90_000, 3)).astype(np.float32)
    s.rot[:] = np.array([0, 0, 0, 1], np.float32)
    s.scale[:] = 1.0
    s.mesh[:40_000] = 0                 # one class: a single run of ~37k keys (every 16th is invisible)
    s.shader[:40_000] = 0
    s.flags = gpu.make_flags(np.arange(n) % 16 != 15, np.full(n, 5, np.uint8))
    return s


SORT_MODES = {"auto": {}, "force_lsd": {"ASTRE_SORT_FORCE_LSD": "1"}, "bins8": {"ASTRE_SORT_BIN_BITS": "8"},
              "bins12": {"ASTRE_SORT_BIN_BITS": "12"}, "bins22": {"ASTRE_SORT_BIN_BITS": "22"}}


@pytest.mark.parametrize("mode", list(SORT_MODES))
def test_sort_paths(oracle, mode, monkeypatch):
    """The key sort: bucket sort (any bucket-digit width sorts correctly) and the LSD fallback, which the device
    picks by itself when buckets are crowded.  Same list bit for bit on every path."""
    for k in ("ASTRE_SORT_FORCE_LSD", "ASTRE_SORT_BIN_BITS"):
        monkeypatch.delenv(k, raising=False)
    for k, v in SORT_MODES[mode].items():
        monkeypatch.setenv(k, v)
    for scene in (scenes.config3(300_000), scenes.config5(n_worlds=24, entities_per_world=4096), _stacked_scene()):
        o_world, o_keys, o_counts = oracle.frame(scene)
        with GpuContext(scene.n, max_views=scene.views_per_world) as ctx:
            scenes.load_into(ctx, scene)
            for _ in range(3):           # frame 2 and 3 size their buckets from the previous frame's key count
                ctx.frame()
                keys, path = ctx.read_all_keys(), ctx.last_sort_path()
                assert np.array_equal(keys, o_keys), (mode, scene.name, path)
                assert np.array_equal(ctx.read_counts(), o_counts)
                if mode == "force_lsd" or scene.name == "config1-flat-10k":     # the stacked scene keeps config1's name
                    assert path == 2, (mode, scene.name, path)                   # crowded buckets: the fallback, chosen on the device
                elif mode in ("auto", "bins22"):
                    assert path == 1, (mode, scene.name, path)


def _spread(values, mask):
    """pdep: deposits the low bits of `values` into the set bits of `mask` (LSB first)."""
    out = np.zeros(len(values), np.uint64)
    j = 0
    for b in range(64):
        if (mask >> b) & 1:
            out |= ((values >> np.uint64(j)) & np.uint64(1)) << np.uint64(b)
            j += 1
    return out


def test_sort_fuzz_arbitrary_key_sets():
    """The key sort on key sets no scene produces (astre_gpu_debug_sort_keys): every size around the tile / chunk /
    bucket boundaries, every bucket width, uniform / clustered / sequential / one-bucket distributions, fragmented
    live masks -- against numpy's sort.  Crowded sets must take the LSD fallback by themselves."""
    rng = np.random.default_rng(5)
    masks = [0x38_1C_07_FF_FC_FF_FF_FF, (1 << 46) - 1, 0xFFF0_0000_0FFF_F0FF, 0x0000_0003_FFFF_FFFF, 0x8000_0000_0000_0001 | (0xFFFFF << 20)]
    sizes = [0, 1, 2, 3, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 2047, 2048, 2049, 3071, 3072, 3073, 4095, 4096, 4097,
             5000, 8191, 8192, 8193, 50_000, 262_144, 1_000_003]
    with GpuContext(1024, max_views=1, max_keys=1_100_000) as ctx:
        for mask in masks:
            live = bin(mask).count("1")
            for n in sizes:
                if n > (1 << min(live, 40)) or (n > 8193 and mask not in masks[:2]):
                    continue
                for kind in ("uniform", "clustered", "sequential", "one_bucket"):
                    if kind == "uniform":
                        v = rng.choice(1 << min(live, 40), size=n, replace=False).astype(np.uint64) if n < 300_000 else \
                            np.unique(rng.integers(0, 1 << min(live, 40), size=2 * n, dtype=np.uint64))[:n]
                        if live > 40:
                            v = v << np.uint64(live - 40) | rng.integers(0, 1 << (live - 40), size=len(v), dtype=np.uint64)
                    elif kind == "clustered":                   # most keys share their top bits: big buckets
                        v = np.unique((rng.integers(0, 7, size=2 * n + 8, dtype=np.uint64) << np.uint64(live - 3)) |
                                      rng.integers(0, max(4 * n, 16), size=2 * n + 8, dtype=np.uint64))[:n]
                    elif kind == "sequential":
                        v = np.arange(n, dtype=np.uint64)[::-1].copy()
                    else:                                        # all keys differ only in their lowest bits
                        v = rng.permutation(n).astype(np.uint64)
                    if len(v) < n:
                        continue
                    keys = _spread(v, mask)
                    rng.shuffle(keys)
                    want = np.sort(keys)
                    for bits in ((2, 8, 13, 17, 22) if n <= 8193 else (8, 17)):
                        got, path = ctx.debug_sort_keys(keys, mask, bits)
                        assert np.array_equal(got, want), (hex(mask), n, kind, bits, path)
                        if n > 1:
                            assert path in (1, 2)
                        if kind == "one_bucket" and n > 2048 and bits <= live - 22:
                            assert path == 2, (n, bits)        # > 2048 keys in one bucket: the device picks the fallback
        # frames still work on the same context afterwards
        s = scenes.config1(1000)
        scenes.load_into(ctx, s)
        ctx.frame()
        assert np.all(np.diff(ctx.read_all_keys().astype(np.int64)) > 0)


def test_sort_bucket_count_follows_the_key_count(oracle):
    """One context, frames with very different key counts (the bucket count is a hint from the last frame)."""
    s = scenes.config1(200_000)
    with GpuContext(s.n, max_views=2) as ctx:
        scenes.load_into(ctx, s)
        dense = s.pos * np.float32(0.02)
        dense[:, 2] -= np.float32(40.0)
        for pos in (s.pos, dense.astype(np.float32), s.pos, dense.astype(np.float32)):
            s.pos = np.ascontiguousarray(pos, np.float32)
            ctx.update_trs_range(0, s.n, s.pos, s.rot, s.scale)
            _, o_keys, o_counts = oracle.frame(s)
            for _ in range(2):
                ctx.frame()
                assert np.array_equal(ctx.read_all_keys(), o_keys) and np.array_equal(ctx.read_counts(), o_counts)


def test_frame_graph_replays_equal_plain_frames(oracle):
    """ASTRE_FRAME_GRAPH: the frame's launches captured once and replayed as one CUDA graph.  Same results as plain
    launches while entities move (no recapture), views change, the key count changes the bucket plan and the entity
    count changes (recapture / exec update), for flat, hierarchical (cooperative launch) and batched-world frames."""
    G = gpu.FRAME_ALL | gpu.FRAME_GRAPH
    for scene in (scenes.config1(), scenes.config3(300_000), scenes.config2(150_000, levels=6), scenes.config5(n_worlds=16, entities_per_world=2048)):
        with GpuContext(scene.n, max_views=scene.views_per_world) as ctx:
            scenes.load_into(ctx, scene)
            o_world, o_keys, o_counts = oracle.frame(scene)
            for _ in range(4):
                ctx.frame(G)
                assert np.array_equal(ctx.read_all_keys(), o_keys) and np.array_equal(ctx.read_counts(), o_counts), scene.name
            assert np.array_equal(oracle_lib.bits(ctx.read_world(0, scene.n)), oracle_lib.bits(o_world))
            with pytest.raises(AstreError):
                ctx.last_cull_ms()                                      # no per-kernel events inside a graph
            # entities move: same graph, new data
            scene.pos = (scene.pos * np.float32(0.5)).astype(np.float32)
            ctx.update_trs_range(0, scene.n, scene.pos, scene.rot, scene.scale)
            o_world, o_keys, o_counts = oracle.frame(scene)
            for _ in range(2):
                ctx.frame(G)
                assert np.array_equal(ctx.read_all_keys(), o_keys) and np.array_equal(ctx.read_counts(), o_counts), scene.name
            # the camera turns: the view records are kernel parameters of the captured launches
            if scene.n_worlds == 1:
                scene.clip = scene.clip.copy()
                scene.clip[0, 0] = scenes.camera_clip(35.0)
                ctx.set_views(scene.clip[0], scene.masks)
                o_world, o_keys, o_counts = oracle.frame(scene)
                ctx.frame(G)
                assert np.array_equal(ctx.read_all_keys(), o_keys) and np.array_equal(ctx.read_counts(), o_counts), scene.name
            # plain launches and graph replays interleave
            ctx.frame()
            assert np.array_equal(ctx.read_all_keys(), o_keys)
            assert ctx.last_cull_ms() > 0
            ctx.frame(G)
            assert np.array_equal(ctx.read_all_keys(), o_keys)
            assert np.array_equal(oracle_lib.bits(ctx.read_world(0, scene.n)), oracle_lib.bits(o_world))


def test_count_exchange_single_rank(oracle):
    """multi.CountExchange without peers: the counts mirror (two device slots written by the frame's last kernel), the
    side-stream table and the device-side segment offsets of every frame, while frames keep being enqueued."""
    import torch
    from astre_b200 import multi
    for scene, rank_major in ((scenes.config3(200_000), False), (scenes.config5(n_worlds=32, entities_per_world=1024), True)):
        n_counts = scene.n_worlds * scene.views_per_world
        with GpuContext(scene.n, max_views=scene.views_per_world) as ctx:
            scenes.load_into(ctx, scene)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                ctx.set_stream(stream.cuda_stream)
                ex = multi.CountExchange(ctx, n_counts, rank_major=rank_major)
                for k in range(5):
                    if k == 3:                                   # the scene changes between frames
                        scene.pos = (scene.pos * np.float32(0.5)).astype(np.float32)
                        ctx.update_trs_range(0, scene.n, scene.pos, scene.rot, scene.scale)
                    ex.before_frame()
                    ctx.frame(gpu.FRAME_ALL | (gpu.FRAME_GRAPH if k % 2 else 0))
                    ex.after_frame()
                    if k in (2, 4):
                        table, off, tot = ex.last()
                        _, _, o_counts = oracle.frame(scene)
                        assert np.array_equal(table[0], o_counts)
                        h_off, h_tot = gpu.global_offsets(table, 0, rank_major)
                        assert np.array_equal(off, h_off) and np.array_equal(tot, h_tot)
                        assert np.array_equal(ctx.read_counts(), o_counts)
                ex.close()
            ctx.set_stream(None)


def test_count_boards_single_rank(oracle):
    """multi.PeerCountExchange without peers: the side-stream kernel pushes the mirrored counts into the rank's own board
    and turns them into table + offsets; more exchanges than board slots, plain and graph frames mixed."""
    import torch
    from astre_b200 import multi
    for scene, rank_major in ((scenes.config3(200_000), False), (scenes.config5(n_worlds=32, entities_per_world=1024), True),
                              (scenes.config5(n_worlds=600, entities_per_world=64), True)):      # 600 counts: more than one block pass
        n_counts = scene.n_worlds * scene.views_per_world
        with GpuContext(scene.n, max_views=scene.views_per_world) as ctx:
            scenes.load_into(ctx, scene)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                ctx.set_stream(stream.cuda_stream)
                ex = multi.PeerCountExchange(ctx, n_counts, rank_major=rank_major)
                for k in range(11):
                    if k == 6:                                   # the scene changes between frames
                        scene.pos = (scene.pos * np.float32(0.5)).astype(np.float32)
                        ctx.update_trs_range(0, scene.n, scene.pos, scene.rot, scene.scale)
                    ex.before_frame()
                    ctx.frame(gpu.FRAME_ALL | (gpu.FRAME_GRAPH if k % 3 else 0))
                    ex.after_frame()
                    if k in (2, 5, 10):
                        table, off, tot = ex.last()
                        _, _, o_counts = oracle.frame(scene)
                        assert np.array_equal(table[0], o_counts)
                        h_off, h_tot = gpu.global_offsets(table, 0, rank_major)
                        assert np.array_equal(off, h_off) and np.array_equal(tot, h_tot)
                        assert np.array_equal(ctx.read_counts(), o_counts)
                ex.close()
                assert ctx.count_board_status() == (11, 0)
            ctx.set_stream(None)


def test_count_boards_two_ranks_on_one_gpu(oracle):
    """Two contexts of one process act as two ranks (their boards are plain device pointers to each other): the whole
    protocol -- stores into the peer's board, tagged words, acknowledgements, slot reuse after 4 exchanges, one rank
    running ahead of the other -- with every checked frame's table and offsets compared against the host formula."""
    import torch
    from astre_b200 import multi
    halves = [scenes.config4(240_000, 0, 120_000), scenes.config4(240_000, 120_000, 240_000)]
    n_counts = halves[0].views_per_world
    want = [oracle.frame(h)[2] for h in halves]
    ctxs = [GpuContext(h.n, max_views=n_counts) for h in halves]
    streams = [torch.cuda.Stream() for _ in ctxs]
    try:
        boards = []
        for c, h, st in zip(ctxs, halves, streams):
            scenes.load_into(c, h)
            c.set_stream(st.cuda_stream)
            boards.append(c.count_board_create(n_counts)[1])
        ex = []
        for r, (c, st) in enumerate(zip(ctxs, streams)):
            with torch.cuda.stream(st):
                ex.append(multi.PeerCountExchange(c, n_counts, boards=boards, rank=r, world=2))

        def frames(r, k, graph):
            with torch.cuda.stream(streams[r]):
                for _ in range(k):
                    ex[r].before_frame()
                    ctxs[r].frame(gpu.FRAME_ALL | (gpu.FRAME_GRAPH if graph else 0))
                    ex[r].after_frame()

        def check(r):
            table, off, tot = ex[r].last()
            assert np.array_equal(table[0], want[0]) and np.array_equal(table[1], want[1]), (r, table)
            h_off, h_tot = gpu.global_offsets(table, r, False)
            assert np.array_equal(off, h_off) and np.array_equal(tot, h_tot)

        frames(0, 3, False); frames(1, 3, False)           # rank 0 enqueues three frames before rank 1 has any
        check(0); check(1)
        frames(1, 2, True); frames(0, 4, True); frames(1, 3, True); frames(0, 1, True)      # 8 each: every slot reused
        check(0); check(1)
        # one rank's scene changes: the other must see the new counts in the very next table
        halves[1].pos = (halves[1].pos * np.float32(0.25)).astype(np.float32)
        ctxs[1].update_trs_range(0, halves[1].n, halves[1].pos, halves[1].rot, halves[1].scale)
        want[1] = oracle.frame(halves[1])[2]
        frames(0, 1, True); frames(1, 1, True)
        check(0); check(1)
        frames(1, 2, False); frames(0, 2, True)
        check(0); check(1)
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                ex[r].finish()
        torch.cuda.synchronize()
        for c in ctxs:
            assert c.count_board_status() == (11, 0)
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                ex[r].close()
    finally:
        for c in ctxs:
            c.set_stream(None)
            c.close()


def test_count_boards_eight_ranks_on_one_gpu(oracle):
    """Eight contexts of one process as the eight ranks of a node (whole-world shards, 300 counts per rank: more than
    one pass of the exchange kernel's block): the host enqueues the ranks' frames in a different random order every
    round, one rank's worlds change half-way, every rank's last table / offsets / totals are checked."""
    import torch
    from astre_b200 import multi
    world, wpr = 8, 300
    full = scenes.config5(n_worlds=world * wpr, entities_per_world=64)
    shards = [full.shard(r, world) for r in range(world)]
    want = np.stack([oracle.frame(s)[2] for s in shards])
    ctxs = [GpuContext(s.n, max_views=1) for s in shards]
    streams = [torch.cuda.Stream() for _ in ctxs]
    rng = np.random.default_rng(11)
    try:
        boards = []
        for c, s, st in zip(ctxs, shards, streams):
            scenes.load_into(c, s)
            c.set_stream(st.cuda_stream)
            boards.append(c.count_board_create(wpr)[1])
        ex = []
        for r, (c, st) in enumerate(zip(ctxs, streams)):
            with torch.cuda.stream(st):
                ex.append(multi.PeerCountExchange(c, wpr, boards=boards, rank=r, world=world, rank_major=True))

        def check():
            for r in range(world):
                table, off, tot = ex[r].last()
                assert np.array_equal(table, want), (r, np.argwhere(table != want)[:4])
                h_off, h_tot = gpu.global_offsets(table, r, True)
                assert np.array_equal(off, h_off) and np.array_equal(tot, h_tot), r

        for k in range(10):
            if k == 5:
                s = shards[3]
                s.pos = (s.pos * np.float32(0.5)).astype(np.float32)
                ctxs[3].update_trs_range(0, s.n, s.pos, s.rot, s.scale)
                want[3] = oracle.frame(s)[2]
            for r in rng.permutation(world):
                with torch.cuda.stream(streams[r]):
                    ex[r].before_frame()
                    ctxs[r].frame(gpu.FRAME_ALL | (gpu.FRAME_GRAPH if k % 2 else 0))
                    ex[r].after_frame()
            if k in (4, 5, 9):
                check()
        torch.cuda.synchronize()
        for c in ctxs:
            assert c.count_board_status() == (10, 0)
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ex[r].close()
    finally:
        for c in ctxs:
            c.set_stream(None)
            c.close()


def test_gather_inside_the_frame(oracle):
    """ASTRE_FRAME_GATHER: the list's world matrices are produced by the frame itself; short lists (<= 4096 keys) arrive
    in host memory with no copy, longer ones need two copies and no launch, lists beyond the gather buffer fall back."""
    for scene, flags in ((scenes.config1(), gpu.FRAME_ALL | gpu.FRAME_GATHER),                     # 1137 keys: inline
                         (scenes.config1(), gpu.FRAME_ALL | gpu.FRAME_GATHER | gpu.FRAME_GRAPH),
                         (scenes.config3(400_000), gpu.FRAME_ALL | gpu.FRAME_GATHER | gpu.FRAME_GRAPH),  # ~17k keys: device gather
                         (scenes.config5(n_worlds=16, entities_per_world=2048), gpu.FRAME_ALL | gpu.FRAME_GATHER)):
        o_world, o_keys, _ = oracle.frame(scene)
        local_bits = 26 - max(int(np.ceil(np.log2(scene.n_worlds))), 0) if scene.n_worlds > 1 else 26
        wshift = 64 - (26 - local_bits)
        slots = (o_keys & np.uint64((1 << local_bits) - 1)).astype(np.int64)
        if scene.n_worlds > 1:
            slots += (o_keys >> np.uint64(wshift)).astype(np.int64) * scene.entities_per_world
        with GpuContext(scene.n, max_views=scene.views_per_world) as ctx:
            scenes.load_into(ctx, scene)
            k, w = np.zeros(len(o_keys) + 8, np.uint64), np.zeros((len(o_keys) + 8, 16), np.float32)
            for _ in range(3):
                ctx.frame(flags)
                n = ctx.read_draw_list(k, w)
                assert n == len(o_keys) and np.array_equal(k[:n], o_keys)
                assert np.array_equal(oracle_lib.bits(w[:n]), oracle_lib.bits(o_world[slots]))
            launches = ctx.launch_count
            ctx.frame(flags)
            ctx.read_draw_list(k, w)
            per_frame = ctx.launch_count - launches
            ctx.frame(flags & ~gpu.FRAME_GATHER)
            n = ctx.read_draw_list(k, w)                                # without the flag: gathered on read, same result
            assert n == len(o_keys) and np.array_equal(oracle_lib.bits(w[:n]), oracle_lib.bits(o_world[slots]))
            assert per_frame == 7                                       # reset, fused, scan, scatter, rank, fallback/counts, gather


def test_read_draw_list(oracle):
    s = scenes.config3(200_000)
    with GpuContext(s.n, max_views=5) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        keys = ctx.read_all_keys()
        vis = ctx.read_visible_world(0, len(keys))
        k2, w2 = np.zeros(len(keys) + 10, np.uint64), np.zeros((len(keys) + 10, 16), np.float32)
        n = ctx.read_draw_list(k2, w2)
        assert n == len(keys) and np.array_equal(k2[:n], keys) and np.array_equal(w2[:n].view(np.uint32), vis.view(np.uint32))
        k3, w3 = np.zeros(100, np.uint64), np.zeros((100, 16), np.float32)
        assert ctx.read_draw_list(k3, w3) == len(keys)                  # capped copy, full length reported
        assert np.array_equal(k3, keys[:100]) and np.array_equal(w3.view(np.uint32), vis[:100].view(np.uint32))
    _, o_keys, _ = oracle.frame(s)
    assert np.array_equal(keys, o_keys)


def test_deferred_sort_equals_one_call(oracle):
    """frame(CULL | SORT_LATER) + sort() reads exactly like frame(ALL); before sort() the keys are there, unsorted."""
    s = scenes.config3(300_000)
    _, o_keys, o_counts = oracle.frame(s)
    with GpuContext(s.n, max_views=s.views_per_world) as ctx:
        scenes.load_into(ctx, s)
        for _ in range(2):
            ctx.frame(gpu.FRAME_TRANSFORM | gpu.FRAME_CULL | gpu.FRAME_SORT_LATER)
            assert np.array_equal(ctx.read_counts(), o_counts)
            assert np.array_equal(np.sort(ctx.read_all_keys()), o_keys)
            with pytest.raises(AstreError):
                ctx.read_keys(0)                       # per-view lists need the sort
            ctx.sort()
            assert np.array_equal(ctx.read_all_keys(), o_keys)
            assert np.array_equal(np.concatenate([ctx.read_keys(v) for v in range(s.views_per_world)]), o_keys)
            with pytest.raises(AstreError):
                ctx.sort()                             # nothing pending any more
        ctx.frame()
        assert np.array_equal(ctx.read_all_keys(), o_keys)


def _random_scene(rng, i):
    """A small adversarial scene: random size, view count, cameras (perspective, orthographic, degenerate),
    masks, flags, optional parent forest, a sprinkle of special float values, random mesh bounds."""
    n = int(rng.choice([1, 2, 31, 33, 127, 129, 500, 1025, 3000]))
    s = scenes.config1(n)
    s.name = f"fuzz-{i}"
    s.pos = (rng.standard_normal((n, 3)) * rng.choice([0.1, 5.0, 80.0])).astype(np.float32)
    s.pos[:, 2] -= np.float32(rng.choice([0.0, 20.0, 200.0]))
    q = rng.standard_normal((n, 4)).astype(np.float32)
    if rng.random() < 0.7:
        q /= np.sqrt((q * q).sum(axis=1, keepdims=True, dtype=np.float32)) + np.float32(1e-20)
    s.rot = q.astype(np.float32)
    s.scale = np.exp(rng.standard_normal((n, 3))).astype(np.float32) * np.float32(rng.choice([1.0, 0.01, 50.0]))
    if rng.random() < 0.5:
        specials = np.array([0.0, -0.0, 1e-42, np.inf, -np.inf, np.nan, 3e38, -1.0, 0.5], np.float32)
        for arr in (s.pos, s.rot, s.scale):
            pick = rng.random(arr.shape) < 0.03
            arr[pick] = specials[rng.integers(0, len(specials), pick.sum())]
    n_meshes = int(rng.integers(1, 9))
    s.bounds = scenes.PREFAB_BOUNDS[:n_meshes].copy()
    s.bounds[:, 3:7] = np.abs(rng.standard_normal((n_meshes, 4))).astype(np.float32) * (rng.random((n_meshes, 4)) < 0.7)
    s.mesh = rng.integers(0, n_meshes + (1 if rng.random() < 0.3 else 0), n).astype(np.uint32)    # sometimes out of range: never drawn
    s.shader = rng.integers(0, int(rng.choice([1, 4, 256])), n).astype(np.uint32)
    s.flags = gpu.make_flags(rng.random(n) < 0.9, rng.integers(0, 16, n).astype(np.uint8))
    F = int(rng.integers(1, 9))
    clips = []
    for _ in range(F):
        kind = rng.random()
        eye = rng.standard_normal(3) * 10.0
        fwd = rng.standard_normal(3)
        fwd /= np.linalg.norm(fwd) + 1e-9
        if kind < 0.5:
            clips.append(gpu.camera_matrices(eye, fwd, (0.0, 1.0, 0.0), float(rng.uniform(20, 120)), float(rng.uniform(0.5, 2.5)),
                                             float(rng.choice([0.01, 0.1, 1.0])), float(rng.choice([50.0, 1000.0, 1e5])))[2])
        elif kind < 0.9:
            h = float(rng.choice([5.0, 100.0, 2000.0]))
            clips.append(gpu.ortho_view_matrix(eye, eye + fwd, (0.0, 1.0, 0.0), -h, h, -h, h, 0.1, float(rng.choice([100.0, 5000.0]))))
        else:
            clips.append((rng.standard_normal(16) * rng.choice([1.0, 0.0])).astype(np.float32))    # arbitrary / all-zero matrix
    s.clip = np.stack(clips)[None].astype(np.float32)
    s.masks = rng.integers(0, 16, F).astype(np.uint8)
    if n > 4 and rng.random() < 0.4:                                       # random forest, parents anywhere (not slot-ordered)
        depth_cap = int(rng.integers(1, 7))
        parent = np.full(n, 0xFFFFFFFF, np.uint32)
        depth = np.zeros(n, np.int64)
        order = rng.permutation(n)
        for k in range(1, n):
            if rng.random() < 0.8:
                p = order[rng.integers(0, k)]
                if depth[p] < depth_cap:
                    parent[order[k]] = p
                    depth[order[k]] = depth[p] + 1
        s.parent = parent
    return s


def test_fuzz_small_scenes(oracle, small_flat_kernel):
    """150 random small scenes (sizes around tile boundaries, 1..8 views of every kind including degenerate
    matrices, random masks / flags / bounds / parents, special values): world matrices, counts and sorted keys
    bit-exact against the oracle."""
    rng = np.random.default_rng(20261019)
    for i in range(150):
        s = _random_scene(rng, i)
        try:
            check_scene(oracle, s)
        except AssertionError as e:
            raise AssertionError(f"scene {i} (n={s.n}, views={s.views_per_world}, hier={s.parent is not None}): {e}") from None


def test_fuzz_batched_worlds(oracle, small_flat_kernel):
    """Random batches of independent worlds: world sizes on and off the 128-entity warp tile (cached and
    uncached view records), 1..3 views per world, random cameras per world."""
    rng = np.random.default_rng(424242)
    for i in range(40):
        n_worlds = int(rng.choice([2, 3, 5, 16, 64]))
        epw = int(rng.choice([4, 36, 128, 132, 256, 1000]))
        F = int(rng.integers(1, 4))
        n = n_worlds * epw
        s = scenes.config1(n)
        s.name = f"fuzz-worlds-{i}"
        s.pos = (rng.standard_normal((n, 3)) * rng.choice([5.0, 60.0])).astype(np.float32)
        s.scale = np.exp(rng.standard_normal((n, 3)) * 0.5).astype(np.float32)
        s.flags = gpu.make_flags(rng.random(n) < 0.9, rng.integers(0, 16, n).astype(np.uint8))
        clips = []
        for _ in range(n_worlds * F):
            eye = rng.standard_normal(3) * 20.0
            fwd = rng.standard_normal(3)
            fwd /= np.linalg.norm(fwd) + 1e-9
            if rng.random() < 0.6:
                clips.append(gpu.camera_matrices(eye, fwd, (0.0, 1.0, 0.0), float(rng.uniform(30, 110)), 1.7, 0.1, 1000.0)[2])
            else:
                h = float(rng.choice([20.0, 300.0]))
                clips.append(gpu.ortho_view_matrix(eye, eye + fwd, (0.0, 1.0, 0.0), -h, h, -h, h, 0.1, 500.0))
        s.clip = np.stack(clips).reshape(n_worlds, F, 16).astype(np.float32)
        s.masks = rng.integers(1, 16, F).astype(np.uint8)
        s.n_worlds, s.entities_per_world = n_worlds, epw
        try:
            check_scene(oracle, s)
        except AssertionError as e:
            raise AssertionError(f"scene {i} (worlds={n_worlds} x {epw}, views={F}): {e}") from None
