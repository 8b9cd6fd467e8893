"""The oracle against the hand-derived known-answer vectors and against the
independent numpy restatement.  CPU only."""
import json
import os

import numpy as np
import pytest

import np_reference as npr
import oracle_lib
from astre_b200 import gpu, scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat_hand_derived.json")) as f:
        return json.load(f)


def test_kat_trs(oracle, kat):
    for case in kat["trs"]:
        got = oracle.trs_matrix(case["pos"], case["rot"], case["scale"])
        assert np.array_equal(got, np.array(case["world"], np.float32)), case["name"]


def test_kat_basis(oracle, kat):
    for case in kat["basis"]:
        f, u, r = oracle.basis(case["rot"])
        assert np.array_equal(f, np.array(case["forward"], np.float32)), case["name"]
        assert np.array_equal(u, np.array(case["up"], np.float32)), case["name"]
        assert np.array_equal(r, np.array(case["right"], np.float32)), case["name"]


def test_kat_view_matrices(oracle, kat):
    for case in kat["look_at"]:
        assert np.array_equal(oracle.look_at(case["eye"], case["center"], case["up"]), np.array(case["view"], np.float32))
    for case in kat["ortho"]:
        assert np.array_equal(oracle.ortho(*case["args"]), np.array(case["proj"], np.float32))
    for case in kat["perspective"]:
        fov, aspect, zn, zf = case["args_deg"]
        m = oracle.perspective(float(np.float32(np.deg2rad(fov))), aspect, zn, zf)
        assert m[10] == case["m10"] and m[11] == case["m11"] and m[14] == case["m14"]
        assert abs(m[0] - 1.0) < 1e-6 and abs(m[5] - 1.0) < 1e-6
        assert np.count_nonzero(m) == 5
    for case in kat["mat4_mul"]:
        assert np.array_equal(oracle.mat4_mul(case["a"], case["b"]), np.array(case["out"], np.float32))


def test_kat_planes(oracle, kat):
    for case in kat["planes"]:
        v = oracle.make_views(np.array(case["clip"], np.float32).reshape(1, 1, 16), [1])[0]
        for i, name in enumerate(["left", "right", "bottom", "top", "near", "far"]):
            assert list(v.plane[i]) == case[name], name
        assert list(v.depth) == case["depth"] and v.depth_scale == case["depth_scale"]


def test_kat_cull_and_key(oracle, kat):
    for group in kat["cull"]:
        clip = np.array(group["clip"], np.float32).reshape(1, 1, 16)
        for case in group["cases"]:
            world = oracle.trs_matrix(case["pos"], [0, 0, 0, 1], [1, 1, 1]).reshape(1, 16)
            keys, counts = oracle.cull_keys(world, [0], [0], gpu.make_flags([1], [1]), [group["bounds"]], clip, [1])
            assert (counts[0] == 1) == case["visible"], (group["name"], case)
            if case["visible"]:
                assert (int(keys[0]) >> 26) & 16383 == case["depth14"], (group["name"], case)
    for case in kat["key"]:
        # place one entity so that it is visible, then check the packing of every field
        n = case["slot"] + 1
        clip = np.array(kat["planes"][0]["clip"], np.float32).reshape(1, 1, 16)
        clips = np.repeat(clip, case["view"] + 1, axis=1)
        pos = np.zeros((n, 3), np.float32)
        pos[:, 2] = -1.0 - 2.0 * (case["depth"] + 0.5) / 16384.0
        world = oracle.transform_flat(pos, np.tile(np.array([0, 0, 0, 1], np.float32), (n, 1)), np.ones((n, 3), np.float32))
        flags = gpu.make_flags(np.arange(n) == case["slot"], np.ones(n))
        bounds = np.zeros((case["mesh"] + 1, 7), np.float32)
        keys, _ = oracle.cull_keys(world, np.full(n, case["mesh"]), np.full(n, case["shader"]), flags, bounds, clips,
                                   [0] * case["view"] + [1])
        assert [hex(int(k)) for k in keys] == [case["key_hex"]]


def test_oracle_matches_numpy_flat(oracle):
    s = scenes.config1(20_000)
    assert np.array_equal(oracle_lib.bits(oracle.transform_flat(s.pos, s.rot, s.scale)),
                          oracle_lib.bits(npr.trs(s.pos, s.rot, s.scale)))


def test_oracle_matches_numpy_edge_values(oracle):
    s = scenes.config1(4096)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 1e-42, -1e-42, 3e38, -3e38, np.inf, -np.inf, np.nan,
                         0.70710678, -0.70710678, 0.5, -0.5, 2.0], np.float32)
    rng = np.random.default_rng(11)
    for arr in (s.pos, s.rot, s.scale):
        pick = rng.random(arr.shape) < 0.35
        arr[pick] = specials[rng.integers(0, len(specials), pick.sum())]
    assert np.array_equal(oracle_lib.bits(oracle.transform_flat(s.pos, s.rot, s.scale)),
                          oracle_lib.bits(npr.trs(s.pos, s.rot, s.scale)))


def test_oracle_matches_numpy_hierarchy(oracle):
    s = scenes.config2(n=5000)
    world, levels = oracle.transform_hierarchy(s.pos, s.rot, s.scale, s.parent)
    assert levels == 8
    assert np.array_equal(oracle_lib.bits(world), oracle_lib.bits(npr.hierarchy(s.pos, s.rot, s.scale, s.parent)))


@pytest.mark.parametrize("make", [lambda: scenes.config1(), lambda: scenes.config3(60_000),
                                  lambda: scenes.config5(n_worlds=6, entities_per_world=2048)])
def test_oracle_matches_numpy_cull_keys(oracle, make):
    s = make()
    world, keys, counts = oracle.frame(s)
    nk, nc = npr.cull_keys(world, s.mesh, s.shader, s.flags, s.bounds, s.clip, s.masks, s.n_worlds, s.entities_per_world)
    assert np.array_equal(counts, nc)
    assert np.array_equal(keys, nk)
    assert len(keys) > 0


def test_keys_unique_sorted_and_counts_consistent(oracle):
    s = scenes.config3(200_000)
    _, keys, counts = oracle.frame(s)
    assert np.all(keys[1:] > keys[:-1])
    views = (keys >> np.uint64(59)).astype(np.int64)
    assert np.array_equal(np.bincount(views, minlength=5), counts)


def test_sort_matches_numpy(oracle):
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 1000, 100_003):
        k = rng.integers(0, 2 ** 63, n, dtype=np.uint64) | (rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63))
        assert np.array_equal(oracle.sort_keys(k), np.sort(k))


def test_cull_keys_mt_equals_serial(oracle):
    """orc_frame uses the chunk-parallel variant for one world; same list as the (world,view)-parallel one."""
    import ctypes as C
    s = scenes.config3(150_000)
    world = oracle.transform_flat(s.pos, s.rot, s.scale)
    keys, counts = oracle.cull_keys(world, s.mesh, s.shader, s.flags, s.bounds, s.clip, s.masks, sort=False)
    brecs, nb = oracle.make_bounds(s.bounds)
    vrecs = oracle.make_views(s.clip, s.masks)
    k2 = np.empty(s.n * 5, np.uint64)
    c2 = np.zeros(5, np.uint32)
    p = oracle_lib._p
    total = oracle.lib.orc_cull_keys_mt(C.c_uint32(s.n), p(world, C.c_float), p(s.mesh, C.c_uint32), p(s.shader, C.c_uint32),
                                        p(s.flags, C.c_uint8), brecs, C.c_uint32(nb), C.c_uint32(5), vrecs,
                                        p(k2, C.c_uint64), C.c_uint64(len(k2)), p(c2, C.c_uint32))
    assert total == len(keys) and np.array_equal(c2, counts) and np.array_equal(k2[:total], keys)


def test_bad_hierarchy_rejected(oracle):
    parent = np.array([1, 0, 0xFFFFFFFF], np.uint32)
    with pytest.raises(ValueError):
        oracle.transform_hierarchy(np.zeros((3, 3)), np.zeros((3, 4)), np.ones((3, 3)), parent)
    with pytest.raises(ValueError):
        oracle.transform_hierarchy(np.zeros((3, 3)), np.zeros((3, 4)), np.ones((3, 3)), np.array([7, 0, 0], np.uint32))


def test_interpolate_endpoints(oracle):
    """render.cpp:26-39: with equal unit rotations slerp takes its lerp branch, so alpha 0 / 1
    reproduce the endpoint matrices exactly; alpha is clamped to [0,1] first (render.cpp:10)."""
    a, b = scenes.config1(2000), scenes.config1(2000)
    b.pos = b.pos + np.float32(1.0)
    b.scale = b.scale * np.float32(2.0)
    exact = np.array([[0, 0, 0, 1], [0, 1, 0, 0], [0.5, 0.5, 0.5, 0.5], [1, 0, 0, 0]], np.float32)
    a.rot = exact[np.arange(2000) % 4].copy()      # |q|^2 == 1 exactly -> dot == 1 -> lerp branch
    b.rot = a.rot.copy()
    unit = np.ones(2000, bool)
    A, B = (a.pos, a.rot, a.scale), (b.pos, b.rot, b.scale)
    w0, w1 = oracle.interpolate(0.0, A, B), oracle.interpolate(1.0, A, B)
    assert np.array_equal(oracle_lib.bits(w0[unit]), oracle_lib.bits(oracle.transform_flat(*A)[unit]))
    assert np.array_equal(oracle_lib.bits(w1[unit]), oracle_lib.bits(oracle.transform_flat(*B)[unit]))
    assert np.array_equal(oracle_lib.bits(oracle.interpolate(-3.0, A, B)), oracle_lib.bits(w0))
    assert np.array_equal(oracle_lib.bits(oracle.interpolate(7.5, A, B)), oracle_lib.bits(w1))
    half = oracle.interpolate(0.5, A, B)
    assert np.allclose(half[unit][:, 12:15], (a.pos[unit] + b.pos[unit]) / 2, rtol=1e-6, atol=1e-5)


def test_view_box_contains_the_view_volume(oracle):
    """Spec step 1 operand: the world-space box of a view.  ortho(-2,2,-1,1,1,3) with an identity view is the box
    x[-2,2] y[-1,1] z[-3,-1]; the box is widened by 1e-4 of its size + 1e-6 and rounded outwards."""
    clip = np.array([0.5, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, -2, 1], np.float32).reshape(1, 1, 16)
    v = oracle.make_views(clip, [1])[0]
    lo, hi = np.array(list(v.lo)[:3]), np.array(list(v.hi)[:3])
    assert np.all(lo <= [-2, -1, -3]) and np.all(lo >= np.array([-2, -1, -3]) - 1e-3)
    assert np.all(hi >= [2, 1, -1]) and np.all(hi <= np.array([2, 1, -1]) + 1e-3)
    nlo, nhi = npr.view_box(clip[0, 0])
    assert np.array_equal(nlo, lo.astype(np.float32)) and np.array_equal(nhi, hi.astype(np.float32))
    # a singular clip matrix disables the box test instead of culling everything
    v = oracle.make_views(np.zeros((1, 1, 16), np.float32), [1])[0]
    assert list(v.lo)[:3] == [-np.inf] * 3 and list(v.hi)[:3] == [np.inf] * 3
    # the level_0 camera: far corners reach +-1000*tan(30deg)*1.7 in x
    c = scenes.camera_clip()
    lo, hi = npr.view_box(c)
    assert 980 < hi[0] < 983 and -983 < lo[0] < -980 and -986 < lo[2] < -984 and 14.9 < hi[2] < 15.1


def test_trs_matrix_conventions_against_scipy(oracle):
    """Tolerance-level check of the conventions (not of the bits): quaternion order x,y,z,w, right-handed active
    rotation, column-major storage, M = T * R * S -- against scipy's independent Rotation class."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    for _ in range(200):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        p, s = rng.uniform(-50, 50, 3), rng.uniform(0.2, 4, 3)
        m = oracle.trs_matrix(p, q, s).reshape(4, 4).T               # column-major -> math layout [row, col]
        want = np.eye(4)
        want[:3, :3] = Rotation.from_quat(q).as_matrix() @ np.diag(s)
        want[:3, 3] = p
        assert np.allclose(m, want, rtol=0, atol=2e-5)
        f, u, r = oracle.basis(q.astype(np.float32))
        rot = Rotation.from_quat(q)
        assert np.allclose(f, rot.apply([0, 0, -1]), atol=1e-6) and np.allclose(u, rot.apply([0, 1, 0]), atol=1e-6)
        assert np.allclose(r, rot.apply([1, 0, 0]), atol=1e-6)


def test_box_pretest_only_removes_what_lies_outside_the_view_volume(oracle):
    """The cull spec's box pre-test against the plain six-plane test of SURVEY Appendix B:
    (a) box + planes keeps a SUBSET of planes-only; (b) every entity the box removes is outside the view volume --
    checked in float64 against the exact box of the volume's eight corners, not the widened float box the spec uses."""
    import ctypes as C
    s = scenes.config3(400_000)
    s.pos = (s.pos * np.float32(0.1)).astype(np.float32)            # +-200 around the camera: many borderline entities
    s.scale = (s.scale * np.float32(3.0)).astype(np.float32)         # big bounds: plane-test false positives beyond the corners
    world = oracle.transform_flat(s.pos, s.rot, s.scale)
    F = s.views_per_world
    keys_box, counts_box = oracle.cull_keys(world, s.mesh, s.shader, s.flags, s.bounds, s.clip, s.masks)
    # planes only: the same views with an unbounded box
    brecs, nb = oracle.make_bounds(s.bounds)
    vrecs = oracle.make_views(s.clip, s.masks)
    for v in range(F):
        for i in range(3):
            vrecs[v].lo[i], vrecs[v].hi[i] = -np.inf, np.inf
    cap = s.n * F
    keys = np.empty(cap, np.uint64)
    counts = np.zeros(F, np.uint32)
    p = oracle_lib._p
    total = oracle.lib.orc_cull_keys(C.c_uint32(s.n), p(world, C.c_float), p(s.mesh, C.c_uint32), p(s.shader, C.c_uint32),
                                     p(s.flags, C.c_uint8), brecs, C.c_uint32(nb), C.c_uint32(1), C.c_uint32(0), C.c_uint32(F), vrecs,
                                     p(keys, C.c_uint64), C.c_uint64(cap), p(counts, C.c_uint32))
    keys_planes = oracle.sort_keys(keys[:total])
    assert np.all(counts_box <= counts)
    removed = np.setdiff1d(keys_planes, keys_box, assume_unique=True)
    assert len(np.setdiff1d(keys_box, keys_planes, assume_unique=True)) == 0          # (a) subset
    assert len(removed) > 50, "the scene must contain entities only the box removes"

    # (b) float64 geometry: the exact corner box of each view volume, and each removed entity's bounds box
    w64 = world.astype(np.float64).reshape(-1, 4, 4)                                   # [n, column, row]
    bnd = s.bounds.astype(np.float64)
    for key in removed:
        v = int(key >> np.uint64(59)) & 31
        slot = int(key & np.uint64((1 << 26) - 1))
        clip = s.clip[0, v].astype(np.float64).reshape(4, 4).T                         # math layout [row, col]
        corners = np.array([[x, y, z, 1.0] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)])
        pts = (np.linalg.inv(clip) @ corners.T).T
        pts = pts[:, :3] / pts[:, 3:4]
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        m = w64[slot]
        b = bnd[s.mesh[slot]]
        c = m[0, :3] * b[0] + m[1, :3] * b[1] + m[2, :3] * b[2] + m[3, :3]
        e = np.abs(m[0, :3]) * b[3] + np.abs(m[1, :3]) * b[4] + np.abs(m[2, :3]) * b[5]
        r = b[6] * np.sqrt(max((m[j, :3] ** 2).sum() for j in range(3)))
        t = e + r
        assert np.any(c + t < lo) or np.any(c - t > hi), (hex(int(key)), c, t, lo, hi)


def test_draw_runs_and_indirect_commands_hand_derived():
    """f2 expectation (np_reference.draw_runs / indirect_commands, the checker of the GPU's astre_gpu_read_runs /
    astre_gpu_build_indirect) on a key list small enough to cut by hand.  Key = view<<59 | shader<<51 | mesh<<40 |
    depth<<26 | slot."""
    def key(view, shader, mesh, depth, slot):
        return (view << 59) | (shader << 51) | (mesh << 40) | (depth << 26) | slot
    keys = np.array([key(0, 0, 1, 5, 3), key(0, 0, 1, 9, 1), key(0, 0, 4, 2, 7), key(0, 2, 1, 0, 2),
                     key(1, 0, 1, 1, 3), key(1, 0, 1, 1, 9), key(1, 0, 1, 7, 0), key(1, 0, 9, 0, 4)], np.uint64)
    assert all(int(a) < int(b) for a, b in zip(keys[:-1], keys[1:]))     # a sorted list, as the frame produces it
    first, length, cls = npr.draw_runs(keys)
    assert first.tolist() == [0, 2, 3, 4, 7] and length.tolist() == [2, 1, 1, 3, 1]
    assert (cls & np.uint64(2047)).tolist() == [1, 4, 1, 1, 9] and (cls >> np.uint64(19)).tolist() == [0, 0, 0, 1, 1]
    info = [[36, 0, 0], [6, 36, 24], [12, 42, 28], [24, 54, 40], [60, 78, 52]]          # mesh ids 0..4; 9 has no entry
    assert npr.indirect_commands(keys, info).tolist() == [[6, 2, 36, 24, 0], [60, 1, 78, 52, 2], [6, 1, 36, 24, 3],
                                                          [6, 3, 36, 24, 4], [0, 1, 0, 0, 7]]
    assert npr.indirect_commands(keys[:0], info).shape == (0, 5)
    # batched worlds: the world index sits above the view field and widens the class
    wk = np.array([(0 << 62) | key(0, 0, 1, 0, 1) >> 2, (1 << 62) | key(0, 0, 1, 0, 1) >> 2], np.uint64)
    assert npr.draw_runs(wk, world_bits=2)[0].tolist() == [0, 1]
