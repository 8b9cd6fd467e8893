"""Committed golden fixtures (tests/golden/make_fixtures.py).  CPU: the oracle still reproduces them.
GPU: the CUDA path reproduces them without the oracle being involved at test time."""
import json
import os

import numpy as np
import pytest

from astre_b200 import gpu, scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def unhex(lst, shape):
    return np.array([int(x, 16) for x in lst], np.uint32).view(np.float32).reshape(shape)


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def level0_inputs(d):
    n = len(d["ids"])
    return (unhex(d["pos"], (n, 3)), unhex(d["rot"], (n, 4)), unhex(d["scale"], (n, 3)),
            np.array(d["mesh"], np.uint32), np.array(d["shader"], np.uint32), np.array(d["flags"], np.uint8))


def test_level0_oracle_reproduces_fixture(oracle):
    d = load("level0_fixture.json")
    pos, rot, scl, mesh, shader, flags = level0_inputs(d)
    n = len(pos)
    assert n == 14 and d["camera"]["entity"] == 5          # game/resources/worlds/levels/level_0.json
    world, f, u, r = oracle.transform_flat(pos, rot, scl, basis=True)
    assert np.array_equal(world.view(np.uint32), unhex(d["world"], (n, 16)).view(np.uint32))
    for got, key in ((f, "forward"), (u, "up"), (r, "right")):
        assert np.array_equal(got.view(np.uint32), unhex(d[key], (n, 3)).view(np.uint32))
    c = d["camera"]
    view, proj = oracle.camera_view_proj(pos[c["index"]], f[c["index"]], u[c["index"]], c["fov"], c["aspect"], c["near"], c["far"])
    assert np.array_equal(view.view(np.uint32), unhex(d["view"], 16).view(np.uint32))
    assert np.array_equal(proj.view(np.uint32), unhex(d["proj"], 16).view(np.uint32))
    keys, counts = oracle.cull_keys(world, mesh, shader, flags, scenes.PREFAB_BOUNDS, oracle.mat4_mul(proj, view).reshape(1, 1, 16), [1])
    assert [f"{int(k):016x}" for k in keys] == d["camera_keys"] and counts[0] == d["camera_count"]


def test_config1_oracle_reproduces_fixture(oracle):
    d = load("config1_fixture.json")
    s = scenes.config1(d["n"])
    assert np.array_equal(s.pos.view(np.uint32), unhex(d["pos"], (d["n"], 3)).view(np.uint32))      # the generator is stable
    assert np.array_equal(s.rot.view(np.uint32), unhex(d["rot"], (d["n"], 4)).view(np.uint32))
    world, keys, counts = oracle.frame(s)
    assert np.array_equal(world.view(np.uint32), unhex(d["world"], (d["n"], 16)).view(np.uint32))
    assert [f"{int(k):016x}" for k in keys] == d["keys"] and counts.tolist() == d["counts"]


@pytest.mark.gpu
def test_level0_on_gpu_matches_fixture():
    d = load("level0_fixture.json")
    pos, rot, scl, mesh, shader, flags = level0_inputs(d)
    n, c = len(pos), d["camera"]
    with gpu.GpuContext(n, max_views=1) as ctx:
        ctx.upload_entities(0, pos, rot, scl, None, mesh, shader, flags, entity_id=np.array(d["ids"], np.uint64))
        ctx.set_mesh_bounds(scenes.PREFAB_BOUNDS)
        ctx.frame(gpu.FRAME_TRANSFORM | gpu.FRAME_BASIS)
        f, u, r = ctx.read_basis(0, n)
        _, _, clip = gpu.camera_matrices(pos[c["index"]], f[c["index"]], u[c["index"]], c["fov"], c["aspect"], c["near"], c["far"])
        ctx.set_views(clip.reshape(1, 16), [gpu.PHASE_OPAQUE])
        ctx.frame()
        world, keys = ctx.read_world(0, n), ctx.read_all_keys()
        assert ctx.slot_entity(4) == 5
    assert np.array_equal(world.view(np.uint32), unhex(d["world"], (n, 16)).view(np.uint32))
    for got, key in ((f, "forward"), (u, "up"), (r, "right")):
        assert np.array_equal(got.view(np.uint32), unhex(d[key], (n, 3)).view(np.uint32))
    assert [f"{int(k):016x}" for k in keys] == d["camera_keys"]


@pytest.mark.gpu
def test_config1_on_gpu_matches_fixture():
    d = load("config1_fixture.json")
    s = scenes.config1(d["n"])
    with gpu.GpuContext(s.n, max_views=1) as ctx:
        scenes.load_into(ctx, s)
        ctx.frame()
        world, keys, counts = ctx.read_world(0, s.n), ctx.read_all_keys(), ctx.read_counts()
    assert np.array_equal(world.view(np.uint32), unhex(d["world"], (s.n, 16)).view(np.uint32))
    assert [f"{int(k):016x}" for k in keys] == d["keys"] and counts.tolist() == d["counts"]


# ---- vectors computed by the reference's own sources (tests/golden/make_ref_fixtures.py, oracle/_ref) ---------

def unhexs(s, shape):
    """string of 8-hex-digit float32 bit patterns -> array"""
    return np.array([int(s[i:i + 8], 16) for i in range(0, len(s), 8)], np.uint32).view(np.float32).reshape(shape)


def _same(a, b):
    import oracle_lib
    return np.array_equal(oracle_lib.bits(a), oracle_lib.bits(b))


def test_ref_fixture_transform_oracle(oracle):
    t = load("ref_fixture.json")["transform"]
    n = t["n"]
    pos, rot, scl = unhexs(t["pos"], (n, 3)), unhexs(t["rot"], (n, 4)), unhexs(t["scale"], (n, 3))
    world, f, u, r = oracle.transform_flat(pos, rot, scl, basis=True)
    assert _same(world, unhexs(t["world"], (n, 16)))
    assert _same(f, unhexs(t["forward"], (n, 3))) and _same(u, unhexs(t["up"], (n, 3))) and _same(r, unhexs(t["right"], (n, 3)))


def test_ref_fixture_camera_and_lights_oracle_and_host_math(oracle, gpu_lib):
    """The oracle AND the product's host-side view builders (astre_camera_matrices / astre_light_space_matrix:
    plain C++, no device needed) reproduce the matrices the reference's CameraSystem / calculateLightSpaceMatrices
    computed."""
    d = load("ref_fixture.json")
    for c in d["camera"]:
        pos, fwd, up = unhexs(c["pos"], 3), unhexs(c["forward"], 3), unhexs(c["up"], 3)
        view, proj = unhexs(c["view"], 16), unhexs(c["proj"], 16)
        oview, oproj = oracle.camera_view_proj(pos, fwd, up, c["fov"], c["aspect"], c["near"], c["far"])
        assert _same(oview, view) and _same(oproj, proj)
        pview, pproj, pclip = gpu.camera_matrices(pos, fwd, up, c["fov"], c["aspect"], c["near"], c["far"])
        assert _same(pview, view) and _same(pproj, proj) and _same(pclip, oracle.mat4_mul(proj, view))
    for l in d["lights"]:
        pos, direction, want = unhexs(l["position"], 3), unhexs(l["direction"], 3), unhexs(l["light_space"], 16)
        assert _same(oracle.light_space(pos, direction, l["type"], l["outer_cutoff"]), want), l
        assert _same(gpu.light_space_matrix(pos, direction, l["type"], l["outer_cutoff"]), want), l


def test_ref_fixture_interpolate_oracle(oracle):
    t = load("ref_fixture.json")["interpolate"]
    n = t["n"]
    a = tuple(unhexs(t[k], (n, w)) for k, w in (("pos_a", 3), ("rot_a", 4), ("scale_a", 3)))
    b = tuple(unhexs(t[k], (n, w)) for k, w in (("pos_b", 3), ("rot_b", 4), ("scale_b", 3)))
    for fr in t["frames"]:
        assert _same(oracle.interpolate(fr["alpha"], a, b), unhexs(fr["model"], (n, 16))), fr["alpha"]


@pytest.mark.gpu
def test_ref_fixture_transform_on_gpu():
    """The CUDA path reproduces what the reference's TransformSystem::run computed -- no oracle involved."""
    t = load("ref_fixture.json")["transform"]
    n = t["n"]
    pos, rot, scl = unhexs(t["pos"], (n, 3)), unhexs(t["rot"], (n, 4)), unhexs(t["scale"], (n, 3))
    with gpu.GpuContext(n, max_views=1) as ctx:
        ctx.upload_entities(0, pos, rot, scl)
        ctx.set_entity_count(n)
        ctx.frame(gpu.FRAME_TRANSFORM | gpu.FRAME_BASIS)
        world = ctx.read_world(0, n)
        f, u, r = ctx.read_basis(0, n)
    assert _same(world, unhexs(t["world"], (n, 16)))
    assert _same(f, unhexs(t["forward"], (n, 3))) and _same(u, unhexs(t["up"], (n, 3))) and _same(r, unhexs(t["right"], (n, 3)))
