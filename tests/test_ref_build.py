"""Pins oracle/astre_oracle.c to the reference's OWN sources.

oracle/_ref/libastre_ref.so is TransformSystem / CameraSystem / VisualSystem / LightSystem /
calculateLightSpaceMatrices / interpolateFrame / math::serialize compiled unmodified from /root/reference
(oracle/ref_recipe/Makefile).  Every test here runs the reference's code and the oracle's restatement on the same
inputs and demands identical bits (NaN payloads canonicalised).  Skipped only when the library neither exists
nor can be built (no reference tree); tests/test_golden.py then still checks the committed vectors it produced.
"""
import json
import os

import numpy as np
import pytest

import ref_lib
from oracle_lib import bits
from astre_b200 import scenes

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built and /root/reference absent")

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

SPECIALS = np.array([0.0, -0.0, 1.0, -1.0, 1e-42, -1e-42, 3e38, -3e38, np.inf, -np.inf, np.nan,
                     0.70710678, -0.70710678, 0.5, -0.5, 2.0], np.float32)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


def edge_scene(n=4096, seed=11):
    """config-1 entities with special values sprinkled in, plus the axis-aligned / mirrored -0 cases."""
    s = scenes.config1(n)
    rng = np.random.default_rng(seed)
    for arr in (s.pos, s.rot, s.scale):
        pick = rng.random(arr.shape) < 0.35
        arr[pick] = SPECIALS[rng.integers(0, len(SPECIALS), pick.sum())]
    s.rot[:64] = 0
    s.rot[:64, 1] = np.float32(-0.70710678)
    s.rot[:64, 3] = np.float32(0.70710678)
    s.scale[:64:2, 0] = np.float32(-1.0)
    return s


def ref_transforms(pos, rot, scale, has=None):
    ids = np.arange(1, len(pos) + 1, dtype=np.uint64)            # entity 0 is INVALID_ENTITY (ecs/entity.hpp:12)
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, pos, rot, scale, has)
        w.run_transform()
        return w.transforms(ids)


def check_flat(oracle, pos, rot, scale):
    m, f, u, r = ref_transforms(pos, rot, scale)
    om, of, ou, orr = oracle.transform_flat(pos, rot, scale, basis=True)
    assert same(m, om), "transform_matrix: oracle differs from the reference build"
    assert same(f, of) and same(u, ou) and same(r, orr), "forward/up/right: oracle differs from the reference build"


def test_config1_transform_and_basis(oracle):
    s = scenes.config1()
    check_flat(oracle, s.pos, s.rot, s.scale)


def test_other_streams(oracle):
    for s in (scenes.config3(20_000), scenes.config2(4096), scenes.config5(2, 4096)):
        check_flat(oracle, s.pos, s.rot, s.scale)


def test_edge_values(oracle):
    s = edge_scene()
    check_flat(oracle, s.pos, s.rot, s.scale)


def test_absent_components_take_the_reference_defaults(oracle):
    """transform_system.cpp:24-49: position 0, rotation (w=1), scale 1 when the sub-message is absent."""
    s = scenes.config1(512)
    has = (np.arange(512) % 8).astype(np.uint8)
    m, f, u, r = ref_transforms(s.pos, s.rot, s.scale, has)
    pos, rot, scale = s.pos.copy(), s.rot.copy(), s.scale.copy()
    pos[(has & ref_lib.HAS_POSITION) == 0] = 0
    rot[(has & ref_lib.HAS_ROTATION) == 0] = (0, 0, 0, 1)
    scale[(has & ref_lib.HAS_SCALE) == 0] = 1
    om, of, ou, orr = oracle.transform_flat(pos, rot, scale, basis=True)
    assert same(m, om) and same(f, of) and same(u, ou) and same(r, orr)


def test_level0_fixture_matches_reference_build(oracle):
    """The reference's one shipped scene (game/resources/worlds/levels/level_0.json), inputs from the fixture."""
    d = json.load(open(os.path.join(GOLDEN, "level0_fixture.json")))
    n = len(d["ids"])

    def unhex(lst, shape):
        return np.array([int(x, 16) for x in lst], np.uint32).view(np.float32).reshape(shape)

    pos, rot, scale = unhex(d["pos"], (n, 3)), unhex(d["rot"], (n, 4)), unhex(d["scale"], (n, 3))
    ids = np.array(d["ids"], np.uint64)
    c = d["camera"]
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, pos, rot, scale)
        w.add_camera(c["entity"], c["fov"], c["near"], c["far"], c["aspect"])
        w.run_transform()
        m, f, u, r = w.transforms(ids)
        view, proj, cam_pos = w.run_camera(c["entity"])
    assert same(m, unhex(d["world"], (n, 16)))
    assert same(f, unhex(d["forward"], (n, 3))) and same(u, unhex(d["up"], (n, 3))) and same(r, unhex(d["right"], (n, 3)))
    assert same(view, unhex(d["view"], 16)) and same(proj, unhex(d["proj"], 16))
    assert same(cam_pos, pos[c["index"]])


def test_camera_system(oracle):
    """camera_system.cpp:24-37 for many camera poses, including un-normalised rotations."""
    s = scenes.config1(300)
    ids = np.arange(1, s.n + 1, dtype=np.uint64)
    rng = np.random.default_rng(5)
    fov = rng.uniform(20, 120, s.n).astype(np.float32)
    aspect = rng.uniform(0.5, 2.4, s.n).astype(np.float32)
    near = rng.uniform(0.01, 1.0, s.n).astype(np.float32)
    far = rng.uniform(50, 5000, s.n).astype(np.float32)
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, s.pos, s.rot, s.scale)
        for i, e in enumerate(ids):
            w.add_camera(e, fov[i], near[i], far[i], aspect[i])
        w.run_transform()
        _, f, u, _ = w.transforms(ids)
        for i, e in enumerate(ids):
            view, proj, cam = w.run_camera(e)
            oview, oproj = oracle.camera_view_proj(s.pos[i], f[i], u[i], fov[i], aspect[i], near[i], far[i])
            assert same(view, oview) and same(proj, oproj) and same(cam, s.pos[i]), i
    # no active camera / an entity without a CameraComponent: identity matrices, origin (camera_system.cpp:15-19)
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids[:2], s.pos[:2], s.rot[:2], s.scale[:2])
        w.run_transform()
        view, proj, cam = w.run_camera(ids[0])
        assert same(view, np.eye(4, dtype=np.float32).ravel()) and same(proj, np.eye(4, dtype=np.float32).ravel()) and same(cam, np.zeros(3))


def test_light_space_matrices(oracle):
    """LightSystem::run + calculateLightSpaceMatrices (light_system.cpp:24-160): directional, spot over the whole
    cut-off range (both clamps), point, a light that casts no shadow and one of unknown type."""
    s = scenes.config1(64)
    n = 14                                                       # <= MAX_SHADOW_CASTERS (16)
    ids = np.arange(1, n + 1, dtype=np.uint64)
    types = [ref_lib.DIRECTIONAL, ref_lib.POINT] + [ref_lib.SPOT] * 9 + [ref_lib.DIRECTIONAL, 0, ref_lib.SPOT]
    outer = [0.0, 0.0, -2.0, -1.0, -0.9995, -0.5, 0.0, 0.3, 0.9, 0.9993, 1.5, 0.0, 0.0, 0.82]
    casts = [1] * 11 + [0, 1, 1]
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, s.pos[:n], s.rot[:n], s.scale[:n])
        for i, e in enumerate(ids):
            w.add_light(e, types[i], casts[i], outer_cutoff=outer[i], inner_cutoff=0.95)
        w.run_transform()
        _, fwd, _, _ = w.transforms(ids)
        w.run_light(0)
        mats = w.light_space_matrices(0)
        assert w.shadow_casters_count(0) == sum(casts) == len(mats)
        seen = set()
        for i, e in enumerate(ids):
            light = w.light(e)
            # light_system.cpp:44-49: the light sits 0.05 along its direction
            want_pos = (s.pos[i] + fwd[i] * np.float32(0.05)).astype(np.float32)
            assert same(light[0:3], want_pos) and same(light[4:7], fwd[i]) and light[7] == types[i]
            assert light[18] == casts[i]
            if not casts[i]:
                continue
            slot = int(light[19])
            assert slot not in seen
            seen.add(slot)
            want = oracle.light_space(light[0:3], light[4:7], types[i], outer[i])
            if want is None:                                     # unknown type: the matrix is left value-initialised
                assert types[i] == 0
                continue
            assert same(mats[slot], want), (i, types[i], outer[i])


def test_visual_system_proxy_rules():
    """visual_system.cpp:20-58: which entities get a proxy and what it carries."""
    s = scenes.config1(40)
    ids = np.arange(1, s.n + 1, dtype=np.uint64)
    has = np.full(s.n, 7, np.uint8)
    has[3], has[4], has[5] = 6, 5, 3                             # no position / no rotation / no scale
    with ref_lib.RefWorld() as w:
        w.register_vertex_buffer("cube_prefab", 11)
        w.register_vertex_buffer("cone_prefab", 12)
        w.register_shader("deferred_shader", 21)
        w.add_transforms(ids, s.pos, s.rot, s.scale, has)
        for i, e in enumerate(ids):
            if i % 10 == 9:
                continue                                         # no VisualComponent: no proxy
            vb = "no_such_mesh" if i % 10 == 7 else ("cone_prefab" if i % 2 else "cube_prefab")
            sh = "no_such_shader" if i % 10 == 8 else "deferred_shader"
            w.add_visual(e, i % 4 != 0, vb, sh, None if i % 3 == 0 else (0.1 * i, 0.2, 0.3, 1.0))
        w.run_transform()
        m, _, _, _ = w.transforms(ids)
        w.run_visual(0)
        want_ids = [int(e) for i, e in enumerate(ids) if i % 10 not in (7, 8, 9)]
        assert w.proxy_ids(0).tolist() == want_ids
        for i, e in enumerate(ids):
            p = w.proxy(e)
            if int(e) not in want_ids:
                assert p is None
                continue
            assert p["visible"] == (i % 4 != 0)                 # the authored flag, copied (no test of any kind)
            assert p["phases"] == 1 | 4                          # Opaque | ShadowCaster, always
            assert p["vertex_buffer"] == (12 if i % 2 else 11) and p["shader"] == 21 and not p["useTexture"]
            assert same(p["uModel"], m[i])                       # the matrix TransformSystem wrote
            color = (1, 0, 1, 1) if i % 3 == 0 else (np.float32(0.1 * i), 0.2, 0.3, 1.0)
            assert same(p["uColor"], np.array(color, np.float32))
            # raw proto reads without has_*: an absent field is zeros here, not TransformSystem's default
            assert same(p["position"], s.pos[i] if has[i] & 1 else np.zeros(3))
            assert same(p["rotation"], s.rot[i] if has[i] & 2 else np.zeros(4))
            assert same(p["scale"], s.scale[i] if has[i] & 4 else np.zeros(3))


@pytest.mark.parametrize("alpha", [-0.5, 0.0, 0.3, 0.999, 1.0, 1.7])
def test_interpolate_frame(oracle, alpha):
    """render::interpolateFrame (render.cpp:8-68): per-proxy mix / slerp / mix -> T*R*S, plus camera mixes."""
    a = scenes.config1(3000)
    b = scenes.config3(3000)
    b.pos = (b.pos * np.float32(0.1)).astype(np.float32)
    b.rot[:500] = a.rot[:500]                                    # identical ends: the lerp branch of slerp
    b.rot[500:600] = -a.rot[500:600]                             # antipodal ends
    b.rot[600:700] = (a.rot[600:700] + np.float32(1e-4)).astype(np.float32)
    ids = np.arange(1, a.n + 1, dtype=np.uint64)
    with ref_lib.RefWorld() as w:
        w.register_vertex_buffer("cube_prefab", 1)
        w.register_shader("deferred_shader", 1)
        w.add_transforms(ids, a.pos, a.rot, a.scale)
        w.add_camera(ids[0], 60.0, 0.1, 1000.0, 1.7)
        for e in ids:
            w.add_visual(e, True, "cube_prefab", "deferred_shader")
        w.run_transform()
        va, pa, ca = w.run_camera(ids[0], 0)
        w.run_visual(0)
        w.set_trs(ids, b.pos, b.rot, b.scale)
        w.run_transform()
        vb, pb, cb = w.run_camera(ids[0], 1)
        w.run_visual(1)
        w.interpolate(0, 1, alpha, 2)
        got = w.models(ids, 2)
        view, proj, cam = w.frame_camera(2)
    want = oracle.interpolate(alpha, (a.pos, a.rot, a.scale), (b.pos, b.rot, b.scale))
    assert same(got, want), "interpolated uModel: oracle differs from the reference build"
    t = np.float32(min(max(alpha, 0.0), 1.0))
    mix = lambda x, y: (x * (np.float32(1) - t) + y * t).astype(np.float32)      # glm::mix, one rounding per op
    assert same(view, mix(va, vb)) and same(proj, mix(pa, pb)) and same(cam, mix(ca, cb))


def test_product_tree_does_not_touch_the_reference_build():
    """oracle/_ref is a checker: nothing under astre_b200/ or include/ may name it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base in ("astre_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            for fn in files:
                if fn.endswith((".so", ".o", ".pyc", ".log")):
                    continue
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "libastre_ref" not in text and "oracle/_ref" not in text and "ref_lib" not in text, os.path.join(dirpath, fn)
