"""The product's host-side view construction (astre_b200/csrc/host_math.hpp, reached
through the C ABI) against the oracle's separately written GLM restatement."""
import numpy as np
import pytest

import oracle_lib
from astre_b200 import gpu, scenes


def test_camera_matrices_match_oracle(oracle):
    rng = np.random.default_rng(2)
    for _ in range(50):
        pos = rng.uniform(-50, 50, 3).astype(np.float32)
        fwd = rng.normal(size=3).astype(np.float32)
        fwd /= np.linalg.norm(fwd)
        up = np.array([0, 1, 0], np.float32)
        fov, aspect, zn, zf = float(rng.uniform(20, 120)), float(rng.uniform(0.5, 2.5)), 0.1, float(rng.uniform(50, 2000))
        view, proj, clip = gpu.camera_matrices(pos, fwd, up, fov, aspect, zn, zf)
        o_view, o_proj = oracle.camera_view_proj(pos, fwd, up, fov, aspect, zn, zf)
        assert np.array_equal(oracle_lib.bits(view), oracle_lib.bits(o_view))
        assert np.array_equal(oracle_lib.bits(proj), oracle_lib.bits(o_proj))
        assert np.array_equal(oracle_lib.bits(clip), oracle_lib.bits(oracle.mat4_mul(o_proj, o_view)))


def test_level0_camera(oracle):
    """The reference's shipped camera: game/resources/worlds/levels/level_0.json:120-141."""
    c = scenes.CAMERA
    _, _, clip = gpu.camera_matrices(c["pos"], c["forward"], c["up"], c["fov_deg"], c["aspect"], c["z_near"], c["z_far"])
    v, p = oracle.camera_view_proj(c["pos"], c["forward"], c["up"], c["fov_deg"], c["aspect"], c["z_near"], c["z_far"])
    assert np.array_equal(oracle_lib.bits(clip), oracle_lib.bits(oracle.mat4_mul(p, v)))


@pytest.mark.parametrize("light_type,cutoff", [(1, 0.0), (2, 0.0), (3, 0.85), (3, 1.5), (3, -0.2), (0, 0.0), (7, 0.0)])
def test_light_space_matches_oracle(oracle, light_type, cutoff):
    pos, direction = (3.0, 12.0, -4.0), (0.3, -1.0, 0.2)
    got = gpu.light_space_matrix(pos, direction, light_type, cutoff)
    exp = oracle.light_space(pos, direction, light_type, cutoff)
    if exp is None:
        assert got is None
    else:
        assert np.array_equal(oracle_lib.bits(got), oracle_lib.bits(exp))


def test_ortho_view_matches_oracle(oracle):
    eye, center, up = (10.0, 40.0, 5.0), (0.0, 0.0, -20.0), (0.0, 1.0, 0.0)
    got = gpu.ortho_view_matrix(eye, center, up, -30, 30, -20, 20, 0.1, 200.0)
    exp = oracle.mat4_mul(oracle.ortho(-30, 30, -20, 20, 0.1, 200.0), oracle.look_at(eye, center, up))
    assert np.array_equal(oracle_lib.bits(got), oracle_lib.bits(exp))


def test_oracle_backed_view_builder_matches_the_product(oracle):
    """bench.py's reference arm builds the synthetic views with the oracle so that its process never maps the product
    library; both builders must give the same clip matrices bit for bit."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from astre_b200 import scenes
    a = scenes.config3(16)
    a5 = scenes.config5(n_worlds=8, entities_per_world=16)
    scenes.set_matrix_provider(bench.OracleMatrices)
    try:
        b = scenes.config3(16)
        b5 = scenes.config5(n_worlds=8, entities_per_world=16)
    finally:
        scenes.set_matrix_provider(None)
    assert np.array_equal(a.clip.view(np.uint32), b.clip.view(np.uint32))
    assert np.array_equal(a5.clip.view(np.uint32), b5.clip.view(np.uint32))
