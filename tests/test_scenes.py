import numpy as np

from astre_b200 import gpu, scenes


def test_scenes_are_deterministic_and_shaped():
    a, b = scenes.config1(), scenes.config1()
    for name in ("pos", "rot", "scale", "mesh", "shader", "flags"):
        assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True)
    assert a.n == 10_000 and a.clip.shape == (1, 1, 16) and a.pos.dtype == np.float32
    assert np.all(np.abs(a.pos) <= 200) and np.all(a.scale >= 0.25) and np.all(a.scale < 4)
    assert np.all(a.rot[1023] == 0) and not np.isclose(np.linalg.norm(a.rot[63]), 1.0, atol=1e-3)
    assert np.all((a.flags[15::16] & 1) == 0) and np.all((a.flags[0::16] & 1) == 1)
    assert a.mesh.max() == 7 and a.shader.max() == 3


def test_ranges_reproduce_the_whole_stream():
    whole = scenes.config4(50_000)
    part = scenes.config4(50_000, 12_000, 30_000)
    assert np.array_equal(whole.pos[12_000:30_000], part.pos) and np.array_equal(whole.rot[12_000:30_000], part.rot)
    sh = [whole.shard(r, 4) for r in range(4)]
    assert sum(s.n for s in sh) == whole.n
    assert np.array_equal(np.concatenate([s.pos for s in sh]), whole.pos)


def test_config2_levels():
    s = scenes.config2(n=2550)
    roots = s.parent == gpu.NO_PARENT
    assert roots.sum() == 10
    child = np.nonzero(~roots)[0]
    assert np.all(s.parent[child] < child)      # parents precede children: level-sorted slots


def test_config5_world_shards():
    s = scenes.config5(n_worlds=8, entities_per_world=256)
    assert s.n == 8 * 256 and s.clip.shape == (8, 1, 16)
    half = s.shard(1, 2)
    assert half.n_worlds == 4 and np.array_equal(half.pos, s.pos[4 * 256:])
    assert np.array_equal(half.clip, s.clip[4:])


def test_c_generator_matches_the_numpy_definition(monkeypatch):
    """tools/scenegen (harness speed-up) and the numpy definition produce the same bits."""
    if not scenes._scenegen():
        import pytest
        pytest.skip("tools/_build/libscenegen.so not built")
    fast = [scenes.entity_block(7, 1000, 9000, 200.0, (0.25, 4.0)), scenes.entity_block(2, 0, 3000, 10.0, (0.8, 1.25), 5, 3)]
    monkeypatch.setattr(scenes, "_SCENEGEN", False)
    slow = [scenes.entity_block(7, 1000, 9000, 200.0, (0.25, 4.0)), scenes.entity_block(2, 0, 3000, 10.0, (0.8, 1.25), 5, 3)]
    for f, s in zip(fast, slow):
        for a, b in zip(f, s):
            assert a.dtype == b.dtype and a.shape == b.shape
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
