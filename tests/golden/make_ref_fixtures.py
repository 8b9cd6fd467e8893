#!/usr/bin/env python
"""Generates tests/golden/ref_fixture.json from the REFERENCE ITSELF (run in the build container):

oracle/_ref/libastre_ref.so is the reference's own TransformSystem / CameraSystem / LightSystem /
calculateLightSpaceMatrices / interpolateFrame sources compiled from /root/reference (oracle/ref_recipe/).
This script feeds it seeded inputs and stores what it computes, so that the pin travels with the repository:
tests/test_golden.py checks the oracle (CPU) and the CUDA path (GPU) against these vectors without needing the
reference tree or the _ref library at test time.  float32 arrays are stored as strings of hex bit patterns (8 digits per value).

  transform      384 entities: config-1 stream with special values (+-0, denormals, Inf, NaN, huge), the
                 axis-aligned / mirrored -0 cases, un-normalised and zero quaternions -> transform_matrix,
                 forward, up, right
  camera         8 camera entities -> view, proj
  lights         directional / point / spot (both clamps) -> light-space matrices
  interpolate    256 proxies at alpha in {0, 0.3, 1} (identical, antipodal and nearly-equal rotations included)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_lib  # noqa: E402
from astre_b200 import scenes  # noqa: E402
from test_ref_build import edge_scene  # noqa: E402


def hexf(a):
    """float32 array -> one string of 8 hex digits per value (bit patterns)."""
    return "".join(f"{int(x):08x}" for x in np.ascontiguousarray(a, np.float32).view(np.uint32).ravel())


def transform_block():
    s = edge_scene(384, seed=23)
    ids = np.arange(1, s.n + 1, dtype=np.uint64)
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, s.pos, s.rot, s.scale)
        w.run_transform()
        m, f, u, r = w.transforms(ids)
    return dict(n=s.n, pos=hexf(s.pos), rot=hexf(s.rot), scale=hexf(s.scale), world=hexf(m), forward=hexf(f), up=hexf(u), right=hexf(r))


def camera_block():
    s = scenes.config1(8)
    ids = np.arange(1, 9, dtype=np.uint64)
    params = [(60.0, 0.1, 1000.0, 1.7), (45.0, 0.5, 200.0, 1.0), (90.0, 0.01, 5000.0, 2.39), (30.0, 1.0, 50.0, 0.75),
              (120.0, 0.1, 100.0, 1.7777778), (75.0, 0.25, 800.0, 1.3333334), (20.0, 2.0, 4000.0, 1.6), (100.0, 0.05, 60.0, 1.25)]
    out = []
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, s.pos, s.rot, s.scale)
        for e, (fov, near, far, aspect) in zip(ids, params):
            w.add_camera(e, fov, near, far, aspect)
        w.run_transform()
        _, f, u, _ = w.transforms(ids)
        for i, e in enumerate(ids):
            view, proj, _ = w.run_camera(e)
            fov, near, far, aspect = params[i]
            out.append(dict(pos=hexf(s.pos[i]), forward=hexf(f[i]), up=hexf(u[i]), fov=fov, near=near, far=far, aspect=aspect,
                            view=hexf(view), proj=hexf(proj)))
    return out


def light_block():
    s = scenes.config1(16)
    types = [1, 2] + [3] * 8
    outer = [0.0, 0.0, -2.0, -0.9995, -0.5, 0.0, 0.82, 0.9993, 1.5, 0.3]
    ids = np.arange(1, len(types) + 1, dtype=np.uint64)
    out = []
    with ref_lib.RefWorld() as w:
        w.add_transforms(ids, s.pos[:len(ids)], s.rot[:len(ids)], s.scale[:len(ids)])
        for i, e in enumerate(ids):
            w.add_light(e, types[i], True, outer_cutoff=outer[i])
        w.run_transform()
        w.run_light(0)
        mats = w.light_space_matrices(0)
        for i, e in enumerate(ids):
            light = w.light(e)
            out.append(dict(type=types[i], outer_cutoff=outer[i], position=hexf(light[0:3]), direction=hexf(light[4:7]),
                            light_space=hexf(mats[int(light[19])])))
    return out


def interpolate_block():
    n = 256
    a, b = scenes.config1(n), scenes.config3(n)
    b.pos = (b.pos * np.float32(0.1)).astype(np.float32)
    b.rot[:32] = a.rot[:32]
    b.rot[32:64] = -a.rot[32:64]
    b.rot[64:96] = (a.rot[64:96] + np.float32(1e-4)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint64)
    out = dict(n=n, pos_a=hexf(a.pos), rot_a=hexf(a.rot), scale_a=hexf(a.scale), pos_b=hexf(b.pos), rot_b=hexf(b.rot), scale_b=hexf(b.scale), frames=[])
    with ref_lib.RefWorld() as w:
        w.register_vertex_buffer("cube_prefab", 1)
        w.register_shader("deferred_shader", 1)
        w.add_transforms(ids, a.pos, a.rot, a.scale)
        for e in ids:
            w.add_visual(e, True, "cube_prefab", "deferred_shader")
        w.run_transform()
        w.run_visual(0)
        w.set_trs(ids, b.pos, b.rot, b.scale)
        w.run_transform()
        w.run_visual(1)
        for alpha in (0.0, 0.3, 1.0):
            w.interpolate(0, 1, alpha, 2)
            out["frames"].append(dict(alpha=alpha, model=hexf(w.models(ids, 2))))
    return out


if __name__ == "__main__":
    sha = open(os.path.join(ROOT, "oracle", "_ref", "SOURCES.sha256")).read().split("\n")
    doc = dict(generator="tests/golden/make_ref_fixtures.py", library="oracle/_ref/libastre_ref.so (reference sources compiled by oracle/ref_recipe/Makefile)",
               reference_sources_sha256=[l for l in sha if l], transform=transform_block(), camera=camera_block(), lights=light_block(),
               interpolate=interpolate_block())
    json.dump(doc, open(os.path.join(HERE, "ref_fixture.json"), "w"), indent=1)
    print("written")
