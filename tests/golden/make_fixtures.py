#!/usr/bin/env python
"""Generates the committed golden fixtures (run in the build container, where /root/reference exists):

  level0_fixture.json   inputs taken from the reference's only shipped scene,
                        game/resources/worlds/levels/level_0.json (14 entities: position / rotation / scale with
                        TransformSystem's defaults for absent fields, mesh and shader names, visible flag, and
                        the camera entity's parameters); outputs = what the CPU oracle computes for them
                        (world matrices, basis vectors, camera view/proj, the camera's draw list).
  config1_fixture.json  the first 96 entities of synthetic configuration 1 with the oracle's outputs.

The reference ships no expected outputs for this path, so the outputs are the oracle's: these
files pin the oracle against regressions and give the GPU tests a vector that does not depend on the oracle
binary at test time.  float32 values are stored as hex bit patterns.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from astre_b200 import gpu, scenes  # noqa: E402

MESHES = ["cube_prefab", "quad_prefab", "cone_prefab", "cylinder_prefab", "icosphere1_prefab", "icosphere2_prefab",
          "icosphere3_prefab", "dome_prefab"]
SHADERS = ["deferred_shader", "basic_shader", "light_shader", "simple_NDC"]


def hexf(a):
    return [f"{int(x):08x}" for x in np.ascontiguousarray(a, np.float32).view(np.uint32).ravel()]


def level0():
    src = "/root/reference/game/resources/worlds/levels/level_0.json"
    data = json.load(open(src))
    ents = [e for c in data["chunks"] for e in c["entities"]]
    ents.sort(key=lambda e: e["id"])
    n = len(ents)
    pos, rot, scl = np.zeros((n, 3), np.float32), np.tile(np.array([0, 0, 0, 1], np.float32), (n, 1)), np.ones((n, 3), np.float32)
    mesh, shader, flags = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint8)
    camera = None
    for i, e in enumerate(ents):
        t = e.get("transform", {})
        if "position" in t: pos[i] = [float(t["position"].get(k, 0)) for k in "xyz"]
        if "rotation" in t: rot[i] = [float(t["rotation"].get(k, 0)) for k in "xyzw"]
        if "scale" in t: scl[i] = [float(t["scale"].get(k, 0)) for k in "xyz"]
        v = e.get("visual")
        if v and v.get("vertex_buffer_name") in MESHES and v.get("shader_name") in SHADERS:
            mesh[i], shader[i] = MESHES.index(v["vertex_buffer_name"]), SHADERS.index(v["shader_name"])
            flags[i] = gpu.make_flags([bool(v.get("visible", False))], [gpu.PHASE_OPAQUE | gpu.PHASE_SHADOW_CASTER])[0]
        if "camera" in e:
            c = e["camera"]
            camera = dict(entity=e["id"], index=i, fov=float(c["fov"]), near=float(c["near_plane"]), far=float(c["far_plane"]), aspect=float(c["aspect"]))
    o = oracle_lib.load()
    world, f, u, r = o.transform_flat(pos, rot, scl, basis=True)
    ci = camera["index"]
    view, proj = o.camera_view_proj(pos[ci], f[ci], u[ci], camera["fov"], camera["aspect"], camera["near"], camera["far"])
    clip = o.mat4_mul(proj, view).reshape(1, 1, 16)
    keys, counts = o.cull_keys(world, mesh, shader, flags, scenes.PREFAB_BOUNDS, clip, [gpu.PHASE_OPAQUE])
    return dict(source=src, ids=[e["id"] for e in ents], names=[e.get("name", "") for e in ents],
                pos=hexf(pos), rot=hexf(rot), scale=hexf(scl), mesh=mesh.tolist(), shader=shader.tolist(), flags=flags.tolist(),
                camera=camera, world=hexf(world), forward=hexf(f), up=hexf(u), right=hexf(r), view=hexf(view), proj=hexf(proj),
                camera_keys=[f"{int(k):016x}" for k in keys], camera_count=int(counts[0]))


def config1():
    s = scenes.config1(96)
    o = oracle_lib.load()
    world, keys, counts = o.frame(s)
    return dict(n=s.n, seed=1, world=hexf(world), keys=[f"{int(k):016x}" for k in keys], counts=counts.tolist(),
                pos=hexf(s.pos), rot=hexf(s.rot), scale=hexf(s.scale), mesh=s.mesh.tolist(), shader=s.shader.tolist(), flags=s.flags.tolist())


if __name__ == "__main__":
    json.dump(level0(), open(os.path.join(HERE, "level0_fixture.json"), "w"), indent=0)
    json.dump(config1(), open(os.path.join(HERE, "config1_fixture.json"), "w"), indent=0)
    print("written")
