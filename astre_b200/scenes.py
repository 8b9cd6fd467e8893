"""Deterministic synthetic scenes for the five BASELINE.json configurations.

Every value comes from an integer hash (splitmix64 finaliser) turned into a
24-bit-mantissa float, so the oracle and the GPU path consume identical float
bits on any machine.  The reference ships one 14-entity level
(game/resources/worlds/levels/level_0.json); its camera (fov 60, aspect 1.7,
near 0.1, far 1000 at (0,5,15)) is the camera used here.

Mesh ids 0..7 stand for the reference's prefab meshes (cube, quad, cone,
cylinder, icosphere1-3, dome -- engine/modules/Render/src/vertex.cpp); their
local bounds are +-0.5 boxes and unit spheres.
"""
import ctypes as _C
import os as _os
from dataclasses import dataclass, field

import numpy as np

from . import gpu as _gpu     # constants and (lazily) the product's host-side view builders; importing it loads nothing

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def _hash(seed, index, channel):
    """uint64 hash of (seed, entity index, channel); index is a uint64 array."""
    with np.errstate(over="ignore"):
        base = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        x = base + index * np.uint64(16) + np.uint64(channel)
        return _mix(_mix(x) + np.uint64(0x9E3779B97F4A7C15))


def _unit(seed, index, channel):
    """float32 uniform in [0,1) with a 24-bit mantissa (exact)."""
    return ((_hash(seed, index, channel) >> np.uint64(40)).astype(np.float32)) * np.float32(2.0 ** -24)


def _uniform(seed, index, channel, lo, hi):
    u = _unit(seed, index, channel)
    return np.float32(lo) + np.float32(hi - lo) * u


def _randint(seed, index, channel, n):
    return ((_hash(seed, index, channel) >> np.uint64(33)) % np.uint64(n)).astype(np.uint32)


PREFAB_BOUNDS = np.array([
    # center xyz, extent xyz, radius
    [0, 0, 0, 0.5, 0.5, 0.5, 0.0],   # cube
    [0, 0, 0, 0.5, 0.5, 0.0, 0.0],   # quad
    [0, 0, 0, 0.5, 0.5, 0.5, 0.0],   # cone
    [0, 0, 0, 0.5, 0.5, 0.5, 0.0],   # cylinder
    [0, 0, 0, 0.0, 0.0, 0.0, 1.0],   # icosphere1
    [0, 0, 0, 0.0, 0.0, 0.0, 1.0],   # icosphere2
    [0, 0, 0, 0.0, 0.0, 0.0, 1.0],   # icosphere3
    [0, 0, 0, 0.0, 0.0, 0.0, 1.0],   # dome
], dtype=np.float32)


@dataclass
class Scene:
    name: str
    pos: np.ndarray
    rot: np.ndarray            # x,y,z,w
    scale: np.ndarray
    mesh: np.ndarray
    shader: np.ndarray
    flags: np.ndarray
    parent: np.ndarray = None  # uint32 or None (flat)
    bounds: np.ndarray = field(default_factory=lambda: PREFAB_BOUNDS.copy())
    clip: np.ndarray = None    # [n_worlds, views_per_world, 16]
    masks: np.ndarray = None   # [views_per_world]
    n_worlds: int = 1
    entities_per_world: int = 0
    global_index: np.ndarray = None   # subtree shards: global entity index of every local slot

    @property
    def n(self):
        return len(self.pos)

    @property
    def views_per_world(self):
        return len(self.masks)

    def shard(self, rank, world_size):
        """Contiguous entity-range shard (flat scenes), whole-world shard (batched worlds) or
        root-subtree shard (hierarchies; `.global_index` maps local slots back)."""
        if self.parent is not None:
            from . import multi
            mine, local_parent = multi.shard_subtrees(self.parent, world_size)[rank]
            sh = Scene(f"{self.name}[{rank}/{world_size}]", self.pos[mine], self.rot[mine], self.scale[mine],
                       self.mesh[mine], self.shader[mine], self.flags[mine], local_parent, self.bounds, self.clip,
                       self.masks, 1, self.entities_per_world)
            sh.global_index = mine
            return sh
        if self.n_worlds > 1:
            assert self.n_worlds % world_size == 0
            wl = self.n_worlds // world_size
            lo, hi = rank * wl * self.entities_per_world, (rank + 1) * wl * self.entities_per_world
            clip = self.clip[rank * wl:(rank + 1) * wl]
            n_worlds = wl
        else:
            per = -(-self.n // world_size)
            per = (per + 3) & ~3
            lo, hi = min(rank * per, self.n), min((rank + 1) * per, self.n)
            clip, n_worlds = self.clip, 1
        return Scene(f"{self.name}[{rank}/{world_size}]", self.pos[lo:hi], self.rot[lo:hi], self.scale[lo:hi],
                     self.mesh[lo:hi], self.shader[lo:hi], self.flags[lo:hi], None, self.bounds, clip, self.masks,
                     n_worlds, self.entities_per_world)


_SCENEGEN = None


def _scenegen():
    """tools/_build/libscenegen.so (tools/scenegen/scenegen.c): the same generator in C with OpenMP -- 64M entities
    in about a second instead of a minute.  Harness only; absent or ASTRE_SCENEGEN=numpy -> the numpy definition below."""
    global _SCENEGEN
    if _SCENEGEN is None:
        _SCENEGEN = False
        so = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tools", "_build", "libscenegen.so")
        if _os.environ.get("ASTRE_SCENEGEN", "") != "numpy" and _os.path.exists(so):
            lib = _C.CDLL(so)
            lib.scenegen_block.restype = None
            lib.scenegen_block.argtypes = [_C.c_uint64, _C.c_uint64, _C.c_uint64, _C.c_float, _C.c_float, _C.c_float, _C.c_float,
                                           _C.c_uint32, _C.c_uint32, _C.c_uint8] + [_C.c_void_p] * 6
            _SCENEGEN = lib
    return _SCENEGEN


def entity_block(seed, lo, hi, pos_range, scale_range, n_mesh=8, n_shader=4):
    """TRS + visual attributes of entity indices [lo, hi) of the stream `seed`."""
    lib = _scenegen()
    if lib:
        n = int(hi) - int(lo)
        pos, rot, scale = np.empty((n, 3), np.float32), np.empty((n, 4), np.float32), np.empty((n, 3), np.float32)
        mesh, shader, flags = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.uint8)
        lib.scenegen_block(int(seed), int(lo), int(hi), np.float32(-pos_range), np.float32(pos_range - -pos_range),
                           np.float32(scale_range[0]), np.float32(scale_range[1] - scale_range[0]), n_mesh, n_shader,
                           _gpu.PHASE_OPAQUE | _gpu.PHASE_SHADOW_CASTER,
                           *[a.ctypes.data for a in (pos, rot, scale, mesh, shader, flags)])
        return pos, rot, scale, mesh, shader, flags
    idx = np.arange(lo, hi, dtype=np.uint64)
    pos = np.stack([_uniform(seed, idx, c, -pos_range, pos_range) for c in range(3)], axis=1)
    q = np.stack([_uniform(seed, idx, 3 + c, -1.0, 1.0) for c in range(4)], axis=1)
    n2 = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + (q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    with np.errstate(invalid="ignore", divide="ignore"):
        qn = q / np.sqrt(n2)[:, None]
    qn = np.where(np.isfinite(qn), qn, np.float32(0)).astype(np.float32)
    rot = qn
    raw = (idx % np.uint64(64)) == np.uint64(63)         # every 64th left un-normalised
    rot[raw] = q[raw]
    rot[(idx % np.uint64(1024)) == np.uint64(1023)] = 0  # every 1024th all-zero
    scale = np.stack([_uniform(seed, idx, 7 + c, scale_range[0], scale_range[1]) for c in range(3)], axis=1)
    mesh = _randint(seed, idx, 10, n_mesh)
    shader = _randint(seed, idx, 11, n_shader)
    visible = (idx % np.uint64(16)) != np.uint64(15)     # visible flag clear on every 16th
    phases = np.full(len(idx), _gpu.PHASE_OPAQUE | _gpu.PHASE_SHADOW_CASTER, np.uint8)   # visual_system.cpp:55
    flags = _gpu.make_flags(visible, phases)
    return (np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(rot, np.float32),
            np.ascontiguousarray(scale, np.float32), mesh, shader, flags)


CAMERA = dict(pos=(0.0, 5.0, 15.0), forward=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0),
              fov_deg=60.0, aspect=1.7, z_near=0.1, z_far=1000.0)   # level_0.json:120-141


class _ProductMatrices:
    """View builders of the product library (host-only C++, astre_camera_matrices / astre_ortho_view_matrix)."""

    @staticmethod
    def camera_clip(pos, fwd, up, fov_deg, aspect, z_near, z_far):
        return _gpu.camera_matrices(pos, fwd, up, fov_deg, aspect, z_near, z_far)[2]

    @staticmethod
    def ortho_view(eye, center, up, l, r, b, t, n, f):
        return _gpu.ortho_view_matrix(eye, center, up, l, r, b, t, n, f)


_MATRICES = _ProductMatrices


def set_matrix_provider(provider):
    """Who builds the clip matrices of the synthetic views.  Default: the product's host math.  bench.py's
    reference arm installs an oracle-backed provider so that its process never maps the product library
    (both produce identical bits: tests/test_host_math.py)."""
    global _MATRICES
    _MATRICES = provider or _ProductMatrices


def camera_clip(yaw_deg=0.0):
    """Clip matrix of the level_0 camera, optionally yawed about +Y."""
    a = np.float32(np.deg2rad(np.float32(yaw_deg)))
    fwd = (float(-np.sin(a)), 0.0, float(-np.cos(a)))
    return _MATRICES.camera_clip(CAMERA["pos"], fwd, CAMERA["up"], CAMERA["fov_deg"], CAMERA["aspect"],
                                 CAMERA["z_near"], CAMERA["z_far"])


def cascade_clips(splits=(10.0, 40.0, 160.0, 1000.0)):
    """Four orthographic shadow frusta along the camera axis, built like the
    directional branch of calculateLightSpaceMatrices (light_system.cpp:122-132):
    lookAt(eye, eye+dir, +Y) and a symmetric ortho box around one depth slice."""
    direction = np.array([-0.5, -1.0, -0.3], np.float32)
    direction = direction / np.sqrt((direction * direction).sum(dtype=np.float32))
    cam = np.array(CAMERA["pos"], np.float32)
    fwd = np.array(CAMERA["forward"], np.float32)
    clips, near = [], np.float32(CAMERA["z_near"])
    for far in splits:
        far = np.float32(far)
        centre = cam + fwd * ((near + far) * np.float32(0.5))
        half = (far - near) * np.float32(0.75)
        dist = half + np.float32(50.0)
        eye = centre - direction * dist
        clips.append(_MATRICES.ortho_view(eye, centre, (0.0, 1.0, 0.0), -half, half, -half, half,
                                          0.1, float(dist + half)))
        near = far
    return clips


def _flat(name, n, seed, pos_range, views):
    pos, rot, scale, mesh, shader, flags = entity_block(seed, 0, n, pos_range, (0.25, 4.0))
    if views == "camera":
        clip = np.stack([camera_clip()])[None]
        masks = np.array([_gpu.PHASE_OPAQUE], np.uint8)
    else:
        clip = np.stack([camera_clip()] + cascade_clips())[None]
        masks = np.array([_gpu.PHASE_OPAQUE] + [_gpu.PHASE_SHADOW_CASTER] * 4, np.uint8)
    return Scene(name, pos, rot, scale, mesh, shader, flags, None, PREFAB_BOUNDS.copy(), clip.astype(np.float32), masks)


def config1(n=10_000):
    """Headless flat scene, single camera (BASELINE.json configs[0])."""
    return _flat("config1-flat-10k", n, 1, 200.0, "camera")


def config2(n=1_048_576, levels=8):
    """8-deep hierarchy, level k has (n // (2^levels - 1)) * 2^k entities (the last
    level takes the remainder); children are contiguous per parent."""
    pos, rot, scale, mesh, shader, flags = entity_block(2, 0, n, 10.0, (0.8, 1.25))
    base = n // ((1 << levels) - 1)
    sizes = [base << k for k in range(levels)]
    sizes[-1] += n - sum(sizes)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    parent = np.full(n, _gpu.NO_PARENT, np.uint32)
    for k in range(1, levels):
        j = np.arange(sizes[k], dtype=np.int64)
        parent[starts[k]:starts[k + 1]] = (starts[k - 1] + np.minimum(j // 2, sizes[k - 1] - 1)).astype(np.uint32)
    clip = np.stack([camera_clip()])[None].astype(np.float32)
    return Scene("config2-hier-1M", pos, rot, scale, mesh, shader, flags, parent, PREFAB_BOUNDS.copy(), clip,
                 np.array([_gpu.PHASE_OPAQUE], np.uint8))


def config3(n=16_777_216):
    """16M flat entities, camera + 4 shadow-cascade frusta (the headline metric)."""
    return _flat("config3-flat-16M-5views", n, 3, 2000.0, "cascades")


def config4(n=67_108_864, lo=None, hi=None):
    """64M flat entities; [lo, hi) selects one rank's range without building the rest."""
    lo, hi = (0 if lo is None else lo), (n if hi is None else hi)
    pos, rot, scale, mesh, shader, flags = entity_block(4, lo, hi, 2000.0, (0.25, 4.0))
    clip = np.stack([camera_clip()] + cascade_clips())[None].astype(np.float32)
    masks = np.array([_gpu.PHASE_OPAQUE] + [_gpu.PHASE_SHADOW_CASTER] * 4, np.uint8)
    return Scene(f"config4-flat-64M[{lo}:{hi}]", pos, rot, scale, mesh, shader, flags, None, PREFAB_BOUNDS.copy(), clip, masks)


def config5(n_worlds=4096, entities_per_world=16_384, w_lo=None, w_hi=None):
    """Independent worlds, one yawed camera each; [w_lo, w_hi) selects one rank's worlds."""
    w_lo, w_hi = (0 if w_lo is None else w_lo), (n_worlds if w_hi is None else w_hi)
    parts = [entity_block(5 + w, 0, entities_per_world, 200.0, (0.25, 4.0)) for w in range(w_lo, w_hi)]
    pos, rot, scale, mesh, shader, flags = (np.concatenate([p[i] for p in parts]) for i in range(6))
    clip = np.stack([camera_clip(w * 360.0 / n_worlds) for w in range(w_lo, w_hi)])[:, None, :].astype(np.float32)
    return Scene(f"config5-worlds[{w_lo}:{w_hi}]x{entities_per_world}", pos, rot, scale, mesh, shader, flags, None,
                 PREFAB_BOUNDS.copy(), clip, np.array([_gpu.PHASE_OPAQUE], np.uint8), w_hi - w_lo, entities_per_world)


def load_into(ctx, scene):
    """Uploads a scene into a GpuContext (mirror + bounds + views)."""
    ctx.upload_entities(0, scene.pos, scene.rot, scene.scale, scene.parent, scene.mesh, scene.shader, scene.flags)
    ctx.set_entity_count(scene.n)
    ctx.set_mesh_bounds(scene.bounds)
    if scene.n_worlds > 1:
        ctx.set_world_views(scene.n_worlds, scene.entities_per_world, scene.clip, scene.masks)
    else:
        ctx.set_views(scene.clip[0], scene.masks)
