"""ctypes binding of oracle/astre_oracle.c -- the CPU checker.  Imported by tests,
__graft_entry__.smoke() and bench.py's CPU-baseline legs only; never by astre_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "_build", "libastre_oracle.so")

_F = C.POINTER(C.c_float)
_U32 = C.POINTER(C.c_uint32)
_U64 = C.POINTER(C.c_uint64)
_U8 = C.POINTER(C.c_uint8)


class Bounds(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("extent", C.c_float * 3), ("radius", C.c_float), ("reserved", C.c_uint32)]


class View(C.Structure):
    _fields_ = [("plane", (C.c_float * 4) * 6), ("aplane", (C.c_float * 4) * 6), ("depth", C.c_float * 4),
                ("depth_scale", C.c_float), ("phase_mask", C.c_uint32), ("pad", C.c_uint32 * 2),
                ("lo", C.c_float * 4), ("hi", C.c_float * 4)]


def build():
    src = os.path.join(ORACLE_DIR, "astre_oracle.c")
    if os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(src):
        return SO
    subprocess.run(["make", "-C", ORACLE_DIR, "_build/libastre_oracle.so"], check=True, capture_output=True)
    return SO


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_cull_keys.restype = C.c_uint64
        lib.orc_cull_keys_mt.restype = C.c_uint64
        lib.orc_frame.restype = C.c_uint64
        lib.orc_sort_keys.argtypes = [_U64, C.c_uint64]
        lib.orc_perspective.argtypes = [C.c_float] * 4 + [_F]
        lib.orc_ortho.argtypes = [C.c_float] * 6 + [_F]
        lib.orc_camera_view_proj.argtypes = [_F, _F, _F] + [C.c_float] * 4 + [_F, _F]
        lib.orc_light_space.argtypes = [_F, _F, C.c_int, C.c_float, _F]
        lib.orc_interpolate_trs.argtypes = [C.c_uint32, C.c_float] + [_F] * 7

    # -- GLM restatements --
    def trs_matrix(self, pos, rot, scale):
        pos, rot, scale = (np.ascontiguousarray(a, np.float32) for a in (pos, rot, scale))
        out = np.empty(16, np.float32)
        self.lib.orc_trs_matrix(_p(pos, C.c_float), _p(rot, C.c_float), _p(scale, C.c_float), _p(out, C.c_float))
        return out

    def basis(self, rot):
        rot = np.ascontiguousarray(rot, np.float32)
        f, u, r = (np.empty(3, np.float32) for _ in range(3))
        self.lib.orc_basis(_p(rot, C.c_float), _p(f, C.c_float), _p(u, C.c_float), _p(r, C.c_float))
        return f, u, r

    def mat4_mul(self, a, b):
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        out = np.empty(16, np.float32)
        self.lib.orc_mat4_mul(_p(a, C.c_float), _p(b, C.c_float), _p(out, C.c_float))
        return out

    def look_at(self, eye, center, up):
        e, c, u = (np.ascontiguousarray(a, np.float32) for a in (eye, center, up))
        out = np.empty(16, np.float32)
        self.lib.orc_look_at(_p(e, C.c_float), _p(c, C.c_float), _p(u, C.c_float), _p(out, C.c_float))
        return out

    def perspective(self, fovy, aspect, zn, zf):
        out = np.empty(16, np.float32)
        self.lib.orc_perspective(fovy, aspect, zn, zf, _p(out, C.c_float))
        return out

    def ortho(self, l, r, b, t, n, f):
        out = np.empty(16, np.float32)
        self.lib.orc_ortho(l, r, b, t, n, f, _p(out, C.c_float))
        return out

    def camera_view_proj(self, pos, fwd, up, fov_deg, aspect, zn, zf):
        p, f, u = (np.ascontiguousarray(a, np.float32) for a in (pos, fwd, up))
        view, proj = np.empty(16, np.float32), np.empty(16, np.float32)
        self.lib.orc_camera_view_proj(_p(p, C.c_float), _p(f, C.c_float), _p(u, C.c_float), fov_deg, aspect, zn, zf,
                                      _p(view, C.c_float), _p(proj, C.c_float))
        return view, proj

    def light_space(self, pos, direction, light_type, outer_cutoff=0.0):
        p, d = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(direction, np.float32)
        out = np.empty(16, np.float32)
        ok = self.lib.orc_light_space(_p(p, C.c_float), _p(d, C.c_float), int(light_type), float(outer_cutoff), _p(out, C.c_float))
        return out if ok else None

    # -- path --
    def transform_flat(self, pos, rot, scale, basis=False):
        pos, rot, scale = (np.ascontiguousarray(a, np.float32) for a in (pos, rot, scale))
        n = len(pos)
        world = np.empty((n, 16), np.float32)
        f = u = r = None
        if basis:
            f, u, r = (np.empty((n, 3), np.float32) for _ in range(3))
        self.lib.orc_transform_flat(C.c_uint32(n), _p(pos, C.c_float), _p(rot, C.c_float), _p(scale, C.c_float),
                                    _p(world, C.c_float), _p(f, C.c_float), _p(u, C.c_float), _p(r, C.c_float))
        return (world, f, u, r) if basis else world

    def transform_hierarchy(self, pos, rot, scale, parent):
        pos, rot, scale = (np.ascontiguousarray(a, np.float32) for a in (pos, rot, scale))
        parent = np.ascontiguousarray(parent, np.uint32)
        n = len(pos)
        world = np.empty((n, 16), np.float32)
        levels = self.lib.orc_transform_hierarchy(C.c_uint32(n), _p(pos, C.c_float), _p(rot, C.c_float),
                                                  _p(scale, C.c_float), _p(parent, C.c_uint32), _p(world, C.c_float))
        if levels < 0:
            raise ValueError("bad hierarchy")
        return world, levels

    def make_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, np.float32).reshape(-1, 7)
        recs = (Bounds * max(len(b), 1))()
        for i, row in enumerate(b):
            recs[i].center[:] = row[0:3].tolist()
            recs[i].extent[:] = row[3:6].tolist()
            recs[i].radius = float(row[6])
        return recs, len(b)

    def make_views(self, clip, masks):
        """clip [n_worlds, F, 16], masks [F] -> View array in [world][view] order."""
        clip = np.ascontiguousarray(clip, np.float32).reshape(-1, len(masks), 16)
        recs = (View * max(clip.shape[0] * clip.shape[1], 1))()
        k = 0
        for w in range(clip.shape[0]):
            for v in range(clip.shape[1]):
                m = np.ascontiguousarray(clip[w, v])
                self.lib.orc_extract_view(_p(m, C.c_float), C.c_uint32(int(masks[v]) & 15), C.byref(recs[k]))
                k += 1
        return recs

    def cull_keys(self, world, mesh, shader, flags, bounds, clip, masks, n_worlds=1, entities_per_world=0, sort=True):
        world = np.ascontiguousarray(world, np.float32)
        mesh, shader = np.ascontiguousarray(mesh, np.uint32), np.ascontiguousarray(shader, np.uint32)
        flags = np.ascontiguousarray(flags, np.uint8)
        n, F = len(world), len(masks)
        brecs, nb = self.make_bounds(bounds)
        vrecs = self.make_views(clip, masks)
        cap = max(n * F, 1)
        keys = np.empty(cap, np.uint64)
        counts = np.zeros(max(n_worlds * F, 1), np.uint32)
        total = self.lib.orc_cull_keys(C.c_uint32(n), _p(world, C.c_float), _p(mesh, C.c_uint32), _p(shader, C.c_uint32),
                                       _p(flags, C.c_uint8), brecs, C.c_uint32(nb), C.c_uint32(n_worlds),
                                       C.c_uint32(entities_per_world), C.c_uint32(F), vrecs,
                                       _p(keys, C.c_uint64), C.c_uint64(cap), _p(counts, C.c_uint32))
        keys = keys[:total].copy()
        if sort:
            self.lib.orc_sort_keys(_p(keys, C.c_uint64), C.c_uint64(len(keys)))
        return keys, counts[:n_worlds * F]

    def frame(self, scene, sort=True):
        """Oracle result of a whole scene: (world [n,16], sorted keys, counts)."""
        if scene.parent is not None:
            world, _ = self.transform_hierarchy(scene.pos, scene.rot, scene.scale, scene.parent)
        else:
            world = self.transform_flat(scene.pos, scene.rot, scene.scale)
        keys, counts = self.cull_keys(world, scene.mesh, scene.shader, scene.flags, scene.bounds, scene.clip,
                                      scene.masks, scene.n_worlds, scene.entities_per_world, sort)
        return world, keys, counts

    def interpolate(self, alpha, a, b):
        pa, ra, sa = (np.ascontiguousarray(x, np.float32) for x in a)
        pb, rb, sb = (np.ascontiguousarray(x, np.float32) for x in b)
        n = len(pa)
        world = np.empty((n, 16), np.float32)
        self.lib.orc_interpolate_trs(C.c_uint32(n), C.c_float(alpha), _p(pa, C.c_float), _p(ra, C.c_float), _p(sa, C.c_float),
                                     _p(pb, C.c_float), _p(rb, C.c_float), _p(sb, C.c_float), _p(world, C.c_float))
        return world

    def merge_lists(self, lists, slot_bits=26):
        """f3: G sorted lists (list r = shard r) -> global list ordered (key >> slot_bits, shard, slot)."""
        lists = [np.ascontiguousarray(a, np.uint64) for a in lists]
        g = len(lists)
        ptrs = (C.POINTER(C.c_uint64) * g)(*[_p(a, C.c_uint64) for a in lists])
        counts = np.array([len(a) for a in lists], np.uint64)
        total = int(counts.sum())
        keys = np.empty(total, np.uint64)
        src = np.empty(total, np.uint8)
        self.lib.orc_merge_lists(C.c_uint32(g), ptrs, _p(counts, C.c_uint64), C.c_uint32(slot_bits),
                                 _p(keys, C.c_uint64), _p(src, C.c_uint8))
        return keys, src

    def sort_keys(self, keys):
        k = np.ascontiguousarray(keys, np.uint64).copy()
        self.lib.orc_sort_keys(_p(k, C.c_uint64), C.c_uint64(len(k)))
        return k


_ORACLE = None


def load():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle(C.CDLL(build()))
    return _ORACLE


def bits(a):
    """uint32 view of float32 data, with every NaN mapped to one pattern (x86 and the
    GPU produce different NaN payloads; both are 'NaN')."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).copy()
    u[np.isnan(a)] = 0x7FC00000
    return u


def ulp_diff(a, b):
    """Per-element distance in units of least precision (sign-magnitude ordering)."""
    ia = bits(a).astype(np.int64)
    ib = bits(b).astype(np.int64)
    ia = np.where(ia & 0x80000000, 0x80000000 - ia, ia)
    ib = np.where(ib & 0x80000000, 0x80000000 - ib, ib)
    return np.abs(ia - ib)
