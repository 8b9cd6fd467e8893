"""ctypes binding of oracle/_ref/libastre_ref.so -- the reference's OWN transform / camera / visual / light /
interpolateFrame sources compiled from /root/reference by oracle/ref_recipe/Makefile.  Test infrastructure:
imported by tests/ and bench.py's CPU legs only, never by astre_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECIPE = os.path.join(ROOT, "oracle", "ref_recipe")
SO = os.path.join(ROOT, "oracle", "_ref", "libastre_ref.so")
REFERENCE = "/root/reference"

HAS_POSITION, HAS_ROTATION, HAS_SCALE = 1, 2, 4
DIRECTIONAL, POINT, SPOT = 1, 2, 3

_F = C.POINTER(C.c_float)
_U64 = C.POINTER(C.c_uint64)
_U8 = C.POINTER(C.c_uint8)


def available():
    """True when the library exists or can be built here (the reference tree is present)."""
    return os.path.exists(SO) or os.path.isdir(REFERENCE)


def build():
    """Build from the reference tree when it is present (this container); on the GPU box the prebuilt
    library travels with the snapshot and the reference tree does not exist."""
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", RECIPE], check=True, capture_output=True)
    if not os.path.exists(SO):
        raise FileNotFoundError(SO + " (needs /root/reference to build)")
    return SO


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(build())
        lib.ref_world_create.restype = C.c_void_p
        lib.ref_world_destroy.argtypes = [C.c_void_p]
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_last_error.argtypes = [C.c_void_p]
        lib.ref_add_transforms.argtypes = [C.c_void_p, C.c_uint32, _U64, _U8, _F, _F, _F]
        lib.ref_set_trs.argtypes = [C.c_void_p, C.c_uint32, _U64, _F, _F, _F]
        lib.ref_add_visual.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_char_p, _F]
        lib.ref_add_camera.argtypes = [C.c_void_p, C.c_uint64] + [C.c_float] * 4
        lib.ref_add_light.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, _F] + [C.c_float] * 5
        lib.ref_register_vertex_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        lib.ref_register_shader.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        lib.ref_run_transform.argtypes = [C.c_void_p]
        lib.ref_get_transforms.argtypes = [C.c_void_p, C.c_uint32, _U64, _F, _F, _F, _F]
        lib.ref_run_camera.argtypes = [C.c_void_p, C.c_uint64, C.c_int, _F, _F, _F]
        lib.ref_run_visual.argtypes = [C.c_void_p, C.c_int]
        lib.ref_proxy_count.restype = C.c_uint64
        lib.ref_proxy_count.argtypes = [C.c_void_p, C.c_int]
        lib.ref_proxy_ids.argtypes = [C.c_void_p, C.c_int, _U64, C.c_uint64]
        lib.ref_get_proxy.argtypes = [C.c_void_p, C.c_int, C.c_uint64, _U64, _F]
        lib.ref_get_models.argtypes = [C.c_void_p, C.c_int, C.c_uint32, _U64, _F]
        lib.ref_run_light.argtypes = [C.c_void_p, C.c_int]
        lib.ref_shadow_casters_count.restype = C.c_uint32
        lib.ref_shadow_casters_count.argtypes = [C.c_void_p, C.c_int]
        lib.ref_get_light.argtypes = [C.c_void_p, C.c_int, C.c_uint64, _F]
        lib.ref_light_space_matrices.argtypes = [C.c_void_p, C.c_int, _F, C.c_uint32]
        lib.ref_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
        lib.ref_get_frame_camera.argtypes = [C.c_void_p, C.c_int, _F, _F, _F]
        lib.ref_time_tick.restype = C.c_double
        lib.ref_time_tick.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _LIB = lib
    return _LIB


class RefWorld:
    """One reference `ecs::Registry` + the four systems + four `render::Frame` slots."""

    def __init__(self):
        self.lib = _lib()
        self.h = self.lib.ref_world_create()
        if not self.h:
            raise RuntimeError("ref_world_create failed")

    def close(self):
        if self.h:
            self.lib.ref_world_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError(self.lib.ref_last_error(self.h).decode())
        return rc

    # -- world building (EntityLoader::load equivalent) --
    def add_transforms(self, ids, pos, rot, scale, has=None):
        ids = np.ascontiguousarray(ids, np.uint64)
        pos, rot, scale = _f32(pos), _f32(rot), _f32(scale)
        has = None if has is None else np.ascontiguousarray(has, np.uint8)
        self._check(self.lib.ref_add_transforms(self.h, len(ids), _p(ids, C.c_uint64), _p(has, C.c_uint8),
                                                _p(pos, C.c_float), _p(rot, C.c_float), _p(scale, C.c_float)))

    def set_trs(self, ids, pos=None, rot=None, scale=None):
        ids = np.ascontiguousarray(ids, np.uint64)
        pos, rot, scale = _f32(pos), _f32(rot), _f32(scale)
        self._check(self.lib.ref_set_trs(self.h, len(ids), _p(ids, C.c_uint64), _p(pos, C.c_float), _p(rot, C.c_float),
                                         _p(scale, C.c_float)))

    def add_visual(self, eid, visible, vertex_buffer_name, shader_name, color=None):
        color = _f32(color)
        self._check(self.lib.ref_add_visual(self.h, int(eid), int(bool(visible)), vertex_buffer_name.encode(),
                                            shader_name.encode(), _p(color, C.c_float)))

    def add_camera(self, eid, fov_deg, near, far, aspect):
        self._check(self.lib.ref_add_camera(self.h, int(eid), fov_deg, near, far, aspect))

    def add_light(self, eid, light_type, cast_shadows, color=(1, 1, 1, 1), constant=1.0, linear=0.0, quadratic=0.0,
                  inner_cutoff=0.0, outer_cutoff=0.0):
        color = _f32(color)
        self._check(self.lib.ref_add_light(self.h, int(eid), int(light_type), int(bool(cast_shadows)), _p(color, C.c_float),
                                           constant, linear, quadratic, inner_cutoff, outer_cutoff))

    def register_vertex_buffer(self, name, vb_id):
        self.lib.ref_register_vertex_buffer(self.h, name.encode(), int(vb_id))

    def register_shader(self, name, sh_id):
        self.lib.ref_register_shader(self.h, name.encode(), int(sh_id))

    # -- the reference's systems --
    def run_transform(self):
        self._check(self.lib.ref_run_transform(self.h))

    def transforms(self, ids):
        """-> (transform_matrix [n,16], forward [n,3], up [n,3], right [n,3]) as TransformSystem::run left them."""
        ids = np.ascontiguousarray(ids, np.uint64)
        n = len(ids)
        m, f, u, r = np.empty((n, 16), np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        self._check(self.lib.ref_get_transforms(self.h, n, _p(ids, C.c_uint64), _p(m, C.c_float), _p(f, C.c_float),
                                                _p(u, C.c_float), _p(r, C.c_float)))
        return m, f, u, r

    def run_camera(self, camera_entity, frame=0):
        v, p, pos = np.empty(16, np.float32), np.empty(16, np.float32), np.empty(3, np.float32)
        self._check(self.lib.ref_run_camera(self.h, int(camera_entity), frame, _p(v, C.c_float), _p(p, C.c_float), _p(pos, C.c_float)))
        return v, p, pos

    def run_visual(self, frame=0):
        self._check(self.lib.ref_run_visual(self.h, frame))

    def proxy_ids(self, frame=0):
        n = int(self.lib.ref_proxy_count(self.h, frame))
        ids = np.empty(n, np.uint64)
        self._check(self.lib.ref_proxy_ids(self.h, frame, _p(ids, C.c_uint64), n))
        return ids

    def proxy(self, eid, frame=0):
        """-> dict of the RenderProxy fields, or None when VisualSystem emitted no proxy for `eid`."""
        u, f = np.empty(5, np.uint64), np.empty(30, np.float32)
        rc = self._check(self.lib.ref_get_proxy(self.h, frame, int(eid), _p(u, C.c_uint64), _p(f, C.c_float)))
        if rc == 1:
            return None
        return {"visible": bool(u[0]), "phases": int(u[1]), "vertex_buffer": int(u[2]), "shader": int(u[3]),
                "useTexture": bool(u[4]), "position": f[0:3].copy(), "rotation": f[3:7].copy(), "scale": f[7:10].copy(),
                "uModel": f[10:26].copy(), "uColor": f[26:30].copy()}

    def models(self, ids, frame=0):
        ids = np.ascontiguousarray(ids, np.uint64)
        m = np.empty((len(ids), 16), np.float32)
        self._check(self.lib.ref_get_models(self.h, frame, len(ids), _p(ids, C.c_uint64), _p(m, C.c_float)))
        return m

    def run_light(self, frame=0):
        self._check(self.lib.ref_run_light(self.h, frame))

    def shadow_casters_count(self, frame=0):
        return int(self.lib.ref_shadow_casters_count(self.h, frame))

    def light(self, eid, frame=0):
        out = np.empty(20, np.float32)
        self._check(self.lib.ref_get_light(self.h, frame, int(eid), _p(out, C.c_float)))
        return out

    def light_space_matrices(self, frame=0):
        out = np.empty((16, 16), np.float32)
        n = self._check(self.lib.ref_light_space_matrices(self.h, frame, _p(out, C.c_float), 16))
        return out[:n].copy()

    def interpolate(self, a, b, alpha, dst):
        self._check(self.lib.ref_interpolate(self.h, a, b, alpha, dst))

    def frame_camera(self, frame):
        v, p, pos = np.empty(16, np.float32), np.empty(16, np.float32), np.empty(3, np.float32)
        self.lib.ref_get_frame_camera(self.h, frame, _p(v, C.c_float), _p(p, C.c_float), _p(pos, C.c_float))
        return v, p, pos

    def time_tick(self, iters=1, with_visual=True):
        """Seconds for `iters` x (TransformSystem::run [+ VisualSystem::run]) -- one thread, as the reference runs them."""
        s = self.lib.ref_time_tick(self.h, iters, int(with_visual))
        if s < 0:
            raise RuntimeError(self.lib.ref_last_error(self.h).decode())
        return s
