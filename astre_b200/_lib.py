"""ctypes binding of include/astre_gpu.h.  No fallback: a missing library is an error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class AstreError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"astre_gpu error {code}: {message}")
        self.code = code


def library_path():
    return os.path.join(_HERE, "libastre_gpu.so")


class DrawRun(C.Structure):
    _fields_ = [("first", C.c_uint32), ("count", C.c_uint32), ("world_view", C.c_uint32), ("shader_mesh", C.c_uint32)]


class MeshDrawInfo(C.Structure):
    _fields_ = [("index_count", C.c_uint32), ("first_index", C.c_uint32), ("base_vertex", C.c_int32)]


class DrawIndirect(C.Structure):     # OpenGL's DrawElementsIndirectCommand
    _fields_ = [("count", C.c_uint32), ("instance_count", C.c_uint32), ("first_index", C.c_uint32), ("base_vertex", C.c_int32),
                ("base_instance", C.c_uint32)]


class MeshBounds(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("extent", C.c_float * 3), ("radius", C.c_float), ("reserved", C.c_uint32)]


_P = C.c_void_p
_F = C.POINTER(C.c_float)
_U32 = C.POINTER(C.c_uint32)
_U64 = C.POINTER(C.c_uint64)
_U8 = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol include/astre_gpu.h declares
SIGNATURES = {
    "astre_gpu_create": (_P, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64]),
    "astre_gpu_destroy": (None, [_P]),
    "astre_gpu_last_error": (C.c_char_p, [_P]),
    "astre_gpu_version": (C.c_char_p, []),
    "astre_gpu_upload_entities": (C.c_int, [_P, C.c_uint32, C.c_uint32, _U64, _F, _F, _F, _U32, _U32, _U32, _U8]),
    "astre_gpu_set_entity_count": (C.c_int, [_P, C.c_uint32]),
    "astre_gpu_update_trs": (C.c_int, [_P, C.c_uint32, _U32, _F, _F, _F]),
    "astre_gpu_update_trs_range": (C.c_int, [_P, C.c_uint32, C.c_uint32, _F, _F, _F]),
    "astre_gpu_update_flags": (C.c_int, [_P, C.c_uint32, _U32, _U8]),
    "astre_gpu_set_mesh_bounds": (C.c_int, [_P, C.c_uint32, C.POINTER(MeshBounds)]),
    "astre_gpu_set_views": (C.c_int, [_P, C.c_uint32, _F, _U8]),
    "astre_gpu_set_world_views": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _F, _U8]),
    "astre_gpu_frame": (C.c_int, [_P, C.c_uint32, _P]),
    "astre_gpu_sort": (C.c_int, [_P, _P]),
    "astre_gpu_last_sort_path": (C.c_int, [_P, _U32]),
    "astre_gpu_graph_stats": (C.c_int, [_P, _U32, _U32, _U32]),
    "astre_gpu_debug_sort_keys": (C.c_int, [_P, _U64, C.c_uint64, C.c_uint64, C.c_uint32, _U64, _U32]),
    "astre_gpu_set_counts_mirror": (C.c_int, [_P, _P, _P]),
    "astre_gpu_frame_index": (C.c_int, [_P, _U64]),
    "astre_gpu_count_board_create": (C.c_int, [_P, C.c_uint32, _P, C.POINTER(_P)]),
    "astre_gpu_count_board_attach": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "astre_gpu_count_board_exchange": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "astre_gpu_count_board_status": (C.c_int, [_P, _U32, _U32]),
    "astre_gpu_global_offsets_dev": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P]),
    "astre_gpu_debug_guard_check": (C.c_int64, []),
    "astre_gpu_debug_guard_selftest": (C.c_int64, []),
    "astre_gpu_sync": (C.c_int, [_P]),
    "astre_gpu_set_stream": (C.c_int, [_P, _P]),
    "astre_gpu_launch_count": (C.c_uint64, [_P]),
    "astre_gpu_last_frame_ms": (C.c_int, [_P, _F]),
    "astre_gpu_last_cull_ms": (C.c_int, [_P, _F]),
    "astre_gpu_snapshot_trs": (C.c_int, [_P, _P]),
    "astre_gpu_interpolate": (C.c_int, [_P, C.c_float, _P]),
    "astre_gpu_read_world": (C.c_int, [_P, C.c_uint32, C.c_uint32, _F]),
    "astre_gpu_read_basis": (C.c_int, [_P, C.c_uint32, C.c_uint32, _F, _F, _F]),
    "astre_gpu_read_counts": (C.c_int, [_P, _U32]),
    "astre_gpu_read_total": (C.c_int, [_P, _U64]),
    "astre_gpu_read_keys": (C.c_int, [_P, C.c_uint32, C.c_uint32, _U64, C.c_uint64, _U64]),
    "astre_gpu_read_all_keys": (C.c_int, [_P, _U64, C.c_uint64, _U64]),
    "astre_gpu_read_visible_world": (C.c_int, [_P, C.c_uint64, C.c_uint64, _F]),
    "astre_gpu_read_draw_list": (C.c_int, [_P, _U64, _F, C.c_uint64, _U64]),
    "astre_gpu_read_runs": (C.c_int, [_P, C.POINTER(DrawRun), C.c_uint32, _U32]),
    "astre_gpu_set_mesh_draw_info": (C.c_int, [_P, C.c_uint32, C.POINTER(MeshDrawInfo)]),
    "astre_gpu_build_indirect": (C.c_int, [_P, _U32, C.POINTER(_P)]),
    "astre_gpu_read_indirect": (C.c_int, [_P, C.POINTER(DrawIndirect), C.c_uint32, _U32]),
    "astre_gpu_slot_entity": (C.c_int, [_P, C.c_uint32, _U64]),
    "astre_gpu_device_world": (_P, [_P]),
    "astre_gpu_device_counts": (_P, [_P]),
    "astre_gpu_device_sorted_keys": (_P, [_P]),
    "astre_gpu_publish_keys": (C.c_int, [_P, C.c_uint32, _P]),
    "astre_gpu_device_published_keys": (_P, [_P, C.c_uint32]),
    "astre_gpu_export_handle": (C.c_int, [_P, C.c_uint32, _P]),
    "astre_gpu_import_handle": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "astre_gpu_release_import": (C.c_int, [_P, _P]),
    "astre_gpu_merge_lists": (C.c_int, [_P, C.c_uint32, C.POINTER(_P), _U64, C.c_uint64, C.c_uint64, _P]),
    "astre_gpu_merge_prepare": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(_P), _U64, _P]),
    "astre_gpu_merge_published": (C.c_int, [_P, C.c_uint32, C.POINTER(_P), _U64, C.c_uint64, C.c_uint64, _P]),
    "astre_gpu_merge_prepare_dev": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(_P), _P, C.c_uint32, _P]),
    "astre_gpu_merge_published_dev": (C.c_int, [_P, C.c_uint32, C.POINTER(_P), C.c_uint32, C.c_uint32, C.c_uint64, _P]),
    "astre_gpu_merge_p2p": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(_P), C.POINTER(_P), C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_uint64, _P]),
    "astre_gpu_merged_count": (C.c_int, [_P, _U64]),
    "astre_gpu_device_merged_keys": (_P, [_P]),
    "astre_gpu_device_merged_sources": (_P, [_P]),
    "astre_gpu_read_merged": (C.c_int, [_P, C.c_uint64, C.c_uint64, _U64, C.POINTER(C.c_uint8)]),
    "astre_gpu_last_merge_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "astre_camera_matrices": (None, [_F, _F, _F, C.c_float, C.c_float, C.c_float, C.c_float, _F, _F, _F]),
    "astre_light_space_matrix": (C.c_int, [_F, _F, C.c_int, C.c_float, _F]),
    "astre_ortho_view_matrix": (None, [_F, _F, _F, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _F]),
    "astre_global_offsets": (None, [_U32, C.c_uint32, C.c_uint32, C.c_uint32, _U64, _U64]),
    "astre_global_offsets_rank_major": (None, [_U32, C.c_uint32, C.c_uint32, C.c_uint32, _U64, _U64]),
}


def load_library():
    """Loads libastre_gpu.so (built by astre_b200/csrc/Makefile).  Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `make -C astre_b200/csrc` (or __graft_entry__.build()). "
            "astre_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
