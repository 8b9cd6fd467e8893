"""GpuContext: numpy-facing wrapper over the C ABI (one context = one GPU's entity shard)."""
import ctypes as C

import numpy as np

from ._lib import AstreError, DrawIndirect, DrawRun, MeshBounds, MeshDrawInfo, load_library

FRAME_TRANSFORM, FRAME_CULL, FRAME_SORT, FRAME_BASIS, FRAME_SORT_LATER, FRAME_GRAPH, FRAME_GATHER = 1, 2, 4, 8, 16, 32, 64
FRAME_ALL = 7

PHASE_OPAQUE, PHASE_TRANSPARENT, PHASE_SHADOW_CASTER, PHASE_DEBUG = 1, 2, 4, 8
NO_PARENT = 0xFFFFFFFF


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def _arr(a, dtype, shape_tail=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape_tail is not None and a.ndim > 1:
        assert a.shape[1:] == shape_tail, (a.shape, shape_tail)
    return a


def make_flags(visible, phases):
    """Per-entity flags byte: bit0 visible, bits 1..4 RenderPhase (visual_system.cpp:32,55)."""
    return (np.asarray(visible).astype(np.uint8) & 1) | ((np.asarray(phases).astype(np.uint8) & 15) << 1)


class GpuContext:
    def __init__(self, max_entities, max_views=8, max_keys=0, device=0):
        self._lib = load_library()
        self.bound_stream = None      # cudaStream_t the context enqueues on (None: its own non-blocking stream)
        self._h = self._lib.astre_gpu_create(int(device), int(max_entities), int(max_views), int(max_keys))
        if not self._h:
            raise AstreError(-1, self._lib.astre_gpu_last_error(None).decode())
        self.max_entities = int(max_entities)
        self.device = int(device)
        self.n_worlds, self.views_per_world = 1, 0

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise AstreError(rc, self._lib.astre_gpu_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.astre_gpu_destroy(self._h)
            self._h = None
            # ASTRE_GUARD=1 (the GPU tests): a kernel that stored outside one of the library's buffers fails here
            damaged = self._lib.astre_gpu_debug_guard_check()
            if damaged > 0:
                raise RuntimeError(f"astre_gpu: {damaged} guard bytes around device buffers were overwritten")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mirror --
    def upload_entities(self, first, pos=None, rot=None, scale=None, parent=None, mesh=None, shader=None, flags=None,
                        entity_id=None, count=None):
        pos, rot, scale = _arr(pos, np.float32), _arr(rot, np.float32), _arr(scale, np.float32)
        parent, mesh, shader = _arr(parent, np.uint32), _arr(mesh, np.uint32), _arr(shader, np.uint32)
        flags, entity_id = _arr(flags, np.uint8), _arr(entity_id, np.uint64)
        if count is None:
            for a, per in ((pos, 3), (rot, 4), (scale, 3), (parent, 1), (mesh, 1), (shader, 1), (flags, 1), (entity_id, 1)):
                if a is not None:
                    count = a.size // per
                    break
        self._keep = (pos, rot, scale, parent, mesh, shader, flags, entity_id)
        self._check(self._lib.astre_gpu_upload_entities(
            self._h, int(first), int(count), _ptr(entity_id, C.c_uint64), _ptr(pos, C.c_float), _ptr(rot, C.c_float),
            _ptr(scale, C.c_float), _ptr(parent, C.c_uint32), _ptr(mesh, C.c_uint32), _ptr(shader, C.c_uint32),
            _ptr(flags, C.c_uint8)))

    def set_entity_count(self, n):
        self._check(self._lib.astre_gpu_set_entity_count(self._h, int(n)))

    def update_trs(self, slots, pos=None, rot=None, scale=None):
        slots = _arr(slots, np.uint32)
        pos, rot, scale = _arr(pos, np.float32), _arr(rot, np.float32), _arr(scale, np.float32)
        self._check(self._lib.astre_gpu_update_trs(self._h, slots.size, _ptr(slots, C.c_uint32), _ptr(pos, C.c_float),
                                                    _ptr(rot, C.c_float), _ptr(scale, C.c_float)))

    def update_trs_range(self, first, count, pos=None, rot=None, scale=None):
        pos, rot, scale = _arr(pos, np.float32), _arr(rot, np.float32), _arr(scale, np.float32)
        self._check(self._lib.astre_gpu_update_trs_range(self._h, int(first), int(count), _ptr(pos, C.c_float),
                                                          _ptr(rot, C.c_float), _ptr(scale, C.c_float)))

    def update_flags(self, slots, flags):
        slots, flags = _arr(slots, np.uint32), _arr(flags, np.uint8)
        self._check(self._lib.astre_gpu_update_flags(self._h, slots.size, _ptr(slots, C.c_uint32), _ptr(flags, C.c_uint8)))

    def set_mesh_bounds(self, bounds):
        """bounds: float array [n_meshes, 7] = center xyz, extent xyz, radius."""
        b = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 7)
        recs = (MeshBounds * len(b))()
        for i, row in enumerate(b):
            recs[i].center[:] = row[0:3].tolist()
            recs[i].extent[:] = row[3:6].tolist()
            recs[i].radius = float(row[6])
        self._check(self._lib.astre_gpu_set_mesh_bounds(self._h, len(b), recs))

    # -- views --
    def set_views(self, clip, phase_masks):
        clip = np.ascontiguousarray(clip, dtype=np.float32).reshape(-1, 16)
        masks = np.ascontiguousarray(phase_masks, dtype=np.uint8)
        assert len(masks) == len(clip)
        self._check(self._lib.astre_gpu_set_views(self._h, len(clip), _ptr(clip, C.c_float), _ptr(masks, C.c_uint8)))
        self.n_worlds, self.views_per_world = 1, len(clip)

    def set_world_views(self, n_worlds, entities_per_world, clip, phase_masks):
        masks = np.ascontiguousarray(phase_masks, dtype=np.uint8)
        clip = np.ascontiguousarray(clip, dtype=np.float32).reshape(n_worlds, len(masks), 16)
        self._check(self._lib.astre_gpu_set_world_views(self._h, int(n_worlds), int(entities_per_world), len(masks),
                                                         _ptr(clip, C.c_float), _ptr(masks, C.c_uint8)))
        self.n_worlds, self.views_per_world = int(n_worlds), len(masks)

    # -- frame --
    def frame(self, flags=FRAME_ALL, stream=None):
        self._check(self._lib.astre_gpu_frame(self._h, int(flags), C.c_void_p(stream) if stream else None))

    def sort(self, stream=None):
        """Sort the keys of a frame that ran with FRAME_CULL | FRAME_SORT_LATER (lets the count exchange overlap the sort)."""
        self._check(self._lib.astre_gpu_sort(self._h, stream))

    def set_counts_mirror(self, ptr0, ptr1):
        """Frame k also writes its per-view counts to device buffer ptr[k & 1] (k = frame_index()); 0, 0 switches it off."""
        self._check(self._lib.astre_gpu_set_counts_mirror(self._h, int(ptr0) or None, int(ptr1) or None))

    # ---- count boards (8e as stores into peer memory) ----
    def count_board_create(self, max_counts):
        """Allocates this context's count board -> (CUDA IPC handle as bytes, local device pointer)."""
        buf = (C.c_ubyte * 64)()
        out = C.c_void_p()
        self._check(self._lib.astre_gpu_count_board_create(self._h, int(max_counts), buf, C.byref(out)))
        return bytes(buf), out.value

    def count_board_attach(self, world_size, rank, board_ptrs):
        """board_ptrs[t]: device pointer of rank t's board (own entry ignored)."""
        ptrs = (C.c_void_p * int(world_size))(*[int(p) if p else None for p in board_ptrs])
        self._check(self._lib.astre_gpu_count_board_attach(self._h, int(world_size), int(rank), ptrs))

    def count_board_exchange(self, counts_ptr, n_counts, offsets_ptr, totals_ptr=0, table_ptr=0, stream=None, rank_major=False):
        """Enqueues one exchange: counts (device) -> every rank's board -> table [world, n_counts], offsets, totals (device)."""
        self._check(self._lib.astre_gpu_count_board_exchange(self._h, int(counts_ptr) or None, int(n_counts), 1 if rank_major else 0,
                                                             int(table_ptr) or None, int(offsets_ptr) or None, int(totals_ptr) or None, stream))

    def count_board_status(self):
        """(exchanges finished, error) of this rank's board; synchronises the device."""
        a, b = C.c_uint32(), C.c_uint32()
        self._check(self._lib.astre_gpu_count_board_status(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def frame_index(self):
        v = C.c_uint64()
        self._check(self._lib.astre_gpu_frame_index(self._h, C.byref(v)))
        return int(v.value)

    def debug_sort_keys(self, keys, live_mask, bin_bits):
        """Test hook: the frame's key sort on an arbitrary set of unique keys -> (sorted keys, path)."""
        keys = np.ascontiguousarray(keys, np.uint64)
        out = np.empty(len(keys), np.uint64)
        path = C.c_uint32(0)
        self._check(self._lib.astre_gpu_debug_sort_keys(self._h, _ptr(keys, C.c_uint64), len(keys), int(live_mask), int(bin_bits),
                                                        _ptr(out, C.c_uint64), C.byref(path)))
        return out, int(path.value)

    def graph_stats(self):
        """(captures or exec updates, view-record patches, bucket-digit bits of the last plan) of ASTRE_FRAME_GRAPH frames."""
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib.astre_gpu_graph_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def last_sort_path(self):
        """0 = no sort ran, 1 = bucket sort, 2 = LSD radix fallback (decided on the device)."""
        v = C.c_uint32(0)
        self._check(self._lib.astre_gpu_last_sort_path(self._h, C.byref(v)))
        return int(v.value)

    def global_offsets_dev(self, all_counts_ptr, world_size, rank, n_views, offsets_ptr, totals_ptr=0, stream=None, rank_major=False):
        """Segment offsets of this rank from an all-gathered count table, all in device memory (stream-ordered)."""
        self._check(self._lib.astre_gpu_global_offsets_dev(self._h, int(all_counts_ptr), int(world_size), int(rank), int(n_views),
                                                           1 if rank_major else 0, int(offsets_ptr) or None, int(totals_ptr) or None, stream))

    def sync(self):
        self._check(self._lib.astre_gpu_sync(self._h))

    def set_stream(self, stream):
        """stream: a cudaStream_t handle as int (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        self._check(self._lib.astre_gpu_set_stream(self._h, C.c_void_p(stream) if stream else None))
        self.bound_stream = int(stream) if stream else None

    @property
    def launch_count(self):
        return int(self._lib.astre_gpu_launch_count(self._h))

    def last_frame_ms(self):
        ms = C.c_float()
        self._check(self._lib.astre_gpu_last_frame_ms(self._h, C.byref(ms)))
        return ms.value

    def last_cull_ms(self):
        ms = C.c_float()
        self._check(self._lib.astre_gpu_last_cull_ms(self._h, C.byref(ms)))
        return ms.value

    def snapshot_trs(self, stream=None):
        self._check(self._lib.astre_gpu_snapshot_trs(self._h, C.c_void_p(stream) if stream else None))

    def interpolate(self, alpha, stream=None):
        self._check(self._lib.astre_gpu_interpolate(self._h, float(alpha), C.c_void_p(stream) if stream else None))

    # -- results --
    def read_world(self, first, count, out=None):
        out = np.empty((count, 16), np.float32) if out is None else out
        self._check(self._lib.astre_gpu_read_world(self._h, int(first), int(count), _ptr(out, C.c_float)))
        return out

    def read_basis(self, first, count):
        f, u, r = (np.empty((count, 3), np.float32) for _ in range(3))
        self._check(self._lib.astre_gpu_read_basis(self._h, int(first), int(count), _ptr(f, C.c_float),
                                                    _ptr(u, C.c_float), _ptr(r, C.c_float)))
        return f, u, r

    def read_counts(self, out=None):
        out = np.zeros(self.n_worlds * self.views_per_world, np.uint32) if out is None else out
        self._check(self._lib.astre_gpu_read_counts(self._h, _ptr(out, C.c_uint32)))
        return out

    def read_total(self):
        t = C.c_uint64()
        self._check(self._lib.astre_gpu_read_total(self._h, C.byref(t)))
        return t.value

    def read_keys(self, view, world=0):
        n = C.c_uint64()
        self._check(self._lib.astre_gpu_read_keys(self._h, int(world), int(view), None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint64)
        if n.value:
            self._check(self._lib.astre_gpu_read_keys(self._h, int(world), int(view), _ptr(out, C.c_uint64), n.value, C.byref(n)))
        return out

    def read_all_keys(self, out=None):
        n = C.c_uint64()
        if out is None:
            out = np.empty(self.read_total(), np.uint64)
        self._check(self._lib.astre_gpu_read_all_keys(self._h, _ptr(out, C.c_uint64), out.size, C.byref(n)))
        return out[:min(n.value, out.size)]

    def read_visible_world(self, first_key, n, out=None):
        out = np.empty((n, 16), np.float32) if out is None else out
        self._check(self._lib.astre_gpu_read_visible_world(self._h, int(first_key), int(n), _ptr(out, C.c_float)))
        return out

    def read_draw_list(self, keys_out, world_out):
        """Sorted keys and the world matrices of their entities in list order, into caller buffers (uint64 [cap], float32
        [cap, 16]); returns the list length.  One synchronisation for the frame, one for both copies."""
        n = C.c_uint64()
        cap = min(len(keys_out), len(world_out))
        self._check(self._lib.astre_gpu_read_draw_list(self._h, _ptr(keys_out, C.c_uint64), _ptr(world_out, C.c_float), cap, C.byref(n)))
        return int(n.value)

    def read_runs(self):
        """Draw runs of the sorted list: array [n, 4] = first, count, world<<5|view, shader<<11|mesh."""
        n = C.c_uint32()
        self._check(self._lib.astre_gpu_read_runs(self._h, None, 0, C.byref(n)))
        runs = (DrawRun * max(n.value, 1))()
        self._check(self._lib.astre_gpu_read_runs(self._h, runs, n.value, C.byref(n)))
        return np.array([[r.first, r.count, r.world_view, r.shader_mesh] for r in runs[:n.value]], np.uint32).reshape(-1, 4)

    def set_mesh_draw_info(self, info):
        """info: [n_meshes, 3] = index_count, first_index, base_vertex of every mesh id (the renderer's index ranges)."""
        info = np.asarray(info, np.int64).reshape(-1, 3)
        arr = (MeshDrawInfo * max(len(info), 1))(*[MeshDrawInfo(int(a), int(b), int(c)) for a, b, c in info])
        self._check(self._lib.astre_gpu_set_mesh_draw_info(self._h, len(info), arr))

    def build_indirect(self):
        """One DrawElementsIndirectCommand per draw run, left on the device -> (n, device pointer)."""
        n, p = C.c_uint32(), C.c_void_p()
        self._check(self._lib.astre_gpu_build_indirect(self._h, C.byref(n), C.byref(p)))
        return int(n.value), p.value

    def read_indirect(self):
        """The commands on the host: array [n, 5] = count, instance_count, first_index, base_vertex, base_instance."""
        n = C.c_uint32()
        self._check(self._lib.astre_gpu_read_indirect(self._h, None, 0, C.byref(n)))
        cmds = (DrawIndirect * max(n.value, 1))()
        self._check(self._lib.astre_gpu_read_indirect(self._h, cmds, n.value, C.byref(n)))
        return np.array([[c.count, c.instance_count, c.first_index, c.base_vertex, c.base_instance] for c in cmds[:n.value]],
                        np.int64).reshape(-1, 5)

    def slot_entity(self, slot):
        e = C.c_uint64()
        self._check(self._lib.astre_gpu_slot_entity(self._h, int(slot), C.byref(e)))
        return e.value

    # ---- f3: global key-sorted draw list across GPUs ----
    def publish_keys(self, parity, stream=None):
        """Copy the frame's sorted list into publish buffer `parity` (fixed address, what peers map)."""
        self._check(self._lib.astre_gpu_publish_keys(self._h, int(parity), stream))

    def device_published_keys_ptr(self, parity):
        return self._lib.astre_gpu_device_published_keys(self._h, int(parity))

    def device_sorted_keys_ptr(self):
        return self._lib.astre_gpu_device_sorted_keys(self._h)

    def export_handle(self, parity):
        """CUDA IPC handle (bytes) of publish buffer `parity`."""
        buf = (C.c_ubyte * 64)()
        self._check(self._lib.astre_gpu_export_handle(self._h, int(parity), buf))
        return bytes(buf)

    def import_handle(self, handle):
        """Map a peer's publish buffer; returns the device pointer (int)."""
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        self._check(self._lib.astre_gpu_import_handle(self._h, buf, C.byref(out)))
        return out.value

    def release_import(self, ptr):
        self._check(self._lib.astre_gpu_release_import(self._h, C.c_void_p(ptr)))

    def merge_lists(self, list_ptrs, counts, out_first=0, out_count=None, stream=None):
        """Merge the ranks' sorted lists (device pointers, list r = rank r) into the global draw list,
        keeping positions [out_first, out_first + out_count)."""
        g = len(list_ptrs)
        ptrs = (C.c_void_p * g)(*[int(p) if p else None for p in list_ptrs])
        cnt = np.ascontiguousarray(counts, np.uint64)
        oc = 0xFFFFFFFFFFFFFFFF if out_count is None else int(out_count)
        self._check(self._lib.astre_gpu_merge_lists(self._h, g, ptrs, _ptr(cnt, C.c_uint64), int(out_first), oc, stream))

    def merge_prepare(self, rank, publish_ptrs, counts, stream=None):
        """Peer-to-peer flow, step 1: rank every sample of every published list inside this rank's own list."""
        g = len(publish_ptrs)
        ptrs = (C.c_void_p * g)(*[int(p) if p else None for p in publish_ptrs])
        cnt = np.ascontiguousarray(counts, np.uint64)
        self._check(self._lib.astre_gpu_merge_prepare(self._h, g, int(rank), ptrs, _ptr(cnt, C.c_uint64), stream))

    def merge_published(self, publish_ptrs, counts, out_first=0, out_count=None, stream=None):
        """Peer-to-peer flow, step 2 (after a barrier): tile table from all cut tables + the merge."""
        g = len(publish_ptrs)
        ptrs = (C.c_void_p * g)(*[int(p) if p else None for p in publish_ptrs])
        cnt = np.ascontiguousarray(counts, np.uint64)
        oc = 0xFFFFFFFFFFFFFFFF if out_count is None else int(out_count)
        self._check(self._lib.astre_gpu_merge_published(self._h, g, ptrs, _ptr(cnt, C.c_uint64), int(out_first), oc, stream))

    def merge_prepare_dev(self, rank, publish_ptrs, all_counts_dev_ptr, n_views, stream=None):
        """Step 1 with the counts left on the device ([n_lists][n_views] uint32 table, e.g. the all-gathered tensor)."""
        g = len(publish_ptrs)
        ptrs = (C.c_void_p * g)(*[int(p) if p else None for p in publish_ptrs])
        self._check(self._lib.astre_gpu_merge_prepare_dev(self._h, g, int(rank), ptrs, C.c_void_p(int(all_counts_dev_ptr)),
                                                          int(n_views), stream))

    def merge_published_dev(self, publish_ptrs, slice_rank=0, slice_world=0, out_cap=0, stream=None):
        """Step 2 with device-side sizes: slice `slice_rank` of `slice_world` (0 = the whole list)."""
        g = len(publish_ptrs)
        ptrs = (C.c_void_p * g)(*[int(p) if p else None for p in publish_ptrs])
        self._check(self._lib.astre_gpu_merge_published_dev(self._h, g, ptrs, int(slice_rank), int(slice_world), int(out_cap), stream))

    def merge_p2p(self, rank, publish0, publish1, epoch, slice_rank=0, slice_world=0, out_cap=0, stream=None):
        """The whole exchange over NVLink peer memory (no NCCL, no host round trip): publish, wait for the peers,
        cut table, wait, merge.  publish0 / publish1: all ranks' publish buffers of parity 0 / 1; epoch: frame number from 1."""
        g = len(publish0)
        p0 = (C.c_void_p * g)(*[int(p) if p else None for p in publish0])
        p1 = (C.c_void_p * g)(*[int(p) if p else None for p in publish1])
        self._check(self._lib.astre_gpu_merge_p2p(self._h, g, int(rank), p0, p1, int(epoch), int(slice_rank), int(slice_world),
                                                  int(out_cap), stream))

    def merged_count(self):
        n = C.c_uint64()
        self._check(self._lib.astre_gpu_merged_count(self._h, C.byref(n)))
        return n.value

    def read_merged(self, first, n):
        keys = np.empty(int(n), np.uint64)
        src = np.empty(int(n), np.uint8)
        self._check(self._lib.astre_gpu_read_merged(self._h, int(first), int(n), _ptr(keys, C.c_uint64), _ptr(src, C.c_uint8)))
        return keys, src

    def device_merged_ptrs(self):
        return self._lib.astre_gpu_device_merged_keys(self._h), self._lib.astre_gpu_device_merged_sources(self._h)

    def last_merge_ms(self):
        ms = C.c_float()
        self._check(self._lib.astre_gpu_last_merge_ms(self._h, C.byref(ms)))
        return ms.value

    def device_counts_ptr(self):
        return self._lib.astre_gpu_device_counts(self._h)

    def device_world_ptr(self):
        return self._lib.astre_gpu_device_world(self._h)


# ---- host-side helpers ---------------------------------------------------------

def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def camera_matrices(pos, forward, up, fov_deg, aspect, z_near, z_far):
    """CameraSystem::run's view/proj (camera_system.cpp:24-37) and clip = proj*view."""
    lib = load_library()
    view, proj, clip = (np.empty(16, np.float32) for _ in range(3))
    lib.astre_camera_matrices(_f3(pos), _f3(forward), _f3(up), fov_deg, aspect, z_near, z_far,
                              _ptr(view, C.c_float), _ptr(proj, C.c_float), _ptr(clip, C.c_float))
    return view, proj, clip


def light_space_matrix(pos, direction, light_type, outer_cutoff=0.0):
    lib = load_library()
    clip = np.empty(16, np.float32)
    ok = lib.astre_light_space_matrix(_f3(pos), _f3(direction), int(light_type), float(outer_cutoff), _ptr(clip, C.c_float))
    return clip if ok else None


def ortho_view_matrix(eye, center, up, l, r, b, t, n, f):
    lib = load_library()
    clip = np.empty(16, np.float32)
    lib.astre_ortho_view_matrix(_f3(eye), _f3(center), _f3(up), l, r, b, t, n, f, _ptr(clip, C.c_float))
    return clip


def global_offsets(all_counts, rank, rank_major=False):
    """all_counts [world_size, n_views] -> (offsets[n_views] of this rank, totals[n_views]).
    rank_major: whole-world shards (the global list is the concatenation of the ranks' lists); totals = own segment lengths."""
    lib = load_library()
    a = np.ascontiguousarray(all_counts, dtype=np.uint32)
    ws, nv = a.shape
    off, tot = np.zeros(nv, np.uint64), np.zeros(nv, np.uint64)
    fn = lib.astre_global_offsets_rank_major if rank_major else lib.astre_global_offsets
    fn(_ptr(a, C.c_uint32), ws, int(rank), nv, _ptr(off, C.c_uint64), _ptr(tot, C.c_uint64))
    return off, tot
