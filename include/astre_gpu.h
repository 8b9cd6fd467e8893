/*
 * astre_gpu.h -- C ABI of the B200-native transform + visibility path for the
 * Astre engine (libastre_gpu.so, sm_100a).
 *
 * The reference engine has no FFI / plugin seam for this path;
 * each entry point below cites the reference interface it replaces, relative to
 * the reference root.  Path shorthand: ECS/, Render/, Pipeline/, Math/, Loader/
 * are under engine/modules/.
 *
 * Contract
 *   - opaque context, int status returns (ASTRE_OK == 0), no exceptions, no
 *     torch / C++ types in any signature;
 *   - caller-owned host buffers (pinned memory makes the copies asynchronous),
 *     library-owned device buffers;
 *   - all calls on one context come from one thread at a time, which matches
 *     "exactly one TransformSystem::run in flight" (Pipeline/src/logic_pipelines.cpp:30-33);
 *   - there is no CPU fallback: without a CUDA device astre_gpu_create returns NULL
 *     and astre_gpu_last_error(NULL) says why.
 *
 * Layout conventions (same as the reference): mat4 is 16 floats, column-major,
 * m[c*4+r] == glm M[c][r] (Math/src/math.cpp:70-93); quaternions are x,y,z,w
 * (Math/src/math.cpp:124-133); vec3 arrays are packed xyz.
 */
#ifndef ASTRE_GPU_H
#define ASTRE_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ASTRE_API __attribute__((visibility("default")))
#else
#define ASTRE_API
#endif

typedef struct astre_gpu_ctx astre_gpu_ctx;

enum {
    ASTRE_OK            = 0,
    ASTRE_ERR_INVALID   = 1,  /* bad argument (range, alignment, null)            */
    ASTRE_ERR_CUDA      = 2,  /* a CUDA call failed; see astre_gpu_last_error     */
    ASTRE_ERR_CAPACITY  = 3,  /* more visible keys than max_keys; lists truncated */
    ASTRE_ERR_HIERARCHY = 4,  /* parent index out of range or cyclic              */
    ASTRE_ERR_STATE     = 5   /* call made before its prerequisites               */
};

/* RenderPhase bits, Render/include/render/render.hpp:29-36 */
enum {
    ASTRE_PHASE_OPAQUE        = 1u << 0,
    ASTRE_PHASE_TRANSPARENT   = 1u << 1,
    ASTRE_PHASE_SHADOW_CASTER = 1u << 2,
    ASTRE_PHASE_DEBUG         = 1u << 3
};

/* Per-entity flags byte: bit0 = VisualComponent.visible
 * (Proto/ECS/components/visual_component.proto:9, ECS/src/system/visual_system.cpp:32),
 * bits 1..4 = RenderProxy.phases (visual_system.cpp:55).  An entity for which
 * VisualSystem would emit no proxy (no Visual component, unresolved mesh or
 * shader name -- visual_system.cpp:26-30) is uploaded with phases == 0. */
#define ASTRE_FLAG_VISIBLE        0x01u
#define ASTRE_FLAGS(visible, phases) ((uint8_t)(((visible) ? 1u : 0u) | (((phases) & 15u) << 1)))

#define ASTRE_NO_PARENT 0xFFFFFFFFu

/* frame stages */
enum {
    ASTRE_FRAME_TRANSFORM = 1u << 0,  /* TransformSystem::run matrices (always implied)     */
    ASTRE_FRAME_CULL      = 1u << 1,  /* frustum test + compaction + key emission            */
    ASTRE_FRAME_SORT      = 1u << 2,  /* device sort of the emitted keys                     */
    ASTRE_FRAME_BASIS     = 1u << 3,  /* forward/up/right of every entity                    */
    ASTRE_FRAME_SORT_LATER = 1u << 4, /* prepare the sort (bucket counts) but run it in astre_gpu_sort */
    ASTRE_FRAME_GRAPH     = 1u << 5,  /* replay the frame's launches as ONE CUDA graph (captured again when the   */
                                      /* plan, views, hierarchy or stream change); astre_gpu_last_cull_ms is then  */
                                      /* unavailable.  For small frames, where launch latency is the frame.        */
    ASTRE_FRAME_GATHER    = 1u << 6,  /* with CULL | SORT: also gather the world matrices of the draw list's entities in */
                                      /* list order inside the frame (astre_gpu_read_draw_list then launches nothing;  */
                                      /* lists of <= 4096 keys are left in host memory by the frame itself)            */
    ASTRE_FRAME_ALL       = 7u
};

/* Local-space bounds of one mesh: a box of half extents `extent` around
 * `center`, inflated by `radius` (AABB: radius 0; sphere: extent 0).
 * Spec-defined -- the reference keeps no bounds. */
typedef struct {
    float center[3];
    float extent[3];
    float radius;
    uint32_t reserved;
} astre_mesh_bounds;

/* Draw key, ascending = draw order (spec-defined):
 *   world(Wb) | view(5) | shader(8) | mesh(11) | depth(14) | local slot(26-Wb)
 * Wb = ceil(log2(n_worlds)) (0 for a single world, which gives
 * view[63:59] shader[58:51] mesh[50:40] depth[39:26] slot[25:0]). */
#define ASTRE_KEY_SLOT_BITS   26
#define ASTRE_KEY_DEPTH_BITS  14
#define ASTRE_KEY_MESH_BITS   11
#define ASTRE_KEY_SHADER_BITS 8
#define ASTRE_KEY_VIEW_BITS   5
#define ASTRE_MAX_VIEWS       32

/* ---- lifetime ------------------------------------------------------------- */

/* Replaces the Registry + ComponentStorage<T> hash maps as the iteration source
 * (ECS/include/ecs/registry.hpp:81-91, ECS/include/ecs/component_manager.hpp:34-90):
 * a dense slot <-> entity mirror in device structure-of-arrays buffers.
 * max_keys == 0 means max_entities * max_views. */
ASTRE_API astre_gpu_ctx *astre_gpu_create(int device, uint32_t max_entities, uint32_t max_views, uint64_t max_keys);
ASTRE_API void astre_gpu_destroy(astre_gpu_ctx *ctx);
/* Last error text of ctx, or of the last failed astre_gpu_create when ctx is NULL. */
ASTRE_API const char *astre_gpu_last_error(const astre_gpu_ctx *ctx);
ASTRE_API const char *astre_gpu_version(void);

/* ---- mirror maintenance --------------------------------------------------- */

/* Spawn / full refresh of slots [first, first+count): what EntityLoader::load ->
 * Registry::addComponent feeds the component storages (Loader/src/entity_loader.cpp:11-76).
 * Null pos3/rot4/scale3 take TransformSystem's defaults (transform_system.cpp:24-49:
 * position 0, rotation w=1, scale 1).  parent_slot may be NULL (all roots);
 * entity_id may be NULL (id == slot).  mesh_id < 2048, shader_id < 256. */
ASTRE_API int astre_gpu_upload_entities(astre_gpu_ctx *ctx, uint32_t first, uint32_t count,
                                        const uint64_t *entity_id,
                                        const float *pos3, const float *rot_xyzw4, const float *scale3,
                                        const uint32_t *parent_slot,
                                        const uint32_t *mesh_id, const uint32_t *shader_id,
                                        const uint8_t *flags);
/* Number of live slots becomes n (slots >= n are ignored by frames). */
ASTRE_API int astre_gpu_set_entity_count(astre_gpu_ctx *ctx, uint32_t n);

/* Dirty-set update of position/rotation/scale: the writes ScriptSystem / Lua
 * bindings make between ticks (ECS/src/system/script_system.cpp:31-35,
 * Script/include/script/component_binding.hpp:26-115).  Any of the three arrays
 * may be NULL (left unchanged). */
ASTRE_API int astre_gpu_update_trs(astre_gpu_ctx *ctx, uint32_t n, const uint32_t *slots,
                                   const float *pos3, const float *rot_xyzw4, const float *scale3);
/* Same for a contiguous slot range (no index list). */
ASTRE_API int astre_gpu_update_trs_range(astre_gpu_ctx *ctx, uint32_t first, uint32_t count,
                                         const float *pos3, const float *rot_xyzw4, const float *scale3);
/* VisualComponent.visible / phases changes. */
ASTRE_API int astre_gpu_update_flags(astre_gpu_ctx *ctx, uint32_t n, const uint32_t *slots, const uint8_t *flags);

ASTRE_API int astre_gpu_set_mesh_bounds(astre_gpu_ctx *ctx, uint32_t n_meshes, const astre_mesh_bounds *bounds);

/* ---- views ---------------------------------------------------------------- */

/* One clip matrix (proj*view, column-major) and one RenderPhase mask per view:
 * view 0 is normally the camera (CameraSystem::run, ECS/src/system/camera_system.cpp:24-37,
 * mask Opaque -- Pipeline/src/deferred_shading.cpp:117-120), views 1.. the shadow
 * casters (calculateLightSpaceMatrices, ECS/src/system/light_system.cpp:105-160,
 * mask ShadowCaster -- deferred_shading.cpp:153-156). */
ASTRE_API int astre_gpu_set_views(astre_gpu_ctx *ctx, uint32_t n_views,
                                  const float *clip_colmajor16, const uint8_t *phase_mask_per_view);
/* Batched independent worlds: world w owns slots [w*entities_per_world,
 * (w+1)*entities_per_world) and has its own views_per_world clip matrices
 * (clip layout [world][view][16], masks [view]).  entities_per_world % 4 == 0. */
ASTRE_API int astre_gpu_set_world_views(astre_gpu_ctx *ctx, uint32_t n_worlds, uint32_t entities_per_world,
                                        uint32_t views_per_world,
                                        const float *clip_colmajor16, const uint8_t *phase_mask_per_view);

/* ---- the frame ------------------------------------------------------------ */

/* TransformSystem::run + VisualSystem::run + the per-pass filters of
 * deferredShadingStage in one enqueue (transform_system.cpp:11-70,
 * visual_system.cpp:14-61, deferred_shading.cpp:115-121,145-161).
 * stream: a cudaStream_t cast to void* (NULL = the context's own stream).
 * Asynchronous; results are read with the astre_gpu_read_* calls, which
 * synchronise. */
ASTRE_API int astre_gpu_frame(astre_gpu_ctx *ctx, uint32_t stage_flags, void *stream);
/* Sorts the keys of the last frame, which must have run with ASTRE_FRAME_CULL |
 * ASTRE_FRAME_SORT_LATER.  Splitting the frame lets a caller start work that only
 * needs the per-view counts -- the multi-GPU count exchange -- while
 * the sort runs.  Afterwards the frame reads exactly like one run with ASTRE_FRAME_SORT. */
ASTRE_API int astre_gpu_sort(astre_gpu_ctx *ctx, void *stream);
/* Which way the last frame's keys were sorted (decided on the device from the bucket sizes):
 * 0 = no sort ran, 1 = bucket sort, 2 = LSD radix fallback.  Spec-defined stage (SURVEY.md
 * Appendix B): the reference has no draw-order sort (deferred_shading.cpp:115,151 walk a hash map).
 * Synchronises. */
ASTRE_API int astre_gpu_last_sort_path(astre_gpu_ctx *ctx, uint32_t *path);
/* Diagnostics of ASTRE_FRAME_GRAPH: how often the frame was captured (or its executable updated), how often only the
 * view records were patched into it, and the bucket-digit width of the last frame's sort plan. */
ASTRE_API int astre_gpu_graph_stats(const astre_gpu_ctx *ctx, uint32_t *captures, uint32_t *view_patches, uint32_t *bucket_bits);
/* Test hook of the spec-defined key sort (no reference counterpart): sorts n UNIQUE 64-bit keys whose varying bits lie
 * inside live_mask with a bucket digit of bin_bits bits, through exactly the kernels a frame uses (bucket path or,
 * chosen on the device, the LSD fallback); *path_out as astre_gpu_last_sort_path.  Invalidates the last frame's
 * results.  Lets tests drive the sort with key sets no scene produces. */
ASTRE_API int astre_gpu_debug_sort_keys(astre_gpu_ctx *ctx, const uint64_t *keys, uint64_t n, uint64_t live_mask,
                                        uint32_t bin_bits, uint64_t *sorted_out, uint32_t *path_out);
/* Test hook: with ASTRE_GUARD=1 in the environment every device buffer of the library carries a 4 KB canary on both
 * sides.  Checks the canaries of all live buffers (synchronises the device) and returns the number of canary bytes found
 * overwritten so far, including those seen when buffers were freed; -1 when guards are off.  No reference counterpart. */
ASTRE_API int64_t astre_gpu_debug_guard_check(void);
/* Overruns a scratch buffer on purpose (three bytes past its end, one before its start) and returns the damage the
 * detector reports for it (4); does not count towards astre_gpu_debug_guard_check.  -1 when guards are off. */
ASTRE_API int64_t astre_gpu_debug_guard_selftest(void);
ASTRE_API int astre_gpu_sync(astre_gpu_ctx *ctx);
/* Makes every later call on ctx (mirror updates, frames with stream == NULL,
 * read-backs) enqueue on the caller's cudaStream_t instead of the context's own
 * stream; NULL restores the internal stream.  Synchronises the old stream. */
ASTRE_API int astre_gpu_set_stream(astre_gpu_ctx *ctx, void *stream);
/* Number of kernel launches the library has enqueued on this context so far. */
ASTRE_API uint64_t astre_gpu_launch_count(const astre_gpu_ctx *ctx);
/* Duration in ms of the most recent astre_gpu_frame, measured with CUDA events
 * on the launching stream (synchronises). */
ASTRE_API int astre_gpu_last_frame_ms(astre_gpu_ctx *ctx, float *ms);
/* Same, for the transform+cull kernels only (the dominant kernel). */
ASTRE_API int astre_gpu_last_cull_ms(astre_gpu_ctx *ctx, float *ms);

/* f1: interpolateFrame's per-proxy step (Render/src/render.cpp:26-39): the TRS of
 * the previous tick is kept on the device by astre_gpu_snapshot_trs; this writes
 * translate(mix(p)) * toMat4(slerp(q)) * scale(mix(s)) into the world buffer for every
 * slot that existed at the snapshot (later slots get their plain transform, like a proxy
 * that is missing from frame `a`, render.cpp:28-29).  alpha is clamped to [0,1] (render.cpp:10).
 * slerp uses acosf/sinf: results agree with the CPU to a few ULP, not bit for bit. */
ASTRE_API int astre_gpu_snapshot_trs(astre_gpu_ctx *ctx, void *stream);
ASTRE_API int astre_gpu_interpolate(astre_gpu_ctx *ctx, float alpha, void *stream);

/* ---- results -------------------------------------------------------------- */

/* transform_matrix of slots [first, first+count) (TransformComponent field 7). */
ASTRE_API int astre_gpu_read_world(astre_gpu_ctx *ctx, uint32_t first, uint32_t count, float *mat4_colmajor);
/* forward / up / right of slots [first, first+count) (fields 4,6,5); needs ASTRE_FRAME_BASIS. */
ASTRE_API int astre_gpu_read_basis(astre_gpu_ctx *ctx, uint32_t first, uint32_t count,
                                   float *forward3, float *up3, float *right3);
/* counts[world*views_per_world + view].  Returns ASTRE_ERR_CAPACITY (after filling
 * counts) if the frame produced more keys than max_keys. */
ASTRE_API int astre_gpu_read_counts(astre_gpu_ctx *ctx, uint32_t *counts);
/* Total number of keys of the last frame. */
ASTRE_API int astre_gpu_read_total(astre_gpu_ctx *ctx, uint64_t *total);
/* The sorted draw list of (world, view): up to cap keys, *n_out = its length. */
ASTRE_API int astre_gpu_read_keys(astre_gpu_ctx *ctx, uint32_t world, uint32_t view,
                                  uint64_t *keys, uint64_t cap, uint64_t *n_out);
/* The whole key buffer (all worlds and views; sorted iff ASTRE_FRAME_SORT ran). */
ASTRE_API int astre_gpu_read_all_keys(astre_gpu_ctx *ctx, uint64_t *keys, uint64_t cap, uint64_t *n_out);
/* World matrices of the first n keys of the sorted list, in list order (what
 * VisualSystem copies into RenderProxy.inputs["uModel"], visual_system.cpp:47). */
ASTRE_API int astre_gpu_read_visible_world(astre_gpu_ctx *ctx, uint64_t first_key, uint64_t n, float *mat4_colmajor);
/* The whole draw list in one call -- sorted keys and, in the same order, the world matrices of their entities
 * (what VisualSystem::run leaves in frame.render_proxies[e].inputs.in_mat4["uModel"], visual_system.cpp:47):
 * one synchronisation for the frame and one for both copies.  keys / mat4_colmajor may be NULL; at most cap
 * entries are written, *n_out is the list's length. */
ASTRE_API int astre_gpu_read_draw_list(astre_gpu_ctx *ctx, uint64_t *keys, float *mat4_colmajor, uint64_t cap, uint64_t *n_out);
/* f2: the sorted list of the last frame cut into draw runs -- maximal stretches of keys with the
 * same (world, view, shader, mesh), i.e. what one instanced / indirect draw covers instead of one
 * IRenderer::render per proxy (Pipeline/src/deferred_shading.cpp:115-132, Render/src/opengl/opengl.cpp:563-670).
 * `first` indexes the whole sorted key list (astre_gpu_read_all_keys); runs come back in list order. */
typedef struct {
    uint32_t first;
    uint32_t count;
    uint32_t world_view;    /* world << 5 | view   */
    uint32_t shader_mesh;   /* shader << 11 | mesh */
} astre_draw_run;
ASTRE_API int astre_gpu_read_runs(astre_gpu_ctx *ctx, astre_draw_run *runs, uint32_t cap, uint32_t *n_out);
/* f2, one step further: the runs as indirect draw commands that stay on the GPU.  One record per run, in list
 * order, in the layout of OpenGL's DrawElementsIndirectCommand (glMultiDrawElementsIndirect; the reference issues one
 * glDrawElements per proxy, Render/src/opengl/opengl.cpp:563-670): count / first_index / base_vertex from the mesh's
 * entry of astre_gpu_set_mesh_draw_info (the renderer's index ranges per vertex buffer, Render/src/vertex.cpp), one
 * instance per key of the run, base_instance = the run's first position in the sorted list = the row of its first
 * entity in the gathered matrices (ASTRE_FRAME_GATHER / astre_gpu_read_draw_list), so a vertex shader reads
 * uModel[gl_BaseInstance + gl_InstanceID].  A mesh id without an entry gives an empty command (count 0).  The command
 * of run i belongs to astre_gpu_read_runs' run i (world, view, shader).  *commands_dev: library-owned device buffer
 * (a GL buffer can alias it through CUDA-GL interop), valid until the next frame; stream-ordered after the frame.
 * Spec-defined. */
typedef struct { uint32_t index_count, first_index; int32_t base_vertex; } astre_mesh_draw_info;
typedef struct { uint32_t count, instance_count, first_index; int32_t base_vertex; uint32_t base_instance; } astre_draw_indirect;
ASTRE_API int astre_gpu_set_mesh_draw_info(astre_gpu_ctx *ctx, uint32_t n_meshes, const astre_mesh_draw_info *info);
ASTRE_API int astre_gpu_build_indirect(astre_gpu_ctx *ctx, uint32_t *n_out, const void **commands_dev);
ASTRE_API int astre_gpu_read_indirect(astre_gpu_ctx *ctx, astre_draw_indirect *commands, uint32_t cap, uint32_t *n_out);

/* Entity id of a slot (the id given at upload). */
ASTRE_API int astre_gpu_slot_entity(const astre_gpu_ctx *ctx, uint32_t slot, uint64_t *entity_id);

/* Device pointers for zero-copy consumers (CUDA-GL interop, NCCL exchange of
 * counts).  Valid until astre_gpu_destroy. */
ASTRE_API void *astre_gpu_device_world(astre_gpu_ctx *ctx);        /* float[max_entities][16] */
ASTRE_API void *astre_gpu_device_counts(astre_gpu_ctx *ctx);       /* uint32[n_worlds*views]   */
ASTRE_API void *astre_gpu_device_sorted_keys(astre_gpu_ctx *ctx);  /* uint64[]; synchronises  */

/* ---- f3: one globally key-sorted draw list across GPUs --
 *
 * Spec-defined; the reference has no multi-GPU path.  Rank r owns a contiguous
 * range of the scene's entities, so the global list is the ranks' sorted lists
 * merged under the order (key >> slot_bits, rank, key & slot_mask): class and
 * depth first, then global entity order.  Batched worlds shard by whole worlds;
 * their global list is the rank-major concatenation already described by
 * astre_global_offsets and needs no merge.
 *
 * Two ways to feed the merge:
 *  (a) astre_gpu_merge_lists on device pointers this GPU can read at full speed
 *      (its own list, rows of an all-gathered buffer).
 *  (b) peer to peer, nothing gathered: every rank publishes its list at a fixed
 *      device address (astre_gpu_publish_keys: keys, then every 128th key as a
 *      sample), the addresses are exchanged once as CUDA IPC handles, and per
 *      frame
 *        frame(ALL) -> publish_keys -> all-gather of counts (8e; also the barrier
 *        that completes the peers' publish) -> astre_gpu_merge_prepare (each
 *        rank ranks all samples inside ITS OWN list: no search crosses NVLink)
 *        -> any collective as a barrier -> astre_gpu_merge_published (tile table
 *        from the peers' cut tables, then the merge kernel pulls the peers' keys
 *        over NVLink itself; with out_first/out_count only this rank's slice).
 *      Every rank must have been created with the same max_keys. */
#define ASTRE_MAX_LISTS 8
#define ASTRE_IPC_HANDLE_BYTES 64

/* Copies the last frame's sorted list (and its samples) into publish buffer
 * `parity` (0/1, fixed device addresses, capacity max_keys); asynchronous, no
 * host round trip.
 * Alternate the parity from frame to frame: a peer may still be reading the
 * previous frame's buffer. */
ASTRE_API int astre_gpu_publish_keys(astre_gpu_ctx *ctx, uint32_t parity, void *stream);
ASTRE_API void *astre_gpu_device_published_keys(astre_gpu_ctx *ctx, uint32_t parity);
/* CUDA IPC handle (ASTRE_IPC_HANDLE_BYTES bytes) of publish buffer `parity`,
 * for another process on the same node. */
ASTRE_API int astre_gpu_export_handle(astre_gpu_ctx *ctx, uint32_t parity, void *handle_out);
/* Maps a peer's publish buffer into this process; *dev_ptr stays valid until
 * astre_gpu_release_import or astre_gpu_destroy. */
ASTRE_API int astre_gpu_import_handle(astre_gpu_ctx *ctx, const void *handle, void **dev_ptr);
ASTRE_API int astre_gpu_release_import(astre_gpu_ctx *ctx, void *dev_ptr);

/* Merges n_lists (<= ASTRE_MAX_LISTS) sorted key lists, list r being rank r's
 * (device pointers: local, peer-mapped, or rows of a gathered buffer; counts on
 * the host), and keeps positions [out_first, out_first+out_count) of the
 * global list (out_count = UINT64_MAX: everything from out_first on).
 * Asynchronous on `stream`. */
ASTRE_API int astre_gpu_merge_lists(astre_gpu_ctx *ctx, uint32_t n_lists, const void *const *lists,
                                    const uint64_t *counts, uint64_t out_first, uint64_t out_count, void *stream);
/* Peer-to-peer flow, see (b) above.  publish[r] = base of rank r's publish
 * buffer of this frame's parity (own: astre_gpu_device_published_keys, peers:
 * astre_gpu_import_handle); counts[r] = rank r's key total. */
ASTRE_API int astre_gpu_merge_prepare(astre_gpu_ctx *ctx, uint32_t n_lists, uint32_t rank, const void *const *publish,
                                      const uint64_t *counts, void *stream);
ASTRE_API int astre_gpu_merge_published(astre_gpu_ctx *ctx, uint32_t n_lists, const void *const *publish,
                                        const uint64_t *counts, uint64_t out_first, uint64_t out_count, void *stream);
/* The same two steps with the key counts left on the device: all_counts_dev is
 * the all-gathered [n_lists][n_views] table of per-view counts (uint32, device
 * memory) -- what the 8e exchange produces.  Nothing waits for the host, so the
 * frame -> publish -> exchange -> merge chain can be enqueued ahead.  The merge
 * keeps slice slice_rank of slice_world equal slices of the list (slice_world 0:
 * the whole list), at most out_cap keys (0: worst case from max_keys);
 * n_lists * max_keys must stay below 2^31. */
ASTRE_API int astre_gpu_merge_prepare_dev(astre_gpu_ctx *ctx, uint32_t n_lists, uint32_t rank, const void *const *publish,
                                          const uint32_t *all_counts_dev, uint32_t n_views, void *stream);
ASTRE_API int astre_gpu_merge_published_dev(astre_gpu_ctx *ctx, uint32_t n_lists, const void *const *publish,
                                            uint32_t slice_rank, uint32_t slice_world, uint64_t out_cap, void *stream);
/* (b) without NCCL and without a host round trip: counts and "ready" flags
 * travel in the publish buffers' headers and are polled over NVLink by one-warp
 * kernels.  One call per frame and rank does publish -> wait for the peers ->
 * cut table -> wait -> merge.  publish0 / publish1: the publish buffers of all
 * ranks for parity 0 / 1 (own and imported); epoch: the frame number, the same
 * on every rank, from 1, +1 per call.  A peer that never signals makes
 * astre_gpu_merged_count / _read_merged fail after ~2 s instead of hanging. */
ASTRE_API int astre_gpu_merge_p2p(astre_gpu_ctx *ctx, uint32_t n_lists, uint32_t rank, const void *const *publish0,
                                  const void *const *publish1, uint32_t epoch, uint32_t slice_rank, uint32_t slice_world,
                                  uint64_t out_cap, void *stream);
/* Keys kept by the last merge (synchronises if the sizes were computed on the device). */
ASTRE_API int astre_gpu_merged_count(astre_gpu_ctx *ctx, uint64_t *n_out);
/* Result of the last merge: keys[i], sources[i] (= rank) for global position
 * out_first + i.  read_merged synchronises; `first` is relative to out_first. */
ASTRE_API void *astre_gpu_device_merged_keys(astre_gpu_ctx *ctx);     /* uint64[] */
ASTRE_API void *astre_gpu_device_merged_sources(astre_gpu_ctx *ctx);  /* uint8[]  */
ASTRE_API int astre_gpu_read_merged(astre_gpu_ctx *ctx, uint64_t first, uint64_t n, uint64_t *keys, uint8_t *sources);
ASTRE_API int astre_gpu_last_merge_ms(astre_gpu_ctx *ctx, float *ms);

/* ---- host-side helpers (no device work) ------------------------------------ */

/* CameraSystem::run's matrices (camera_system.cpp:24-37): view = lookAt(pos, pos+forward, up),
 * proj = perspective(radians(fov_deg), aspect, near, far); clip = proj*view. */
ASTRE_API void astre_camera_matrices(const float *pos3, const float *forward3, const float *up3,
                                     float fov_deg, float aspect, float z_near, float z_far,
                                     float *view16, float *proj16, float *clip16);
/* calculateLightSpaceMatrices for one light (light_system.cpp:105-160);
 * type 1 directional, 2 point, 3 spot.  Returns 0 for an unknown type. */
ASTRE_API int astre_light_space_matrix(const float *pos3, const float *dir3, int type, float outer_cutoff, float *clip16);
/* lookAt + ortho clip matrix, for cascade-style shadow frusta. */
ASTRE_API void astre_ortho_view_matrix(const float *eye3, const float *center3, const float *up3,
                                       float l, float r, float b, float t, float n, float f, float *clip16);
/* Exclusive scan over (view, rank) of all ranks' per-view counts: offset of
 * this rank's segment of each view in the global draw list.
 * all_counts is [world_size][n_views]; offsets/view_totals are [n_views]. */
ASTRE_API void astre_global_offsets(const uint32_t *all_counts, uint32_t world_size, uint32_t rank,
                                    uint32_t n_views, uint64_t *offsets, uint64_t *view_totals);
/* Batched worlds sharded by WHOLE WORLDS: unit u of rank r is (world, view) number u of that rank's worlds, and
 * the global list is the rank-major concatenation of the ranks' lists (= world-major):
 * offsets[u] = keys of all earlier ranks + keys of this rank's earlier units; segment_len[u] = all_counts[rank][u]. */
ASTRE_API void astre_global_offsets_rank_major(const uint32_t *all_counts, uint32_t world_size, uint32_t rank,
                                               uint32_t n_units, uint64_t *offsets, uint64_t *segment_len);
/* The same on the device and stream-ordered, for an all-gathered count table that never leaves the GPU
 * (all pointers are device pointers; offsets_dev / view_totals_dev may be NULL): the per-frame exchange of
 * SURVEY.md 8e then needs no host round trip.  rank_major != 0: the layout of astre_global_offsets_rank_major.  stream: cudaStream_t cast to void* (NULL = the context's). */
/* ---- count boards: the exchange of SURVEY.md 8e as ONE kernel over peer memory, no collective library ----
 * Every rank owns a board in its own HBM that its peers map (CUDA IPC, or the plain device pointer inside one
 * process).  astre_gpu_count_board_exchange enqueues a one-block kernel that stores this rank's counts (counts_dev:
 * device memory, e.g. a slot of astre_gpu_set_counts_mirror) into the board of every rank over NVLink -- each count
 * as one 8-byte word tagged with the exchange number, so no flag, fence or ordering is needed --, waits (bounded,
 * ~2 s) until the counts of all ranks have landed in its own board, writes the table [world_size][n_counts]
 * (table_dev, nullable), this rank's segment offsets and the per-view totals (as astre_gpu_global_offsets_dev) and
 * acknowledges.  All ranks call it the same number of times with the same n_counts; a rank may be up to 4 exchanges
 * ahead of the slowest.  The kernel waits for peers: enqueue it on a stream other than the one the frames run on.
 * Spec-defined (the reference is single-process).
 *
 * create: allocates the board for up to max_counts counts per rank; ipc_handle_out (ASTRE_IPC_HANDLE_BYTES,
 *   nullable) receives the handle peers open with astre_gpu_import_handle, *board_dev_out (nullable) the local pointer.
 * attach: boards[t] = device pointer of rank t's board (boards[rank] is ignored).  Every rank must have created
 *   its board before any rank exchanges.
 * status: exchanges this rank has finished and the sticky error (0, or 1 + the rank that did not answer in time);
 *   synchronises the device. */
ASTRE_API int astre_gpu_count_board_create(astre_gpu_ctx *ctx, uint32_t max_counts, void *ipc_handle_out, void **board_dev_out);
ASTRE_API int astre_gpu_count_board_attach(astre_gpu_ctx *ctx, uint32_t world_size, uint32_t rank, void *const *boards);
ASTRE_API int astre_gpu_count_board_exchange(astre_gpu_ctx *ctx, const uint32_t *counts_dev, uint32_t n_counts, uint32_t rank_major,
                                             uint32_t *table_dev, uint64_t *offsets_dev, uint64_t *view_totals_dev, void *stream);
ASTRE_API int astre_gpu_count_board_status(astre_gpu_ctx *ctx, uint32_t *exchanges, uint32_t *error);
/* Frame k (astre_gpu_frame_index, counted from 1 and wrapping at 2^27) also leaves its per-(world, view) counts in
 * dev_slot[k & 1] (device memory of the caller, uint32[n_worlds * views] each): the exchange of SURVEY.md 8e can
 * then read them on another stream while frame k+1 -- which zeroes the live counters -- is already running.
 * Two NULLs switch the mirror off.  Synchronises. */
ASTRE_API int astre_gpu_set_counts_mirror(astre_gpu_ctx *ctx, void *dev_slot0, void *dev_slot1);
ASTRE_API int astre_gpu_frame_index(const astre_gpu_ctx *ctx, uint64_t *index);
ASTRE_API int astre_gpu_global_offsets_dev(astre_gpu_ctx *ctx, const uint32_t *all_counts_dev, uint32_t world_size,
                                           uint32_t rank, uint32_t n_views, uint32_t rank_major, uint64_t *offsets_dev,
                                           uint64_t *view_totals_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ASTRE_GPU_H */
