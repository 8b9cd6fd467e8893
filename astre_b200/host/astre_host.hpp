// astre_host.hpp -- engine-side glue above the C ABI (include/astre_gpu.h).
//
// Mirrors, for the transform + visibility path only, the interface the reference
// engine's own systems present, so that the engine's call sites keep their shape:
//
//   reference                                                     here
//   ------------------------------------------------------------  ------------------------------
//   ecs::Registry iteration + ComponentStorage<T>                  astre::gpu::EntityMirror
//     (ECS/include/ecs/registry.hpp:81-91,                           dense slot <-> entity table,
//      ECS/include/ecs/component_manager.hpp:34-90)                  free list, dirty set
//   ecs::system::TransformSystem::run(float dt)                    GpuTransformSystem::run(float dt)
//     (ECS/src/system/transform_system.cpp:11-70)
//   ecs::system::VisualSystem::run(float dt, render::Frame&)        GpuVisualSystem::run(float dt, Frame&)
//     (ECS/src/system/visual_system.cpp:14-61)
//   ecs::system::CameraSystem::run(float dt, render::Frame&)        GpuCameraSystem::run(float dt, Frame&)
//     (ECS/src/system/camera_system.cpp:13-40)
//   ecs::system::calculateLightSpaceMatrices(const Frame&)          calculateLightSpaceMatrices(const Frame&)
//     (ECS/src/system/light_system.cpp:105-160)
//   render::RenderProxy / render::Frame                             astre::gpu::RenderProxy / Frame
//     (Render/include/render/render.hpp:45-86)
//   render::interpolateFrame(const Frame&, const Frame&, float)     GpuFrameInterpolator::interpolateFrame
//     (Render/src/render.cpp:8-68)
//   EntityLoader::load / Registry::destroyEntity                    EntityMirror::spawn / despawn
//     (Loader/src/entity_loader.cpp:11-76, ECS/src/registry.cpp:52-57)
//   LuaBinding<TransformComponent> setters                          EntityMirror::setPosition/Rotation/Scale
//     (Script/include/script/component_binding.hpp:26-115)
//
// The reference's components are protobuf messages (engine/modules/Proto/ECS/components/*.proto);
// protobuf is not available in this build, so plain structs with std::optional
// stand in for "has_x()" presence.  Paths above are relative to engine/modules/.
//
// Error behaviour follows the reference: run() returns nothing; a failure of
// the device library throws std::runtime_error, which propagates out of the
// caller's parallel group exactly like an exception thrown inside a reference
// system (Pipeline/src/logic_pipelines.cpp:38-39).
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/astre_gpu.h"

namespace astre::gpu {

using Entity = std::uint64_t;                  // ecs/entity.hpp:10-13, 0 is invalid
struct Vec3 { float x = 0, y = 0, z = 0; };
struct Vec4 { float x = 0, y = 0, z = 0, w = 0; };
struct Quat { float x = 0, y = 0, z = 0, w = 1; };
using Mat4 = std::array<float, 16>;            // column-major, m[c*4+r] (Math/src/math.cpp:70-93)

// proto::ecs::TransformComponent (transform_component.proto:7-18)
struct TransformComponent {
    std::optional<Vec3> position;
    std::optional<Quat> rotation;
    std::optional<Vec3> scale;
    Vec3 forward, right, up;
    Mat4 transform_matrix{};
};
// proto::ecs::VisualComponent (visual_component.proto:6-12)
struct VisualComponent {
    bool visible = true;
    std::string vertex_buffer_name, shader_name;
    std::optional<Vec4> color;
};
// proto::ecs::CameraComponent
struct CameraComponent { float fov = 60.0f, near_plane = 0.1f, far_plane = 1000.0f, aspect = 1.7f; };

// the two name lookups VisualSystem makes on the renderer (render.hpp:212,243)
struct IResourceNames {
    virtual ~IResourceNames() = default;
    virtual std::optional<std::size_t> getVertexBuffer(const std::string &name) const = 0;
    virtual std::optional<std::size_t> getShader(const std::string &name) const = 0;
};

// render::RenderProxy (render.hpp:45-61); inputs reduced to the three VisualSystem writes
struct RenderProxy {
    bool visible = false;
    std::uint8_t phases = 0;
    std::size_t vertex_buffer = 0, shader = 0;
    Vec3 position, scale;
    Quat rotation;
    Mat4 uModel{};
    Vec4 uColor{1.0f, 0.0f, 1.0f, 1.0f};
    bool useTexture = false;
};

struct GPULight {     // render.hpp:63-72, the fields calculateLightSpaceMatrices reads
    Vec4 position, direction;   // direction.w = LightType
    float outer_cutoff = 0.0f;
    bool cast_shadows = false;
    std::uint32_t shadow_map_index = 0;
};

// render::Frame (render.hpp:74-86) + the sorted draw lists the GPU path adds
struct Frame {
    Vec3 camera_position;
    Mat4 view_matrix{}, proj_matrix{};
    std::unordered_map<std::size_t, RenderProxy> render_proxies;
    std::unordered_map<std::size_t, GPULight> gpu_lights;
    std::vector<Mat4> light_space_matrices;
    std::uint32_t shadow_casters_count = 0;
    // view 0 = camera (Opaque), views 1.. = shadow casters (ShadowCaster): entity ids in draw order
    std::vector<std::vector<Entity>> draw_lists;
    // f2: the lists cut into runs of one (view, shader, mesh): one instanced draw each instead of one
    // glDrawElements per proxy and pass (Pipeline/src/deferred_shading.cpp:115-174).  `first` indexes draw_lists[view].
    struct DrawRun { std::uint32_t view, first, count; std::size_t shader, vertex_buffer; };
    std::vector<DrawRun> draw_runs;
    // ... and the same runs as GL DrawElementsIndirectCommand records in device memory (draw_commands_dev[i] belongs to
    // draw_runs[i]; nullptr unless GpuVisualSystem::setMeshDrawInfo was called): what glMultiDrawElementsIndirect
    // consumes through CUDA-GL interop instead of one glDrawElements per proxy (Render/src/opengl/opengl.cpp:563-670)
    const void *draw_commands_dev = nullptr;
    std::uint32_t draw_command_count = 0;
};

// Dense slot <-> entity table with spawn/despawn and a dirty set: what keeps
// host->device traffic proportional to changes.
class EntityMirror {
public:
    explicit EntityMirror(std::uint32_t max_entities);

    // EntityLoader::load: returns the slot.  parent == 0 means no parent.
    std::uint32_t spawn(Entity e, const TransformComponent &t, const VisualComponent *v, const IResourceNames *names,
                        Entity parent = 0);
    void despawn(Entity e);                               // Registry::destroyEntity
    bool contains(Entity e) const { return _slot_of.count(e) != 0; }
    std::uint32_t slotOf(Entity e) const { return _slot_of.at(e); }
    Entity entityAt(std::uint32_t slot) const { return _entity_at[slot]; }
    std::uint32_t slotCount() const { return _high_water; }
    std::uint32_t liveCount() const { return (std::uint32_t)_slot_of.size(); }

    // script-side writes (mark the slot dirty)
    void setPosition(Entity e, Vec3 p);
    void setRotation(Entity e, Quat q);
    void setScale(Entity e, Vec3 s);
    void setVisible(Entity e, bool visible);

    // Pushes spawns, despawns and dirty TRS / flags to the device; returns the number of slots touched.
    std::uint32_t flush(astre_gpu_ctx *ctx);
    std::uint32_t pendingCount() const { return (std::uint32_t)(_dirty.size() + _fresh.size()); }

    const float *position(std::uint32_t slot) const { return &_pos[3 * (std::size_t)slot]; }
    const float *rotation(std::uint32_t slot) const { return &_rot[4 * (std::size_t)slot]; }
    const float *scale(std::uint32_t slot) const { return &_scl[3 * (std::size_t)slot]; }
    std::uint8_t flags(std::uint32_t slot) const { return _flags[slot]; }
    std::uint32_t mesh(std::uint32_t slot) const { return _mesh[slot]; }
    std::uint32_t shader(std::uint32_t slot) const { return _shader[slot]; }
    const Vec4 &color(std::uint32_t slot) const { return _color[slot]; }
    // which of position / rotation / scale the component carries (proto has_*): bit 0 / 1 / 2.  TransformSystem
    // substitutes defaults for absent ones (transform_system.cpp:24-49); VisualSystem copies the raw -- zero --
    // proto values into the proxy (visual_system.cpp:50-52).  Both behaviours are kept.
    std::uint8_t present(std::uint32_t slot) const { return _has[slot]; }

private:
    void markDirty(std::uint32_t slot);
    std::uint32_t _max, _high_water = 0;
    std::unordered_map<Entity, std::uint32_t> _slot_of;
    std::vector<Entity> _entity_at;
    std::vector<std::uint32_t> _free;
    std::vector<float> _pos, _rot, _scl;
    std::vector<std::uint32_t> _mesh, _shader, _parent;
    std::vector<std::uint8_t> _flags, _state, _has;   // _state: bit0 dirty TRS, bit1 dirty flags, bit2 fresh
    std::vector<Vec4> _color;
    std::vector<std::uint32_t> _dirty, _fresh;
};

class GpuTransformSystem {
public:
    GpuTransformSystem(EntityMirror &mirror, astre_gpu_ctx *ctx) : _mirror(mirror), _ctx(ctx) {}
    // Same contract as TransformSystem::run: afterwards every entity's transform_matrix
    // (and forward/up/right) is current -- on the device; read() copies one back.
    void run(float dt);
    // Fills transform_matrix / forward / up / right of one component, like the in-place write at
    // transform_system.cpp:51-65.
    void read(Entity e, TransformComponent &out);
    bool wantBasis = true;
private:
    EntityMirror &_mirror;
    astre_gpu_ctx *_ctx;
};

class GpuCameraSystem {
public:
    GpuCameraSystem(EntityMirror &mirror, GpuTransformSystem &transforms) : _mirror(mirror), _transforms(transforms) {}
    std::optional<Entity> active_camera_entity;
    CameraComponent camera;
    void run(float dt, Frame &frame);           // camera_system.cpp:13-40
private:
    EntityMirror &_mirror;
    GpuTransformSystem &_transforms;
};

// light_system.cpp:105-160
std::vector<Mat4> calculateLightSpaceMatrices(const Frame &interpolated_frame);

// render::interpolateFrame (Render/src/render.cpp:8-68).  The per-proxy part (:22-39) -- uModel =
// translate(mix(p)) * toMat4(slerp(q)) * scale(mix(s)) for every proxy present in both frames -- runs on the
// device for all entities at once: snapshot() at the end of logic tick `a` keeps that tick's TRS,
// interpolateFrame() evaluates the mix between it and the current TRS (tick `b`).  Camera position and the
// view / projection matrices are mixed on the host like the reference does (:16-20).
class GpuFrameInterpolator {
public:
    GpuFrameInterpolator(EntityMirror &mirror, astre_gpu_ctx *ctx) : _mirror(mirror), _ctx(ctx) {}
    void snapshot();                                                     // astre_gpu_snapshot_trs
    Frame interpolateFrame(const Frame &a, const Frame &b, float alpha);
private:
    EntityMirror &_mirror;
    astre_gpu_ctx *_ctx;
};

class GpuVisualSystem {
public:
    GpuVisualSystem(EntityMirror &mirror, astre_gpu_ctx *ctx) : _mirror(mirror), _ctx(ctx) {}
    // Uploads the frame's views (camera + light-space matrices), runs cull + sort and
    // rebuilds frame.render_proxies -- for the entities that survive any view only --
    // plus frame.draw_lists.
    void run(float dt, Frame &frame);
    // Index range of every vertex buffer id (IRenderer's buffers: index count, first index, base vertex): switches on
    // frame.draw_commands_dev.
    void setMeshDrawInfo(const std::vector<astre_mesh_draw_info> &info);
private:
    EntityMirror &_mirror;
    astre_gpu_ctx *_ctx;
    bool _indirect = false;
};

}  // namespace astre::gpu
