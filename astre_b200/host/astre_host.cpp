// astre_host.cpp -- see astre_host.hpp.  Host-side only; every number that reaches
// a draw list or a matrix is computed by libastre_gpu.so.
#include "astre_host.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "../csrc/host_math.hpp"

namespace astre::gpu {

namespace {

constexpr std::uint8_t kDirtyTrs = 1, kDirtyFlags = 2, kFresh = 4;
constexpr std::uint8_t kHasPos = 1, kHasRot = 2, kHasScale = 4;

void check(astre_gpu_ctx *ctx, int rc, const char *what)
{
    if (rc != ASTRE_OK)
        throw std::runtime_error(std::string(what) + ": " + astre_gpu_last_error(ctx));
}

}  // namespace

EntityMirror::EntityMirror(std::uint32_t max_entities)
    : _max(max_entities), _entity_at(max_entities, 0), _pos(3 * (std::size_t)max_entities, 0.0f),
      _rot(4 * (std::size_t)max_entities, 0.0f), _scl(3 * (std::size_t)max_entities, 1.0f),
      _mesh(max_entities, 0), _shader(max_entities, 0), _parent(max_entities, ASTRE_NO_PARENT),
      _flags(max_entities, 0), _state(max_entities, 0), _has(max_entities, 0), _color(max_entities)
{
}

void EntityMirror::markDirty(std::uint32_t slot)
{
    if (_state[slot] & (kFresh | kDirtyTrs | kDirtyFlags)) return;
    _dirty.push_back(slot);
}

std::uint32_t EntityMirror::spawn(Entity e, const TransformComponent &t, const VisualComponent *v,
                                  const IResourceNames *names, Entity parent)
{
    if (e == 0) throw std::invalid_argument("entity 0 is invalid (ecs/entity.hpp:10-13)");
    if (_slot_of.count(e)) throw std::invalid_argument("entity already mirrored");
    std::uint32_t slot;
    if (!_free.empty()) { slot = _free.back(); _free.pop_back(); }
    else {
        if (_high_water == _max) throw std::length_error("EntityMirror is full");
        slot = _high_water++;
    }
    _slot_of[e] = slot;
    _entity_at[slot] = e;
    // transform_system.cpp:24-49: absent sub-messages take position 0, rotation w=1, scale 1
    const Vec3 p = t.position.value_or(Vec3{0, 0, 0}), s = t.scale.value_or(Vec3{1, 1, 1});
    const Quat q = t.rotation.value_or(Quat{0, 0, 0, 1});
    float *P = &_pos[3 * (std::size_t)slot], *R = &_rot[4 * (std::size_t)slot], *S = &_scl[3 * (std::size_t)slot];
    P[0] = p.x; P[1] = p.y; P[2] = p.z;
    R[0] = q.x; R[1] = q.y; R[2] = q.z; R[3] = q.w;
    S[0] = s.x; S[1] = s.y; S[2] = s.z;
    _has[slot] = (std::uint8_t)((t.position ? 1u : 0u) | (t.rotation ? 2u : 0u) | (t.scale ? 4u : 0u));
    _parent[slot] = parent ? _slot_of.at(parent) : ASTRE_NO_PARENT;
    // visual_system.cpp:23-30: a proxy exists only if both names resolve
    std::uint8_t flags = 0;
    _mesh[slot] = 0; _shader[slot] = 0;
    _color[slot] = Vec4{1.0f, 0.0f, 1.0f, 1.0f};                     // visual_system.cpp:44
    if (v && names) {
        const auto vb = names->getVertexBuffer(v->vertex_buffer_name);
        const auto sh = names->getShader(v->shader_name);
        if (vb && sh) {
            _mesh[slot] = (std::uint32_t)*vb;
            _shader[slot] = (std::uint32_t)*sh;
            flags = ASTRE_FLAGS(v->visible, ASTRE_PHASE_OPAQUE | ASTRE_PHASE_SHADOW_CASTER);   // visual_system.cpp:32,55
            if (v->color) _color[slot] = *v->color;
        }
    }
    _flags[slot] = flags;
    if (!(_state[slot] & kFresh)) _fresh.push_back(slot);
    _state[slot] = kFresh;
    return slot;
}

void EntityMirror::despawn(Entity e)
{
    const auto it = _slot_of.find(e);
    if (it == _slot_of.end()) return;
    const std::uint32_t slot = it->second;
    _slot_of.erase(it);
    _entity_at[slot] = 0;
    _flags[slot] = 0;                       // never drawn again; the slot keeps being transformed until reused
    _parent[slot] = ASTRE_NO_PARENT;
    if (!(_state[slot] & kFresh)) _fresh.push_back(slot);
    _state[slot] = kFresh;
    _free.push_back(slot);
}

void EntityMirror::setPosition(Entity e, Vec3 p)
{
    const std::uint32_t s = _slot_of.at(e);
    float *P = &_pos[3 * (std::size_t)s];
    P[0] = p.x; P[1] = p.y; P[2] = p.z;
    _has[s] |= 1u;
    markDirty(s); _state[s] |= kDirtyTrs;
}
void EntityMirror::setRotation(Entity e, Quat q)
{
    const std::uint32_t s = _slot_of.at(e);
    float *R = &_rot[4 * (std::size_t)s];
    R[0] = q.x; R[1] = q.y; R[2] = q.z; R[3] = q.w;
    _has[s] |= 2u;
    markDirty(s); _state[s] |= kDirtyTrs;
}
void EntityMirror::setScale(Entity e, Vec3 v)
{
    const std::uint32_t s = _slot_of.at(e);
    float *S = &_scl[3 * (std::size_t)s];
    S[0] = v.x; S[1] = v.y; S[2] = v.z;
    _has[s] |= 4u;
    markDirty(s); _state[s] |= kDirtyTrs;
}
void EntityMirror::setVisible(Entity e, bool visible)
{
    const std::uint32_t s = _slot_of.at(e);
    _flags[s] = (std::uint8_t)((_flags[s] & ~1u) | (visible ? 1u : 0u));
    markDirty(s); _state[s] |= kDirtyFlags;
}

std::uint32_t EntityMirror::flush(astre_gpu_ctx *ctx)
{
    std::uint32_t touched = 0;
    // spawns / despawns: whole records, one call per run of consecutive slots
    std::sort(_fresh.begin(), _fresh.end());
    for (std::size_t i = 0; i < _fresh.size();) {
        std::size_t j = i + 1;
        while (j < _fresh.size() && _fresh[j] == _fresh[j - 1] + 1) ++j;
        const std::uint32_t first = _fresh[i], count = (std::uint32_t)(j - i);
        check(ctx, astre_gpu_upload_entities(ctx, first, count, &_entity_at[first], &_pos[3 * (std::size_t)first],
                                             &_rot[4 * (std::size_t)first], &_scl[3 * (std::size_t)first],
                                             &_parent[first], &_mesh[first], &_shader[first], &_flags[first]),
              "astre_gpu_upload_entities");
        for (std::size_t k = i; k < j; ++k) _state[_fresh[k]] = 0;
        touched += count;
        i = j;
    }
    _fresh.clear();
    // dirty sets
    std::vector<std::uint32_t> trs_slots, flag_slots;
    for (std::uint32_t s : _dirty) {
        if (_state[s] & kDirtyTrs) trs_slots.push_back(s);
        if (_state[s] & kDirtyFlags) flag_slots.push_back(s);
        _state[s] = 0;
    }
    _dirty.clear();
    if (!trs_slots.empty()) {
        std::vector<float> p(3 * trs_slots.size()), r(4 * trs_slots.size()), sc(3 * trs_slots.size());
        for (std::size_t i = 0; i < trs_slots.size(); ++i) {
            std::memcpy(&p[3 * i], &_pos[3 * (std::size_t)trs_slots[i]], 12);
            std::memcpy(&r[4 * i], &_rot[4 * (std::size_t)trs_slots[i]], 16);
            std::memcpy(&sc[3 * i], &_scl[3 * (std::size_t)trs_slots[i]], 12);
        }
        check(ctx, astre_gpu_update_trs(ctx, (std::uint32_t)trs_slots.size(), trs_slots.data(), p.data(), r.data(), sc.data()),
              "astre_gpu_update_trs");
        touched += (std::uint32_t)trs_slots.size();
    }
    if (!flag_slots.empty()) {
        std::vector<std::uint8_t> f(flag_slots.size());
        for (std::size_t i = 0; i < flag_slots.size(); ++i) f[i] = _flags[flag_slots[i]];
        check(ctx, astre_gpu_update_flags(ctx, (std::uint32_t)flag_slots.size(), flag_slots.data(), f.data()),
              "astre_gpu_update_flags");
        touched += (std::uint32_t)flag_slots.size();
    }
    check(ctx, astre_gpu_set_entity_count(ctx, _high_water), "astre_gpu_set_entity_count");
    return touched;
}

void GpuTransformSystem::run(float)
{
    _mirror.flush(_ctx);
    check(_ctx, astre_gpu_frame(_ctx, ASTRE_FRAME_TRANSFORM | (wantBasis ? ASTRE_FRAME_BASIS : 0u), nullptr), "astre_gpu_frame");
}

void GpuTransformSystem::read(Entity e, TransformComponent &out)
{
    const std::uint32_t slot = _mirror.slotOf(e);
    check(_ctx, astre_gpu_read_world(_ctx, slot, 1, out.transform_matrix.data()), "astre_gpu_read_world");
    if (wantBasis)
        check(_ctx, astre_gpu_read_basis(_ctx, slot, 1, &out.forward.x, &out.up.x, &out.right.x), "astre_gpu_read_basis");
}

void GpuCameraSystem::run(float, Frame &frame)
{
    // camera_system.cpp:15-19
    frame.camera_position = Vec3{0, 0, 0};
    astre_host::identity(frame.view_matrix.data());
    astre_host::identity(frame.proj_matrix.data());
    if (!active_camera_entity || !_mirror.contains(*active_camera_entity)) return;
    TransformComponent t;
    _transforms.read(*active_camera_entity, t);
    const float *p = _mirror.position(_mirror.slotOf(*active_camera_entity));
    frame.camera_position = Vec3{p[0], p[1], p[2]};
    astre_camera_matrices(p, &t.forward.x, &t.up.x, camera.fov, camera.aspect, camera.near_plane, camera.far_plane,
                          frame.view_matrix.data(), frame.proj_matrix.data(), nullptr);
}

std::vector<Mat4> calculateLightSpaceMatrices(const Frame &f)
{
    std::vector<Mat4> out(f.shadow_casters_count);
    for (const auto &[id, light] : f.gpu_lights) {
        if (!light.cast_shadows) continue;                                   // light_system.cpp:113
        if (light.shadow_map_index >= f.shadow_casters_count) continue;      // :117
        Mat4 m{};
        if (!astre_light_space_matrix(&light.position.x, &light.direction.x, (int)light.direction.w, light.outer_cutoff, m.data()))
            continue;                                                         // unknown type: `default: continue`
        out[light.shadow_map_index] = m;
    }
    return out;
}

void GpuFrameInterpolator::snapshot()
{
    check(_ctx, astre_gpu_snapshot_trs(_ctx, nullptr), "astre_gpu_snapshot_trs");
}

static float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }   // glm::mix

Frame GpuFrameInterpolator::interpolateFrame(const Frame &a, const Frame &b, float alpha)
{
    alpha = alpha < 0.0f ? 0.0f : (alpha > 1.0f ? 1.0f : alpha);                  // render.cpp:10
    Frame result = b;                                                              // proxies, lights, lists of b (:22,41,66)
    result.camera_position = Vec3{mixf(a.camera_position.x, b.camera_position.x, alpha),
                                  mixf(a.camera_position.y, b.camera_position.y, alpha),
                                  mixf(a.camera_position.z, b.camera_position.z, alpha)};
    for (int i = 0; i < 16; ++i) {                                                 // math::interpolate, math.hpp:124-135
        result.view_matrix[i] = mixf(a.view_matrix[i], b.view_matrix[i], alpha);
        result.proj_matrix[i] = mixf(a.proj_matrix[i], b.proj_matrix[i], alpha);
    }
    check(_ctx, astre_gpu_interpolate(_ctx, alpha, nullptr), "astre_gpu_interpolate");
    const std::uint32_t n = _mirror.slotCount();
    std::vector<float> world(16 * (std::size_t)n);
    if (n) check(_ctx, astre_gpu_read_world(_ctx, 0, n, world.data()), "astre_gpu_read_world");
    for (auto &[id, proxy] : result.render_proxies) {
        if (a.render_proxies.find(id) == a.render_proxies.end()) continue;        // :28-29
        if (!_mirror.contains((Entity)id)) continue;
        std::memcpy(proxy.uModel.data(), &world[16 * (std::size_t)_mirror.slotOf((Entity)id)], 64);
    }
    for (auto &[id, light] : result.gpu_lights) {                                 // :45-62, the fields this glue carries
        auto it = a.gpu_lights.find(id);
        if (it == a.gpu_lights.end()) continue;
        const GPULight &la = it->second, &lb = b.gpu_lights.at(id);
        light.position = Vec4{mixf(la.position.x, lb.position.x, alpha), mixf(la.position.y, lb.position.y, alpha),
                              mixf(la.position.z, lb.position.z, alpha), mixf(la.position.w, lb.position.w, alpha)};
        light.direction = Vec4{mixf(la.direction.x, lb.direction.x, alpha), mixf(la.direction.y, lb.direction.y, alpha),
                               mixf(la.direction.z, lb.direction.z, alpha), lb.direction.w};
    }
    return result;
}

void GpuVisualSystem::run(float, Frame &frame)
{
    frame.render_proxies.clear();                                            // visual_system.cpp:16
    frame.draw_lists.clear();
    // views: camera first (Opaque), then one per shadow caster (ShadowCaster) -- deferred_shading.cpp:115-121,145-161
    const std::uint32_t n_views = 1u + (std::uint32_t)frame.light_space_matrices.size();
    std::vector<float> clip(16 * (std::size_t)n_views);
    std::vector<std::uint8_t> masks(n_views, ASTRE_PHASE_SHADOW_CASTER);
    masks[0] = ASTRE_PHASE_OPAQUE;
    astre_host::mat_mul(frame.proj_matrix.data(), frame.view_matrix.data(), &clip[0]);
    for (std::uint32_t v = 1; v < n_views; ++v) std::memcpy(&clip[16 * (std::size_t)v], frame.light_space_matrices[v - 1].data(), 64);
    check(_ctx, astre_gpu_set_views(_ctx, n_views, clip.data(), masks.data()), "astre_gpu_set_views");
    // one graph launch; the list's uModel matrices are gathered inside the frame (short lists arrive in host memory)
    check(_ctx, astre_gpu_frame(_ctx, ASTRE_FRAME_ALL | ASTRE_FRAME_GATHER | ASTRE_FRAME_GRAPH, nullptr), "astre_gpu_frame");

    std::vector<std::uint32_t> counts(n_views);
    check(_ctx, astre_gpu_read_counts(_ctx, counts.data()), "astre_gpu_read_counts");
    std::uint64_t total = 0;
    check(_ctx, astre_gpu_read_total(_ctx, &total), "astre_gpu_read_total");
    std::vector<std::uint64_t> keys(total);
    std::vector<float> world(16 * (std::size_t)total);
    std::uint64_t got = 0;
    // sorted keys and, in the same order, uModel of their entities: one call, one copy each
    check(_ctx, astre_gpu_read_draw_list(_ctx, keys.data(), world.data(), total, &got), "astre_gpu_read_draw_list");

    frame.draw_lists.resize(n_views);
    std::size_t k = 0;
    for (std::uint32_t v = 0; v < n_views; ++v) {
        frame.draw_lists[v].reserve(counts[v]);
        for (std::uint32_t i = 0; i < counts[v]; ++i, ++k) {
            const std::uint32_t slot = (std::uint32_t)(keys[k] & ((1ull << ASTRE_KEY_SLOT_BITS) - 1ull));
            const Entity e = _mirror.entityAt(slot);
            frame.draw_lists[v].push_back(e);
            auto [it, fresh] = frame.render_proxies.try_emplace((std::size_t)e);
            if (!fresh) continue;
            RenderProxy &p = it->second;                                      // visual_system.cpp:32-55
            p.visible = _mirror.flags(slot) & 1u;
            p.phases = (std::uint8_t)(_mirror.flags(slot) >> 1);
            p.vertex_buffer = _mirror.mesh(slot);
            p.shader = _mirror.shader(slot);
            p.useTexture = false;
            p.uColor = _mirror.color(slot);
            std::memcpy(p.uModel.data(), &world[16 * k], 64);
            // "used for interpolation" (visual_system.cpp:49-52): raw proto reads, so an absent field is zeros here
            // (not TransformSystem's default) -- kept as the reference has it
            const float *P = _mirror.position(slot), *R = _mirror.rotation(slot), *S = _mirror.scale(slot);
            const std::uint8_t has = _mirror.present(slot);
            p.position = (has & 1u) ? Vec3{P[0], P[1], P[2]} : Vec3{0, 0, 0};
            p.rotation = (has & 2u) ? Quat{R[0], R[1], R[2], R[3]} : Quat{0, 0, 0, 0};
            p.scale = (has & 4u) ? Vec3{S[0], S[1], S[2]} : Vec3{0, 0, 0};
        }
    }

    // f2: runs of one (view, shader, mesh) in the sorted list
    frame.draw_runs.clear();
    std::uint32_t n_runs = 0;
    check(_ctx, astre_gpu_read_runs(_ctx, nullptr, 0, &n_runs), "astre_gpu_read_runs");
    std::vector<astre_draw_run> runs(n_runs);
    if (n_runs) check(_ctx, astre_gpu_read_runs(_ctx, runs.data(), n_runs, &n_runs), "astre_gpu_read_runs");
    std::vector<std::uint64_t> view_first(n_views + 1, 0);
    for (std::uint32_t v = 0; v < n_views; ++v) view_first[v + 1] = view_first[v] + counts[v];
    frame.draw_runs.reserve(n_runs);
    for (const astre_draw_run &r : runs) {
        const std::uint32_t v = r.world_view & 31u;
        frame.draw_runs.push_back(Frame::DrawRun{v, (std::uint32_t)(r.first - view_first[v]), r.count,
                                                 (std::size_t)(r.shader_mesh >> 11), (std::size_t)(r.shader_mesh & 2047u)});
    }
    frame.draw_commands_dev = nullptr;
    frame.draw_command_count = 0;
    if (_indirect) check(_ctx, astre_gpu_build_indirect(_ctx, &frame.draw_command_count, &frame.draw_commands_dev), "astre_gpu_build_indirect");
}

void GpuVisualSystem::setMeshDrawInfo(const std::vector<astre_mesh_draw_info> &info)
{
    check(_ctx, astre_gpu_set_mesh_draw_info(_ctx, (std::uint32_t)info.size(), info.data()), "astre_gpu_set_mesh_draw_info");
    _indirect = !info.empty();
}

}  // namespace astre::gpu
