// ref_store.cpp -- CPU baseline B0: the oracle arithmetic run over the
// reference's *storage shape* (SURVEY.md 8d).  TEST / BENCH INFRASTRUCTURE ONLY.
//
// The reference keeps components as protobuf messages in absl::flat_hash_map
// keyed by entity (engine/modules/ECS/include/ecs/component_manager.hpp:34-90),
// iterates the entity->mask map (engine/modules/ECS/include/ecs/registry.hpp:81-91)
// and runs each system as one coroutine, i.e. on one thread
// (engine/modules/Pipeline/src/logic_pipelines.cpp:31,49).  protobuf and abseil
// are not in this image, so std::unordered_map and plain structs stand in: the
// hash-map iteration, the two lookups per component, the heap-backed 16-float
// matrix field and the string-keyed shader inputs are reproduced; protobuf's
// own accessor overhead is not.
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <bitset>
#include <typeindex>
#include <memory>

extern "C" {
void orc_trs_matrix(const float *pos, const float *rot_xyzw, const float *scl, float *out16);
void orc_basis(const float *rot_xyzw, float *fwd, float *up, float *right);
}

namespace {

struct Vec3S { float x = 0, y = 0, z = 0; };
struct QuatS { float x = 0, y = 0, z = 0, w = 0; };

// stand-in for proto::ecs::TransformComponent (transform_component.proto:7-18)
struct TransformComponent {
    std::unique_ptr<Vec3S> position, scale, forward, right, up;   // optional sub-messages
    std::unique_ptr<QuatS> rotation;
    std::vector<float> transform_matrix;                          // repeated float data
};

// stand-in for proto::ecs::VisualComponent (visual_component.proto:6-12)
struct VisualComponent {
    bool visible = true;
    std::string vertex_buffer_name, shader_name;
};

struct ShaderInputs {   // render/shader.hpp:20-41, the three maps VisualSystem touches
    std::unordered_map<std::string, bool> in_bool;
    std::unordered_map<std::string, std::vector<float>> in_vec4;
    std::unordered_map<std::string, std::vector<float>> in_mat4;
};

struct RenderProxy {    // render/render.hpp:45-61
    bool visible = false;
    uint8_t phases = 0;
    size_t vertex_buffer = 0, shader = 0;
    Vec3S position, scale;
    QuatS rotation;
    ShaderInputs inputs;
};

struct Store {
    std::unordered_map<uint64_t, std::bitset<256>> masks;                     // EntityManager::_masks
    std::unordered_map<std::type_index, std::unordered_map<uint64_t, TransformComponent>> transforms;
    std::unordered_map<std::type_index, std::unordered_map<uint64_t, VisualComponent>> visuals;
    std::unordered_map<std::string, size_t> vertex_buffers, shaders;         // IRenderer name tables
    std::unordered_map<size_t, RenderProxy> render_proxies;                  // Frame::render_proxies
};

const char *kMeshNames[8] = { "cube_prefab", "quad_prefab", "cone_prefab", "cylinder_prefab",
                              "icosphere1_prefab", "icosphere2_prefab", "icosphere3_prefab", "dome_prefab" };

}  // namespace

extern "C" {

__attribute__((visibility("default")))
void *refstore_create(uint32_t n, const float *pos3, const float *rot4, const float *scl3,
                      const uint32_t *mesh, const uint32_t *shader, const uint8_t *flags)
{
    Store *s = new Store();
    auto &tr = s->transforms[std::type_index(typeid(TransformComponent))];
    auto &vs = s->visuals[std::type_index(typeid(VisualComponent))];
    for (int m = 0; m < 8; ++m) s->vertex_buffers[kMeshNames[m]] = (size_t)m + 1;
    for (int sh = 0; sh < 256; ++sh) s->shaders["shader_" + std::to_string(sh)] = (size_t)sh + 1;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t e = (uint64_t)i + 1;   // entity 0 is invalid (ecs/entity.hpp:10-13)
        s->masks[e].set(0).set(1);
        TransformComponent &t = tr[e];
        t.position.reset(new Vec3S{ pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2] });
        t.rotation.reset(new QuatS{ rot4[4 * i], rot4[4 * i + 1], rot4[4 * i + 2], rot4[4 * i + 3] });
        t.scale.reset(new Vec3S{ scl3[3 * i], scl3[3 * i + 1], scl3[3 * i + 2] });
        VisualComponent &v = vs[e];
        v.visible = flags[i] & 1u;
        v.vertex_buffer_name = kMeshNames[mesh[i] & 7u];
        v.shader_name = "shader_" + std::to_string(shader[i] & 255u);
    }
    return s;
}

__attribute__((visibility("default")))
void refstore_destroy(void *p) { delete (Store *)p; }

// TransformSystem::run (transform_system.cpp:21-67) over the hash-map store.
__attribute__((visibility("default")))
void refstore_transform(void *p)
{
    Store *s = (Store *)p;
    for (auto &em : s->masks) {
        if (!em.second.test(0)) continue;
        auto &storage = s->transforms.find(std::type_index(typeid(TransformComponent)))->second;
        TransformComponent &t = storage.find(em.first)->second;
        float pos[3] = { 0, 0, 0 }, rot[4] = { 0, 0, 0, 1 }, scl[3] = { 1, 1, 1 };
        if (t.position) { pos[0] = t.position->x; pos[1] = t.position->y; pos[2] = t.position->z; }
        if (t.rotation) { rot[0] = t.rotation->x; rot[1] = t.rotation->y; rot[2] = t.rotation->z; rot[3] = t.rotation->w; }
        if (t.scale) { scl[0] = t.scale->x; scl[1] = t.scale->y; scl[2] = t.scale->z; }
        float m[16];
        orc_trs_matrix(pos, rot, scl, m);
        std::vector<float> ser;                       // math::serialize builds a temporary message
        for (int i = 0; i < 16; ++i) ser.push_back(m[i]);
        t.transform_matrix = ser;                     // CopyFrom
        float f[3], u[3], r[3];
        orc_basis(rot, f, u, r);
        t.forward.reset(new Vec3S{ f[0], f[1], f[2] });
        t.up.reset(new Vec3S{ u[0], u[1], u[2] });
        t.right.reset(new Vec3S{ r[0], r[1], r[2] });
    }
}

// VisualSystem::run (visual_system.cpp:14-61).  Returns the number of proxies.
__attribute__((visibility("default")))
uint64_t refstore_visual(void *p)
{
    Store *s = (Store *)p;
    s->render_proxies.clear();
    for (auto &em : s->masks) {
        if (!em.second.test(0) || !em.second.test(1)) continue;
        const uint64_t e = em.first;
        TransformComponent &t = s->transforms.find(std::type_index(typeid(TransformComponent)))->second.find(e)->second;
        VisualComponent &v = s->visuals.find(std::type_index(typeid(VisualComponent)))->second.find(e)->second;
        auto vb = s->vertex_buffers.find(std::string(v.vertex_buffer_name));
        auto sh = s->shaders.find(std::string(v.shader_name));
        if (vb == s->vertex_buffers.end() || sh == s->shaders.end()) continue;
        s->render_proxies[e].visible = v.visible;
        s->render_proxies[e].vertex_buffer = vb->second;
        s->render_proxies[e].shader = sh->second;
        s->render_proxies[e].inputs.in_bool["useTexture"] = false;
        s->render_proxies[e].inputs.in_vec4["uColor"] = std::vector<float>{ 1.0f, 0.0f, 1.0f, 1.0f };
        s->render_proxies[e].inputs.in_mat4["uModel"] = t.transform_matrix;
        s->render_proxies[e].position = *t.position;
        s->render_proxies[e].rotation = *t.rotation;
        s->render_proxies[e].scale = *t.scale;
        s->render_proxies[e].phases = 1 | 4;
    }
    return s->render_proxies.size();
}

// Copies entity i's transform_matrix out (for checking B0 against the oracle).
__attribute__((visibility("default")))
void refstore_read_world(void *p, uint32_t i, float *out16)
{
    Store *s = (Store *)p;
    const auto &t = s->transforms.find(std::type_index(typeid(TransformComponent)))->second.find((uint64_t)i + 1)->second;
    std::memcpy(out16, t.transform_matrix.data(), 16 * sizeof(float));
}

}  // extern "C"
