// Driver around the reference's OWN sources, compiled unmodified from /root/reference into oracle/_ref
// (see Makefile in this directory).  TEST INFRASTRUCTURE ONLY -- used by tests/ to pin oracle/astre_oracle.c
// and by bench.py's CPU legs; never by the product library.
//
// What runs here is the reference's code, not a restatement:
//   ecs::Registry / EntityManager / ComponentManager   engine/modules/ECS/src/{registry,entity_manager,component_manager}.cpp
//   TransformSystem::run                                engine/modules/ECS/src/system/transform_system.cpp:11-70
//   CameraSystem::run                                   engine/modules/ECS/src/system/camera_system.cpp:13-40
//   VisualSystem::run                                   engine/modules/ECS/src/system/visual_system.cpp:14-61
//   LightSystem::run, calculateLightSpaceMatrices       engine/modules/ECS/src/system/light_system.cpp:24-160
//   render::interpolateFrame                            engine/modules/Render/src/render.cpp:8-68
//   math::serialize / deserialize                       engine/modules/Math/src/math.cpp
// This file only moves flat arrays in and out through a C ABI (ctypes-friendly) and stands in for the
// OpenGL renderer with a name -> id table (the two lookups VisualSystem makes, visual_system.cpp:23-24).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "ecs/ecs_min.hpp"

using namespace astre;

namespace
{
    // IRenderer with just the two name tables VisualSystem reads; everything else is unreachable on this path.
    class TableRenderer final : public render::IRenderer
    {
    public:
        std::unordered_map<std::string, std::size_t> vertex_buffers, shaders;

        [[noreturn]] static void off_path() { throw std::logic_error("TableRenderer: call outside the transform/visibility path"); }

        void move(IRenderer *) override { off_path(); }
        bool good() const override { return true; }
        asio::awaitable<void> close() override { co_return; }
        void join() override {}
        async::AsyncContext<asio::io_context> & getAsyncContext() override { off_path(); }
        asio::awaitable<void> clearScreen(math::Vec4, std::optional<std::size_t>) override { co_return; }
        asio::awaitable<render::FrameStats> render(std::size_t, std::size_t, render::ShaderInputs, render::RenderOptions,
                                                   std::optional<std::size_t>) override { co_return render::FrameStats{}; }
        asio::awaitable<void> present() override { co_return; }
        asio::awaitable<void> updateViewportSize(unsigned int, unsigned int) override { co_return; }
        std::pair<unsigned int, unsigned int> getViewportSize() const override { return {1280u, 720u}; }
        asio::awaitable<void> enableVSync() override { co_return; }
        asio::awaitable<void> disableVSync() override { co_return; }
        asio::awaitable<std::optional<std::size_t>> createVertexBuffer(std::string name, const render::Mesh &) override
        {
            std::size_t id = vertex_buffers.size() + 1;
            vertex_buffers[name] = id;
            co_return id;
        }
        asio::awaitable<bool> eraseVertexBuffer(std::size_t) override { co_return false; }
        std::optional<std::size_t> getVertexBuffer(std::string name) const override
        {
            auto it = vertex_buffers.find(name);
            if (it == vertex_buffers.end()) return std::nullopt;
            return it->second;
        }
        asio::awaitable<std::optional<std::size_t>> createShader(std::string, std::vector<std::string>) override { co_return std::nullopt; }
        asio::awaitable<std::optional<std::size_t>> createShader(std::string, std::vector<std::string>, std::vector<std::string>) override { co_return std::nullopt; }
        asio::awaitable<bool> eraseShader(std::size_t) override { co_return false; }
        std::optional<std::size_t> getShader(std::string name) const override
        {
            auto it = shaders.find(name);
            if (it == shaders.end()) return std::nullopt;
            return it->second;
        }
        asio::awaitable<std::optional<std::size_t>> createShaderStorageBuffer(std::string, unsigned int, const std::size_t, const void *) override { co_return std::nullopt; }
        asio::awaitable<bool> updateShaderStorageBuffer(std::size_t, const std::size_t, const void *) override { co_return false; }
        asio::awaitable<bool> eraseShaderStorageBuffer(std::size_t) override { co_return false; }
        std::optional<std::size_t> getShaderStorageBuffer(std::string) const override { return std::nullopt; }
        asio::awaitable<std::optional<std::size_t>> createFrameBufferObject(std::string, std::pair<unsigned int, unsigned int>,
                                                                            std::initializer_list<render::FBOAttachment>) override { co_return std::nullopt; }
        asio::awaitable<bool> eraseFrameBufferObject(std::size_t) override { co_return false; }
        std::optional<std::size_t> getFrameBufferObject(std::string) const override { return std::nullopt; }
        std::vector<std::size_t> getFrameBufferObjectTextures(std::size_t) const override { return {}; }
        asio::awaitable<std::optional<std::uint64_t>> readPixelUint64(std::size_t, unsigned, int, int) override { co_return std::nullopt; }
    };

    constexpr int N_FRAMES = 4;

    struct World
    {
        process::IProcess process;
        ecs::Registry registry{process};
        TableRenderer renderer;
        ecs::system::TransformSystem transform{registry};
        ecs::system::CameraSystem camera{registry};
        ecs::system::VisualSystem visual{renderer, registry};
        ecs::system::LightSystem light{registry};
        render::Frame frames[N_FRAMES];
        std::string error;
    };

    void put_mat4(const math::Mat4 & m, float * out)   // column-major, the order Math/src/math.cpp:70-93 serialises
    {
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) out[c * 4 + r] = m[c][r];
    }

    template<class F>
    int guarded(World * w, F && body)
    {
        try { body(); return 0; }
        catch (const std::exception & e) { if (w) w->error = e.what(); return -1; }
    }
}

extern "C"
{
    // Bit flags of `has` in ref_add_transforms: which of the three proto sub-messages are present.
    enum { REF_HAS_POSITION = 1, REF_HAS_ROTATION = 2, REF_HAS_SCALE = 4 };

    void * ref_world_create() { try { return new World(); } catch (...) { return nullptr; } }
    void ref_world_destroy(void * h) { delete static_cast<World *>(h); }
    const char * ref_last_error(void * h) { return static_cast<World *>(h)->error.c_str(); }

    // Spawns n entities with a TransformComponent each (Registry::spawnEntity + addComponent, as EntityLoader does,
    // Loader/src/entity_loader.cpp:11-37).  rot is (x,y,z,w) per entity; `has` may be NULL (= all present).
    int ref_add_transforms(void * h, std::uint32_t n, const std::uint64_t * ids, const std::uint8_t * has,
                           const float * pos3, const float * rot_xyzw4, const float * scale3)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            for (std::uint32_t i = 0; i < n; ++i)
            {
                proto::ecs::EntityDefinition def;
                def.set_id(ids[i]);
                def.set_name("e");
                if (!w->registry.spawnEntity(def).get()) throw std::runtime_error("spawnEntity failed");
                const unsigned m = has ? has[i] : 7u;
                proto::ecs::TransformComponent t;
                if (m & REF_HAS_POSITION)
                {
                    auto * p = t.mutable_position();
                    p->set_x(pos3[3 * i]); p->set_y(pos3[3 * i + 1]); p->set_z(pos3[3 * i + 2]);
                }
                if (m & REF_HAS_ROTATION)
                {
                    auto * q = t.mutable_rotation();
                    q->set_x(rot_xyzw4[4 * i]); q->set_y(rot_xyzw4[4 * i + 1]); q->set_z(rot_xyzw4[4 * i + 2]); q->set_w(rot_xyzw4[4 * i + 3]);
                }
                if (m & REF_HAS_SCALE)
                {
                    auto * s = t.mutable_scale();
                    s->set_x(scale3[3 * i]); s->set_y(scale3[3 * i + 1]); s->set_z(scale3[3 * i + 2]);
                }
                w->registry.addComponent<proto::ecs::TransformComponent>(ids[i], std::move(t)).get();
            }
        });
    }

    // Overwrites position / rotation / scale of existing entities (what a script tick does,
    // Script/include/script/component_binding.hpp:26-115).  NULL arrays are left untouched.
    int ref_set_trs(void * h, std::uint32_t n, const std::uint64_t * ids, const float * pos3, const float * rot_xyzw4, const float * scale3)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            for (std::uint32_t i = 0; i < n; ++i)
                w->registry.runOnSingleWithComponents<proto::ecs::TransformComponent>(ids[i],
                    [&](const ecs::Entity, proto::ecs::TransformComponent & t) {
                        if (pos3) { auto * p = t.mutable_position(); p->set_x(pos3[3 * i]); p->set_y(pos3[3 * i + 1]); p->set_z(pos3[3 * i + 2]); }
                        if (rot_xyzw4) { auto * q = t.mutable_rotation(); q->set_x(rot_xyzw4[4 * i]); q->set_y(rot_xyzw4[4 * i + 1]); q->set_z(rot_xyzw4[4 * i + 2]); q->set_w(rot_xyzw4[4 * i + 3]); }
                        if (scale3) { auto * s = t.mutable_scale(); s->set_x(scale3[3 * i]); s->set_y(scale3[3 * i + 1]); s->set_z(scale3[3 * i + 2]); }
                    });
        });
    }

    int ref_add_visual(void * h, std::uint64_t id, int visible, const char * vertex_buffer_name, const char * shader_name, const float * color4)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            proto::ecs::VisualComponent v;
            v.set_visible(visible != 0);
            v.set_vertex_buffer_name(vertex_buffer_name);
            v.set_shader_name(shader_name);
            if (color4)
            {
                auto * c = v.mutable_color();
                c->set_x(color4[0]); c->set_y(color4[1]); c->set_z(color4[2]); c->set_w(color4[3]);
            }
            w->registry.addComponent<proto::ecs::VisualComponent>(id, std::move(v)).get();
        });
    }

    int ref_add_camera(void * h, std::uint64_t id, float fov_deg, float near_plane, float far_plane, float aspect)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            proto::ecs::CameraComponent c;
            c.set_fov(fov_deg); c.set_near_plane(near_plane); c.set_far_plane(far_plane); c.set_aspect(aspect);
            w->registry.addComponent<proto::ecs::CameraComponent>(id, std::move(c)).get();
        });
    }

    // type: proto::ecs::LightType (1 directional, 2 point, 3 spot).  cutoffs are cosines.
    int ref_add_light(void * h, std::uint64_t id, int type, int cast_shadows, const float * color4,
                      float constant, float linear, float quadratic, float inner_cutoff, float outer_cutoff)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            proto::ecs::LightComponent l;
            l.set_type(static_cast<proto::ecs::LightType>(type));
            l.set_cast_shadows(cast_shadows != 0);
            if (color4)
            {
                auto * c = l.mutable_color();
                c->set_x(color4[0]); c->set_y(color4[1]); c->set_z(color4[2]); c->set_w(color4[3]);
            }
            l.set_constant(constant); l.set_linear(linear); l.set_quadratic(quadratic);
            l.set_inner_cutoff(inner_cutoff); l.set_outer_cutoff(outer_cutoff);
            w->registry.addComponent<proto::ecs::LightComponent>(id, std::move(l)).get();
        });
    }

    int ref_register_vertex_buffer(void * h, const char * name, std::uint64_t id)
    {
        static_cast<World *>(h)->renderer.vertex_buffers[name] = static_cast<std::size_t>(id);
        return 0;
    }
    int ref_register_shader(void * h, const char * name, std::uint64_t id)
    {
        static_cast<World *>(h)->renderer.shaders[name] = static_cast<std::size_t>(id);
        return 0;
    }

    // ---- the reference's systems -------------------------------------------------------------------------
    int ref_run_transform(void * h)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] { w->transform.run(0.1f).get(); });
    }

    // transform_matrix (16 floats, column-major as serialised) and forward / up / right (3 floats each) per id.
    // Any output pointer may be NULL.
    int ref_get_transforms(void * h, std::uint32_t n, const std::uint64_t * ids, float * mat16, float * forward3, float * up3, float * right3)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            for (std::uint32_t i = 0; i < n; ++i)
            {
                bool found = false;
                w->registry.runOnSingleWithComponents<proto::ecs::TransformComponent>(ids[i],
                    [&](const ecs::Entity, proto::ecs::TransformComponent & t) {
                        found = true;
                        if (mat16)
                        {
                            if (t.transform_matrix().data_size() != 16) throw std::runtime_error("transform_matrix not computed");
                            for (int k = 0; k < 16; ++k) mat16[16 * std::size_t(i) + k] = t.transform_matrix().data(k);
                        }
                        if (forward3) { forward3[3 * i] = t.forward().x(); forward3[3 * i + 1] = t.forward().y(); forward3[3 * i + 2] = t.forward().z(); }
                        if (up3) { up3[3 * i] = t.up().x(); up3[3 * i + 1] = t.up().y(); up3[3 * i + 2] = t.up().z(); }
                        if (right3) { right3[3 * i] = t.right().x(); right3[3 * i + 1] = t.right().y(); right3[3 * i + 2] = t.right().z(); }
                    });
                if (!found) throw std::runtime_error("no such entity / no TransformComponent");
            }
        });
    }

    int ref_run_camera(void * h, std::uint64_t camera_entity, int frame, float * view16, float * proj16, float * position3)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            render::Frame & f = w->frames[frame];
            w->camera.setActiveCameraEntity(camera_entity);
            w->camera.run(0.1f, f);
            if (view16) put_mat4(f.view_matrix, view16);
            if (proj16) put_mat4(f.proj_matrix, proj16);
            if (position3) { position3[0] = f.camera_position.x; position3[1] = f.camera_position.y; position3[2] = f.camera_position.z; }
        });
    }

    int ref_run_visual(void * h, int frame)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] { w->visual.run(0.1f, w->frames[frame]).get(); });
    }

    std::uint64_t ref_proxy_count(void * h, int frame) { return static_cast<World *>(h)->frames[frame].render_proxies.size(); }

    // ids of the frame's proxies, ascending (the map's own order is unspecified).
    int ref_proxy_ids(void * h, int frame, std::uint64_t * ids, std::uint64_t cap)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            std::vector<std::uint64_t> v;
            for (auto & [id, p] : w->frames[frame].render_proxies) v.push_back(id);
            std::sort(v.begin(), v.end());
            if (v.size() > cap) throw std::runtime_error("id buffer too small");
            std::memcpy(ids, v.data(), v.size() * sizeof(std::uint64_t));
        });
    }

    // One proxy, flattened: out_u64 = {visible, phases, vertex_buffer, shader, useTexture},
    // out_f32 = position3 | rotation xyzw | scale3 | uModel 16 (column-major) | uColor 4  (30 floats).
    // Returns 1 when the frame has no proxy for `id`.
    int ref_get_proxy(void * h, int frame, std::uint64_t id, std::uint64_t * out_u64, float * out_f32)
    {
        World * w = static_cast<World *>(h);
        int missing = 0;
        int rc = guarded(w, [&] {
            auto & map = w->frames[frame].render_proxies;
            auto it = map.find(static_cast<std::size_t>(id));
            if (it == map.end()) { missing = 1; return; }
            const render::RenderProxy & p = it->second;
            out_u64[0] = p.visible ? 1 : 0;
            out_u64[1] = static_cast<std::uint8_t>(p.phases);
            out_u64[2] = p.vertex_buffer;
            out_u64[3] = p.shader;
            auto tex = p.inputs.in_bool.find("useTexture");
            out_u64[4] = (tex != p.inputs.in_bool.end() && tex->second) ? 1 : 0;
            float * o = out_f32;
            *o++ = p.position.x; *o++ = p.position.y; *o++ = p.position.z;
            *o++ = p.rotation.x; *o++ = p.rotation.y; *o++ = p.rotation.z; *o++ = p.rotation.w;
            *o++ = p.scale.x; *o++ = p.scale.y; *o++ = p.scale.z;
            put_mat4(p.inputs.in_mat4.at("uModel"), o); o += 16;
            const math::Vec4 & c = p.inputs.in_vec4.at("uColor");
            *o++ = c.x; *o++ = c.y; *o++ = c.z; *o++ = c.w;
        });
        return rc ? rc : missing;
    }

    // uModel of many proxies at once (16 floats each, in the order of `ids`); used for interpolateFrame checks.
    int ref_get_models(void * h, int frame, std::uint32_t n, const std::uint64_t * ids, float * mat16)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            auto & map = w->frames[frame].render_proxies;
            for (std::uint32_t i = 0; i < n; ++i) put_mat4(map.at(static_cast<std::size_t>(ids[i])).inputs.in_mat4.at("uModel"), mat16 + 16 * std::size_t(i));
        });
    }

    int ref_run_light(void * h, int frame)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] { w->light.run(0.1f, w->frames[frame]).get(); });
    }

    std::uint32_t ref_shadow_casters_count(void * h, int frame) { return static_cast<World *>(h)->frames[frame].shadow_casters_count; }

    // GPULight of entity `id`: position4 | direction4 | color4 | attenuation4 | cutoff2 | castShadows2 (20 floats).
    int ref_get_light(void * h, int frame, std::uint64_t id, float * out20)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] {
            const render::GPULight & l = w->frames[frame].gpu_lights.at(static_cast<std::size_t>(id));
            static_assert(sizeof(render::GPULight) == 20 * sizeof(float));
            std::memcpy(out20, &l, sizeof l);
        });
    }

    // calculateLightSpaceMatrices(frame): writes shadow_casters_count matrices (16 floats each), returns the count
    // (negative on error).
    int ref_light_space_matrices(void * h, int frame, float * out16, std::uint32_t cap)
    {
        World * w = static_cast<World *>(h);
        int count = 0;
        int rc = guarded(w, [&] {
            std::vector<math::Mat4> ms = ecs::system::calculateLightSpaceMatrices(w->frames[frame]);
            if (ms.size() > cap) throw std::runtime_error("matrix buffer too small");
            for (std::size_t i = 0; i < ms.size(); ++i) put_mat4(ms[i], out16 + 16 * i);
            count = static_cast<int>(ms.size());
        });
        return rc ? rc : count;
    }

    // frames[dst] = render::interpolateFrame(frames[a], frames[b], alpha)
    int ref_interpolate(void * h, int a, int b, float alpha, int dst)
    {
        World * w = static_cast<World *>(h);
        return guarded(w, [&] { w->frames[dst] = render::interpolateFrame(w->frames[a], w->frames[b], alpha); });
    }

    int ref_get_frame_camera(void * h, int frame, float * view16, float * proj16, float * position3)
    {
        World * w = static_cast<World *>(h);
        const render::Frame & f = w->frames[frame];
        put_mat4(f.view_matrix, view16);
        put_mat4(f.proj_matrix, proj16);
        position3[0] = f.camera_position.x; position3[1] = f.camera_position.y; position3[2] = f.camera_position.z;
        return 0;
    }

    // Seconds for `iters` calls of TransformSystem::run followed by VisualSystem::run on frame 0 -- the reference's
    // per-tick work on this path, on its own storage (one coroutine = one thread, logic_pipelines.cpp:31,49).
    double ref_time_tick(void * h, int iters, int with_visual)
    {
        World * w = static_cast<World *>(h);
        double secs = -1.0;
        guarded(w, [&] {
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < iters; ++i)
            {
                w->transform.run(0.1f).get();
                if (with_visual) w->visual.run(0.1f, w->frames[0]).get();
            }
            secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
        return secs;
    }
}
