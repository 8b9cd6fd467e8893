// Tests of the engine-side glue (astre_b200/host) in the style of the reference's gtest
// suite (engine/tests): build a small world like game/resources/worlds/levels/level_0.json,
// run the systems, compare with the CPU oracle.  `--cpu` runs only the checks that need no device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../astre_b200/host/astre_host.hpp"

using namespace astre::gpu;

extern "C" {
struct orc_bounds { float center[3], extent[3], radius; uint32_t reserved; };
struct orc_view { float plane[6][4], aplane[6][4], depth[4], depth_scale; uint32_t phase_mask, pad[2]; float lo[4], hi[4]; };
void orc_trs_matrix(const float *pos, const float *rot, const float *scl, float *out16);
void orc_basis(const float *rot, float *fwd, float *up, float *right);
void orc_camera_view_proj(const float *pos, const float *fwd, const float *up, float fov, float aspect, float zn, float zf,
                          float *view16, float *proj16);
int orc_light_space(const float *pos3, const float *dir3, int type, float outer_cutoff, float *out16);
void orc_mat4_mul(const float *a, const float *b, float *out);
void orc_extract_view(const float *clip16, uint32_t phase_mask, orc_view *v);
void orc_interpolate_trs(uint32_t n, float alpha, const float *pos_a, const float *rot_a, const float *scl_a,
                         const float *pos_b, const float *rot_b, const float *scl_b, float *world16);
uint64_t orc_frame(uint32_t n, const float *pos3, const float *rot4, const float *scl3, const uint32_t *parent,
                   const uint32_t *mesh_id, const uint32_t *shader_id, const uint8_t *flags, const orc_bounds *bounds,
                   uint32_t n_meshes, uint32_t n_worlds, uint32_t epw, uint32_t vpw, const orc_view *views, float *world16,
                   uint64_t *keys, uint64_t cap, uint32_t *counts);
}

static int g_failures = 0, g_checks = 0;
#define EXPECT(cond)                                                                          \
    do {                                                                                      \
        ++g_checks;                                                                           \
        if (!(cond)) { ++g_failures; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static bool same_bits(const float *a, const float *b, int n)
{
    for (int i = 0; i < n; ++i) {
        if (std::isnan(a[i]) && std::isnan(b[i])) continue;
        if (std::memcmp(&a[i], &b[i], 4) != 0) return false;
    }
    return true;
}

struct Names : IResourceNames {   // stands in for IRenderer's name tables
    std::optional<std::size_t> getVertexBuffer(const std::string &n) const override
    {
        static const char *k[] = {"cube_prefab", "quad_prefab", "cone_prefab", "cylinder_prefab",
                                  "icosphere1_prefab", "icosphere2_prefab", "icosphere3_prefab", "dome_prefab"};
        for (std::size_t i = 0; i < 8; ++i) if (n == k[i]) return i;
        return std::nullopt;
    }
    std::optional<std::size_t> getShader(const std::string &n) const override
    {
        if (n == "deferred_shader") return 0;
        if (n == "basic_shader") return 1;
        return std::nullopt;
    }
};

static const astre_mesh_bounds kBounds[8] = {
    {{0, 0, 0}, {0.5f, 0.5f, 0.5f}, 0, 0}, {{0, 0, 0}, {0.5f, 0.5f, 0}, 0, 0}, {{0, 0, 0}, {0.5f, 0.5f, 0.5f}, 0, 0},
    {{0, 0, 0}, {0.5f, 0.5f, 0.5f}, 0, 0}, {{0, 0, 0}, {0, 0, 0}, 1, 0},       {{0, 0, 0}, {0, 0, 0}, 1, 0},
    {{0, 0, 0}, {0, 0, 0}, 1, 0},          {{0, 0, 0}, {0, 0, 0}, 1, 0}};

static void test_mirror_cpu()
{
    Names names;
    EntityMirror m(8);
    TransformComponent t;
    t.position = Vec3{1, 2, 3};
    VisualComponent v;
    v.vertex_buffer_name = "cube_prefab"; v.shader_name = "deferred_shader";
    EXPECT(m.spawn(10, t, &v, &names) == 0);
    EXPECT(m.spawn(11, t, nullptr, nullptr) == 1);
    VisualComponent bad = v;
    bad.shader_name = "no_such_shader";
    EXPECT(m.spawn(12, t, &bad, &names) == 2);
    EXPECT(m.flags(0) == ASTRE_FLAGS(1, 5));          // visible | Opaque | ShadowCaster
    EXPECT(m.flags(1) == 0 && m.flags(2) == 0);       // no Visual / unresolved name: no proxy (visual_system.cpp:26-30)
    EXPECT(m.rotation(0)[3] == 1.0f && m.scale(0)[1] == 1.0f);   // defaults (transform_system.cpp:36-49)
    EXPECT(m.slotCount() == 3 && m.liveCount() == 3 && m.pendingCount() == 3);
    m.despawn(11);
    EXPECT(!m.contains(11) && m.liveCount() == 2);
    EXPECT(m.spawn(13, t, &v, &names, /*parent*/ 10) == 1);      // freed slot is reused
    EXPECT(m.entityAt(1) == 13);
    bool threw = false;
    try { m.spawn(13, t, &v, &names); } catch (const std::invalid_argument &) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { m.spawn(0, t, &v, &names); } catch (const std::invalid_argument &) { threw = true; }
    EXPECT(threw);
}

struct World {
    Names names;
    EntityMirror mirror;
    astre_gpu_ctx *ctx;
    GpuTransformSystem transforms;
    GpuCameraSystem cameras;
    GpuVisualSystem visuals;
    explicit World(uint32_t cap)
        : mirror(cap), ctx(astre_gpu_create(0, cap, 8, 0)), transforms(mirror, ctx), cameras(mirror, transforms), visuals(mirror, ctx)
    {
        if (!ctx) { std::printf("astre_gpu_create failed: %s\n", astre_gpu_last_error(nullptr)); std::exit(2); }
        astre_gpu_set_mesh_bounds(ctx, 8, kBounds);
    }
    ~World() { astre_gpu_destroy(ctx); }
};

// Oracle result for the mirror's current host state and the frame's views.
static void oracle_frame(const EntityMirror &m, const Frame &f, std::vector<float> &world, std::vector<uint64_t> &keys,
                         std::vector<uint32_t> &counts)
{
    const uint32_t n = m.slotCount(), F = 1 + (uint32_t)f.light_space_matrices.size();
    std::vector<uint32_t> mesh(n), shader(n);
    std::vector<uint8_t> flags(n);
    for (uint32_t i = 0; i < n; ++i) { mesh[i] = m.mesh(i); shader[i] = m.shader(i); flags[i] = m.flags(i); }
    std::vector<orc_view> views(F);
    float clip[16];
    orc_mat4_mul(f.proj_matrix.data(), f.view_matrix.data(), clip);
    orc_extract_view(clip, ASTRE_PHASE_OPAQUE, &views[0]);
    for (uint32_t v = 1; v < F; ++v) orc_extract_view(f.light_space_matrices[v - 1].data(), ASTRE_PHASE_SHADOW_CASTER, &views[v]);
    orc_bounds b[8];
    std::memcpy(b, kBounds, sizeof b);
    world.resize(16 * (size_t)n); keys.resize((size_t)n * F); counts.assign(F, 0);
    const uint64_t total = orc_frame(n, m.position(0), m.rotation(0), m.scale(0), nullptr, mesh.data(), shader.data(), flags.data(),
                                     b, 8, 1, 0, F, views.data(), world.data(), keys.data(), keys.size(), counts.data());
    keys.resize(total);
}

static void check_against_oracle(World &w, Frame &frame)
{
    std::vector<float> o_world; std::vector<uint64_t> o_keys; std::vector<uint32_t> o_counts;
    oracle_frame(w.mirror, frame, o_world, o_keys, o_counts);
    EXPECT(frame.draw_lists.size() == o_counts.size());
    size_t k = 0;
    bool lists_ok = true, models_ok = true;
    for (size_t v = 0; v < o_counts.size() && v < frame.draw_lists.size(); ++v) {
        if (frame.draw_lists[v].size() != o_counts[v]) { lists_ok = false; break; }
        for (uint32_t i = 0; i < o_counts[v]; ++i, ++k) {
            const uint32_t slot = (uint32_t)(o_keys[k] & ((1ull << 26) - 1));
            if (frame.draw_lists[v][i] != w.mirror.entityAt(slot)) lists_ok = false;
            const auto it = frame.render_proxies.find((size_t)w.mirror.entityAt(slot));
            if (it == frame.render_proxies.end() || !same_bits(it->second.uModel.data(), &o_world[16 * (size_t)slot], 16)) models_ok = false;
        }
    }
    EXPECT(lists_ok);
    EXPECT(models_ok);
}

// f2: the runs tile every draw list exactly and are homogeneous in (shader, mesh)
static void check_draw_runs(World &w, const Frame &frame)
{
    std::vector<size_t> covered(frame.draw_lists.size(), 0);
    for (const Frame::DrawRun &r : frame.draw_runs) {
        EXPECT(r.view < frame.draw_lists.size());
        EXPECT(r.first == covered[r.view]);                          // runs come in list order, no gaps
        EXPECT(r.count > 0 && r.first + r.count <= frame.draw_lists[r.view].size());
        for (uint32_t i = 0; i < r.count; ++i) {
            const uint32_t slot = w.mirror.slotOf(frame.draw_lists[r.view][r.first + i]);
            if (w.mirror.shader(slot) != r.shader || w.mirror.mesh(slot) != r.vertex_buffer) { EXPECT(false); break; }
        }
        covered[r.view] += r.count;
    }
    for (size_t v = 0; v < frame.draw_lists.size(); ++v) EXPECT(covered[v] == frame.draw_lists[v].size());
}

// f2 on the device: command i of frame.draw_commands_dev belongs to draw_runs[i]
static void check_draw_commands(World &w, const Frame &frame, const std::vector<astre_mesh_draw_info> &info)
{
    EXPECT(frame.draw_commands_dev != nullptr && frame.draw_command_count == frame.draw_runs.size());
    std::vector<astre_draw_indirect> cmds(frame.draw_command_count);
    uint32_t n = 0;
    EXPECT(astre_gpu_read_indirect(w.ctx, cmds.data(), (uint32_t)cmds.size(), &n) == ASTRE_OK && n == cmds.size());
    std::vector<size_t> view_first(frame.draw_lists.size() + 1, 0);
    for (size_t v = 0; v < frame.draw_lists.size(); ++v) view_first[v + 1] = view_first[v] + frame.draw_lists[v].size();
    for (size_t i = 0; i < cmds.size() && i < frame.draw_runs.size(); ++i) {
        const Frame::DrawRun &r = frame.draw_runs[i];
        const bool known = r.vertex_buffer < info.size();
        EXPECT(cmds[i].instance_count == r.count && cmds[i].base_instance == view_first[r.view] + r.first);
        EXPECT(cmds[i].count == (known ? info[r.vertex_buffer].index_count : 0u));
        EXPECT(cmds[i].first_index == (known ? info[r.vertex_buffer].first_index : 0u));
        EXPECT(cmds[i].base_vertex == (known ? info[r.vertex_buffer].base_vertex : 0));
    }
}

static void test_level0_like_world()
{
    World w(64);
    // 13 props + a camera, in the spirit of level_0.json (one rotated entity, authored scales)
    const char *meshes[] = {"cube_prefab", "icosphere2_prefab", "cone_prefab", "cylinder_prefab", "quad_prefab"};
    for (int i = 0; i < 13; ++i) {
        TransformComponent t;
        t.position = Vec3{(float)(i % 5) * 4.0f - 8.0f, (float)(i / 5), -(float)i * 3.0f};
        t.scale = Vec3{1.0f + 0.25f * i, 1.0f, 2.0f};
        if (i == 3) t.rotation = Quat{0.0f, 0.38268343f, 0.0f, 0.92387953f};
        VisualComponent v;
        v.vertex_buffer_name = meshes[i % 5];
        v.shader_name = (i % 2) ? "deferred_shader" : "basic_shader";
        v.visible = (i != 7);
        w.mirror.spawn(100 + i, t, &v, &w.names);
    }
    TransformComponent cam;
    cam.position = Vec3{0, 5, 15};
    cam.scale = Vec3{1, 1, 1};
    w.mirror.spawn(5, cam, nullptr, nullptr);          // level_0.json:118-141: entity 5 is the camera
    w.cameras.active_camera_entity = 5;

    Frame frame;
    w.transforms.run(0.1f);
    // TransformSystem: transform_matrix + forward/up/right of every entity
    bool mats = true, basis = true;
    for (uint32_t s = 0; s < w.mirror.slotCount(); ++s) {
        TransformComponent out;
        w.transforms.read(w.mirror.entityAt(s), out);
        float m[16], f[3], u[3], r[3];
        orc_trs_matrix(w.mirror.position(s), w.mirror.rotation(s), w.mirror.scale(s), m);
        orc_basis(w.mirror.rotation(s), f, u, r);
        mats &= same_bits(m, out.transform_matrix.data(), 16);
        basis &= same_bits(f, &out.forward.x, 3) && same_bits(u, &out.up.x, 3) && same_bits(r, &out.right.x, 3);
    }
    EXPECT(mats);
    EXPECT(basis);
    // CameraSystem
    w.cameras.run(0.1f, frame);
    float view[16], proj[16];
    const float fwd[3] = {0, 0, -1}, up[3] = {0, 1, 0}, pos[3] = {0, 5, 15};
    orc_camera_view_proj(pos, fwd, up, 60.0f, 1.7f, 0.1f, 1000.0f, view, proj);
    EXPECT(same_bits(view, frame.view_matrix.data(), 16));
    EXPECT(same_bits(proj, frame.proj_matrix.data(), 16));
    // one directional and one spot shadow caster
    GPULight sun; sun.position = Vec4{0, 30, 0, 1}; sun.direction = Vec4{-0.3f, -1.0f, -0.2f, 1.0f}; sun.cast_shadows = true; sun.shadow_map_index = 0;
    GPULight spot; spot.position = Vec4{0, 8, 4, 1}; spot.direction = Vec4{0, -0.5f, -1.0f, 3.0f}; spot.outer_cutoff = 0.85f; spot.cast_shadows = true; spot.shadow_map_index = 1;
    frame.gpu_lights[1] = sun; frame.gpu_lights[2] = spot;
    frame.shadow_casters_count = 2;
    frame.light_space_matrices = calculateLightSpaceMatrices(frame);
    float ls[16];
    orc_light_space(&sun.position.x, &sun.direction.x, 1, 0.0f, ls);
    EXPECT(same_bits(ls, frame.light_space_matrices[0].data(), 16));
    orc_light_space(&spot.position.x, &spot.direction.x, 3, 0.85f, ls);
    EXPECT(same_bits(ls, frame.light_space_matrices[1].data(), 16));
    // VisualSystem
    w.visuals.run(0.1f, frame);
    EXPECT(!frame.draw_lists.empty() && !frame.draw_lists[0].empty());
    EXPECT(frame.render_proxies.count(107) == 0);      // visible == false never reaches a list
    EXPECT(frame.render_proxies.count(5) == 0);        // the camera has no Visual component
    check_against_oracle(w, frame);

    // script writes between ticks, a despawn and a spawn into the freed slot
    w.mirror.setPosition(103, Vec3{2.0f, 1.0f, -4.0f});
    w.mirror.setRotation(104, Quat{0.5f, 0.5f, 0.5f, 0.5f});
    w.mirror.setScale(105, Vec3{-1.0f, 2.0f, 0.5f});
    w.mirror.setVisible(107, true);
    w.mirror.despawn(101);
    TransformComponent t; t.position = Vec3{1, 1, -6};
    VisualComponent v; v.vertex_buffer_name = "dome_prefab"; v.shader_name = "deferred_shader";
    w.mirror.spawn(200, t, &v, &w.names);
    EXPECT(w.mirror.pendingCount() > 0);
    w.transforms.run(0.1f);
    EXPECT(w.mirror.pendingCount() == 0);
    w.visuals.run(0.1f, frame);
    EXPECT(frame.render_proxies.count(101) == 0 && frame.render_proxies.count(200) == 1 && frame.render_proxies.count(107) == 1);
    check_against_oracle(w, frame);
}

static void test_random_world(uint32_t n)
{
    World w(n + 1);
    std::mt19937 rng(42);
    std::uniform_real_distribution<float> P(-60.0f, 60.0f), Q(-1.0f, 1.0f), S(0.25f, 3.0f);
    const char *meshes[] = {"cube_prefab", "icosphere2_prefab", "cone_prefab", "dome_prefab", "quad_prefab", "ghost_mesh"};
    for (uint32_t i = 0; i < n; ++i) {
        TransformComponent t;
        if (i % 7) t.position = Vec3{P(rng), P(rng) * 0.2f, P(rng) - 50.0f};
        if (i % 3) { Quat q{Q(rng), Q(rng), Q(rng), Q(rng)}; const float l = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); t.rotation = Quat{q.x / l, q.y / l, q.z / l, q.w / l}; }
        if (i % 5) t.scale = Vec3{S(rng), S(rng), S(rng)};
        VisualComponent v;
        v.vertex_buffer_name = meshes[i % 6];
        v.shader_name = (i % 4) ? "deferred_shader" : "basic_shader";
        v.visible = (i % 11) != 0;
        w.mirror.spawn(1000 + i, t, (i % 13) ? &v : nullptr, &w.names);
    }
    TransformComponent cam; cam.position = Vec3{0, 5, 15};
    w.mirror.spawn(5, cam, nullptr, nullptr);
    w.cameras.active_camera_entity = 5;
    Frame frame;
    w.transforms.run(0.1f);
    w.cameras.run(0.1f, frame);
    GPULight sun; sun.position = Vec4{0, 40, -40, 1}; sun.direction = Vec4{-0.3f, -1.0f, -0.2f, 1.0f}; sun.cast_shadows = true; sun.shadow_map_index = 0;
    frame.gpu_lights[1] = sun;
    frame.shadow_casters_count = 1;
    frame.light_space_matrices = calculateLightSpaceMatrices(frame);
    w.visuals.run(0.1f, frame);
    EXPECT(frame.draw_lists[0].size() > n / 50);
    check_against_oracle(w, frame);
    check_draw_runs(w, frame);
    EXPECT(frame.draw_runs.size() >= 4 && frame.draw_runs.size() < frame.draw_lists[0].size());
    EXPECT(frame.draw_commands_dev == nullptr);                     // off until the renderer's index ranges are known
    const std::vector<astre_mesh_draw_info> info = {{36, 0, 0}, {6, 36, 24}, {96, 42, 28}, {192, 138, 61}, {60, 330, 94}};
    w.visuals.setMeshDrawInfo(info);
    w.visuals.run(0.1f, frame);
    check_draw_runs(w, frame);
    check_draw_commands(w, frame, info);
    for (uint32_t i = 0; i < n; i += 17) w.mirror.setPosition(1000 + i, Vec3{P(rng), 1.0f, P(rng) - 50.0f});
    for (uint32_t i = 0; i < n; i += 29) w.mirror.despawn(1000 + i);
    w.transforms.run(0.1f);
    w.visuals.run(0.1f, frame);
    check_against_oracle(w, frame);
}

// f1: GpuFrameInterpolator::interpolateFrame against render::interpolateFrame's per-proxy arithmetic (oracle).
static void test_frame_interpolation(uint32_t n)
{
    World w(n + 8);
    GpuFrameInterpolator interp(w.mirror, w.ctx);
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> P(-40.0f, 40.0f), Q(-1.0f, 1.0f), S(0.5f, 2.0f);
    auto unit_quat = [&]() { Quat q{Q(rng), Q(rng), Q(rng), Q(rng)}; const float l = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); return Quat{q.x / l, q.y / l, q.z / l, q.w / l}; };
    for (uint32_t i = 0; i < n; ++i) {
        TransformComponent t;
        t.position = Vec3{P(rng), P(rng) * 0.2f, P(rng) - 45.0f};
        t.rotation = unit_quat();
        t.scale = Vec3{S(rng), S(rng), S(rng)};
        VisualComponent v; v.vertex_buffer_name = "cube_prefab"; v.shader_name = "deferred_shader"; v.visible = true;
        w.mirror.spawn(2000 + i, t, &v, &w.names);
    }
    TransformComponent cam; cam.position = Vec3{0, 5, 15};
    w.mirror.spawn(5, cam, nullptr, nullptr);
    w.cameras.active_camera_entity = 5;
    Frame fa, fb;
    w.transforms.run(0.1f); w.cameras.run(0.1f, fa); w.visuals.run(0.1f, fa);
    interp.snapshot();                                                   // tick a is what the renderer shows now
    const uint32_t n_a = w.mirror.slotCount();
    std::vector<float> pa(w.mirror.position(0), w.mirror.position(0) + 3 * n_a), ra(w.mirror.rotation(0), w.mirror.rotation(0) + 4 * n_a),
        sa(w.mirror.scale(0), w.mirror.scale(0) + 3 * n_a);
    for (uint32_t i = 0; i < n; i += 2) w.mirror.setPosition(2000 + i, Vec3{P(rng), 1.0f, P(rng) - 45.0f});
    for (uint32_t i = 0; i < n; i += 3) w.mirror.setRotation(2000 + i, unit_quat());
    for (uint32_t i = 0; i < n; i += 5) w.mirror.setScale(2000 + i, Vec3{S(rng), S(rng), S(rng)});
    w.mirror.setPosition(5, Vec3{2, 6, 14});
    TransformComponent late; late.position = Vec3{0, 0, -20};
    VisualComponent lv; lv.vertex_buffer_name = "cube_prefab"; lv.shader_name = "deferred_shader"; lv.visible = true;
    w.mirror.spawn(9000, late, &lv, &w.names);                           // exists in b only
    w.transforms.run(0.1f); w.cameras.run(0.1f, fb); w.visuals.run(0.1f, fb);
    const float alpha = 0.25f;
    Frame r = interp.interpolateFrame(fa, fb, alpha);
    const uint32_t n_b = w.mirror.slotCount();
    std::vector<float> pa2(3 * n_b), ra2(4 * n_b), sa2(3 * n_b), exp(16 * (size_t)n_b);
    std::copy(w.mirror.position(0), w.mirror.position(0) + 3 * n_b, pa2.begin());   // slots new in b: a := b
    std::copy(w.mirror.rotation(0), w.mirror.rotation(0) + 4 * n_b, ra2.begin());
    std::copy(w.mirror.scale(0), w.mirror.scale(0) + 3 * n_b, sa2.begin());
    std::copy(pa.begin(), pa.end(), pa2.begin()); std::copy(ra.begin(), ra.end(), ra2.begin()); std::copy(sa.begin(), sa.end(), sa2.begin());
    orc_interpolate_trs(n_b, alpha, pa2.data(), ra2.data(), sa2.data(), w.mirror.position(0), w.mirror.rotation(0), w.mirror.scale(0), exp.data());
    EXPECT(r.render_proxies.size() == fb.render_proxies.size());
    size_t both = 0, exact_t = 0, close_r = 0;
    for (const auto &[id, proxy] : r.render_proxies) {
        if (!fa.render_proxies.count(id)) {                              // render.cpp:28-29: keeps b's matrix
            EXPECT(same_bits(proxy.uModel.data(), fb.render_proxies.at(id).uModel.data(), 16));
            continue;
        }
        ++both;
        const float *e = &exp[16 * (size_t)w.mirror.slotOf((Entity)id)];
        exact_t += same_bits(proxy.uModel.data() + 12, e + 12, 4);       // translation column: mix only
        bool ok = true;
        for (int i = 0; i < 12; ++i) ok = ok && std::fabs(proxy.uModel[i] - e[i]) <= 2e-5f * (1.0f + std::fabs(e[i]));   // acosf/sinf: glibc vs CUDA
        close_r += ok;
    }
    EXPECT(both > n / 50);
    EXPECT(exact_t == both);
    EXPECT(close_r == both);
    EXPECT(fb.render_proxies.count(9000) == 1 && fa.render_proxies.count(9000) == 0);
    const float cx = fa.camera_position.x * (1.0f - alpha) + fb.camera_position.x * alpha;
    EXPECT(r.camera_position.x == cx);
    EXPECT(r.view_matrix[12] == fa.view_matrix[12] * (1.0f - alpha) + fb.view_matrix[12] * alpha);
}

// ---- against the reference's own systems (oracle/_ref/libastre_ref.so: its sources compiled unmodified) ----------
#include <dlfcn.h>
namespace ref {
using u64 = uint64_t;
void *lib = nullptr;
void *(*world_create)();
void (*world_destroy)(void *);
int (*add_transforms)(void *, uint32_t, const u64 *, const uint8_t *, const float *, const float *, const float *);
int (*set_trs)(void *, uint32_t, const u64 *, const float *, const float *, const float *);
int (*add_visual)(void *, u64, int, const char *, const char *, const float *);
int (*add_camera)(void *, u64, float, float, float, float);
int (*register_vertex_buffer)(void *, const char *, u64);
int (*register_shader)(void *, const char *, u64);
int (*run_transform)(void *);
int (*get_transforms)(void *, uint32_t, const u64 *, float *, float *, float *, float *);
int (*run_camera)(void *, u64, int, float *, float *, float *);
int (*run_visual)(void *, int);
u64 (*proxy_count)(void *, int);
int (*get_proxy)(void *, int, u64, u64 *, float *);
template <class F> bool sym(F &f, const char *name) { f = reinterpret_cast<F>(dlsym(lib, name)); return f != nullptr; }
bool load(const char *argv0)
{
    std::string dir(argv0);
    const size_t slash = dir.rfind('/');
    dir = slash == std::string::npos ? "." : dir.substr(0, slash);
    lib = dlopen((dir + "/../../../oracle/_ref/libastre_ref.so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) return false;
    return sym(world_create, "ref_world_create") && sym(world_destroy, "ref_world_destroy") && sym(add_transforms, "ref_add_transforms") &&
           sym(set_trs, "ref_set_trs") && sym(add_visual, "ref_add_visual") && sym(add_camera, "ref_add_camera") &&
           sym(register_vertex_buffer, "ref_register_vertex_buffer") && sym(register_shader, "ref_register_shader") &&
           sym(run_transform, "ref_run_transform") && sym(get_transforms, "ref_get_transforms") && sym(run_camera, "ref_run_camera") &&
           sym(run_visual, "ref_run_visual") && sym(proxy_count, "ref_proxy_count") && sym(get_proxy, "ref_get_proxy");
}
}  // namespace ref

// The same world in the glue (GPU) and in the reference's Registry; TransformSystem, CameraSystem and VisualSystem on
// both sides.  Every field the reference's systems write must come out identical; the glue's proxies are the
// reference's proxies restricted to the entities that survive a view (the documented narrowing).
static void test_against_reference_build(uint32_t n)
{
    void *rw = ref::world_create();
    EXPECT(rw != nullptr);
    if (!rw) return;
    World w(n + 1);
    const char *meshes[] = {"cube_prefab", "icosphere2_prefab", "cone_prefab", "dome_prefab", "quad_prefab", "ghost_mesh"};
    const char *all_meshes[] = {"cube_prefab", "quad_prefab", "cone_prefab", "cylinder_prefab", "icosphere1_prefab", "icosphere2_prefab",
                                "icosphere3_prefab", "dome_prefab"};
    for (uint64_t i = 0; i < 8; ++i) ref::register_vertex_buffer(rw, all_meshes[i], i);
    ref::register_shader(rw, "deferred_shader", 0);
    ref::register_shader(rw, "basic_shader", 1);
    std::mt19937 rng(99);
    std::uniform_real_distribution<float> P(-60.0f, 60.0f), Q(-1.0f, 1.0f), S(0.25f, 3.0f);
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t id = 1000 + i;
        TransformComponent t;
        float p[3] = {0, 0, 0}, q[4] = {0, 0, 0, 0}, s[3] = {0, 0, 0};
        uint8_t has = 0;
        if (i % 7) { p[0] = P(rng); p[1] = P(rng) * 0.2f; p[2] = P(rng) - 50.0f; t.position = Vec3{p[0], p[1], p[2]}; has |= 1; }
        if (i % 3) {
            q[0] = Q(rng); q[1] = Q(rng); q[2] = Q(rng); q[3] = Q(rng);
            if (i % 8) { const float l = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (float &c : q) c /= l; }   // some stay un-normalised
            t.rotation = Quat{q[0], q[1], q[2], q[3]}; has |= 2;
        }
        if (i % 5) { s[0] = S(rng); s[1] = S(rng); s[2] = (i % 9) ? S(rng) : -S(rng); t.scale = Vec3{s[0], s[1], s[2]}; has |= 4; }
        VisualComponent v;
        v.vertex_buffer_name = meshes[i % 6];
        v.shader_name = (i % 4) ? "deferred_shader" : ((i % 8) ? "basic_shader" : "ghost_shader");
        v.visible = (i % 11) != 0;
        float color[4] = {0.1f * (float)(i % 10), 0.5f, 0.25f, 1.0f};
        if (i % 2) v.color = Vec4{color[0], color[1], color[2], color[3]};
        const bool has_visual = (i % 13) != 0;
        w.mirror.spawn(id, t, has_visual ? &v : nullptr, &w.names);
        EXPECT(ref::add_transforms(rw, 1, &id, &has, p, q, s) == 0);
        if (has_visual) EXPECT(ref::add_visual(rw, id, v.visible ? 1 : 0, v.vertex_buffer_name.c_str(), v.shader_name.c_str(), (i % 2) ? color : nullptr) == 0);
    }
    {   // the camera entity of level_0.json
        const uint64_t id = 5;
        const uint8_t has = 1;
        const float p[3] = {0, 5, 15}, q[4] = {0, 0, 0, 0}, s[3] = {0, 0, 0};
        TransformComponent cam; cam.position = Vec3{0, 5, 15};
        w.mirror.spawn(id, cam, nullptr, nullptr);
        ref::add_transforms(rw, 1, &id, &has, p, q, s);
        ref::add_camera(rw, id, 60.0f, 0.1f, 1000.0f, 1.7f);
        w.cameras.active_camera_entity = 5;
    }
    Frame frame;
    w.transforms.run(0.1f);
    EXPECT(ref::run_transform(rw) == 0);
    bool mats = true, basis = true;
    for (uint32_t s = 0; s < w.mirror.slotCount(); ++s) {
        const uint64_t id = w.mirror.entityAt(s);
        TransformComponent out;
        w.transforms.read(id, out);
        float m[16], f[3], u[3], r[3];
        if (ref::get_transforms(rw, 1, &id, m, f, u, r) != 0) { mats = false; break; }
        mats &= same_bits(m, out.transform_matrix.data(), 16);
        basis &= same_bits(f, &out.forward.x, 3) && same_bits(u, &out.up.x, 3) && same_bits(r, &out.right.x, 3);
    }
    EXPECT(mats);
    EXPECT(basis);
    w.cameras.run(0.1f, frame);
    float view[16], proj[16], cpos[3];
    EXPECT(ref::run_camera(rw, 5, 0, view, proj, cpos) == 0);
    EXPECT(same_bits(view, frame.view_matrix.data(), 16) && same_bits(proj, frame.proj_matrix.data(), 16));
    EXPECT(same_bits(cpos, &frame.camera_position.x, 3));
    w.visuals.run(0.1f, frame);
    EXPECT(ref::run_visual(rw, 0) == 0);
    EXPECT(frame.render_proxies.size() > n / 50 && frame.render_proxies.size() <= ref::proxy_count(rw, 0));
    bool proxies_ok = true;
    for (const auto &[id, p] : frame.render_proxies) {
        uint64_t u[5];
        float f[30];
        if (ref::get_proxy(rw, 0, id, u, f) != 0) { proxies_ok = false; break; }      // the reference has no proxy for it
        proxies_ok &= (u[0] != 0) == p.visible && u[1] == p.phases && u[2] == p.vertex_buffer && u[3] == p.shader && (u[4] != 0) == p.useTexture;
        proxies_ok &= same_bits(f + 0, &p.position.x, 3) && same_bits(f + 3, &p.rotation.x, 4) && same_bits(f + 7, &p.scale.x, 3);
        proxies_ok &= same_bits(f + 10, p.uModel.data(), 16) && same_bits(f + 26, &p.uColor.x, 4);
        if (!proxies_ok) { std::printf("proxy %llu differs from the reference build\n", (unsigned long long)id); break; }
    }
    EXPECT(proxies_ok);
    ref::world_destroy(rw);
}

int main(int argc, char **argv)
{
    const bool cpu_only = argc > 1 && std::string(argv[1]) == "--cpu";
    test_mirror_cpu();
    if (!cpu_only) {
        test_level0_like_world();
        test_random_world(20000);
        test_frame_interpolation(5000);
        if (ref::load(argv[0])) test_against_reference_build(20000);
        else std::printf("note: oracle/_ref/libastre_ref.so not found -- comparison with the reference build skipped\n");
        const long long damaged = astre_gpu_debug_guard_check();     // ASTRE_GUARD=1: canaries around every device buffer
        EXPECT(damaged <= 0);
        if (damaged >= 0) std::printf("device buffer guards: %lld damaged bytes\n", damaged);
    }
    std::printf("%s: %d checks, %d failures%s\n", g_failures ? "FAILED" : "PASSED", g_checks, g_failures, cpu_only ? " (cpu-only subset)" : "");
    return g_failures ? 1 : 0;
}
