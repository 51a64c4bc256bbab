// test_plugin_rtx.cpp — integration driver for the `b200rt` backend plugin, written against the reference's own runtime
// API (Context / Device / Stream / Buffer / Mesh / ProceduralPrimitive / Accel, include/luisa/runtime/...) and linked
// against the reference's unmodified liblc-runtime.  It follows src/tests/test_rtx.cpp (triangle + two instanced
// transforms, :24-28, :101-107; the per-frame loop vertex edit -> mesh.build() -> set_transform_on_update ->
// accel.build(), :122-139) and src/tests/test_procedural.cpp (AABB spheres, :29-40), with the DSL kernels replaced by
// b200rt::TraceCommand.  Every case is dumped (geometry, instance records as the test set them, rays, results) to
// `out_path` so that tests/test_gpu_plugin.py can check the results against the oracle on identical inputs.
//
// usage: test_plugin_rtx <out_path> [resolution]
#include <cstdio>
#include <vector>

#include <luisa/core/logging.h>
#include <luisa/core/mathematics.h>
#include <luisa/runtime/context.h>
#include <luisa/runtime/device.h>
#include <luisa/runtime/stream.h>
#include <luisa/runtime/buffer.h>
#include <luisa/runtime/rtx/accel.h>
#include <luisa/runtime/rtx/mesh.h>
#include <luisa/runtime/rtx/procedural_primitive.h>
#include <luisa/dsl/rtx/aabb.h>
#include <luisa/dsl/rtx/triangle.h>

#include "b200rt_ext.h"

using namespace luisa;
using namespace luisa::compute;
using b200rt::TraceCommand;

struct RayRec {
    float o[3], tmin, d[3], tmax;
};
static_assert(sizeof(RayRec) == 32);

struct InstanceDesc {// what the test asked the runtime for, in Modification terms (command.h:534-552)
    uint32_t blas, mask, opaque, user_id;
    float affine[12];
};

static void affine_of(float4x4 m, float *a) noexcept {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) a[4 * r + c] = m[c][r];
}

struct Dump {
    FILE *f;
    void u32(uint32_t v) noexcept { fwrite(&v, 4, 1, f); }
    void bytes(const void *p, size_t n) noexcept { fwrite(p, 1, n, f); }
    // one case: meshes (float4-strided vertices + index triples), procedural AABB sets, instances, rays, closest hits,
    // any-hit flags, query hits (count 0 = not run)
    void scene(const char *name, const std::vector<std::vector<float4>> &verts, const std::vector<std::vector<uint32_t>> &tris,
               const std::vector<std::vector<AABB>> &boxes, const std::vector<InstanceDesc> &inst, const std::vector<RayRec> &rays,
               uint32_t ray_mask, const std::vector<uint32_t> &closest, const std::vector<uint32_t> &any, const std::vector<uint32_t> &query) noexcept {
        char tag[32] = {};
        snprintf(tag, sizeof(tag), "%s", name);
        bytes(tag, 32);
        u32((uint32_t)verts.size());
        for (size_t i = 0; i < verts.size(); i++) {
            u32((uint32_t)verts[i].size()), bytes(verts[i].data(), verts[i].size() * 16);
            u32((uint32_t)tris[i].size() / 3u), bytes(tris[i].data(), tris[i].size() * 4);
        }
        u32((uint32_t)boxes.size());
        for (auto &b : boxes) u32((uint32_t)b.size()), bytes(b.data(), b.size() * sizeof(AABB));
        u32((uint32_t)inst.size()), bytes(inst.data(), inst.size() * sizeof(InstanceDesc));
        u32((uint32_t)rays.size()), u32(ray_mask), bytes(rays.data(), rays.size() * 32);
        u32((uint32_t)closest.size()), bytes(closest.data(), closest.size() * 4);
        u32((uint32_t)any.size()), bytes(any.data(), any.size() * 4);
        u32((uint32_t)query.size()), bytes(query.data(), query.size() * 4);
    }
};

int main(int argc, char *argv[]) {
    log_level_info();
    if (argc < 2) {
        LUISA_INFO("usage: {} <out_path> [resolution]", argv[0]);
        return 1;
    }
    const uint res = argc > 2 ? (uint)atoi(argv[2]) : 256u;
    Context context{argv[0]};
    bool found = false;
    for (auto &&b : context.installed_backends()) found |= b == "b200rt";
    LUISA_ASSERT(found, "backend b200rt is not installed next to {}", argv[0]);
    for (auto &&n : context.backend_device_names("b200rt")) LUISA_INFO("b200rt device: {}", n);
    Device device = context.create_device("b200rt");
    Stream stream = device.create_stream();
    Dump dump{fopen(argv[1], "wb")};
    LUISA_ASSERT(dump.f != nullptr, "cannot open {}", argv[1]);

    // ---------------------------------------------------------------- case 1: test_rtx.cpp, static + per-frame loop
    std::vector<float4> vertices{make_float4(-0.5f, -0.5f, 0.0f, 0.f), make_float4(0.5f, -0.5f, 0.0f, 0.f), make_float4(0.0f, 0.5f, 0.0f, 0.f)};
    std::vector<uint32_t> indices{0u, 1u, 2u};
    Buffer<float4> vertex_buffer = device.create_buffer<float4>(3u);
    Buffer<Triangle> triangle_buffer = device.create_buffer<Triangle>(1u);
    stream << vertex_buffer.copy_from(vertices.data()) << triangle_buffer.copy_from(indices.data());
    AccelOption dynamic_option;
    dynamic_option.allow_update = true;
    Accel accel = device.create_accel(dynamic_option);
    Mesh mesh = device.create_mesh(vertex_buffer, triangle_buffer, dynamic_option);
    const float4x4 m0 = scaling(1.5f);
    float4x4 m1 = translation(make_float3(-0.25f, 0.0f, 0.1f)) * rotation(make_float3(0.0f, 0.0f, 1.0f), 0.5f);
    accel.emplace_back(mesh, m0);
    accel.emplace_back(mesh, m1);
    stream << mesh.build() << accel.build();

    const size_t n = (size_t)res * res;
    std::vector<RayRec> rays(n);
    for (uint y = 0; y < res; y++)
        for (uint x = 0; x < res; x++) {
            RayRec &r = rays[(size_t)y * res + x];
            r.o[0] = ((float)x + 0.5f) / (float)res * 2.f - 1.f, r.o[1] = -(((float)y + 0.5f) / (float)res * 2.f - 1.f), r.o[2] = 1.f;
            r.d[0] = 0.f, r.d[1] = 0.f, r.d[2] = -1.f, r.tmin = 0.f, r.tmax = std::numeric_limits<float>::max();
        }
    Buffer<float4> ray_buffer = device.create_buffer<float4>(2u * n);
    Buffer<uint> hit_buffer = device.create_buffer<uint>(6u * n);
    Buffer<uint> any_buffer = device.create_buffer<uint>(n);
    std::vector<uint32_t> closest(6u * n), any(n), none;
    auto trace_both = [&](Accel &a, uint32_t mask) {
        stream << luisa::make_unique<TraceCommand>(TraceCommand::Mode::CLOSEST, a.handle(), ray_buffer.handle(), 0u, hit_buffer.handle(), 0u, n, mask)
               << luisa::make_unique<TraceCommand>(TraceCommand::Mode::ANY, a.handle(), ray_buffer.handle(), 0u, any_buffer.handle(), 0u, n, mask)
               << hit_buffer.copy_to(closest.data()) << any_buffer.copy_to(any.data()) << synchronize();
    };
    stream << ray_buffer.copy_from(rays.data());
    trace_both(accel, 0xffu);
    std::vector<InstanceDesc> inst(2);
    inst[0] = {0u, 0xffu, 1u, 0u, {}}, inst[1] = {0u, 0xffu, 1u, 0u, {}};
    affine_of(m0, inst[0].affine), affine_of(m1, inst[1].affine);
    dump.scene("c1_static", {vertices}, {indices}, {}, inst, rays, 0xffu, closest, any, none);

    bool callback_ran = false;
    for (uint i = 0u; i < 4u; i++) {// the per-frame loop of test_rtx.cpp:122-139 (refit of the BLAS, then of the TLAS)
        const float t = (float)(i + 1u) * 0.2f;
        vertices[2].y = 0.5f - 0.2f * t;
        m1 = translation(make_float3(-0.25f + t * 0.15f, 0.0f, 0.1f)) * rotation(make_float3(0.0f, 0.0f, 1.0f), 0.5f + t * 0.5f);
        accel.set_transform_on_update(1u, m1);
        stream << vertex_buffer.copy_from(vertices.data()) << mesh.build() << accel.build();
        trace_both(accel, 0xffu);
        affine_of(m1, inst[1].affine);
        char name[32];
        snprintf(name, sizeof(name), "c1_frame%u", i);
        dump.scene(name, {vertices}, {indices}, {}, inst, rays, 0xffu, closest, any, none);
    }
    {// a third, non-opaque instance with a visibility mask, added without a rebuild request first (test_rtx.cpp:132-137)
        const float4x4 mm = translation(make_float3(0.0f, 0.0f, 0.3f)) * rotation(make_float3(0.0f, 0.0f, 1.0f), radians(180.0f));
        accel.emplace_back(mesh, mm, 0x0fu, true, 7u);
        stream << accel.update_instance_buffer() << accel.build() << [&callback_ran] { callback_ran = true; };
        trace_both(accel, 0x03u);
        LUISA_ASSERT(callback_ran, "stream callbacks must have run before synchronize() returns");
        inst.push_back({0u, 0x0fu, 1u, 7u, {}});
        affine_of(mm, inst[2].affine);
        dump.scene("c1_three_masked", {vertices}, {indices}, {}, inst, rays, 0x03u, closest, any, none);
    }

    // ---------------------------------------------------------------- case 2: procedural AABB spheres + a floor mesh, RayQuery
    {
        std::vector<AABB> boxes(48);
        uint32_t state = 0u;
        auto lcg = [&state] {
            state = 1664525u * state + 1013904223u;
            return (float)(state & 0x00ffffffu) * (1.0f / 16777216.0f);
        };
        for (auto &b : boxes) {// the pattern of test_procedural.cpp:29-40: random centres, radius 0.08
            const float cx = lcg() * 1.6f - 0.8f, cy = lcg() * 1.6f - 0.8f, cz = lcg() * 0.6f - 0.3f, r = 0.08f;
            b.packed_min = {cx - r, cy - r, cz - r}, b.packed_max = {cx + r, cy + r, cz + r};
        }
        Buffer<AABB> aabb_buffer = device.create_buffer<AABB>(boxes.size());
        std::vector<float4> floor_v{make_float4(-1.f, -1.f, -0.5f, 0.f), make_float4(1.f, -1.f, -0.5f, 0.f), make_float4(1.f, 1.f, -0.5f, 0.f), make_float4(-1.f, 1.f, -0.5f, 0.f)};
        std::vector<uint32_t> floor_t{0u, 1u, 2u, 0u, 2u, 3u};
        Buffer<float4> floor_vb = device.create_buffer<float4>(4u);
        Buffer<Triangle> floor_tb = device.create_buffer<Triangle>(2u);
        stream << aabb_buffer.copy_from(boxes.data()) << floor_vb.copy_from(floor_v.data()) << floor_tb.copy_from(floor_t.data());
        ProceduralPrimitive spheres = device.create_procedural_primitive(aabb_buffer);
        Mesh floor = device.create_mesh(floor_vb, floor_tb);
        Accel scene = device.create_accel();
        const float4x4 ms = translation(make_float3(0.05f, -0.02f, 0.0f));
        scene.emplace_back(floor, make_float4x4(1.f), 0xffu, true, 1u);
        scene.emplace_back(spheres, ms, 0xffu, 2u);
        stream << spheres.build() << floor.build() << scene.build();
        std::vector<uint32_t> query(6u * n);
        auto q = luisa::make_unique<TraceCommand>(TraceCommand::Mode::QUERY_ALL, scene.handle(), ray_buffer.handle(), 0u, hit_buffer.handle(), 0u, n, 0xffu);
        q->set_programs(0u /* commit every triangle candidate */, 0u /* ray-sphere inscribed in the AABB */);
        stream << std::move(q) << hit_buffer.copy_to(query.data()) << synchronize();
        trace_both(scene, 0xffu);
        std::vector<InstanceDesc> pin(2);
        pin[0] = {0u, 0xffu, 1u, 1u, {}}, pin[1] = {0x80000000u /* procedural set 0 */, 0xffu, 0u, 2u, {}};
        affine_of(make_float4x4(1.f), pin[0].affine), affine_of(ms, pin[1].affine);
        dump.scene("c5_procedural", {floor_v}, {floor_t}, {boxes}, pin, rays, 0xffu, closest, any, query);
    }
    fclose(dump.f);
    LUISA_INFO("b200rt plugin driver finished: {} rays per trace, results in {}", n, argv[1]);
    return 0;
}
