// ref_optix_bench.cpp — times the REFERENCE's own ray-tracing path (LuisaCompute's `cuda` backend: OptiX, which is driver
// software traversal on B200) on the files tools/dump_scene.py writes, and dumps its hits so that the same inputs can
// be compared record by record with this repository's results.  Written against the reference's public API only
// (Context / Device / Stream / Buffer / Mesh / Accel + a DSL kernel that calls accel.intersect / intersect_any, the
// shape of src/tests/test_rtx.cpp:52-99); nothing of this repository is linked or loaded.  It is a measurement tool
// and a source of reference-run golden data, not part of the product.
//
// usage: ref_optix_bench <dir>     (reads verts.bin tris.bin rays_*.bin, writes ref_hits_*.bin ref_any_*.bin ref_result.json)
#include <cstdio>
#include <string>
#include <vector>

#include <luisa/core/clock.h>
#include <luisa/core/logging.h>
#include <luisa/runtime/context.h>
#include <luisa/runtime/device.h>
#include <luisa/runtime/stream.h>
#include <luisa/runtime/buffer.h>
#include <luisa/runtime/shader.h>
#include <luisa/runtime/rtx/accel.h>
#include <luisa/runtime/rtx/mesh.h>
#include <luisa/dsl/syntax.h>
#include <luisa/dsl/rtx/accel.h>
#include <luisa/dsl/rtx/ray.h>
#include <luisa/dsl/rtx/hit.h>
#include <luisa/dsl/rtx/triangle.h>

using namespace luisa;
using namespace luisa::compute;

template<class T>
static std::vector<T> read_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (f == nullptr) return {};
    fseek(f, 0, SEEK_END);
    const auto bytes = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v(bytes / sizeof(T));
    if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) v.clear();
    fclose(f);
    return v;
}
static void write_file(const std::string &path, const void *p, size_t bytes) {
    FILE *f = fopen(path.c_str(), "wb");
    fwrite(p, 1, bytes, f);
    fclose(f);
}

int main(int argc, char *argv[]) {
    log_level_info();
    if (argc < 2) {
        LUISA_INFO("usage: {} <dir>", argv[0]);
        return 1;
    }
    const std::string dir = argv[1];
    auto verts = read_file<float4>(dir + "/verts.bin");
    auto tris = read_file<Triangle>(dir + "/tris.bin");
    LUISA_ASSERT(!verts.empty() && !tris.empty(), "missing verts.bin / tris.bin in {}", dir);
    Context context{argv[0]};
    Device device = context.create_device("cuda");
    Stream stream = device.create_stream(StreamTag::COMPUTE);
    Buffer<float4> vertex_buffer = device.create_buffer<float4>(verts.size());
    Buffer<Triangle> triangle_buffer = device.create_buffer<Triangle>(tris.size());
    stream << vertex_buffer.copy_from(verts.data()) << triangle_buffer.copy_from(tris.data()) << synchronize();
    Mesh mesh = device.create_mesh(vertex_buffer, triangle_buffer);// default AccelOption: FAST_TRACE, compaction on
    Accel accel = device.create_accel();
    accel.emplace_back(mesh, make_float4x4(1.f));
    stream << mesh.build() << accel.build() << synchronize();// includes OptiX start-up
    double build_ms = 1e30;
    for (int i = 0; i < 5; i++) {
        Clock clk;
        stream << mesh.build(AccelBuildRequest::FORCE_BUILD) << synchronize();
        build_ms = std::min(build_ms, clk.toc());
    }
    stream << accel.build(AccelBuildRequest::FORCE_BUILD) << synchronize();

    Kernel1D closest_kernel = [](BufferVar<Ray> rays, BufferVar<TriangleHit> hits, AccelVar accel) noexcept {
        auto i = dispatch_id().x;
        Var<Ray> ray = rays.read(i);
        Var<TriangleHit> hit = accel.intersect(ray, {});
        hits.write(i, hit);
    };
    Kernel1D any_kernel = [](BufferVar<Ray> rays, BufferVar<uint> occluded, AccelVar accel) noexcept {
        auto i = dispatch_id().x;
        Var<Ray> ray = rays.read(i);
        occluded.write(i, ite(accel.intersect_any(ray, {}), 1u, 0u));
    };
    auto closest = device.compile(closest_kernel);
    auto any = device.compile(any_kernel);

    std::string json = "{\"backend\": \"cuda (OptiX)\", \"triangles\": " + std::to_string(tris.size()) + ", \"mesh_build_ms\": " + std::to_string(build_ms);
    for (const char *name : {"primary", "bounce"}) {
        auto rays = read_file<Ray>(dir + "/rays_" + name + ".bin");
        if (rays.empty()) continue;
        const auto n = rays.size();
        Buffer<Ray> ray_buffer = device.create_buffer<Ray>(n);
        Buffer<TriangleHit> hit_buffer = device.create_buffer<TriangleHit>(n);
        Buffer<uint> any_buffer = device.create_buffer<uint>(n);
        stream << ray_buffer.copy_from(rays.data()) << synchronize();
        double closest_ms = 1e30, any_ms = 1e30;
        for (int i = 0; i < 13; i++) {// 3 warm-ups + best of 10, host clock around dispatch + synchronize
            Clock clk;
            stream << closest(ray_buffer, hit_buffer, accel).dispatch((uint)n) << synchronize();
            if (i >= 3) closest_ms = std::min(closest_ms, clk.toc());
        }
        for (int i = 0; i < 8; i++) {
            Clock clk;
            stream << any(ray_buffer, any_buffer, accel).dispatch((uint)n) << synchronize();
            if (i >= 3) any_ms = std::min(any_ms, clk.toc());
        }
        std::vector<TriangleHit> hits(n);
        std::vector<uint> occ(n);
        stream << hit_buffer.copy_to(hits.data()) << any_buffer.copy_to(occ.data()) << synchronize();
        write_file(dir + "/ref_hits_" + name + ".bin", hits.data(), n * sizeof(TriangleHit));
        write_file(dir + "/ref_any_" + name + ".bin", occ.data(), n * sizeof(uint));
        json += std::string(", \"") + name + "\": {\"rays\": " + std::to_string(n) + ", \"closest_ms\": " + std::to_string(closest_ms) +
                ", \"closest_mrays_s\": " + std::to_string((double)n / closest_ms / 1e3) + ", \"any_ms\": " + std::to_string(any_ms) +
                ", \"any_mrays_s\": " + std::to_string((double)n / any_ms / 1e3) + "}";
    }
    json += "}";
    write_file(dir + "/ref_result.json", json.data(), json.size());
    printf("%s\n", json.c_str());
    return 0;
}
