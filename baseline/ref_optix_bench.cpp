// ref_optix_bench.cpp — times the REFERENCE's own ray-tracing path (LuisaCompute's `cuda` backend: OptiX, which is driver
// software traversal on B200) on the files tools/dump_scene.py writes, and dumps its hits so that the same inputs can
// be compared record by record with this repository's results.  Written against the reference's public API only
// (Context / Device / Stream / Buffer / Mesh / Accel + a DSL kernel that calls accel.intersect / intersect_any, the
// shape of src/tests/test_rtx.cpp:52-99); nothing of this repository is linked or loaded.  It is a measurement tool
// and a source of reference-run golden data, not part of the product.
//
// usage: ref_optix_bench <dir>              reads verts.bin tris.bin rays_*.bin, writes ref_hits_*.bin ref_any_*.bin ref_result.json
//        ref_optix_bench <dir> scene <mask> reads scene.bin (meshes, one optional AABB set, instances with transform / mask /
//                                           opacity / user id) and rays_scene.bin; traces with visibility mask <mask>:
//                                           intersect, intersect_any and a RayQuery (query_all) whose triangle handler commits
//                                           unless prim % 3 == 0 and whose procedural handler intersects the sphere inscribed
//                                           in the AABB; writes ref_scene_hits.bin ref_scene_any.bin ref_scene_query.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
#include <luisa/runtime/rtx/procedural_primitive.h>
#include <luisa/dsl/syntax.h>
#include <luisa/dsl/rtx/accel.h>
#include <luisa/dsl/rtx/ray.h>
#include <luisa/dsl/rtx/hit.h>
#include <luisa/dsl/rtx/triangle.h>
#include <luisa/dsl/rtx/aabb.h>
#include <luisa/dsl/rtx/ray_query.h>
#include <luisa/dsl/sugar.h>

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

// "cuda" = the reference's OptiX backend; LC_BACKEND=b200 selects the hybrid (same backend, this repository's BVH and traversal)
static const char *backend_name() noexcept {
    auto e = getenv("LC_BACKEND");
    return e != nullptr ? e : "cuda";
}

struct InstanceDesc {
    uint32_t blas, mask, opaque, user_id;
    float affine[12];
};

static int run_scene(const std::string &dir, uint32_t ray_mask, Device &device, Stream &stream) {
    auto raw = read_file<uint8_t>(dir + "/scene.bin");
    LUISA_ASSERT(!raw.empty(), "missing scene.bin in {}", dir);
    size_t pos = 0;
    auto u32 = [&] {
        uint32_t v;
        memcpy(&v, raw.data() + pos, 4);
        pos += 4;
        return v;
    };
    std::vector<Buffer<float4>> vbs;
    std::vector<Buffer<Triangle>> tbs;
    std::vector<Mesh> meshes;
    const uint32_t n_meshes = u32();
    for (uint32_t i = 0; i < n_meshes; i++) {
        const uint32_t nv = u32();
        auto &vb = vbs.emplace_back(device.create_buffer<float4>(nv));
        stream << vb.copy_from(raw.data() + pos) << synchronize();
        pos += (size_t)nv * 16;
        const uint32_t nt = u32();
        auto &tb = tbs.emplace_back(device.create_buffer<Triangle>(nt));
        stream << tb.copy_from(raw.data() + pos) << synchronize();
        pos += (size_t)nt * 12;
    }
    for (uint32_t i = 0; i < n_meshes; i++) meshes.emplace_back(device.create_mesh(vbs[i], tbs[i]));
    const uint32_t n_procs = u32();
    LUISA_ASSERT(n_procs <= 1u, "one AABB set at most");
    Buffer<AABB> aabb_buffer = device.create_buffer<AABB>(1u);
    ProceduralPrimitive spheres;
    if (n_procs == 1u) {
        const uint32_t nb = u32();
        aabb_buffer = device.create_buffer<AABB>(nb);
        stream << aabb_buffer.copy_from(raw.data() + pos) << synchronize();
        pos += (size_t)nb * 24;
        spheres = device.create_procedural_primitive(aabb_buffer);
    }
    const uint32_t n_inst = u32();
    std::vector<InstanceDesc> inst(n_inst);
    memcpy(inst.data(), raw.data() + pos, (size_t)n_inst * sizeof(InstanceDesc));
    Accel accel = device.create_accel();
    for (auto &d : inst) {
        float4x4 m = make_float4x4(1.f);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) m[c][r] = d.affine[4 * r + c];
        if (d.blas & 0x80000000u) accel.emplace_back(spheres, m, (uint8_t)d.mask, d.user_id);
        else accel.emplace_back(meshes[d.blas], m, (uint8_t)d.mask, d.opaque != 0u, d.user_id);
    }
    for (auto &m : meshes) stream << m.build();
    if (n_procs == 1u) stream << spheres.build();
    stream << accel.build() << synchronize();
    double tlas_ms = 1e30;
    for (int i = 0; i < 5; i++) {
        Clock clk;
        stream << accel.build(AccelBuildRequest::FORCE_BUILD) << synchronize();
        tlas_ms = std::min(tlas_ms, clk.toc());
    }

    Kernel1D closest_kernel = [](BufferVar<Ray> rays, BufferVar<TriangleHit> hits, AccelVar accel, UInt mask) noexcept {
        auto i = dispatch_id().x;
        AccelTraceOptions options;
        options.visibility_mask = mask;
        hits.write(i, accel.intersect(rays.read(i), options));
    };
    Kernel1D any_kernel = [](BufferVar<Ray> rays, BufferVar<uint> occluded, AccelVar accel, UInt mask) noexcept {
        auto i = dispatch_id().x;
        AccelTraceOptions options;
        options.visibility_mask = mask;
        occluded.write(i, ite(accel.intersect_any(rays.read(i), options), 1u, 0u));
    };
    Kernel1D query_kernel = [](BufferVar<Ray> rays, BufferVar<CommittedHit> hits, AccelVar accel, BufferVar<AABB> aabbs, UInt mask) noexcept {
        auto i = dispatch_id().x;
        AccelTraceOptions options;
        options.visibility_mask = mask;
        Var<Ray> ray = rays.read(i);
        Var<CommittedHit> hit = accel.traverse(ray, options)
                                    .on_surface_candidate([&](SurfaceCandidate &candidate) noexcept {
                                        Var<TriangleHit> h = candidate.hit();
                                        $if (h.prim % 3u != 0u) { candidate.commit(); };
                                    })
                                    .on_procedural_candidate([&](ProceduralCandidate &candidate) noexcept {
                                        Var<ProceduralHit> h = candidate.hit();
                                        Var<Ray> r = candidate.ray();// world ray, t_max = committed distance so far
                                        Var<AABB> box = aabbs.read(h.prim);
                                        Float4x4 m = accel.instance_transform(h.inst);
                                        Float3 c_obj = (box->min() + box->max()) * .5f;
                                        Float r_obj = (box->max().x - box->min().x) * .5f;
                                        Float3 centre = (m * make_float4(c_obj, 1.f)).xyz();
                                        Float radius = r_obj * length(m[0].xyz());
                                        Float3 L = r->origin() - centre;
                                        Float3 d = r->direction();
                                        Float a = dot(d, d), hb = dot(L, d), cc = dot(L, L) - radius * radius;
                                        Float disc = hb * hb - a * cc;
                                        $if (disc >= 0.f) {
                                            Float sq = sqrt(disc);
                                            Float t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
                                            $if (t0 >= r->t_min() & t0 < r->t_max()) { candidate.commit(t0); }
                                            $elif (t1 >= r->t_min() & t1 < r->t_max()) { candidate.commit(t1); };
                                        };
                                    })
                                    .trace();
        hits.write(i, hit);
    };
    auto closest = device.compile(closest_kernel);
    auto any = device.compile(any_kernel);
    auto query = device.compile(query_kernel);
    auto rays = read_file<Ray>(dir + "/rays_scene.bin");
    LUISA_ASSERT(!rays.empty(), "missing rays_scene.bin");
    const auto n = rays.size();
    Buffer<Ray> ray_buffer = device.create_buffer<Ray>(n);
    Buffer<TriangleHit> hit_buffer = device.create_buffer<TriangleHit>(n);
    Buffer<CommittedHit> query_buffer = device.create_buffer<CommittedHit>(n);
    Buffer<uint> any_buffer = device.create_buffer<uint>(n);
    std::vector<TriangleHit> hits(n);
    std::vector<CommittedHit> qhits(n);
    std::vector<uint> occ(n);
    double closest_ms = 1e30;
    stream << ray_buffer.copy_from(rays.data()) << synchronize();
    for (int i = 0; i < 8; i++) {
        Clock clk;
        stream << closest(ray_buffer, hit_buffer, accel, ray_mask).dispatch((uint)n) << synchronize();
        if (i >= 3) closest_ms = std::min(closest_ms, clk.toc());
    }
    stream << any(ray_buffer, any_buffer, accel, ray_mask).dispatch((uint)n)
           << query(ray_buffer, query_buffer, accel, aabb_buffer, ray_mask).dispatch((uint)n)
           << hit_buffer.copy_to(hits.data()) << any_buffer.copy_to(occ.data()) << query_buffer.copy_to(qhits.data()) << synchronize();
    write_file(dir + "/ref_scene_hits.bin", hits.data(), n * sizeof(TriangleHit));
    write_file(dir + "/ref_scene_any.bin", occ.data(), n * sizeof(uint));
    write_file(dir + "/ref_scene_query.bin", qhits.data(), n * sizeof(CommittedHit));
    std::string json = "{\"backend\": \"" + std::string{backend_name()} + "\", \"instances\": " + std::to_string(n_inst) + ", \"rays\": " + std::to_string(n) +
                       ", \"tlas_build_ms\": " + std::to_string(tlas_ms) + ", \"closest_ms\": " + std::to_string(closest_ms) + ", \"closest_mrays_s\": " + std::to_string((double)n / closest_ms / 1e3) + "}";
    write_file(dir + "/ref_scene_result.json", json.data(), json.size());
    printf("%s\n", json.c_str());
    return 0;
}

int main(int argc, char *argv[]) {
    log_level_info();
    if (argc >= 3 && std::string{argv[2]} == "scene") {
        Context context{argv[0]};
        Device device = context.create_device(backend_name());
        Stream stream = device.create_stream(StreamTag::COMPUTE);
        return run_scene(argv[1], argc > 3 ? (uint32_t)strtoul(argv[3], nullptr, 0) : 0xffu, device, stream);
    }
    if (argc < 2) {
        LUISA_INFO("usage: {} <dir>", argv[0]);
        return 1;
    }
    const std::string dir = argv[1];
    auto verts = read_file<float4>(dir + "/verts.bin");
    auto tris = read_file<Triangle>(dir + "/tris.bin");
    LUISA_ASSERT(!verts.empty() && !tris.empty(), "missing verts.bin / tris.bin in {}", dir);
    Context context{argv[0]};
    Device device = context.create_device(backend_name());
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

    std::string json = "{\"backend\": \"" + std::string{backend_name()} + "\", \"triangles\": " + std::to_string(tris.size()) + ", \"mesh_build_ms\": " + std::to_string(build_ms);
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
