// c5_dsl_path_tracer.cpp — BASELINE config C5 as a LuisaCompute DSL program: the Cornell room (8 opaque meshes) plus ONE
// ProceduralPrimitive of AABB spheres intersected in an on_procedural_candidate handler, path traced at depth <= 10 with
// next-event estimation (any-hit shadow RayQuery), MIS and Russian roulette, tea / lcg sampler.  It is the DSL counterpart
// of this repository's tools/c5_path_tracer.cu (same integrator, same constants; the integrator follows the description of
// src/tests/test_path_tracing.cpp:145-256 in SURVEY.md §8d) and runs on whichever backend LC_BACKEND names: `cuda` (the
// reference's OptiX backend) or `b200` (plugin/hybrid).  Scene file: tools/ref_optix_c5.py.  Measurement tooling only.
//
// usage: c5_dsl_path_tracer <dir> [width height spp dispatches]   (reads c5_scene.bin; writes c5_<backend>.bin = float4 image, prints one JSON line)
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
#include <luisa/dsl/sugar.h>
#include <luisa/dsl/rtx/accel.h>
#include <luisa/dsl/rtx/ray.h>
#include <luisa/dsl/rtx/hit.h>
#include <luisa/dsl/rtx/triangle.h>
#include <luisa/dsl/rtx/aabb.h>
#include <luisa/dsl/rtx/ray_query.h>

using namespace luisa;
using namespace luisa::compute;

struct InstanceDesc {
    uint32_t blas, mask, opaque, user_id;
    float affine[12];
};

int main(int argc, char *argv[]) {
    log_level_info();
    if (argc < 2) return 1;
    const std::string dir = argv[1];
    const uint width = argc > 2 ? (uint)atoi(argv[2]) : 1920u, height = argc > 3 ? (uint)atoi(argv[3]) : 1080u;
    const uint spp = argc > 4 ? (uint)atoi(argv[4]) : 16u, dispatches = argc > 5 ? (uint)atoi(argv[5]) : 4u;
    const char *backend = getenv("LC_BACKEND") ? getenv("LC_BACKEND") : "cuda";
    // shadow rays end at d_light * C5_SHADOW_TMAX_SCALE (default 1: the integrator as specified).  Rays that end within rounding of
    // the light quad are where backends may legitimately differ (t <= t_max decided exactly vs in fp32); see tools/ref_optix_c5.py
    const bool flat_loop = getenv("C5_FLAT_LOOP") && atoi(getenv("C5_FLAT_LOOP")) != 0;
    const float shadow_scale = getenv("C5_SHADOW_TMAX_SCALE") ? (float)atof(getenv("C5_SHADOW_TMAX_SCALE")) : 1.0f;
    std::vector<uint8_t> raw;
    {
        FILE *f = fopen((dir + "/c5_scene.bin").c_str(), "rb");
        if (f == nullptr) LUISA_ERROR_WITH_LOCATION("missing c5_scene.bin in {}", dir);
        fseek(f, 0, SEEK_END);
        raw.resize((size_t)ftell(f));
        fseek(f, 0, SEEK_SET);
        const size_t got = fread(raw.data(), 1, raw.size(), f);// (no side effects inside LUISA_ASSERT: it may compile to nothing)
        fclose(f);
        if (got != raw.size()) LUISA_ERROR_WITH_LOCATION("short read of c5_scene.bin");
    }
    size_t pos = 0;
    auto u32 = [&] {
        uint32_t v;
        memcpy(&v, raw.data() + pos, 4);
        pos += 4;
        return v;
    };
    Context context{argv[0]};
    Device device = context.create_device(backend);
    Stream stream = device.create_stream(StreamTag::COMPUTE);
    // ---- scene: meshes, the sphere set, instances (identity transforms), per-instance albedo, flattened triangles for normals
    std::vector<Buffer<float4>> vbs;
    std::vector<Buffer<Triangle>> tbs;
    std::vector<Mesh> meshes;
    std::vector<float4> flat;        // three vertices per triangle, meshes concatenated in instance order
    std::vector<uint> first_of_mesh;
    const uint n_meshes = u32();
    LUISA_INFO("scene: {} bytes, {} meshes", raw.size(), n_meshes);
    for (uint i = 0; i < n_meshes; i++) {
        const uint nv = u32();
        std::vector<float4> v(nv);// (the file offsets are not 16-byte aligned: copy before touching float4s)
        memcpy(v.data(), raw.data() + pos, (size_t)nv * 16);
        auto &vb = vbs.emplace_back(device.create_buffer<float4>(nv));
        stream << vb.copy_from(v.data()) << synchronize();
        pos += (size_t)nv * 16;
        const uint nt = u32();
        std::vector<uint> t(3u * nt);
        memcpy(t.data(), raw.data() + pos, (size_t)nt * 12);
        auto &tb = tbs.emplace_back(device.create_buffer<Triangle>(nt));
        stream << tb.copy_from(t.data()) << synchronize();
        pos += (size_t)nt * 12;
        first_of_mesh.push_back((uint)flat.size() / 3u);
        for (uint k = 0; k < 3u * nt; k++) flat.push_back(v[t[k]]);
    }
    for (uint i = 0; i < n_meshes; i++) meshes.emplace_back(device.create_mesh(vbs[i], tbs[i]));
    const uint n_sets = u32();
    if (n_sets != 1u) LUISA_ERROR_WITH_LOCATION("one AABB set expected, got {}", n_sets);
    const uint n_boxes = u32();
    Buffer<AABB> aabb_buffer = device.create_buffer<AABB>(n_boxes);
    std::vector<AABB> box_data(n_boxes);
    memcpy(box_data.data(), raw.data() + pos, (size_t)n_boxes * 24);
    stream << aabb_buffer.copy_from(box_data.data()) << synchronize();
    pos += (size_t)n_boxes * 24;
    ProceduralPrimitive spheres = device.create_procedural_primitive(aabb_buffer);
    const uint n_inst = u32();
    std::vector<InstanceDesc> inst(n_inst);
    memcpy(inst.data(), raw.data() + pos, (size_t)n_inst * sizeof(InstanceDesc));
    pos += (size_t)n_inst * sizeof(InstanceDesc);
    std::vector<float4> albedo(n_inst);
    memcpy(albedo.data(), raw.data() + pos, (size_t)n_inst * 16);
    pos += (size_t)n_inst * 16;
    const uint light_inst = u32();
    std::vector<uint> tri_first(n_inst, 0u);
    Accel accel = device.create_accel();
    for (uint i = 0; i < n_inst; i++) {
        if (inst[i].blas & 0x80000000u) accel.emplace_back(spheres, make_float4x4(1.f), 0xffu, i);
        else accel.emplace_back(meshes[inst[i].blas], make_float4x4(1.f), 0xffu, true, i), tri_first[i] = first_of_mesh[inst[i].blas];
    }
    LUISA_INFO("{} instances, {} boxes, light instance {}", n_inst, n_boxes, light_inst);
    for (auto &m : meshes) stream << m.build();
    stream << spheres.build() << accel.build() << synchronize();
    LUISA_INFO("scene built");
    Buffer<float4> flat_buffer = device.create_buffer<float4>(flat.size());
    Buffer<uint> first_buffer = device.create_buffer<uint>(n_inst);
    Buffer<float4> albedo_buffer = device.create_buffer<float4>(n_inst);
    Buffer<float4> image = device.create_buffer<float4>((size_t)width * height);
    Buffer<uint> seeds = device.create_buffer<uint>((size_t)width * height);
    Buffer<uint> ray_count = device.create_buffer<uint>(1u);
    std::vector<float4> zeros((size_t)width * height, make_float4(0.f));
    std::vector<uint> zero_u((size_t)width * height, 0u);
    stream << flat_buffer.copy_from(flat.data()) << first_buffer.copy_from(tri_first.data()) << albedo_buffer.copy_from(albedo.data())
           << image.copy_from(zeros.data()) << seeds.copy_from(zero_u.data()) << ray_count.copy_from(zero_u.data()) << synchronize();

    Callable tea = [](UInt v0, UInt v1) noexcept {
        UInt s0 = def(0u);
        for (uint n = 0u; n < 4u; n++) {
            s0 += 0x9e3779b9u;
            v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5u) + 0xc8013ea4u);
            v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5u) + 0x7e95761eu);
        }
        return v0;
    };
    Callable lcg = [](UInt &state) noexcept {
        state = 1664525u * state + 1013904223u;
        return cast<float>(state & 0x00ffffffu) * (1.0f / 16777216.0f);
    };
    // ray / sphere inscribed in the AABB; commits the nearest root inside [t_min, t_max)
    auto sphere_handler = [&](BufferVar<AABB> &boxes) {
        return [&boxes](ProceduralCandidate &candidate) noexcept {
            Var<ProceduralHit> h = candidate.hit();
            Var<Ray> r = candidate.ray();
            Var<AABB> b = boxes.read(h.prim);
            Float3 c = (b->min() + b->max()) * .5f;
            Float rad = (b->max().x - b->min().x) * .5f;
            Float3 L = r->origin() - c;
            Float3 d = r->direction();
            Float a = dot(d, d), hb = dot(L, d), cc = dot(L, L) - rad * rad;
            Float disc = hb * hb - a * cc;
            $if (disc >= 0.f) {
                Float sq = sqrt(disc);
                Float t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
                $if (t0 >= r->t_min() & t0 < r->t_max()) { candidate.commit(t0); }
                $elif (t1 >= r->t_min() & t1 < r->t_max()) { candidate.commit(t1); };
            };
        };
    };
    Kernel2D render = [&](BufferVar<float4> image, BufferVar<uint> seeds, BufferVar<uint> ray_count, AccelVar accel, BufferVar<AABB> boxes,
                          BufferVar<float4> tri_verts, BufferVar<uint> tri_first, BufferVar<float4> albedos, UInt spp, UInt light) noexcept {
        set_block_size(16u, 16u, 1u);
        UInt2 coord = dispatch_id().xy();
        UInt2 size = dispatch_size().xy();
        UInt pix = coord.y * size.x + coord.x;
        UInt state = seeds.read(pix);
        $if (state == 0u) { state = tea(coord.x, coord.y); };
        Float frame = cast<float>(min(size.x, size.y));
        Float3 radiance = def(make_float3(0.f));
        UInt rays = def(0u);
        const float3 light_position = make_float3(-0.25f, 0.995f, -0.25f), light_u = make_float3(0.5f, 0.f, 0.f), light_v = make_float3(0.f, 0.f, 0.5f);
        const float3 light_emission = make_float3(17.f, 12.f, 4.f), light_normal = make_float3(0.f, -1.f, 0.f);
        const float light_area = 0.25f, inv_pi = 0.318309886f;
        auto handler = sphere_handler(boxes);
        Var<Ray> ray = make_ray(make_float3(0.f), make_float3(0.f, 0.f, -1.f), 0.f, 0.f);
        Float3 beta = def(make_float3(1.f));
        Float pdf_bsdf = def(0.f);
        auto camera_ray = [&] {
            Float rx = lcg(state), ry = lcg(state);
            Float px = (cast<float>(coord.x) + rx) / frame * 2.f - cast<float>(size.x) / frame;
            Float py = -((cast<float>(coord.y) + ry) / frame * 2.f - cast<float>(size.y) / frame);
            const float th = std::tan(0.5f * 38.0f * 0.01745329252f);
            ray = make_ray(make_float3(0.f, 0.f, 3.4f), normalize(make_float3(px * th, py * th, -1.f)), 0.f, std::numeric_limits<float>::max());
            beta = make_float3(1.f);
            pdf_bsdf = 0.f;
        };
        // one bounce; a `$break` leaves the innermost DSL loop of the caller with path_ended still true
        auto bounce = [&](UInt depth, Bool &path_ended) {
                Var<CommittedHit> hit = accel.traverse(ray, {}).on_procedural_candidate(handler).trace();
                rays += 1u;
                $if (hit->miss()) { $break; };
                Float3 o = ray->origin(), d = ray->direction();
                Float3 p = o + d * hit.committed_ray_t;
                Float3 n = def(make_float3(0.f, 1.f, 0.f));
                $if (hit->is_procedural()) {
                    Var<AABB> b = boxes.read(hit.prim);
                    n = normalize(p - (b->min() + b->max()) * .5f);
                }
                $else {
                    UInt base = 3u * (tri_first.read(hit.inst) + hit.prim);
                    Float3 p0 = tri_verts.read(base).xyz(), p1 = tri_verts.read(base + 1u).xyz(), p2 = tri_verts.read(base + 2u).xyz();
                    n = normalize(cross(p1 - p0, p2 - p0));
                };
                Float cos_wo = -dot(d, n);
                $if (!hit->is_procedural() & cos_wo < 0.f & hit.inst != light) {// two-sided walls
                    n = -n;
                    cos_wo = -cos_wo;
                };
                $if (cos_wo < 1e-4f) { $break; };
                $if (!hit->is_procedural() & hit.inst == light) {
                    $if (depth == 0u) { radiance += light_emission; }
                    $else {
                        Float pdf_light = dot(p - o, p - o) / (light_area * cos_wo);
                        radiance += beta * light_emission * (pdf_bsdf / max(pdf_bsdf + pdf_light, 1e-4f));
                    };
                    $break;
                };
                // next-event estimation
                Float ux_light = lcg(state), uy_light = lcg(state);
                Float3 p_light = light_position + ux_light * light_u + uy_light * light_v;
                Float3 pp = offset_ray_origin(p, n), pp_light = offset_ray_origin(p_light, light_normal);
                Float d_light = distance(pp, pp_light);
                Float3 wi_light = normalize(pp_light - pp);
                Var<Ray> shadow = make_ray(offset_ray_origin(pp, n), wi_light, 0.f, d_light * shadow_scale);
                Var<CommittedHit> occ = accel.traverse_any(shadow, {}).on_procedural_candidate(handler).trace();
                rays += 1u;
                Float cos_wi_light = dot(wi_light, n), cos_light = -dot(light_normal, wi_light);
                Float3 albedo = albedos.read(hit.inst).xyz();
                $if (occ->miss() & cos_wi_light > 1e-4f & cos_light > 1e-4f) {
                    Float pdf_light = (d_light * d_light) / (light_area * cos_light), pdf_b = cos_wi_light * inv_pi;
                    Float mis = pdf_light / max(pdf_light + pdf_b, 1e-4f);
                    radiance += beta * albedo * light_emission * (inv_pi * cos_wi_light * mis / max(pdf_light, 1e-4f));
                };
                // cosine-weighted bounce
                Float3 binormal = normalize(ite(abs(n.x) > abs(n.z), make_float3(-n.y, n.x, 0.f), make_float3(0.f, -n.z, n.y)));
                Float3 tangent = normalize(cross(binormal, n));
                Float ux = lcg(state), uy = lcg(state);
                Float r = sqrt(ux), phi = 6.283185307f * uy;
                Float3 wl = make_float3(r * cos(phi), r * sin(phi), sqrt(1.f - ux));
                ray = make_ray(pp, tangent * wl.x + binormal * wl.y + n * wl.z, 0.f, std::numeric_limits<float>::max());
                pdf_bsdf = abs(wl.z) * inv_pi;
                beta *= albedo;
                Float l = dot(make_float3(0.212671f, 0.715160f, 0.072169f), beta);
                $if (l == 0.f) { $break; };
                Float q = max(l, 0.05f);
                $if (lcg(state) >= q) { $break; };
                beta *= 1.f / q;
                path_ended = false;
        };
        if (!flat_loop) {// the loop nest of src/tests/test_path_tracing.cpp:165-256
            $for (s, spp) {
                camera_ray();
                $for (depth, 10u) {
                    Bool path_ended = def(true);
                    bounce(depth, path_ended);
                };
            };
        } else {// one flat loop over (sample, bounce): a lane whose path ended starts its next sample at once; same random sequence, same sums
            UInt s = def(0u), depth = def(0u);
            $while (s < spp) {
                $if (depth == 0u) { camera_ray(); };
                Bool path_ended = def(true);
                $loop {
                    bounce(depth, path_ended);
                    $break;
                };
                depth += 1u;
                $if (path_ended | depth == 10u) {
                    depth = 0u;
                    s += 1u;
                };
            };
        }
        seeds.write(pix, state);
        $if (any(isnan(radiance))) { radiance = make_float3(0.f); };
        Float4 acc = image.read(pix);
        Float cap = 30.f * cast<float>(spp);
        image.write(pix, acc + make_float4(min(radiance, make_float3(cap)), cast<float>(spp)));
        ray_count.atomic(0u).fetch_add(rays);
    };
    LUISA_INFO("compiling the kernel");
    auto shader = device.compile(render);
    LUISA_INFO("kernel compiled");
    stream << shader(image, seeds, ray_count, accel, aabb_buffer, flat_buffer, first_buffer, albedo_buffer, spp, light_inst).dispatch(width, height) << synchronize();// warm-up
    stream << image.copy_from(zeros.data()) << ray_count.copy_from(zero_u.data()) << synchronize();
    double best = 1e30;
    for (uint i = 0; i < dispatches; i++) {
        Clock clk;
        stream << shader(image, seeds, ray_count, accel, aabb_buffer, flat_buffer, first_buffer, albedo_buffer, spp, light_inst).dispatch(width, height) << synchronize();
        best = std::min(best, clk.toc());
    }
    std::vector<float4> out((size_t)width * height);
    uint total_rays = 0u;
    stream << image.copy_to(out.data()) << ray_count.copy_to(&total_rays) << synchronize();
    double mean[3] = {0, 0, 0};
    for (auto &p : out)
        for (int c = 0; c < 3; c++) mean[c] += p[c] / std::max(p.w, 1.f);
    FILE *f = fopen((dir + "/c5_" + backend + ".bin").c_str(), "wb");
    fwrite(out.data(), sizeof(float4), out.size(), f);
    fclose(f);
    const double rays_per_dispatch = (double)total_rays / dispatches;// the counter wraps at 2^32: kept below by the caller's sizes
    printf("{\"backend\": \"%s\", \"width\": %u, \"height\": %u, \"spp_per_dispatch\": %u, \"dispatch_ms\": %.4f, \"spp_per_s\": %.3f, \"mrays_s\": %.2f, "
           "\"mean_radiance\": [%.6f, %.6f, %.6f]}\n",
           backend, width, height, spp, best, spp / (best * 1e-3), rays_per_dispatch / best / 1e3, mean[0] / out.size(), mean[1] / out.size(), mean[2] / out.size());
    return 0;
}
