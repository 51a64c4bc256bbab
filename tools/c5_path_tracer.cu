// tools/c5_path_tracer.cu — MEASUREMENT WORKLOAD for BASELINE config C5: the integrator of
// src/tests/test_path_tracing.cpp:165-256 (depth <= 10, next-event estimation with an any-hit shadow ray, MIS,
// Russian roulette, tea/lcg sampler :112-134) written as the CUDA the reference's codegen would emit, over the
// re-emitted RT device library.  Differences dictated by the config: the scene is the Cornell room plus ONE
// ProceduralPrimitive of 64 spheres, so closest hits go through RayQuery (lc_accel_query_all + an outlined
// ray-sphere procedural handler, as src/tests/test_procedural.cpp:101-125) and shadow rays through lc_accel_query_any.
#include <cuda_runtime.h>
#include <stdint.h>

#define LCRT_DEVLIB_STANDALONE
#define LUISA_RAY_QUERY_IMPL_COUNT 2
#ifndef C5_NESTED_LOOPS
#define C5_NESTED_LOOPS 0
#endif
#include "../luisacompute_b200/csrc/lcrt_device_library.cuh"

struct C5Params {
    LCAccel accel;
    const float *aabbs;       // 64 x 24 B (procedural primitive input, captured by the handlers)
    const float4 *tri_verts;  // 3 float4 per triangle, all meshes concatenated
    const uint32_t *tri_first;// first triangle of instance i in tri_verts
    const float4 *albedo;     // per instance
    float4 *image;            // w*h accumulated radiance (rgb) + sample count (w)
    uint32_t *seeds;          // w*h sampler state
    unsigned long long *ray_counter;
    uint32_t w, h, spp, light_inst, sphere_inst;
};

struct LCRayQueryCtx0 {
    const float *aabbs;
};

__device__ inline bool ray_sphere(const LCRay &ray, const float *b, float &t_out) {
    const float cx = 0.5f * (b[0] + b[3]), cy = 0.5f * (b[1] + b[4]), cz = 0.5f * (b[2] + b[5]), rad = 0.5f * (b[3] - b[0]);
    const float Lx = ray.m0[0] - cx, Ly = ray.m0[1] - cy, Lz = ray.m0[2] - cz;
    const float dx = ray.m2[0], dy = ray.m2[1], dz = ray.m2[2];
    const float a = dx * dx + dy * dy + dz * dz, hb = Lx * dx + Ly * dy + Lz * dz, cc = Lx * Lx + Ly * Ly + Lz * Lz - rad * rad;
    const float disc = hb * hb - a * cc;
    if (disc < 0.f) return false;
    const float sq = sqrtf(disc);
    const float t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
    if (t0 >= ray.m1 && t0 < ray.m3) { t_out = t0; return true; }
    if (t1 >= ray.m1 && t1 < ray.m3) { t_out = t1; return true; }
    return false;
}

LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL(0) {
    LCIntersectionResult result{};
    return result;// the room is opaque: triangles are committed without a handler
}
LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL(0) {
    LCIntersectionResult result{};
    auto ctx = *static_cast<LCRayQueryCtx0 *>(ctx_in);
    {
        LCProceduralHit h = LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT(q);
        LCRay ray = LC_RAY_QUERY_WORLD_RAY(q);
        float t;
        if (ray_sphere(ray, ctx.aabbs + 6u * h.m1, t)) { LC_RAY_QUERY_COMMIT_PROCEDURAL(q, t); }
    }
    return result;
}
LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL(1) {
    LCIntersectionResult result{};
    return result;
}
LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL(1) {
    LCIntersectionResult result{};
    auto ctx = *static_cast<LCRayQueryCtx0 *>(ctx_in);
    {
        LCProceduralHit h = LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT(q);
        LCRay ray = LC_RAY_QUERY_WORLD_RAY(q);
        float t;
        if (ray_sphere(ray, ctx.aabbs + 6u * h.m1, t)) { LC_RAY_QUERY_COMMIT_PROCEDURAL(q, t); }
    }
    return result;
}

__device__ inline uint32_t tea(uint32_t v0, uint32_t v1) {
    uint32_t s0 = 0u;
    for (uint32_t n = 0u; n < 4u; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5u) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5u) + 0x7e95761eu);
    }
    return v0;
}
__device__ inline float lcg(uint32_t &state) {
    state = 1664525u * state + 1013904223u;
    return (float)(state & 0x00ffffffu) * (1.0f / (float)0x01000000u);
}
struct V3 {
    float x, y, z;
};
__device__ inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ inline V3 normalize(V3 a) { return a * rsqrtf(dot(a, a)); }
__device__ inline LCRay make_ray(V3 o, V3 d, float tmin, float tmax) {
    LCRay r;
    r.m0[0] = o.x, r.m0[1] = o.y, r.m0[2] = o.z, r.m1 = tmin, r.m2[0] = d.x, r.m2[1] = d.y, r.m2[2] = d.z, r.m3 = tmax;
    return r;
}
// src/dsl/rtx/ray.cpp:16-23
__device__ inline V3 offset_ray_origin(V3 p, V3 n) {
    constexpr float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    auto one = [&](float pc, float nc) {
        const int of_i = (int)(int_scale * nc);
        const float p_i = __int_as_float(__float_as_int(pc) + (pc < 0.f ? -of_i : of_i));
        return fabsf(pc) < origin ? pc + float_scale * nc : p_i;
    };
    return v3(one(p.x, n.x), one(p.y, n.y), one(p.z, n.z));
}

extern "C" __global__ void __launch_bounds__(256, 2) k_c5_path_trace(C5Params params) {
    const uint32_t x = blockIdx.x * 16u + (threadIdx.x & 15u), y = blockIdx.y * 16u + (threadIdx.x >> 4);
    if (x >= params.w || y >= params.h) return;
    const uint32_t live = __activemask();// the lanes of this warp that trace (converged here); used by the ray-count reduction at the end
    const uint32_t pix = y * params.w + x;
    uint32_t state = params.seeds[pix];
    if (state == 0u) state = tea(x, y);
    const float frame = (float)min(params.w, params.h);
    V3 radiance = v3(0.f, 0.f, 0.f);
    uint32_t rays = 0u;
    const V3 light_position = v3(-0.25f, 0.995f, -0.25f), light_u = v3(0.5f, 0.f, 0.f), light_v = v3(0.f, 0.f, 0.5f);
    const V3 light_emission = v3(17.f, 12.f, 4.f), light_normal = v3(0.f, -1.f, 0.f);
    const float light_area = 0.25f, inv_pi = 0.318309886f;
    LCRayQueryCtx0 ctx{params.aabbs};
    // One flat loop over (sample, bounce): a lane whose path ended starts its next sample in the following iteration instead
    // of idling until the longest path of the warp is done (nested loops reconverge after every sample: 6.3 of 32 lanes
    // active on this scene, profiles/r1_c5_ncu.md).  The per-pixel random sequence and the order of the radiance sums are
    // those of the nested loops, so the image is bit-identical.  C5_NESTED_LOOPS=1 builds the nested form.
    LCRay ray;
    V3 beta = v3(1.f, 1.f, 1.f);
    float pdf_bsdf = 0.f;
#if C5_NESTED_LOOPS
    for (uint32_t s = 0; s < params.spp; s++) {
        for (uint32_t depth = 0; depth < 10u; depth++) {
#else
    for (uint32_t s = 0, depth = 0; s < params.spp;) {
        {
#endif
            if (depth == 0u) {
                const float rx = lcg(state), ry = lcg(state);
                const float px = ((float)x + rx) / frame * 2.f - (float)params.w / frame, py = -(((float)y + ry) / frame * 2.f - (float)params.h / frame);
                const V3 origin = v3(0.f, 0.f, 3.4f);
                const float th = tanf(0.5f * 38.0f * 0.01745329252f);
                ray = make_ray(origin, normalize(v3(px * th, py * th, -1.f)), 0.f, 3.402823466e38f);
                beta = v3(1.f, 1.f, 1.f);
                pdf_bsdf = 0.f;
            }
            bool path_ended = true;
            do {
            LCRayQueryAll q = lc_accel_query_all(params.accel, ray, 0xffu);
            {// ray query #0
                lc_ray_query_trace(q, 0, &ctx);
            }
            rays++;
            const LCCommittedHit hit = lc_ray_query_committed_hit(q);
            if (hit.m3 == (lc_uint)LCHitType::MISS) break;
            const V3 o = v3(ray.m0[0], ray.m0[1], ray.m0[2]), d = v3(ray.m2[0], ray.m2[1], ray.m2[2]);
            const V3 p = o + d * hit.m4;
            V3 n;
            if (hit.m3 == (lc_uint)LCHitType::PROCEDURAL) {
                const float *b = params.aabbs + 6u * hit.m1;
                n = normalize(p - v3(0.5f * (b[0] + b[3]), 0.5f * (b[1] + b[4]), 0.5f * (b[2] + b[5])));
            } else {
                const float4 *tv = params.tri_verts + 3u * (params.tri_first[hit.m0] + hit.m1);
                const V3 p0 = v3(tv[0].x, tv[0].y, tv[0].z), p1 = v3(tv[1].x, tv[1].y, tv[1].z), p2 = v3(tv[2].x, tv[2].y, tv[2].z);
                n = normalize(cross(p1 - p0, p2 - p0));
            }
            float cos_wo = -dot(d, n);
            if (hit.m3 != (lc_uint)LCHitType::PROCEDURAL && cos_wo < 0.f && hit.m0 != params.light_inst) n = n * -1.f, cos_wo = -cos_wo;// two-sided walls
            if (cos_wo < 1e-4f) break;
            if (hit.m3 == (lc_uint)LCHitType::BUILTIN && hit.m0 == params.light_inst) {
                if (depth == 0u) radiance = radiance + light_emission;
                else {
                    const V3 po = p - o;
                    const float pdf_light = dot(po, po) / (light_area * cos_wo);
                    radiance = radiance + beta * light_emission * (pdf_bsdf / fmaxf(pdf_bsdf + pdf_light, 1e-4f));
                }
                break;
            }
            // sample light
            const float ux_light = lcg(state), uy_light = lcg(state);
            const V3 p_light = light_position + light_u * ux_light + light_v * uy_light;
            const V3 pp = offset_ray_origin(p, n), pp_light = offset_ray_origin(p_light, light_normal);
            const V3 dl = pp_light - pp;
            const float d_light = sqrtf(dot(dl, dl));
            const V3 wi_light = dl * (1.f / d_light);
            LCRayQueryAny qa = lc_accel_query_any(params.accel, make_ray(offset_ray_origin(pp, n), wi_light, 0.f, d_light), 0xffu);
            {// ray query #1
                lc_ray_query_trace(qa, 1, &ctx);
            }
            rays++;
            const bool occluded = lc_ray_query_committed_hit(qa).m3 != (lc_uint)LCHitType::MISS;
            const float cos_wi_light = dot(wi_light, n), cos_light = -dot(light_normal, wi_light);
            const float4 al = params.albedo[hit.m0];
            const V3 albedo = v3(al.x, al.y, al.z);
            if (!occluded && cos_wi_light > 1e-4f && cos_light > 1e-4f) {
                const float pdf_light = (d_light * d_light) / (light_area * cos_light), pdf_b = cos_wi_light * inv_pi;
                const float mis = pdf_light / fmaxf(pdf_light + pdf_b, 1e-4f);
                radiance = radiance + beta * albedo * light_emission * (inv_pi * cos_wi_light * mis / fmaxf(pdf_light, 1e-4f));
            }
            // sample BSDF
            const V3 binormal = normalize(fabsf(n.x) > fabsf(n.z) ? v3(-n.y, n.x, 0.f) : v3(0.f, -n.z, n.y));
            const V3 tangent = normalize(cross(binormal, n));
            const float ux = lcg(state), uy = lcg(state);
            const float r = sqrtf(ux), phi = 6.283185307f * uy;
            const V3 wl = v3(r * cosf(phi), r * sinf(phi), sqrtf(1.f - ux));
            ray = make_ray(pp, tangent * wl.x + binormal * wl.y + n * wl.z, 0.f, 3.402823466e38f);
            pdf_bsdf = fabsf(wl.z) * inv_pi;
            beta = beta * albedo;
            const float l = 0.212671f * beta.x + 0.715160f * beta.y + 0.072169f * beta.z;
            if (l == 0.f) break;
            const float qq = fmaxf(l, 0.05f);
            if (lcg(state) >= qq) break;
            beta = beta * (1.f / qq);
            path_ended = false;
            } while (false);
#if C5_NESTED_LOOPS
            if (path_ended) break;
#else
            if (path_ended || ++depth == 10u) depth = 0u, s++;
#endif
        }
    }
    params.seeds[pix] = state;
    if (!(radiance.x == radiance.x && radiance.y == radiance.y && radiance.z == radiance.z)) radiance = v3(0.f, 0.f, 0.f);
    float4 acc = params.image[pix];
    acc.x += fminf(radiance.x, 30.f * params.spp), acc.y += fminf(radiance.y, 30.f * params.spp), acc.z += fminf(radiance.z, 30.f * params.spp), acc.w += (float)params.spp;
    params.image[pix] = acc;
    // one atomic per warp: reconverge the lanes that entered together first (__activemask() at this point is whatever subset
    // the scheduler happens to run together after the divergent loop, which silently dropped counts)
    __syncwarp(live);
    uint32_t total = 0u;
    for (uint32_t m = live; m; m &= m - 1u) total += __shfl_sync(live, rays, __ffs((int)m) - 1);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs((int)live) - 1)) atomicAdd(params.ray_counter, (unsigned long long)total);
}

extern "C" __attribute__((visibility("default"))) int c5_render(void *stream, uint64_t handle, void *instances, const void *aabbs, const void *tri_verts,
                                                                  const void *tri_first, const void *albedo, void *image, void *seeds, void *ray_counter,
                                                                  uint32_t w, uint32_t h, uint32_t spp, uint32_t light_inst, uint32_t sphere_inst) {
    C5Params p;
    p.accel.handle = handle, p.accel.instances = static_cast<LCAccelInstance *>(instances);
    p.aabbs = static_cast<const float *>(aabbs), p.tri_verts = static_cast<const float4 *>(tri_verts), p.tri_first = static_cast<const uint32_t *>(tri_first);
    p.albedo = static_cast<const float4 *>(albedo), p.image = static_cast<float4 *>(image), p.seeds = static_cast<uint32_t *>(seeds);
    p.ray_counter = static_cast<unsigned long long *>(ray_counter), p.w = w, p.h = h, p.spp = spp, p.light_inst = light_inst, p.sphere_inst = sphere_inst;
    dim3 grid((w + 15u) / 16u, (h + 15u) / 16u);
    k_c5_path_trace<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return (int)cudaGetLastError();
}
