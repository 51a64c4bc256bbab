// tests/devlib/devlib_kernels.cu — TEST INFRASTRUCTURE: kernels written the way LuisaCompute's CUDA codegen emits them
// (src/backends/cuda/cuda_codegen_ast.cpp), compiled against the re-emitted RT device library
// (luisacompute_b200/csrc/lcrt_device_library.cuh) in its standalone mode.  They exercise every name the codegen can
// emit for the RT builtins: lc_accel_trace_closest / _any, lc_accel_query_all / _any + lc_ray_query_trace with
// outlined candidate handlers, LC_RAY_QUERY_* macros, lc_ray_query_committed_hit and the instance accessors.
//
// Shapes mirrored: src/tests/test_rtx.cpp:60-88 (closest hit per pixel), src/tests/test_procedural.cpp:84-131
// (RayQuery with a triangle cut-out handler and a ray-sphere procedural handler).
#include <cuda_runtime.h>
#include <stdint.h>

#define LCRT_DEVLIB_STANDALONE
#define LUISA_RAY_QUERY_IMPL_COUNT 2
#define LCRT_RAY_QUERY_SITE_COUNT 2
#include "../../luisacompute_b200/csrc/lcrt_device_library.cuh"

// ---- what `Kernel1D` argument structs look like after the codegen (resources passed by value)
struct Params {
    LCAccel accel;
    const LCRay *rays;
    lc_uint n;
    lc_uint mask;
    void *out0;
    void *out1;
    const float *aabbs;// 24-byte records, captured by the procedural handler like aabb_buffer in test_procedural.cpp
    lc_uint tri_prog;  // 0 commit, 1 cut-out (prim % 3 == 0 ignored), 2 ignore, 3 terminate
};

extern "C" __global__ void k_devlib_trace(Params params) {
    const lc_uint i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= params.n) return;
    LCRay ray = params.rays[i];
    LCTriangleHit hit = lc_accel_trace_closest(params.accel, ray, params.mask);
    static_cast<LCTriangleHit *>(params.out0)[i] = hit;
    static_cast<lc_uint *>(params.out1)[i] = lc_accel_trace_any(params.accel, ray, params.mask) ? 1u : 0u;
}

// closest-hit only / any-hit only (tools/devlib_perf.py times them)
extern "C" __global__ void k_devlib_closest(Params params) {
    const lc_uint i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= params.n) return;
    LCRay ray = params.rays[i];
    static_cast<LCTriangleHit *>(params.out0)[i] = lc_accel_trace_closest(params.accel, ray, params.mask);
}
extern "C" __global__ void k_devlib_any(Params params) {
    const lc_uint i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= params.n) return;
    LCRay ray = params.rays[i];
    static_cast<lc_uint *>(params.out1)[i] = lc_accel_trace_any(params.accel, ray, params.mask) ? 1u : 0u;
}

// ---- RayQuery site #0: outlined handlers as emitted by RayQueryLowering::_create_outline_definitions
struct LCRayQueryCtx0 {
    LCAccel accel;
    const float *aabbs;
    lc_uint tri_prog;
};

LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL(0) {
    LCIntersectionResult result{};
    auto ctx = *static_cast<LCRayQueryCtx0 *>(ctx_in);
    {// intersection handling body
        LCTriangleHit h = LC_RAY_QUERY_TRIANGLE_CANDIDATE_HIT(q);
        if (ctx.tri_prog == 0u) {
            LC_RAY_QUERY_COMMIT_TRIANGLE(q);
        } else if (ctx.tri_prog == 1u) {
            if (h.m1 % 3u != 0u) { LC_RAY_QUERY_COMMIT_TRIANGLE(q); }
        } else if (ctx.tri_prog == 3u) {
            LC_RAY_QUERY_COMMIT_TRIANGLE(q);
            LC_RAY_QUERY_TERMINATE(q);
        }
    }
    return result;
}

LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL(0) {
    LCIntersectionResult result{};
    auto ctx = *static_cast<LCRayQueryCtx0 *>(ctx_in);
    {// intersection handling body: sphere inscribed in the AABB, moved to world space by the instance transform;
     // the arithmetic follows the oracle's PROC_SPHERE program (oracle/oracle.cpp) in double so both agree
        LCProceduralHit h = LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT(q);
        LCRay ray = LC_RAY_QUERY_WORLD_RAY(q);
        const float *b = ctx.aabbs + 6u * h.m1;
        lc_float4x4 m = lc_accel_instance_transform(ctx.accel, h.m0);
        const double cxo = 0.5 * ((double)b[0] + (double)b[3]), cyo = 0.5 * ((double)b[1] + (double)b[4]), czo = 0.5 * ((double)b[2] + (double)b[5]);
        const double ro = 0.5 * ((double)b[3] - (double)b[0]);
        const double cx = (double)m[0].x * cxo + (double)m[1].x * cyo + (double)m[2].x * czo + (double)m[3].x;
        const double cy = (double)m[0].y * cxo + (double)m[1].y * cyo + (double)m[2].y * czo + (double)m[3].y;
        const double cz = (double)m[0].z * cxo + (double)m[1].z * cyo + (double)m[2].z * czo + (double)m[3].z;
        const double rad = ro * sqrt((double)m[0].x * m[0].x + (double)m[0].y * m[0].y + (double)m[0].z * m[0].z);
        const double Lx = ray.m0[0] - cx, Ly = ray.m0[1] - cy, Lz = ray.m0[2] - cz;
        const double dx = ray.m2[0], dy = ray.m2[1], dz = ray.m2[2];
        const double a = dx * dx + dy * dy + dz * dz, hb = Lx * dx + Ly * dy + Lz * dz, cc = Lx * Lx + Ly * Ly + Lz * Lz - rad * rad;
        const double disc = hb * hb - a * cc;
        if (disc >= 0.0) {
            const double sq = sqrt(disc);
            const double t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
            if (t0 >= (double)ray.m1 && t0 < (double)ray.m3) {
                LC_RAY_QUERY_COMMIT_PROCEDURAL(q, (float)t0);
            } else if (t1 >= (double)ray.m1 && t1 < (double)ray.m3) {
                LC_RAY_QUERY_COMMIT_PROCEDURAL(q, (float)t1);
            }
        }
    }
    return result;
}

// ---- RayQuery site #1: query_any with empty handlers capturing nothing (the corner case of :417-418)
LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL(1) {
    LCIntersectionResult result{};
    LC_RAY_QUERY_COMMIT_TRIANGLE(q);
    return result;
}
LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL(1) {
    LCIntersectionResult result{};
    return result;
}

extern "C" __global__ void k_devlib_query(Params params) {
    const lc_uint i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= params.n) return;
    LCRay ray = params.rays[i];
    LCRayQueryAll q = lc_accel_query_all(params.accel, ray, params.mask);
    {// ray query #0
        LCRayQueryCtx0 ctx{params.accel, params.aabbs, params.tri_prog};
        lc_ray_query_trace(q, 0, &ctx);
    }// ray query #0
    static_cast<LCCommittedHit *>(params.out0)[i] = lc_ray_query_committed_hit(q);
    LCRayQueryAny qa = lc_accel_query_any(params.accel, ray, params.mask);
    {// ray query #1
        lc_ray_query_trace(qa, 1, nullptr);
    }// ray query #1
    static_cast<LCCommittedHit *>(params.out1)[i] = lc_ray_query_committed_hit(qa);
}

// ---- the same two RayQuery sites lowered INLINE (what plugin/hybrid/b200_ray_query_lowering.inl emits): the handler bodies sit in
// the kernel's scope between proceed and resolve and use the kernel's own variables (`params`) instead of a capture struct
extern "C" __global__ void k_devlib_query_inline(Params params) {
    const lc_uint i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= params.n) return;
    LCRay ray = params.rays[i];
    LCRayQueryAll q = lc_accel_query_all(params.accel, ray, params.mask);
    {// ray query #0, lowered inline
        LCRayQueryInline lc_rq;
        lc_ray_query_inline_begin(lc_rq, q);
        while (lc_ray_query_inline_proceed(lc_rq)) {
            LCIntersectionResult result{};
            const LCRTCandidate &lc_rt_candidate = lc_rq.candidate;
            if (lc_rq.triangle_candidate) {
                LCTriangleHit h = LC_RAY_QUERY_TRIANGLE_CANDIDATE_HIT(q);
                if (params.tri_prog == 0u) {
                    LC_RAY_QUERY_COMMIT_TRIANGLE(q);
                } else if (params.tri_prog == 1u) {
                    if (h.m1 % 3u != 0u) { LC_RAY_QUERY_COMMIT_TRIANGLE(q); }
                } else if (params.tri_prog == 3u) {
                    LC_RAY_QUERY_COMMIT_TRIANGLE(q);
                    LC_RAY_QUERY_TERMINATE(q);
                }
            } else {
                LCProceduralHit h = LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT(q);
                LCRay wray = LC_RAY_QUERY_WORLD_RAY(q);
                const float *b = params.aabbs + 6u * h.m1;
                lc_float4x4 m = lc_accel_instance_transform(params.accel, h.m0);
                const double cxo = 0.5 * ((double)b[0] + (double)b[3]), cyo = 0.5 * ((double)b[1] + (double)b[4]), czo = 0.5 * ((double)b[2] + (double)b[5]);
                const double ro = 0.5 * ((double)b[3] - (double)b[0]);
                const double cx = (double)m[0].x * cxo + (double)m[1].x * cyo + (double)m[2].x * czo + (double)m[3].x;
                const double cy = (double)m[0].y * cxo + (double)m[1].y * cyo + (double)m[2].y * czo + (double)m[3].y;
                const double cz = (double)m[0].z * cxo + (double)m[1].z * cyo + (double)m[2].z * czo + (double)m[3].z;
                const double rad = ro * sqrt((double)m[0].x * m[0].x + (double)m[0].y * m[0].y + (double)m[0].z * m[0].z);
                const double Lx = wray.m0[0] - cx, Ly = wray.m0[1] - cy, Lz = wray.m0[2] - cz;
                const double dx = wray.m2[0], dy = wray.m2[1], dz = wray.m2[2];
                const double a = dx * dx + dy * dy + dz * dz, hb = Lx * dx + Ly * dy + Lz * dz, cc = Lx * Lx + Ly * Ly + Lz * Lz - rad * rad;
                const double disc = hb * hb - a * cc;
                if (disc >= 0.0) {
                    const double sq = sqrt(disc);
                    const double t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
                    if (t0 >= (double)wray.m1 && t0 < (double)wray.m3) {
                        LC_RAY_QUERY_COMMIT_PROCEDURAL(q, (float)t0);
                    } else if (t1 >= (double)wray.m1 && t1 < (double)wray.m3) {
                        LC_RAY_QUERY_COMMIT_PROCEDURAL(q, (float)t1);
                    }
                }
            }
            lc_ray_query_inline_resolve(lc_rq, result);
        }
        lc_ray_query_inline_end(lc_rq, q);
    }// ray query #0
    static_cast<LCCommittedHit *>(params.out0)[i] = lc_ray_query_committed_hit(q);
    LCRayQueryAny qa = lc_accel_query_any(params.accel, ray, params.mask);
    {// ray query #1, lowered inline
        LCRayQueryInline lc_rq;
        lc_ray_query_inline_begin(lc_rq, qa);
        while (lc_ray_query_inline_proceed(lc_rq)) {
            LCIntersectionResult result{};
            const LCRTCandidate &lc_rt_candidate = lc_rq.candidate;
            if (lc_rq.triangle_candidate) {
                LC_RAY_QUERY_COMMIT_TRIANGLE(qa);
            } else {
            }
            lc_ray_query_inline_resolve(lc_rq, result);
        }
        lc_ray_query_inline_end(lc_rq, qa);
    }// ray query #1
    static_cast<LCCommittedHit *>(params.out1)[i] = lc_ray_query_committed_hit(qa);
}

// ---- in-kernel instance edits (cuda_device_resource.h:1470-1510): thread i edits instance i
extern "C" __global__ void k_devlib_instance_ops(Params params) {
    const lc_uint i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= params.n) return;
    lc_float4x4 m = lc_accel_instance_transform(params.accel, i);
    m[3].x += 1.0f;// translate by +1 in x
    lc_accel_set_instance_transform(params.accel, i, m);
    lc_accel_set_instance_visibility(params.accel, i, 0x100u | (i & 0xffu));// masked to 8 bits
    lc_accel_set_instance_user_id(params.accel, i, lc_accel_instance_user_id(params.accel, i) + 1000u);
    lc_accel_set_instance_opacity(params.accel, i, (i & 1u) != 0u);
    static_cast<lc_uint *>(params.out0)[i] = lc_accel_instance_visibility_mask(params.accel, i);
}

extern "C" {
#define DL_EXPORT __attribute__((visibility("default")))
static Params make_params(uint64_t handle, void *instances, const void *rays, uint32_t n, uint32_t mask, void *out0, void *out1, const void *aabbs,
                          uint32_t tri_prog) {
    Params p;
    p.accel.handle = handle, p.accel.instances = static_cast<LCAccelInstance *>(instances);
    p.rays = static_cast<const LCRay *>(rays), p.n = n, p.mask = mask, p.out0 = out0, p.out1 = out1;
    p.aabbs = static_cast<const float *>(aabbs), p.tri_prog = tri_prog;
    return p;
}
DL_EXPORT int dl_trace(void *stream, uint64_t handle, void *instances, const void *rays, uint32_t n, uint32_t mask, void *hits, void *occluded) {
    if (n == 0) return 0;
    k_devlib_trace<<<(n + 127u) / 128u, 128, 0, static_cast<cudaStream_t>(stream)>>>(make_params(handle, instances, rays, n, mask, hits, occluded, nullptr, 0));
    return (int)cudaGetLastError();
}
DL_EXPORT int dl_closest(void *stream, uint64_t handle, void *instances, const void *rays, uint32_t n, uint32_t mask, void *hits, uint32_t block) {
    if (n == 0) return 0;
    k_devlib_closest<<<(n + block - 1u) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(make_params(handle, instances, rays, n, mask, hits, nullptr, nullptr, 0));
    return (int)cudaGetLastError();
}
DL_EXPORT int dl_any(void *stream, uint64_t handle, void *instances, const void *rays, uint32_t n, uint32_t mask, void *occluded, uint32_t block) {
    if (n == 0) return 0;
    k_devlib_any<<<(n + block - 1u) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(make_params(handle, instances, rays, n, mask, nullptr, occluded, nullptr, 0));
    return (int)cudaGetLastError();
}
DL_EXPORT int dl_query(void *stream, uint64_t handle, void *instances, const void *rays, uint32_t n, uint32_t mask, const void *aabbs, uint32_t tri_prog,
                       void *hits_all, void *hits_any) {
    if (n == 0) return 0;
    k_devlib_query<<<(n + 127u) / 128u, 128, 0, static_cast<cudaStream_t>(stream)>>>(make_params(handle, instances, rays, n, mask, hits_all, hits_any, aabbs, tri_prog));
    return (int)cudaGetLastError();
}
DL_EXPORT int dl_query_inline(void *stream, uint64_t handle, void *instances, const void *rays, uint32_t n, uint32_t mask, const void *aabbs, uint32_t tri_prog,
                              void *hits_all, void *hits_any) {
    if (n == 0) return 0;
    k_devlib_query_inline<<<(n + 127u) / 128u, 128, 0, static_cast<cudaStream_t>(stream)>>>(make_params(handle, instances, rays, n, mask, hits_all, hits_any, aabbs, tri_prog));
    return (int)cudaGetLastError();
}
DL_EXPORT int dl_instance_ops(void *stream, uint64_t handle, void *instances, uint32_t n, void *masks_out) {
    if (n == 0) return 0;
    k_devlib_instance_ops<<<(n + 127u) / 128u, 128, 0, static_cast<cudaStream_t>(stream)>>>(make_params(handle, instances, nullptr, n, 0, masks_out, nullptr, nullptr, 0));
    return (int)cudaGetLastError();
}
}
