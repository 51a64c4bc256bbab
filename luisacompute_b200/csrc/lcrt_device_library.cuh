// lcrt_device_library.cuh — the ray-tracing section of LuisaCompute's CUDA device library, re-emitted against the
// BVH8 format of this repository.
//
// The reference's codegen pastes a device-library text in front of every generated kernel
// (src/backends/cuda/cuda_compiler.cpp:195-215, cuda_codegen_ast.cpp:1291-1330) and lowers the DSL's RT builtins to
// calls *by name* (cuda_codegen_ast.cpp:983-1004, 1047-1050).  This file defines every one of those names with the
// argument meaning of src/backends/cuda/cuda_builtin/cuda_device_resource.h:1279-1510, 1553-2387, but implements them
// as ordinary device code over lcrt_traverse.h — no OptiX hit objects, payload registers, IS/AH programs or SBT — so a
// kernel that traces rays is launched like any other kernel (cuLaunchKernel; cuda_shader_native.cpp).
//
//   type / function (reference line)                          here
//   LCRay :1279  LCTriangleHit :1286  LCProceduralHit :1293   same layouts (static_asserted)
//   LCCommittedHit :1304  LCAccelInstance :1361  LCAccel :1371 same layouts; LCAccel.handle = device address of the TLAS blob
//   lc_accel_instance_* / lc_accel_set_instance_* :1450-1510  same, plus the spellings the codegen emits
//   lc_accel_trace_closest / _any (+ _motion_blur) :1783-1842 Traversal<MODE_CLOSEST / MODE_ANY>
//   LCRayQuery, lc_accel_query_all/any, lc_ray_query_trace,
//   lc_ray_query_committed_hit :1896-1967                     Traversal<MODE_QUERY> with the handlers inlined
//   LUISA_DECL_RAY_QUERY_*_IMPL, LC_RAY_QUERY_* :2054-2075    same macro names; the candidate travels as a hidden
//                                                             second parameter instead of OptiX intrinsics
//   __intersection__ray_query / __anyhit__ray_query :2146-2385  not needed (the switch over impl_tag is in
//                                                             lc_ray_query_trace, folded at compile time)
//
// Text order when used by the plugin: reference prelude (lc_uint, lc_float2, lc_float4x4, lc_array, ...) ->
// lcrt_common.h -> lcrt_traverse.h -> this file.  With LCRT_DEVLIB_STANDALONE defined (tests, hand-written kernels)
// a minimal prelude is provided here instead.
#pragma once
#include "lcrt_traverse.h"

#if defined(LCRT_DEVLIB_STANDALONE)
#include <stdio.h>
// ---- minimal prelude: only what the RT section needs from cuda_device_math.h / cuda_device_resource.h
using lc_uint = unsigned int;
using lc_ulong = unsigned long long;
using lc_float = float;
using lc_bool = bool;
using lc_float2 = float2;
using lc_float3 = float3;
using lc_float4 = float4;
template<typename T, size_t N>
struct lc_array {
    T m[N];
    __device__ __host__ T &operator[](size_t i) noexcept { return m[i]; }
    __device__ __host__ const T &operator[](size_t i) const noexcept { return m[i]; }
};
struct lc_float4x4 {// column-major like luisa::float4x4: m[col][row]
    lc_float4 cols[4];
    __device__ __host__ lc_float4 &operator[](size_t i) noexcept { return cols[i]; }
    __device__ __host__ const lc_float4 &operator[](size_t i) const noexcept { return cols[i]; }
};
__device__ __host__ inline float &lc_rt_component(lc_float4 &v, int i) noexcept { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __host__ inline lc_float2 lc_make_float2(float x, float y) noexcept { return make_float2(x, y); }
__device__ __host__ inline lc_float4x4 lc_make_float4x4(float m00, float m01, float m02, float m03, float m10, float m11, float m12, float m13,
                                                        float m20, float m21, float m22, float m23, float m30, float m31, float m32,
                                                        float m33) noexcept {
    lc_float4x4 m;
    m.cols[0] = make_float4(m00, m01, m02, m03), m.cols[1] = make_float4(m10, m11, m12, m13);
    m.cols[2] = make_float4(m20, m21, m22, m23), m.cols[3] = make_float4(m30, m31, m32, m33);
    return m;
}
#define lc_assume(x) ((void)0)
#endif

#ifndef LUISA_RAY_QUERY_IMPL_COUNT
#define LUISA_RAY_QUERY_IMPL_COUNT 0
#endif

// ------------------------------------------------------------------------------------------ wire types
struct alignas(16) LCRay {
    lc_array<float, 3> m0;// origin
    float m1;             // t_min
    lc_array<float, 3> m2;// direction
    float m3;             // t_max
};
static_assert(sizeof(LCRay) == 32u && alignof(LCRay) == 16u, "LCRay layout mismatch");

struct alignas(8) LCTriangleHit {
    lc_uint m0;  // instance index
    lc_uint m1;  // primitive index
    lc_float2 m2;// barycentric coordinates
    lc_float m3; // t_hit
};
static_assert(sizeof(LCTriangleHit) == 24u, "LCTriangleHit layout mismatch");

struct LCProceduralHit {
    lc_uint m0;// instance index
    lc_uint m1;// primitive index
};

enum struct LCHitType { MISS = 0, BUILTIN = 1, PROCEDURAL = 2 };

struct LCCommittedHit {
    lc_uint m0;  // instance index
    lc_uint m1;  // primitive index
    lc_float2 m2;// barycentric coordinates
    lc_uint m3;  // hit type
    lc_float m4; // t_hit
};
static_assert(sizeof(LCCommittedHit) == 24u, "LCCommittedHit size mismatch");
static_assert(alignof(LCCommittedHit) == 8u, "LCCommittedHit align mismatch");

enum LCInstanceFlags : unsigned int {
    LC_INSTANCE_FLAG_NONE = 0u,
    LC_INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING = 1u << 0u,
    LC_INSTANCE_FLAG_FLIP_TRIANGLE_FACING = 1u << 1u,
    LC_INSTANCE_FLAG_DISABLE_ANYHIT = 1u << 2u,
    LC_INSTANCE_FLAG_ENFORCE_ANYHIT = 1u << 3u,
};

struct alignas(16) LCAccelInstance {
    lc_array<lc_float4, 3> m;
    lc_uint user_id;
    lc_uint sbt_offset;
    lc_uint mask;
    lc_uint flags;
    unsigned long long handle;
    lc_uint pad[2];
};
static_assert(sizeof(LCAccelInstance) == 80u, "LCAccelInstance layout mismatch");

struct alignas(16u) LCAccel {
    unsigned long long handle;// device address of the TLAS blob (lcrt::BlobHeader)
    LCAccelInstance *instances;
};
static_assert(sizeof(LCAccel) == 16u, "LCAccel layout mismatch");

// ------------------------------------------------------------------------------------------ instance accessors
// (cuda_device_resource.h:1450-1510).  Setters edit the 80-byte instance record; like with OptiX, the edit reaches
// traversal with the next AccelBuildCommand (refit or rebuild) on the host.
[[nodiscard]] __device__ inline lc_float4x4 lc_accel_instance_transform(LCAccel accel, lc_uint instance_id) noexcept {
    auto m = accel.instances[instance_id].m;
    return lc_make_float4x4(m[0].x, m[1].x, m[2].x, 0.0f, m[0].y, m[1].y, m[2].y, 0.0f, m[0].z, m[1].z, m[2].z, 0.0f, m[0].w, m[1].w, m[2].w, 1.0f);
}
[[nodiscard]] __device__ inline lc_uint lc_accel_instance_user_id(LCAccel accel, lc_uint instance_id) noexcept { return accel.instances[instance_id].user_id; }
[[nodiscard]] __device__ inline lc_uint lc_accel_instance_visibility(LCAccel accel, lc_uint index) noexcept { return accel.instances[index].mask; }
// the spelling the codegen emits for CallOp::RAY_TRACING_INSTANCE_VISIBILITY_MASK (cuda_codegen_ast.cpp:985)
[[nodiscard]] __device__ inline lc_uint lc_accel_instance_visibility_mask(LCAccel accel, lc_uint index) noexcept { return accel.instances[index].mask; }

__device__ inline void lc_accel_set_instance_transform(LCAccel accel, lc_uint index, lc_float4x4 m) noexcept {
    lc_array<lc_float4, 3> p;
    p[0].x = m[0].x, p[0].y = m[1].x, p[0].z = m[2].x, p[0].w = m[3].x;
    p[1].x = m[0].y, p[1].y = m[1].y, p[1].z = m[2].y, p[1].w = m[3].y;
    p[2].x = m[0].z, p[2].y = m[1].z, p[2].z = m[2].z, p[2].w = m[3].z;
    accel.instances[index].m = p;
}
__device__ inline void lc_accel_set_instance_visibility(LCAccel accel, lc_uint index, lc_uint mask) noexcept { accel.instances[index].mask = mask & 0xffu; }
__device__ inline void lc_accel_set_instance_opacity(LCAccel accel, lc_uint index, bool opaque) noexcept {
    auto flags = accel.instances[index].flags;
    // procedural primitives ignore the opaque flag: only triangle-mesh instances carry DISABLE_TRIANGLE_FACE_CULLING
    if (flags & LC_INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING) {
        flags &= ~(LC_INSTANCE_FLAG_DISABLE_ANYHIT | LC_INSTANCE_FLAG_ENFORCE_ANYHIT);
        flags |= opaque ? LC_INSTANCE_FLAG_DISABLE_ANYHIT : LC_INSTANCE_FLAG_ENFORCE_ANYHIT;
        accel.instances[index].flags = flags;
    }
}
__device__ inline void lc_accel_set_instance_user_id(LCAccel accel, lc_uint index, lc_uint user_id) noexcept { accel.instances[index].user_id = user_id; }

// ------------------------------------------------------------------------------------------ motion instances (not supported)
// The code generator lowers the DSL's motion-instance accessors to these names (cuda_codegen_ast.cpp:1047-1050; the reference's own
// library spells them lc_instance_motion_* at this commit, cuda_device_resource.h:1387-1448, so such a kernel does not even compile
// there).  Motion blur is outside this backend's scope: the names resolve, and a kernel that reaches one of them
// stops with a message instead of failing inside NVRTC.
namespace lcrt {
__device__ inline void devlib_motion_unsupported() noexcept {
    printf("[lc-backend-b200] motion-blur instances are not supported by this backend (SURVEY.md 8 f4)\n");
    __trap();
}
struct DevlibAnyValue {// converts to whatever record type the generated code expects (LCMotionSRT is emitted after this library)
    template<class T>
    __device__ operator T() const noexcept {
        T t{};
        return t;
    }
};
}// namespace lcrt
[[nodiscard]] __device__ inline lc_float4x4 lc_accel_instance_motion_matrix(LCAccel, lc_uint, lc_uint) noexcept {
    lcrt::devlib_motion_unsupported();
    return lc_make_float4x4(1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f);
}
[[nodiscard]] __device__ inline lcrt::DevlibAnyValue lc_accel_instance_motion_srt(LCAccel, lc_uint, lc_uint) noexcept {
    lcrt::devlib_motion_unsupported();
    return {};
}
__device__ inline void lc_accel_set_instance_motion_matrix(LCAccel, lc_uint, lc_uint, lc_float4x4) noexcept { lcrt::devlib_motion_unsupported(); }
template<class SRT>
__device__ inline void lc_accel_set_instance_motion_srt(LCAccel, lc_uint, lc_uint, SRT) noexcept { lcrt::devlib_motion_unsupported(); }

// ------------------------------------------------------------------------------------------ intersect / intersect_any
namespace lcrt {
__device__ inline const BlobHeader *devlib_tlas(LCAccel accel) noexcept { return reinterpret_cast<const BlobHeader *>(accel.handle); }
// a ray that ran out of traversal stack (never on a structure this library built: BlobHeader::max_stack) is reported, not hidden
__device__ inline void devlib_report(LCAccel accel, bool overflowed) noexcept {
    if (overflowed) atomicOr(&reinterpret_cast<BlobHeader *>(accel.handle)->trace_error, 1u);
}
}// namespace lcrt

[[nodiscard]] __device__ inline LCTriangleHit lc_accel_trace_closest(LCAccel accel, LCRay ray, lc_uint mask) noexcept {
    lcrt::LocalStore store;
    lcrt::NoCallbacks cb;
    lcrt::Traversal<lcrt::MODE_CLOSEST, false, lcrt::LocalStore, lcrt::NoCallbacks> tr;
    tr.init(lcrt::devlib_tlas(accel), reinterpret_cast<const lcrt::Instance *>(accel.instances), reinterpret_cast<const lcrt::Ray *>(&ray), mask, &store, &cb);
    tr.run();
    lcrt::devlib_report(accel, tr.overflowed());
    LCTriangleHit hit;
    lc_uint type;
    tr.result(hit.m0, hit.m1, type, hit.m3, hit.m2.x, hit.m2.y);
    return hit;// miss <=> m0 == ~0u (include/luisa/dsl/rtx/hit.h:67)
}

[[nodiscard]] __device__ inline bool lc_accel_trace_any(LCAccel accel, LCRay ray, lc_uint mask) noexcept {
    lcrt::LocalStore store;
    lcrt::NoCallbacks cb;
    lcrt::Traversal<lcrt::MODE_ANY, false, lcrt::LocalStore, lcrt::NoCallbacks> tr;
    tr.init(lcrt::devlib_tlas(accel), reinterpret_cast<const lcrt::Instance *>(accel.instances), reinterpret_cast<const lcrt::Ray *>(&ray), mask, &store, &cb);
    tr.run();
    lcrt::devlib_report(accel, tr.overflowed());
    return tr.has_hit();
}

// motion blur is outside this backend's scope: static scenes ignore `time`
[[nodiscard]] __device__ inline LCTriangleHit lc_accel_trace_closest_motion_blur(LCAccel accel, LCRay ray, lc_float, lc_uint mask) noexcept { return lc_accel_trace_closest(accel, ray, mask); }
[[nodiscard]] __device__ inline bool lc_accel_trace_any_motion_blur(LCAccel accel, LCRay ray, lc_float, lc_uint mask) noexcept { return lc_accel_trace_any(accel, ray, mask); }

// ------------------------------------------------------------------------------------------ RayQuery
enum LCRayFlags : unsigned int {// the subset of cuda_device_resource.h:1553-1566 that selects behaviour here
    LC_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 1u << 2u,
    LC_RAY_FLAG_DISABLE_CLOSESTHIT = 1u << 3u,
};

struct LCRayQuery {
    LCAccel accel;
    LCRay ray;
    lc_float time;
    lc_uint mask;
    lc_uint flags;
    LCCommittedHit hit;
};
using LCRayQueryAll = LCRayQuery;
using LCRayQueryAny = LCRayQuery;

[[nodiscard]] __device__ inline LCRayQueryAll lc_accel_query_all(LCAccel accel, LCRay ray, lc_uint mask) noexcept {
    return LCRayQueryAll{accel, ray, 0.f, mask, LC_RAY_FLAG_DISABLE_CLOSESTHIT, LCCommittedHit{}};
}
[[nodiscard]] __device__ inline LCRayQueryAny lc_accel_query_any(LCAccel accel, LCRay ray, lc_uint mask) noexcept {
    return LCRayQueryAny{accel, ray, 0.f, mask, LC_RAY_FLAG_TERMINATE_ON_FIRST_HIT | LC_RAY_FLAG_DISABLE_CLOSESTHIT, LCCommittedHit{}};
}
[[nodiscard]] __device__ inline LCRayQueryAll lc_accel_query_all_motion_blur(LCAccel accel, LCRay ray, lc_float time, lc_uint mask) noexcept {
    return LCRayQueryAll{accel, ray, time, mask, LC_RAY_FLAG_DISABLE_CLOSESTHIT, LCCommittedHit{}};
}
[[nodiscard]] __device__ inline LCRayQueryAny lc_accel_query_any_motion_blur(LCAccel accel, LCRay ray, lc_float time, lc_uint mask) noexcept {
    return LCRayQueryAny{accel, ray, time, mask, LC_RAY_FLAG_TERMINATE_ON_FIRST_HIT | LC_RAY_FLAG_DISABLE_CLOSESTHIT, LCCommittedHit{}};
}
[[nodiscard]] __device__ inline LCCommittedHit lc_ray_query_committed_hit(LCRayQuery q) noexcept { return q.hit; }

struct LCIntersectionResult {
    lc_float t_hit{};
    lc_bool committed{};
    lc_bool terminated{};
};

// What a candidate handler can ask for (the OptiX intrinsics of :1969-2048): the candidate hit and the world-space
// ray whose t_max is the committed distance so far (the candidate's t while a triangle handler runs).
struct LCRTCandidate {
    LCTriangleHit triangle;
    LCProceduralHit procedural;
    LCRay world_ray;
};

// Outlined handlers (cuda_codegen_ast.cpp:519-535 emits their bodies between these macros): same names as the
// reference, one hidden extra parameter.
#define LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL(index) \
    [[nodiscard]] __device__ inline LCIntersectionResult lc_ray_query_procedural_intersection_##index(void *ctx_in, const LCRTCandidate &lc_rt_candidate) noexcept
#define LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL(index) \
    [[nodiscard]] __device__ inline LCIntersectionResult lc_ray_query_triangle_intersection_##index(void *ctx_in, const LCRTCandidate &lc_rt_candidate) noexcept

#define LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT(q) (lc_rt_candidate.procedural)
#define LC_RAY_QUERY_TRIANGLE_CANDIDATE_HIT(q) (lc_rt_candidate.triangle)
#define LC_RAY_QUERY_WORLD_RAY(q) (lc_rt_candidate.world_ray)
#define LC_RAY_QUERY_COMMIT_TRIANGLE(q) static_cast<void>(result.committed = true)
#define LC_RAY_QUERY_COMMIT_PROCEDURAL(q, t) \
    do {                                     \
        result.committed = true;             \
        result.t_hit = t;                    \
    } while (false)
#define LC_RAY_QUERY_TERMINATE(q) static_cast<void>(result.terminated = true)

#if LUISA_RAY_QUERY_IMPL_COUNT > 32
#error "LUISA_RAY_QUERY_IMPL_COUNT must be less than or equal to 32"
#endif

// forward declarations for every RayQuery site of the kernel (the codegen defines LUISA_RAY_QUERY_IMPL_COUNT before
// the device library, cuda_codegen_ast.cpp:552)
#define LCRT_DECL_RQ(i) LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL(i); LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL(i);
#define LCRT_CASE_RQ_TRI(i) case i##u: return lc_ray_query_triangle_intersection_##i(ctx, c);
#define LCRT_CASE_RQ_PROC(i) case i##u: return lc_ray_query_procedural_intersection_##i(ctx, c);
#define LCRT_RQ_LIST_0(X)
#define LCRT_RQ_LIST_1(X) X(0)
#define LCRT_RQ_LIST_2(X) LCRT_RQ_LIST_1(X) X(1)
#define LCRT_RQ_LIST_3(X) LCRT_RQ_LIST_2(X) X(2)
#define LCRT_RQ_LIST_4(X) LCRT_RQ_LIST_3(X) X(3)
#define LCRT_RQ_LIST_5(X) LCRT_RQ_LIST_4(X) X(4)
#define LCRT_RQ_LIST_6(X) LCRT_RQ_LIST_5(X) X(5)
#define LCRT_RQ_LIST_7(X) LCRT_RQ_LIST_6(X) X(6)
#define LCRT_RQ_LIST_8(X) LCRT_RQ_LIST_7(X) X(7)
#define LCRT_RQ_LIST_9(X) LCRT_RQ_LIST_8(X) X(8)
#define LCRT_RQ_LIST_10(X) LCRT_RQ_LIST_9(X) X(9)
#define LCRT_RQ_LIST_11(X) LCRT_RQ_LIST_10(X) X(10)
#define LCRT_RQ_LIST_12(X) LCRT_RQ_LIST_11(X) X(11)
#define LCRT_RQ_LIST_13(X) LCRT_RQ_LIST_12(X) X(12)
#define LCRT_RQ_LIST_14(X) LCRT_RQ_LIST_13(X) X(13)
#define LCRT_RQ_LIST_15(X) LCRT_RQ_LIST_14(X) X(14)
#define LCRT_RQ_LIST_16(X) LCRT_RQ_LIST_15(X) X(15)
#define LCRT_RQ_LIST_17(X) LCRT_RQ_LIST_16(X) X(16)
#define LCRT_RQ_LIST_18(X) LCRT_RQ_LIST_17(X) X(17)
#define LCRT_RQ_LIST_19(X) LCRT_RQ_LIST_18(X) X(18)
#define LCRT_RQ_LIST_20(X) LCRT_RQ_LIST_19(X) X(19)
#define LCRT_RQ_LIST_21(X) LCRT_RQ_LIST_20(X) X(20)
#define LCRT_RQ_LIST_22(X) LCRT_RQ_LIST_21(X) X(21)
#define LCRT_RQ_LIST_23(X) LCRT_RQ_LIST_22(X) X(22)
#define LCRT_RQ_LIST_24(X) LCRT_RQ_LIST_23(X) X(23)
#define LCRT_RQ_LIST_25(X) LCRT_RQ_LIST_24(X) X(24)
#define LCRT_RQ_LIST_26(X) LCRT_RQ_LIST_25(X) X(25)
#define LCRT_RQ_LIST_27(X) LCRT_RQ_LIST_26(X) X(26)
#define LCRT_RQ_LIST_28(X) LCRT_RQ_LIST_27(X) X(27)
#define LCRT_RQ_LIST_29(X) LCRT_RQ_LIST_28(X) X(28)
#define LCRT_RQ_LIST_30(X) LCRT_RQ_LIST_29(X) X(29)
#define LCRT_RQ_LIST_31(X) LCRT_RQ_LIST_30(X) X(30)
#define LCRT_RQ_LIST_32(X) LCRT_RQ_LIST_31(X) X(31)
#define LCRT_RQ_CAT2(a, b) a##b
#define LCRT_RQ_CAT(a, b) LCRT_RQ_CAT2(a, b)
#define LCRT_RQ_LIST(X) LCRT_RQ_CAT(LCRT_RQ_LIST_, LUISA_RAY_QUERY_IMPL_COUNT)(X)

LCRT_RQ_LIST(LCRT_DECL_RQ)

namespace lcrt {

// Traversal callbacks of one RayQuery site: build the candidate record, run the outlined handler selected by
// impl_tag (a literal at every call site, so the switch folds away after inlining).
struct DevlibQueryCallbacks {
    lc_uint impl_tag;
    void *ctx;
    __device__ inline LCIntersectionResult tri(const LCRTCandidate &c) const noexcept {
        switch (impl_tag) {
            LCRT_RQ_LIST(LCRT_CASE_RQ_TRI)
            default: break;
        }
        return LCIntersectionResult{};
    }
    __device__ inline LCIntersectionResult proc(const LCRTCandidate &c) const noexcept {
        switch (impl_tag) {
            LCRT_RQ_LIST(LCRT_CASE_RQ_PROC)
            default: break;
        }
        return LCIntersectionResult{};
    }
    __device__ inline static LCRay to_lc_ray(const Ray &r) noexcept {
        LCRay o;
        o.m0[0] = r.ox, o.m0[1] = r.oy, o.m0[2] = r.oz, o.m1 = r.tmin, o.m2[0] = r.dx, o.m2[1] = r.dy, o.m2[2] = r.dz, o.m3 = r.tmax;
        return o;
    }
    __device__ inline CandidateResult on_triangle(uint32_t inst, uint32_t prim, float u, float v, float t, const Ray &wr) const noexcept {
        LCRTCandidate c{};
        c.triangle = LCTriangleHit{inst, prim, lc_make_float2(u, v), t};
        c.world_ray = to_lc_ray(wr);
        const LCIntersectionResult r = tri(c);
        return CandidateResult{t, r.committed, r.terminated};
    }
    __device__ inline CandidateResult on_procedural(uint32_t inst, uint32_t prim, const Ray &wr, const Instance *, const f4, const f4) const noexcept {
        LCRTCandidate c{};
        c.procedural = LCProceduralHit{inst, prim};
        c.world_ray = to_lc_ray(wr);
        const LCIntersectionResult r = proc(c);
        return CandidateResult{r.t_hit, r.committed, r.terminated};
    }
};

}// namespace lcrt

// cuda_device_resource.h:1941-1948: runs the query to completion and leaves the committed hit in q.hit.
// One out-of-line copy of the traversal serves every RayQuery site of the kernel (the handlers are selected by impl_tag at
// run time, as the reference's __intersection__ / __anyhit__ programs do): a path tracer with a query_all and a query_any
// site is 57 KB of SASS with a traversal per site, beyond the 32 KB L1.5 instruction cache, and spent 56 % of its
// issue slots waiting for instructions.  LCRT_RAY_QUERY_INLINE=1 restores one specialised traversal per site.
// A kernel with a single RayQuery site has nothing to share: there the traversal is inlined into the site, which also lets the
// compiler keep the handlers' capture struct (`LCRayQueryCtx`, cuda_codegen_ast.cpp:446-521) in registers instead of
// round-tripping it through local memory behind the `void *ctx` of an out-of-line call.
#ifndef LCRT_RAY_QUERY_INLINE
#if LUISA_RAY_QUERY_IMPL_COUNT <= 1
#define LCRT_RAY_QUERY_INLINE 1
#else
#define LCRT_RAY_QUERY_INLINE 0
#endif
#endif
#if LCRT_RAY_QUERY_INLINE
__device__ inline
#else
__device__ __noinline__
#endif
void lc_ray_query_trace(LCRayQuery &q, lc_uint impl_tag, void *ctx) noexcept {
    lcrt::LocalStore store;
    lcrt::DevlibQueryCallbacks cb{impl_tag, ctx};
    lcrt::Traversal<lcrt::MODE_QUERY, false, lcrt::LocalStore, lcrt::DevlibQueryCallbacks> tr;
    tr.init(lcrt::devlib_tlas(q.accel), reinterpret_cast<const lcrt::Instance *>(q.accel.instances), reinterpret_cast<const lcrt::Ray *>(&q.ray), q.mask,
            &store, &cb);
    tr.terminate_on_first = (q.flags & LC_RAY_FLAG_TERMINATE_ON_FIRST_HIT) != 0u;
    tr.run_voting();
    lcrt::devlib_report(q.accel, tr.overflowed());
    tr.result(q.hit.m0, q.hit.m1, q.hit.m3, q.hit.m4, q.hit.m2.x, q.hit.m2.y);
    if (q.hit.m3 == static_cast<lc_uint>(LCHitType::MISS)) q.hit.m0 = ~0u;
}

//@inline-rayquery-api-begin  (plugin/hybrid keeps this section out of the embedded library text: its code generator emits it, in
// front of the kernels that use it; the reference's embedded array has a fixed size, plugin/hybrid/make_device_resource.py)
// ------------------------------------------------------------------------------------------ RayQuery, lowered inline
// What the inline lowering of RayQueryStmt calls (plugin/hybrid/b200_ray_query_lowering.inl emits it when the codegen is told
// to; SURVEY.md 7.7, the shape of src/backends/common/hlsl/hlsl_codegen.cpp:555-567):
//
//     { LCRayQueryInline lc_rq;
//       lc_ray_query_inline_begin(lc_rq, q);
//       while (lc_ray_query_inline_proceed(lc_rq)) {
//           LCIntersectionResult result{};
//           const LCRTCandidate &lc_rt_candidate = lc_rq.candidate;
//           if (lc_rq.triangle_candidate) { ...on_triangle_candidate body... } else { ...on_procedural_candidate body... }
//           lc_ray_query_inline_resolve(lc_rq, result);
//       }
//       lc_ray_query_inline_end(lc_rq, q); }
//
// The handler bodies are the statements the outlined form puts into LUISA_DECL_RAY_QUERY_*_IMPL functions, with the same
// LC_RAY_QUERY_* macros — but in the kernel's own scope: what they read and write are the kernel's variables, not copies in a
// capture struct behind a `void *` (cuda_codegen_ast.cpp:446-521, 562-617).  The traversal is suspended at every candidate
// (SuspendingCallbacks) and resumed by the next proceed.  With one RayQuery site per kernel proceed is inlined and the
// traversal state stays in registers across the handler; with several, ONE out-of-line copy serves them all (the instruction-
// cache lesson of profiles/r1_c5_ncu.md) and the state crosses the call in local memory.
struct LCRayQueryInline {
    lcrt::LocalStore store;
    lcrt::SuspendingCallbacks cb;
    lcrt::Traversal<lcrt::MODE_QUERY, false, lcrt::LocalStore, lcrt::SuspendingCallbacks> tr;
    LCRTCandidate candidate;
    lc_bool triangle_candidate;
};
#ifndef LCRT_RAY_QUERY_SITE_COUNT
#define LCRT_RAY_QUERY_SITE_COUNT 1
#endif
__device__ inline void lc_ray_query_inline_begin(LCRayQueryInline &s, LCRayQuery &q) noexcept {
    s.cb.pend.kind = 0;
    s.tr.init(lcrt::devlib_tlas(q.accel), reinterpret_cast<const lcrt::Instance *>(q.accel.instances), reinterpret_cast<const lcrt::Ray *>(&q.ray), q.mask,
              &s.store, &s.cb);
    s.tr.terminate_on_first = (q.flags & LC_RAY_FLAG_TERMINATE_ON_FIRST_HIT) != 0u;
}
#if LCRT_RAY_QUERY_SITE_COUNT <= 1
[[nodiscard]] __device__ inline
#else
[[nodiscard]] __device__ __noinline__
#endif
bool lc_ray_query_inline_proceed(LCRayQueryInline &s) noexcept {
    while (s.tr.step()) {
        if (!s.tr.candidate_pending()) continue;
        const lcrt::PendingCandidate &pc = s.cb.pend;
        s.triangle_candidate = pc.kind == 1;
        s.candidate = LCRTCandidate{};
        if (pc.kind == 1) s.candidate.triangle = LCTriangleHit{pc.inst, pc.prim, lc_make_float2(pc.u, pc.v), pc.t};
        else s.candidate.procedural = LCProceduralHit{pc.inst, pc.prim};
        s.candidate.world_ray = lcrt::DevlibQueryCallbacks::to_lc_ray(s.tr.candidate_ray());
        return true;
    }
    return false;
}
__device__ inline void lc_ray_query_inline_resolve(LCRayQueryInline &s, const LCIntersectionResult &r) noexcept {
    s.tr.resolve(r.committed, r.terminated, r.t_hit);
}
__device__ inline void lc_ray_query_inline_end(LCRayQueryInline &s, LCRayQuery &q) noexcept {
    lcrt::devlib_report(q.accel, s.tr.overflowed());
    s.tr.result(q.hit.m0, q.hit.m1, q.hit.m3, q.hit.m4, q.hit.m2.x, q.hit.m2.y);
    if (q.hit.m3 == static_cast<lc_uint>(LCHitType::MISS)) q.hit.m0 = ~0u;
}
//@inline-rayquery-api-end
