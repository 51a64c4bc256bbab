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

// ---- closest / any-hit for generated kernels: one thread, one ray — and idle lanes that help.
// A generated kernel calls lc_accel_trace_closest per thread; the warp leaves the call when its LONGEST ray is done, and
// on incoherent rays most lanes spend most of that time idle (the batch kernels refill idle lanes with new rays; a
// generated kernel has none to give).  Here an idle lane takes over the OLDEST pending subtree of a busy lane's traversal
// stack — the entry at the bottom, the farthest and usually largest piece of work that ray still has — together with the
// ray, traverses it on its own stack, and reports what it found to the lane that owns the ray:
//   * every lane keeps the result of ITS ray apart from the traversal it happens to be running (SharedBest): exact distance
//     (double, the oracle's formula), barycentrics, ids.  A finished (sub)traversal evaluates its best hit exactly with the
//     ray it holds and hands it to the owner, who keeps the smaller distance, ties to the lower (instance, primitive) —
//     the rule the single-lane traversal applies, so the result does not depend on who traversed what;
//   * a reported hit also shrinks the culling distance of every lane still working on that ray (any-hit: ends them);
//   * entries are given away from the bottom of the stack (Traversal::bottom), so the donor's own order — near subtrees
//     first — is untouched; an entry below the instance boundary restarts in the TLAS with the world ray, one above it
//     carries the donor's object-space ray, shear rows and instance context (24 more shuffles).
// All exchanges are warp shuffles between the lanes that entered the call together; nothing is shared through memory.
#ifndef LCRT_TRACE_SHARED
#define LCRT_TRACE_SHARED 1
#endif
#ifndef LCRT_SHARE_PERIOD
#define LCRT_SHARE_PERIOD 4    // iterations between two attempts to pair idle lanes with donors
#endif
#ifndef LCRT_SHARE_MIN_ENTRIES
#define LCRT_SHARE_MIN_ENTRIES 1// stack entries a lane must have pending to give its oldest one away
#endif
namespace lcrt {
struct SharedBest {
    double td;
    float u, v;
    uint32_t inst, prim;
    bool hit;
};
__device__ inline double devlib_shfl_f64(uint32_t group, double x, int src) noexcept {
    const int hi = __shfl_sync(group, __double2hiint(x), src), lo = __shfl_sync(group, __double2loint(x), src);
    return __hiloint2double(hi, lo);
}
__device__ inline float devlib_shfl_f32(uint32_t group, float x, int src) noexcept { return __shfl_sync(group, x, src); }

template<int MODE>
__device__ __noinline__ void trace_shared(const BlobHeader *tlas, const Instance *instances, const Ray *own_ray, uint32_t mask, SharedBest *out,
                                          bool *overflowed) noexcept {
    using T = Traversal<MODE, false, SharedLocalStore, NoCallbacks>;
    SharedLocalStore store;
    NoCallbacks cb;
    T tr;
    Ray cur = *own_ray;// the ray this lane is working on: its own first, other lanes' later
    tr.init(tlas, instances, &cur, mask, &store, &cb);
    *overflowed = tr.overflowed();
    const uint32_t group = __activemask();
    uint32_t lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t owner = lane;
    SharedBest best;
    best.hit = false, best.td = 0.0, best.u = best.v = 0.f, best.inst = best.prim = ~0u;
    bool task = !tr.done;// this lane is running a (sub)traversal
    for (uint32_t iter = 0u;; iter++) {
        // ---- 1. (sub)traversals that just ended report their best hit to the lane that owns the ray
        const bool fin = task && tr.done;
        bool c_hit = false;
        double c_td = 0.0;
        float c_u = 0.f, c_v = 0.f;
        uint32_t c_inst = ~0u, c_prim = ~0u;
        if (fin) {
            task = false;
            if (tr.has_hit()) {
                c_hit = true, c_inst = tr.hit_inst();
                if (MODE == MODE_CLOSEST) {
                    const InstanceRecord *r = tr.tlas_recs + tr.hit_rec();
                    const f4 *tri = reinterpret_cast<const f4 *>(reinterpret_cast<const char *>(r->prims) + (uint64_t)tr.hit_pi() * sizeof(TriRecord));
                    c_prim = as_uint(tri[0].w);
                    if (!tri_exact_world(&cur, r, tri, c_td, c_u, c_v)) c_td = (double)tr.hit_t(), c_u = c_v = 0.f;// (the fp32 bounds exclude this)
                }
            }
        }
        uint32_t fm = __ballot_sync(group, fin && c_hit);
        while (fm) {
            const int h = __ffs((int)fm) - 1;
            fm &= fm - 1u;
            const uint32_t o = __shfl_sync(group, owner, h);
            if (MODE == MODE_ANY) {
                if (lane == o) best.hit = true;
                if (task && owner == o) tr.done = true, task = false;// any hit ends every lane working on that ray
            } else {
                const double td = devlib_shfl_f64(group, c_td, h);
                const float u = devlib_shfl_f32(group, c_u, h), v = devlib_shfl_f32(group, c_v, h);
                const uint32_t inst = __shfl_sync(group, c_inst, h), prim = __shfl_sync(group, c_prim, h);
                if (lane == o && (!best.hit || td < best.td || (td == best.td && (inst < best.inst || (inst == best.inst && prim < best.prim)))))
                    best.hit = true, best.td = td, best.u = u, best.v = v, best.inst = inst, best.prim = prim;
                if (task && owner == o) {
                    const float t = (float)td;
                    tr.tmax = fminf(tr.tmax, fmaf(fabsf(t), 3.6e-7f, t));// as Traversal::accept: a hair beyond, exact ties must still be visited
                }
            }
        }
        // ---- 2. done when nobody is running anything
        const uint32_t busy = __ballot_sync(group, task);
        if (busy == 0u) break;
        // ---- 3. idle lanes take the oldest pending subtree of busy lanes
        const uint32_t idle = group & ~busy;
        if (idle != 0u && (iter % (uint32_t)LCRT_SHARE_PERIOD) == 0u) {
            const bool can_give = task && !tr.done && tr.sp - tr.bottom >= LCRT_SHARE_MIN_ENTRIES;
            const uint32_t donors = __ballot_sync(group, can_give);
            if (donors != 0u) {
                const int pairs = min(__popc(idle), __popc(donors));
                const int rank = task ? __popc(donors & lt) : __popc(idle & lt);
                const bool taker = !task && rank < pairs, giver = can_give && rank < pairs;
                const int src = taker ? (int)__fns(donors, 0u, rank + 1) : (int)lane;
                u2 e = make_u2(0u, 0u);
                uint32_t in_tlas = 1u;
                if (giver) {
                    e = store.load(tr.bottom);
                    in_tlas = (tr.blas_sp < 0 || tr.bottom < tr.blas_sp) ? 1u : 0u;
                    tr.bottom++;
                }
                e.x = __shfl_sync(group, e.x, src), e.y = __shfl_sync(group, e.y, src);
                in_tlas = __shfl_sync(group, in_tlas, src);
                const uint32_t n_owner = __shfl_sync(group, owner, src), n_mask = __shfl_sync(group, tr.ray_mask, src);
                const float n_tmax = devlib_shfl_f32(group, tr.tmax, src);
                Ray nr;
                nr.ox = devlib_shfl_f32(group, cur.ox, src), nr.oy = devlib_shfl_f32(group, cur.oy, src), nr.oz = devlib_shfl_f32(group, cur.oz, src);
                nr.tmin = devlib_shfl_f32(group, cur.tmin, src), nr.dx = devlib_shfl_f32(group, cur.dx, src), nr.dy = devlib_shfl_f32(group, cur.dy, src);
                nr.dz = devlib_shfl_f32(group, cur.dz, src), nr.tmax = devlib_shfl_f32(group, cur.tmax, src);
                if (taker) {
                    cur = nr, owner = n_owner, task = true;
                    tr.tmin = nr.tmin, tr.tmax = n_tmax, tr.ray_mask = n_mask;
                    tr.sp = 0, tr.bottom = 0, tr.tp = 0u, tr.done = false;
                    if (e.y & 0xff000000u) tr.G = e, tr.Gt = make_u2(0u, 0u);
                    else tr.Gt = e, tr.G = make_u2(0u, 0u);
                    store.stw(CW_HIT_T, as_uint(nr.tmax)), store.stw(CW_HIT_TERR, as_uint(-1.0f));
                    store.stw(CW_HIT_INST, ~0u), store.stw(CW_HIT_PRIM, ~0u), store.stw(CW_HIT_TYPE, HIT_MISS), store.stw(CW_HIT_EXACT, 0u);
                    // a TLAS entry: the world ray is set up when the first node needs it
                    tr.blas_sp = -1, tr.nodes = tr.tlas_nodes, tr.prims = reinterpret_cast<const char *>(tr.tlas_recs);
                    tr.cur_inst = ~0u, tr.cur_flags = 0u, tr.octinv = T::RAY_PENDING;
                    store.stw(CW_REC, 0u);
                }
                if (__any_sync(group, taker && in_tlas == 0u)) {
                    // an entry inside an instance: the giver's object-space ray and instance context come along
                    const float b_ox = devlib_shfl_f32(group, tr.ox, src), b_oy = devlib_shfl_f32(group, tr.oy, src), b_oz = devlib_shfl_f32(group, tr.oz, src);
                    const float b_ix = devlib_shfl_f32(group, tr.idx, src), b_iy = devlib_shfl_f32(group, tr.idy, src), b_iz = devlib_shfl_f32(group, tr.idz, src);
                    const uint32_t b_oct = __shfl_sync(group, tr.octinv, src), b_inst = __shfl_sync(group, tr.cur_inst, src), b_flags = __shfl_sync(group, tr.cur_flags, src);
                    const unsigned long long b_nodes = __shfl_sync(group, (unsigned long long)reinterpret_cast<uintptr_t>(tr.nodes), src);
                    const unsigned long long b_prims = __shfl_sync(group, (unsigned long long)reinterpret_cast<uintptr_t>(tr.prims), src);
                    f4 r1 = store.ld4(0), r2 = store.ld4(1);
                    r1.x = devlib_shfl_f32(group, r1.x, src), r1.y = devlib_shfl_f32(group, r1.y, src), r1.z = devlib_shfl_f32(group, r1.z, src), r1.w = devlib_shfl_f32(group, r1.w, src);
                    r2.x = devlib_shfl_f32(group, r2.x, src), r2.y = devlib_shfl_f32(group, r2.y, src), r2.z = devlib_shfl_f32(group, r2.z, src), r2.w = devlib_shfl_f32(group, r2.w, src);
                    const uint32_t b_sz = __shfl_sync(group, store.ldw(CW_SZ), src), b_kz = __shfl_sync(group, store.ldw(CW_KZ), src), b_rec = __shfl_sync(group, store.ldw(CW_REC), src);
                    if (taker && in_tlas == 0u) {
                        tr.ox = b_ox, tr.oy = b_oy, tr.oz = b_oz, tr.idx = b_ix, tr.idy = b_iy, tr.idz = b_iz, tr.octinv = b_oct;
                        tr.cur_inst = b_inst, tr.cur_flags = b_flags;
                        tr.nodes = reinterpret_cast<const u4 *>(static_cast<uintptr_t>(b_nodes)), tr.prims = reinterpret_cast<const char *>(static_cast<uintptr_t>(b_prims));
                        tr.blas_sp = 0;
                        store.st4(0, r1), store.st4(1, r2), store.stw(CW_SZ, b_sz), store.stw(CW_KZ, b_kz), store.stw(CW_REC, b_rec);
                    }
                }
            }
        }
        if (task) tr.step();
    }
    *out = best;
}
}// namespace lcrt

[[nodiscard]] __device__ inline LCTriangleHit lc_accel_trace_closest(LCAccel accel, LCRay ray, lc_uint mask) noexcept {
#if LCRT_TRACE_SHARED
    lcrt::SharedBest best;
    bool overflowed;
    lcrt::trace_shared<lcrt::MODE_CLOSEST>(lcrt::devlib_tlas(accel), reinterpret_cast<const lcrt::Instance *>(accel.instances),
                                           reinterpret_cast<const lcrt::Ray *>(&ray), mask, &best, &overflowed);
    lcrt::devlib_report(accel, overflowed);
    LCTriangleHit hit;
    hit.m0 = best.hit ? best.inst : ~0u, hit.m1 = best.prim, hit.m2 = lc_make_float2(best.u, best.v), hit.m3 = best.hit ? (float)best.td : ray.m3;
    return hit;// miss <=> m0 == ~0u (include/luisa/dsl/rtx/hit.h:67)
#else
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
#endif
}

[[nodiscard]] __device__ inline bool lc_accel_trace_any(LCAccel accel, LCRay ray, lc_uint mask) noexcept {
#if LCRT_TRACE_SHARED
    lcrt::SharedBest best;
    bool overflowed;
    lcrt::trace_shared<lcrt::MODE_ANY>(lcrt::devlib_tlas(accel), reinterpret_cast<const lcrt::Instance *>(accel.instances),
                                       reinterpret_cast<const lcrt::Ray *>(&ray), mask, &best, &overflowed);
    lcrt::devlib_report(accel, overflowed);
    return best.hit;
#else
    lcrt::LocalStore store;
    lcrt::NoCallbacks cb;
    lcrt::Traversal<lcrt::MODE_ANY, false, lcrt::LocalStore, lcrt::NoCallbacks> tr;
    tr.init(lcrt::devlib_tlas(accel), reinterpret_cast<const lcrt::Instance *>(accel.instances), reinterpret_cast<const lcrt::Ray *>(&ray), mask, &store, &cb);
    tr.run();
    lcrt::devlib_report(accel, tr.overflowed());
    return tr.has_hit();
#endif
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
