// lcrt_common.h — wire layouts, BVH8 / primitive record formats and portable intrinsics.
//
// Everything in this header compiles both under nvcc (device code for sm_100a) and under plain g++
// (tests/hostsim: the same per-thread functions executed serially on the CPU to validate the builder and the
// traversal logic without a GPU).  The host build is test infrastructure only; the product path is CUDA.
//
// Wire layouts follow the reference (paths relative to /root/reference):
//   Ray            include/luisa/runtime/rtx/ray.h:10-16, cuda_device_resource.h:1279-1284
//   SurfaceHit     include/luisa/runtime/rtx/hit.h:30-37,  cuda_device_resource.h:1286-1291
//   CommittedHit   include/luisa/runtime/rtx/hit.h:18-25,  cuda_device_resource.h:1304-1313
//   Modification   include/luisa/runtime/rhi/command.h:534-552
//   Instance       src/backends/cuda/cuda_builtin/cuda_builtin_kernels.cu:3-15 (== optix::Instance, LCAccelInstance)
//   AccelBinding   src/backends/cuda/cuda_accel.h:29-32
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define LCRT_HD __host__ __device__ __forceinline__
#define LCRT_HD_NOINLINE __host__ __device__
#define LCRT_HD_COLD __host__ __device__ __noinline__
#else
#define LCRT_HD inline
#define LCRT_HD_NOINLINE inline
#define LCRT_HD_COLD inline
#endif

namespace lcrt {

// ------------------------------------------------------------------------------------------ vector types
#if defined(__CUDACC__)
typedef ::float4 f4;
typedef ::uint4 u4;
typedef ::uint2 u2;
#else
struct alignas(16) f4 {
    float x, y, z, w;
};
struct alignas(16) u4 {
    uint32_t x, y, z, w;
};
struct alignas(8) u2 {
    uint32_t x, y;
};
#endif
LCRT_HD f4 make_f4(float x, float y, float z, float w) {
    f4 r;
    r.x = x, r.y = y, r.z = z, r.w = w;
    return r;
}
LCRT_HD u4 make_u4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    u4 r;
    r.x = x, r.y = y, r.z = z, r.w = w;
    return r;
}
LCRT_HD u2 make_u2(uint32_t x, uint32_t y) {
    u2 r;
    r.x = x, r.y = y;
    return r;
}

// ------------------------------------------------------------------------------------------ intrinsics
LCRT_HD uint32_t as_uint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
LCRT_HD float as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
LCRT_HD int bfind(uint32_t x) {// index of the most significant set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
LCRT_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
LCRT_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
// replicate the sign bit of each byte across the byte (prmt with sign-replicate selectors)
LCRT_HD uint32_t sign_extend_s8x4(uint32_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
#else
    return ((x >> 7) & 0x01010101u) * 0xffu;
#endif
}
LCRT_HD uint32_t extract_byte(uint32_t x, int i) { return (x >> (8 * i)) & 0xffu; }

// copy n 16-byte vectors (dst, src 16-byte aligned): vector stores on the device, memcpy on the host (aliasing-safe)
LCRT_HD void copy_vec16(void *dst, const void *src, int n) {
#if defined(__CUDA_ARCH__)
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int k = 0; k < n; k++) d[k] = s[k];
#else
    memcpy(dst, src, 16 * (size_t)n);
#endif
}

// no-contraction arithmetic: the watertight edge functions rely on a*b - c*d being exactly antisymmetric
LCRT_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
LCRT_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b;
    return r;
#endif
}

// ------------------------------------------------------------------------------------------ wire layouts
struct alignas(16) Ray {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
};
static_assert(sizeof(Ray) == 32, "Ray must be 32 bytes");

struct alignas(8) SurfaceHit {// miss <=> inst == ~0u
    uint32_t inst, prim;
    float u, v;// weights of v1, v2
    float t;
    uint32_t pad;
};
static_assert(sizeof(SurfaceHit) == 24, "SurfaceHit must be 24 bytes");

enum HitType : uint32_t { HIT_MISS = 0u, HIT_SURFACE = 1u, HIT_PROCEDURAL = 2u };

struct alignas(8) CommittedHit {
    uint32_t inst, prim;
    float u, v;
    uint32_t hit_type;
    float t;
};
static_assert(sizeof(CommittedHit) == 24, "CommittedHit must be 24 bytes");

struct alignas(16) Modification {
    uint32_t index, user_id, flags, vis_mask;
    float affine[12];// row-major 3x4
    uint64_t primitive;
    uint64_t _pad;
};
static_assert(sizeof(Modification) == 80, "Modification must be 80 bytes");

enum ModFlags : uint32_t {
    MOD_PRIMITIVE = 1u << 0,
    MOD_TRANSFORM = 1u << 1,
    MOD_OPAQUE_ON = 1u << 2,
    MOD_OPAQUE_OFF = 1u << 3,
    MOD_VISIBILITY = 1u << 4,
    MOD_USER_ID = 1u << 5,
    // backend-internal (cuda_accel.cpp:186-190): set by the host encoder from the primitive's tag
    MOD_PROCEDURAL = 1u << 8,
};

enum InstanceFlags : uint32_t {// cuda_builtin_kernels.cu:31-37
    INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING = 1u << 0,
    INSTANCE_FLAG_FLIP_TRIANGLE_FACING = 1u << 1,
    INSTANCE_FLAG_DISABLE_ANYHIT = 1u << 2,// opaque
    INSTANCE_FLAG_ENFORCE_ANYHIT = 1u << 3,// non-opaque
};

struct alignas(16) Instance {
    float affine[12];// 3 rows x float4, object -> world
    uint32_t user_id, sbt_offset, mask, flags;
    uint64_t handle;// BLAS blob (device address of its BlobHeader)
    uint64_t pad;
};
static_assert(sizeof(Instance) == 80, "Instance must be 80 bytes");

struct alignas(16) InstanceHandleUpdate {// cuda_builtin_kernels.cu:26-29
    uint64_t index, handle;
};

struct alignas(16) AccelBinding {// what a kernel receives for an Accel argument (LCAccel)
    uint64_t handle;// TLAS blob
    Instance *instances;
};
static_assert(sizeof(AccelBinding) == 16, "AccelBinding must be 16 bytes");

// ------------------------------------------------------------------------------------------ BVH8 node (80 B)
// Compressed wide node after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs" (HPG 2017).  Five 16-byte vectors:
//   n0 = { p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24 }
//   n1 = { child_base, prim_base, meta[0..3], meta[4..7] }
//   n2 = { qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7] }
//   n3 = { qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7] }
//   n4 = { qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7] }
// meta: empty slot 0; internal child (1 << 5) | (24 + slot); leaf (unary count 1/3/7 << 5) | first prim offset (0..23).
// Child box i spans p + q * 2^(e - 127), e stored biased so the scale is as_float(e << 23).
struct alignas(16) Node8 {
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t child_base, prim_base;
    uint8_t meta[8];
    uint8_t qlo_x[8], qlo_y[8], qlo_z[8];
    uint8_t qhi_x[8], qhi_y[8], qhi_z[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");
constexpr uint32_t NODE_BYTES = 80;

// ------------------------------------------------------------------------------------------ primitive records
struct alignas(16) TriRecord {// 48 B: three 16-byte vertex vectors, primitive id in v0.w
    float v0x, v0y, v0z;
    uint32_t prim;
    float v1x, v1y, v1z;
    uint32_t pad1;
    float v2x, v2y, v2z;
    uint32_t pad2;
};
static_assert(sizeof(TriRecord) == 48, "TriRecord must be 48 bytes");

struct alignas(16) BoxRecord {// 32 B: procedural primitive
    float lox, loy, loz;
    uint32_t prim;
    float hix, hiy, hiz;
    uint32_t pad;
};
static_assert(sizeof(BoxRecord) == 32, "BoxRecord must be 32 bytes");

enum BlasKind : uint32_t { KIND_TRIANGLES = 0u, KIND_AABBS = 1u, KIND_INSTANCES = 2u };

struct alignas(16) InstanceRecord {// 256 B = two 128-byte lines: TLAS leaf, everything the traversal needs to enter a BLAS
    // ---- hot line (80 B used): all that an instance entry fetches, five 16-byte vectors
    float inv[12];     // world -> object, row-major 3x4, fp32 (rounded from inv_exact)
    uint64_t nodes;    // device address of the BLAS's Node8 array
    uint64_t prims;    // device address of the BLAS's primitive records
    uint32_t inst;     // index into the 80-byte instance array
    uint32_t mask;     // visibility mask (8 bits)
    uint32_t flags;    // bits 0..7: BlasKind of the target, bit 8: opaque, bit 9: the transform is exactly the identity
    uint32_t pad;
    uint32_t spare[12];// LCRT_RECENTER: spare[0..2] = the instance's world translation (fp32 bits), see Traversal::setup_ray
    // ---- cold line: only the (rare) double-precision fallback of the triangle test reads it
    double inv_exact[12];// world -> object in double: the fallback re-derives the object-space ray from the world ray
                         // with it, so its decisions and t do not inherit fp32 transform rounding
    uint64_t spare2[4];
};
static_assert(sizeof(InstanceRecord) == 256, "InstanceRecord must be 256 bytes");
// LCRT_RECENTER=1 (off by default: not yet measured on a GPU, it changes the traversal kernels): the object-space ray origin is
// computed as inv_linear * (o - T) instead of inv * o + inv_translation.  With an instance far from the origin the two large
// terms of the second form are rounded separately and cancel (error ~ |T| * 2^-24, 0.004 at |T| = 1e5); in the first form o - T is
// exact when o and T are close (Sterbenz) and the error is relative to the SMALL difference.  tests/hostsim builds both forms;
// tests/test_hostsim.py::test_general_affine_instances_far_from_the_origin shows the difference.
#ifndef LCRT_RECENTER
#define LCRT_RECENTER 0
#endif
constexpr uint32_t INSTANCE_FETCH_BYTES = 80;// algorithmic bytes of one instance entry (roofline)
constexpr uint32_t IREC_KIND_MASK = 0xffu, IREC_OPAQUE = 1u << 8, IREC_IDENTITY = 1u << 9;

// ------------------------------------------------------------------------------------------ blob header
// Every BLAS / TLAS lives in ONE caller-provided allocation ("blob"): header, nodes, primitive records, aux.
// All internal references are offsets, so a BLAS blob can be copied / NCCL-broadcast and used at another address.
constexpr uint32_t BLOB_MAGIC = 0x4c435254u;// 'LCRT'
struct alignas(16) BlobHeader {
    uint32_t magic, kind, prim_count, node_capacity;
    uint32_t node_count, prim_written, error, built;// device-written by the build
    float lo[3];
    uint32_t flags;
    float hi[3];
    uint32_t prim_stride;
    uint64_t nodes_offset, prims_offset;
    uint64_t parent_offset, total_bytes;// parent_offset: u32 per node, (parent << 3) | slot, ~0 for the root
    uint32_t plant_magic;// PLANT_MAGIC (lcrt_traverse.h): the inline entry points load it from here so that the compiler can
                         // not fold it, see Traversal::init(const BlobHeader *, ...)
    uint32_t trace_error;// set (atomicOr) by a traversal that had to give up on a ray: bit 0 = stack overflow (lcrt_traverse.h push())
    uint32_t depth;      // BVH8 levels of this structure (device-written by the build)
    uint32_t max_stack;  // TLAS: depth + 2 + deepest instanced BLAS depth = the most stack entries a ray can need; BLAS: depth
    uint64_t reserved[2];
};
static_assert(sizeof(BlobHeader) == 128, "BlobHeader must be 128 bytes");

enum BuildFlags : uint32_t {
    BUILD_FAST_TRACE = 0u,
    BUILD_FAST_BUILD = 1u << 0,
    BUILD_ALLOW_UPDATE = 1u << 1,
    BUILD_ALLOW_COMPACTION = 1u << 2,
};

LCRT_HD uint32_t node_capacity_for(uint32_t prim_count) {
    // W <= S/2 + S/14 + 2 for a collapse in which every non-full wide node has only leaf children
    return (uint32_t)(((uint64_t)prim_count * 4u + 6u) / 7u) + 2u;
}

struct Box {
    float lo[3], hi[3];
};
LCRT_HD Box box_empty() {
    Box b;
    for (int a = 0; a < 3; a++) b.lo[a] = INFINITY, b.hi[a] = -INFINITY;
    return b;
}
LCRT_HD void box_grow(Box &b, const Box &o) {
    for (int a = 0; a < 3; a++) b.lo[a] = fminf(b.lo[a], o.lo[a]), b.hi[a] = fmaxf(b.hi[a], o.hi[a]);
}
LCRT_HD void box_grow(Box &b, float x, float y, float z) {
    b.lo[0] = fminf(b.lo[0], x), b.hi[0] = fmaxf(b.hi[0], x);
    b.lo[1] = fminf(b.lo[1], y), b.hi[1] = fmaxf(b.hi[1], y);
    b.lo[2] = fminf(b.lo[2], z), b.hi[2] = fmaxf(b.hi[2], z);
}
LCRT_HD float box_half_area(const Box &b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}

}// namespace lcrt
