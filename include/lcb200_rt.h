/* lcb200_rt.h — C-ABI of the B200-native ray-tracing acceleration library (liblcb200_rt.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types.  Each entry point replaces one
 * call the reference's CUDA backend makes into OptiX's function table or one of its own built-in kernels
 * (paths relative to /root/reference):
 *
 *   lcrtAccelComputeMemoryUsage  <- optixAccelComputeMemoryUsage   src/backends/cuda/cuda_primitive.cpp:31-36, cuda_accel.cpp:45-50
 *   lcrtBuildTriangles           <- optixAccelBuild (triangles)     cuda_primitive.cpp:57,99 with cuda_mesh.cpp:13-29 build input
 *   lcrtBuildAabbs               <- optixAccelBuild (custom prims)  cuda_primitive.cpp:57,99 with cuda_procedural_primitive.cpp:13-23
 *   lcrtRefitTriangles/Aabbs     <- optixAccelBuild(OPERATION_UPDATE) cuda_primitive.cpp:115-132
 *   lcrtInstancesApplyMods       <- kernel update_accel             cuda_builtin/cuda_builtin_kernels.cu:52-111 (cuda_accel.cpp:234-238)
 *   lcrtInstancesUpdateHandles   <- kernel update_accel_instance_handles  cuda_builtin_kernels.cu:42-50 (cuda_accel.cpp:300-306)
 *   lcrtBuildTlas / lcrtRefitTlas<- optixAccelBuild (instances)     cuda_accel.cpp:72,102 / :136
 *   lcrtTraceClosest / Any       <- optixLaunch of a raygen calling lc_accel_trace_closest / _any
 *                                   cuda_shader_optix.cpp:478-483, cuda_device_resource.h:1783-1829
 *   lcrtQuery                    <- lc_ray_query_trace + __intersection__/__anyhit__ray_query
 *                                   cuda_device_resource.h:1934-1941, 2146-2385 (with the fixed candidate programs below)
 *   lcrtDeviceLibrarySource      <- the embedded device-library text the codegen pastes into every kernel
 *                                   (cuda_compiler.cpp:195-215): here it defines lc_accel_trace_closest & co. over BVH8
 *
 * Memory model (as OptiX): the caller owns all device memory.  A BLAS or TLAS is ONE allocation ("blob") of
 * `blob_bytes`; builds need `build_temp_bytes` of scratch that may be reused after the call's work has completed on
 * `stream`.  All work is stream-ordered; no entry point synchronises the host.  Handles are device addresses of
 * blobs.  BLAS blobs are position independent (may be copied / broadcast); a TLAS stores device addresses of the
 * BLASes it instances and must be (re)built on the device that uses it.
 *
 * Wire layouts (all little endian): Ray 32 B, SurfaceHit 24 B, CommittedHit 24 B, AABB 24 B, Triangle 3 x u32,
 * Modification 80 B, Instance 80 B — identical to the reference.
 *
 * Return value: 0 on success, a negative LCRT_ERROR_* otherwise.  There is NO CPU fallback.
 */
#ifndef LCB200_RT_H
#define LCB200_RT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LCRT_API __declspec(dllexport)
#else
#define LCRT_API __attribute__((visibility("default")))
#endif

typedef void *lcrtStream; /* cudaStream_t / CUstream */

enum {
    LCRT_SUCCESS = 0,
    LCRT_ERROR_INVALID_VALUE = -1,
    LCRT_ERROR_BUFFER_TOO_SMALL = -2,
    LCRT_ERROR_CUDA = -3,
    LCRT_ERROR_NOT_UPDATABLE = -4,
    LCRT_ERROR_MISALIGNED = -5,
    LCRT_ERROR_BUILD_FAILED = -6 /* a build / refit / traversal flagged an error on the device (lcrtBlobInfo.error, .trace_error) */
};

enum { LCRT_KIND_TRIANGLES = 0, LCRT_KIND_AABBS = 1, LCRT_KIND_INSTANCES = 2 };

/* AccelOption (include/luisa/runtime/rhi/resource.h:94-126) as build flags */
enum {
    LCRT_BUILD_FAST_TRACE = 0,
    LCRT_BUILD_FAST_BUILD = 1 << 0,
    LCRT_BUILD_ALLOW_UPDATE = 1 << 1,
    LCRT_BUILD_ALLOW_COMPACTION = 1 << 2
};

/* fixed RayQuery candidate programs for lcrtQuery (arbitrary handlers go through the device library) */
enum { LCRT_TRI_COMMIT_ALL = 0, LCRT_TRI_CUTOUT_MOD3 = 1, LCRT_TRI_IGNORE_ALL = 2, LCRT_TRI_TERMINATE_FIRST = 3 };
enum { LCRT_PROC_SPHERE = 0, LCRT_PROC_BOX = 1, LCRT_PROC_IGNORE_ALL = 2 };

typedef struct lcrtSizes {
    uint64_t blob_bytes;
    uint64_t build_temp_bytes;
    uint64_t update_temp_bytes;
} lcrtSizes;

LCRT_API int lcrtGetVersion(void);
LCRT_API const char *lcrtGetErrorString(int code);
LCRT_API const char *lcrtGetLastCudaError(void);

LCRT_API int lcrtAccelComputeMemoryUsage(uint32_t kind, uint32_t prim_count, uint32_t flags, lcrtSizes *out);

/* vertices: float3 at `vertex_stride` bytes (stride % 4 == 0); triangles: 3 x u32 per triangle (12-byte stride) */
LCRT_API int lcrtBuildTriangles(lcrtStream stream, const void *vertices, uint64_t vertex_stride, const void *triangles,
                                uint32_t triangle_count, uint32_t flags, void *temp, uint64_t temp_bytes, void *blob,
                                uint64_t blob_bytes);
LCRT_API int lcrtRefitTriangles(lcrtStream stream, const void *vertices, uint64_t vertex_stride, const void *triangles,
                                uint32_t triangle_count, void *temp, uint64_t temp_bytes, void *blob);

/* aabbs: 24-byte records {min[3], max[3]} */
LCRT_API int lcrtBuildAabbs(lcrtStream stream, const void *aabbs, uint32_t aabb_count, uint32_t flags, void *temp,
                            uint64_t temp_bytes, void *blob, uint64_t blob_bytes);
LCRT_API int lcrtRefitAabbs(lcrtStream stream, const void *aabbs, uint32_t aabb_count, void *temp, uint64_t temp_bytes, void *blob);

/* instance buffer (80-byte records) maintenance */
LCRT_API int lcrtInstancesInit(lcrtStream stream, void *instances, uint32_t first, uint32_t count);
LCRT_API int lcrtInstancesApplyMods(lcrtStream stream, void *instances, const void *mods_device, uint32_t mod_count);
/* the same for at most 16 modifications in HOST memory: they travel in the kernel's launch parameters, no staging copy */
LCRT_API int lcrtInstancesApplyModsHost(lcrtStream stream, void *instances, const void *mods_host, uint32_t mod_count);
LCRT_API int lcrtInstancesUpdateHandles(lcrtStream stream, void *instances, const void *updates_device, uint32_t count);

LCRT_API int lcrtBuildTlas(lcrtStream stream, const void *instances, uint32_t instance_count, uint32_t flags, void *temp,
                           uint64_t temp_bytes, void *blob, uint64_t blob_bytes);
LCRT_API int lcrtRefitTlas(lcrtStream stream, const void *instances, uint32_t instance_count, void *temp, uint64_t temp_bytes,
                           void *blob);

/* Batch tracing.  masks: device array of `ray_count` u32 if masks != NULL, else `mask` for every ray.
 * hits: SurfaceHit[ray_count] (miss <=> inst == 0xffffffff).  occluded: u32[ray_count] (0 / 1).
 * counters (optional, device, 4 x u64): += BVH8 nodes visited, primitive records tested, instance leaves entered,
 * double-precision triangle evaluations taken by the fp32 filter's fallback. */
LCRT_API int lcrtTraceClosest(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count,
                              const uint32_t *masks, uint32_t mask, void *hits, uint64_t *counters);
LCRT_API int lcrtTraceAny(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count,
                          const uint32_t *masks, uint32_t mask, uint32_t *occluded, uint64_t *counters);
/* lcrtTraceClosest whose final records go into a framebuffer of tiles instead of hits[]: ray i -> dst + (i / tile_rays) *
 * tile_stride_bytes + (i % tile_rays) * 24 (tile_rays a multiple of 128).  scratch_hits: SurfaceHit[ray_count] on this device (the traversal's provisional
 * records).  dst may be another GPU's memory mapped into this process (lcbFramebufferOpen): tile-sharded multi-GPU tracing
 *  writes straight into the gathering rank's ray-ordered framebuffer over NVLink from the same launches, with no
 * gather collective and no reorder copy behind the trace.  Nothing like it exists in the reference (no multi-GPU path at all).
 * flags: 0 = the records of each 128-ray batch are collected (in scratch_hits) and stored together when the batch is complete
 * (512-byte warp stores: what a peer's NVLink ingress handles well when several GPUs write into one);
 * LCRT_TILED_STORE_EACH_RECORD = every record is stored on its own when its ray ends (less work per ray; fine for one or two writers). */
#define LCRT_TILED_STORE_EACH_RECORD 1u
LCRT_API int lcrtTraceClosestTiled(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count,
                                   const uint32_t *masks, uint32_t mask, void *scratch_hits, void *dst, uint64_t tile_rays,
                                   uint64_t tile_stride_bytes, uint32_t flags, uint64_t *counters);
/* RayQuery (query_all, or query_any if terminate_on_first) with fixed candidate programs; hits: CommittedHit[ray_count] */
LCRT_API int lcrtQuery(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count,
                       const uint32_t *masks, uint32_t mask, int terminate_on_first, uint32_t tri_program,
                       uint32_t proc_program, void *hits);

/* blob introspection (stream-ordered device -> host copy of the 128-byte header; synchronises `stream`) */
typedef struct lcrtBlobInfo {
    uint32_t kind, prim_count, node_count, node_capacity;
    uint32_t built, error, flags, prim_stride;
    float lo[3], hi[3];
    uint64_t nodes_offset, prims_offset, total_bytes;
    uint32_t depth;       /* BVH8 levels */
    uint32_t max_stack;   /* TLAS: levels + 2 + deepest instanced BLAS = stack entries a ray can need (<= 96 or error 5) */
    uint32_t trace_error; /* bit 0: a traversal ran out of stack on this TLAS (never on structures built by this library) */
    uint32_t reserved;
} lcrtBlobInfo;
LCRT_API int lcrtBlobGetInfo(lcrtStream stream, const void *blob, lcrtBlobInfo *out);

/* CUDA C++ text of the device library (lc_accel_trace_closest, lc_accel_trace_any, lc_accel_query_*, instance
 * accessors, ...) for a code generator to paste in front of generated kernels (NVRTC). */
LCRT_API const char *lcrtDeviceLibrarySource(void);

/* standalone radix sort of (u64 key, u32 value) pairs — exposed for tests; key_bits <= 64.
 * temp needs lcrtSortTempBytes(n) bytes; keys/values are sorted in place. */
LCRT_API uint64_t lcrtSortTempBytes(uint32_t n);
LCRT_API int lcrtSortPairs(lcrtStream stream, uint64_t *keys, uint32_t *values, uint32_t n, uint32_t key_bits, void *temp,
                           uint64_t temp_bytes);

#ifdef __cplusplus
}
#endif
#endif /* LCB200_RT_H */
