/* lcb200_device.h — flat C API of the host-side device layer (C++ implementation in luisacompute_b200/host/).
 *
 * Mirrors, for the ray-tracing path only, the reference's backend interface and its CUDA implementation
 * (paths relative to /root/reference):
 *
 *   lcbDeviceCreate / Destroy          <- extern "C" create / destroy      src/backends/cuda/cuda_device.cpp:1192-1211
 *   lcbCreateMesh / DestroyMesh        <- DeviceInterface::create_mesh / destroy_mesh          cuda_device.cpp:909-921
 *   lcbCreateProceduralPrimitive / ... <- create_procedural_primitive / destroy_...           cuda_device.cpp:939-951
 *   lcbCreateAccel / DestroyAccel      <- create_accel / destroy_accel                         cuda_device.cpp:969-981
 *   lcbMeshBuild                       <- CUDACommandEncoder::visit(MeshBuildCommand*) -> CUDAMesh::build   cuda_mesh.cpp:31-69
 *   lcbProceduralPrimitiveBuild        <- visit(ProceduralPrimitiveBuildCommand*) -> CUDAProceduralPrimitive::build
 *                                                                                  cuda_procedural_primitive.cpp:26-55
 *   lcbAccelBuild                      <- visit(AccelBuildCommand*) -> CUDAAccel::build        cuda_accel.cpp:152-318
 *   lcbAccelGetBinding                 <- CUDAAccel::binding()                                 cuda_accel.h:29-32
 *   lcbTraceClosest / Any / Query      <- a ShaderDispatchCommand of a kernel that only calls accel.intersect / ...
 *
 * Command structs carry the same fields as include/luisa/runtime/rhi/command.h:407-444 (MeshBuildCommand),
 * :489-510 (ProceduralPrimitiveBuildCommand), :531-616 (AccelBuildCommand); a "buffer handle" is a device address.
 * Like the reference, nothing here returns errors for misuse in normal operation: failures are logged to stderr
 * and reported as a negative code (the plugin shim aborts on them, as LUISA_ERROR_WITH_LOCATION does).
 * All work is stream-ordered; no call synchronises the host except lcbSynchronize and the *Info getters.
 */
#ifndef LCB200_DEVICE_H
#define LCB200_DEVICE_H

#include <stdint.h>

#include "lcb200_rt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lcbDevice lcbDevice;

/* AccelOption, include/luisa/runtime/rhi/resource.h:94-126 */
typedef struct lcbAccelOption {
    uint32_t hint;            /* 0 FAST_TRACE, 1 FAST_BUILD */
    uint32_t allow_compaction;/* default true in the reference; accepted, see DESIGN.md (no host stall) */
    uint32_t allow_update;    /* default false */
} lcbAccelOption;

enum { LCB_BUILD_PREFER_UPDATE = 0, LCB_BUILD_FORCE_BUILD = 1 }; /* AccelBuildRequest, command.h:402-405 */

typedef struct lcbMeshBuildCommand {
    uint64_t handle;
    uint32_t request;
    uint64_t vertex_buffer, vertex_buffer_offset, vertex_buffer_size, vertex_stride;
    uint64_t triangle_buffer, triangle_buffer_offset, triangle_buffer_size;
} lcbMeshBuildCommand;

typedef struct lcbProceduralPrimitiveBuildCommand {
    uint64_t handle;
    uint32_t request;
    uint64_t aabb_buffer, aabb_buffer_offset, aabb_buffer_size;
} lcbProceduralPrimitiveBuildCommand;

typedef struct lcbAccelBuildCommand {
    uint64_t handle;
    uint32_t instance_count;
    uint32_t request;
    uint32_t update_instance_buffer_only;
    uint32_t modification_count;
    const void *modifications; /* host array of 80-byte Modification records; `primitive` holds mesh / procedural handles */
} lcbAccelBuildCommand;

typedef struct lcbAccelBinding {
    uint64_t handle;    /* device address of the TLAS blob */
    uint64_t instances; /* device address of the 80-byte instance array */
} lcbAccelBinding;

LCRT_API lcbDevice *lcbDeviceCreate(int device_index);
LCRT_API void lcbDeviceDestroy(lcbDevice *device);
LCRT_API const char *lcbBackendName(void);
/* backend_device_names (cuda_device.cpp:1213-1228): CUDA devices this process can see, 0 without a GPU */
LCRT_API int lcbDeviceCount(void);
LCRT_API int lcbDeviceName(int device_index, char *name, int capacity);

LCRT_API uint64_t lcbCreateMesh(lcbDevice *device, const lcbAccelOption *option);
LCRT_API void lcbDestroyMesh(lcbDevice *device, uint64_t handle);
LCRT_API uint64_t lcbCreateProceduralPrimitive(lcbDevice *device, const lcbAccelOption *option);
LCRT_API void lcbDestroyProceduralPrimitive(lcbDevice *device, uint64_t handle);
LCRT_API uint64_t lcbCreateAccel(lcbDevice *device, const lcbAccelOption *option);
LCRT_API void lcbDestroyAccel(lcbDevice *device, uint64_t handle);

LCRT_API int lcbMeshBuild(lcbDevice *device, lcrtStream stream, const lcbMeshBuildCommand *command);
LCRT_API int lcbProceduralPrimitiveBuild(lcbDevice *device, lcrtStream stream, const lcbProceduralPrimitiveBuildCommand *command);
LCRT_API int lcbAccelBuild(lcbDevice *device, lcrtStream stream, const lcbAccelBuildCommand *command);
LCRT_API int lcbAccelGetBinding(lcbDevice *device, uint64_t accel, lcbAccelBinding *out);

/* what the last build of a primitive / accel did: 0 nothing yet, 1 full build, 2 refit (update) */
LCRT_API int lcbLastBuildKind(lcbDevice *device, uint64_t handle);
/* native handle (device address of the blob) of a mesh / procedural primitive / accel */
LCRT_API uint64_t lcbNativeHandle(lcbDevice *device, uint64_t handle);
LCRT_API int lcbBlobInfo(lcbDevice *device, lcrtStream stream, uint64_t handle, lcrtBlobInfo *out);

LCRT_API int lcbTraceClosest(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count,
                             const uint32_t *masks, uint32_t mask, void *hits, uint64_t *counters);
LCRT_API int lcbTraceAny(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count,
                         const uint32_t *masks, uint32_t mask, uint32_t *occluded, uint64_t *counters);
/* lcbTraceClosest with the records written into a tiled framebuffer, possibly in a peer GPU's memory (lcrtTraceClosestTiled) */
LCRT_API int lcbTraceClosestTiled(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count,
                                  const uint32_t *masks, uint32_t mask, void *scratch_hits, void *dst, uint64_t tile_rays,
                                  uint64_t tile_stride_bytes, uint32_t flags, uint64_t *counters);
/* A framebuffer other processes of this node can write into (one process per GPU): Create allocates `bytes` on
 * this device and fills the 64-byte handle the owner hands to its peers (any channel: a torch.distributed broadcast, a
 * file); Open maps it into the calling process (peer access over NVLink / PCIe is enabled on first use) and returns the
 * address kernels of this device can store to; Close unmaps; Destroy frees (owner only). */
LCRT_API int lcbFramebufferCreate(lcbDevice *device, uint64_t bytes, uint64_t *address, unsigned char handle[64]);
LCRT_API int lcbFramebufferOpen(lcbDevice *device, const unsigned char handle[64], uint64_t *address);
LCRT_API int lcbFramebufferClose(lcbDevice *device, uint64_t address);
/* `tile_count` tiles of `tile_bytes` each, contiguous at `src` on this device, to dst + j * dst_stride_bytes (j = 0 ..): one
 * strided device-to-device copy on `stream`, done by the copy engines — over NVLink when dst is a peer's framebuffer — so it
 * overlaps with kernels of other streams without taking SMs from them. */
LCRT_API int lcbFramebufferCopyTiles(lcbDevice *device, lcrtStream stream, const void *src, uint64_t dst, uint64_t tile_bytes,
                                     uint64_t dst_stride_bytes, uint64_t tile_count);
LCRT_API int lcbFramebufferDestroy(lcbDevice *device, uint64_t address);
LCRT_API int lcbQuery(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count, const uint32_t *masks,
                      uint32_t mask, int terminate_on_first, uint32_t tri_program, uint32_t proc_program, void *hits);

/* Host-buffer convenience (the end-to-end path bench.py times as `e2e`): rays in pinned host memory -> chunked,
 * pipelined (four chunks in flight) H2D copy / trace / D2H copy on internal streams -> hits in pinned host memory.  Synchronous. */
LCRT_API int lcbTraceClosestHost(lcbDevice *device, uint64_t accel, const void *host_rays, uint64_t ray_count, uint32_t mask,
                                 void *host_hits, uint64_t chunk_rays);

/* Serialisable BLAS: a built mesh / procedural primitive is one position-independent blob.  Export copies it into a
 * caller-provided device buffer (returns the byte count through *bytes; pass dst = NULL to query), import adopts
 * such a blob (e.g. after an NCCL broadcast from the rank that built it).  Stream-ordered device-to-device copies. */
LCRT_API int lcbPrimitiveExport(lcbDevice *device, lcrtStream stream, uint64_t handle, void *dst_device, uint64_t capacity, uint64_t *bytes);
LCRT_API int lcbPrimitiveImport(lcbDevice *device, lcrtStream stream, uint64_t handle, const void *src_device, uint64_t bytes);

LCRT_API int lcbSynchronize(lcbDevice *device, lcrtStream stream);
/* BufferDownloadCommand / BufferUploadCommand analogues for raw device addresses (synchronous) */
LCRT_API int lcbDownload(lcbDevice *device, lcrtStream stream, uint64_t device_address, void *host, uint64_t bytes);
LCRT_API int lcbUpload(lcbDevice *device, lcrtStream stream, uint64_t device_address, const void *host, uint64_t bytes);

/* Plumbing the backend plugin (plugin/lc_backend_b200rt.cpp) builds DeviceInterface's buffer / stream / event methods
 * from, so that the plugin contains no CUDA calls: replaces CUDADevice::create_buffer / create_stream / create_event
 * (src/backends/cuda/cuda_device.cpp:281-340, 460-480, 985-1020), CUDAStream (cuda_stream.cpp:142-199: host callbacks run
 * after the GPU work queued before them) and CUDAEvent (cuda_event.cpp: timeline semaphore with fence values).
 * A buffer handle IS its device address (~0 on failure). */
typedef struct lcbEvent lcbEvent;
LCRT_API uint64_t lcbBufferCreate(lcbDevice *device, uint64_t bytes);
LCRT_API void lcbBufferDestroy(lcbDevice *device, uint64_t address);
LCRT_API lcrtStream lcbStreamCreate(lcbDevice *device);
LCRT_API void lcbStreamDestroy(lcbDevice *device, lcrtStream stream);
LCRT_API int lcbCopyAsync(lcbDevice *device, lcrtStream stream, void *dst, const void *src, uint64_t bytes);
LCRT_API int lcbHostCallback(lcbDevice *device, lcrtStream stream, void (*fn)(void *), void *user);
LCRT_API lcbEvent *lcbEventCreate(lcbDevice *device);
LCRT_API void lcbEventDestroy(lcbDevice *device, lcbEvent *event);
LCRT_API int lcbEventSignal(lcbDevice *device, lcbEvent *event, lcrtStream stream, uint64_t value);
LCRT_API int lcbEventWait(lcbDevice *device, lcbEvent *event, lcrtStream stream, uint64_t value);
LCRT_API int lcbEventSynchronize(lcbDevice *device, lcbEvent *event, uint64_t value);
LCRT_API int lcbEventIsCompleted(lcbDevice *device, lcbEvent *event, uint64_t value);
#ifdef __cplusplus
}
#endif
#endif /* LCB200_DEVICE_H */
