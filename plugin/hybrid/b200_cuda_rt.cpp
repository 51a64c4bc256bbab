// b200_cuda_rt.cpp — the acceleration-structure classes of LuisaCompute's CUDA backend (CUDAPrimitive, CUDAMesh,
// CUDAProceduralPrimitive, CUDAAccel; declared in the reference's src/backends/cuda/cuda_{primitive,mesh,procedural_primitive,
// accel}.h) implemented over this repository's host layer (include/lcb200_device.h) instead of OptiX.
//
// This is the f1 integration of SURVEY.md §8: linked INSTEAD OF the reference's cuda_primitive.cpp, cuda_mesh.cpp,
// cuda_procedural_primitive.cpp and cuda_accel.cpp into an otherwise unmodified build of its CUDA backend (device, streams,
// buffers, textures, the DSL code generator, the NVRTC compiler and the shader classes are the reference's own objects), and
// together with the device library whose RT section plugin/splice_device_library.py replaced, it yields
// `liblc-backend-b200.so`: unmodified LuisaCompute programs — DSL kernels calling accel.intersect / intersect_any / RayQuery
// included — run with this repository's BVH8 builder and traversal.  See plugin/hybrid/build.sh.
//
//   reference (replaced)                         here
//   CUDAMesh::build            cuda_mesh.cpp:31-69          lcbMeshBuild (same rebuild-vs-refit rule, in lcb_device.cpp)
//   CUDAProceduralPrimitive::build  cuda_procedural_primitive.cpp:26-55   lcbProceduralPrimitiveBuild
//   CUDAAccel::build           cuda_accel.cpp:152-318       lcbAccelBuild; the 80-byte Modification records are passed on with
//                                                           their `primitive` fields translated to lcb handles
//   CUDAAccel::binding         cuda_accel.h:29-32           { device address of the TLAS blob, 80-byte instance array }
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include <cuda.h>

#include <luisa/core/logging.h>
#include <luisa/runtime/rhi/command.h>

#include "cuda_error.h"
#include "cuda_buffer.h"
#include "cuda_stream.h"
#include "cuda_command_encoder.h"
#include "cuda_primitive.h"
#include "cuda_mesh.h"
#include "cuda_procedural_primitive.h"
#include "cuda_accel.h"

#include <lcb200_device.h>

namespace luisa::compute::cuda {

namespace {

struct Registry {
    std::mutex mutex;
    std::unordered_map<int, lcbDevice *> devices;               // CUDA ordinal -> host-layer device
    std::unordered_map<const void *, uint64_t> handles;         // backend object -> lcb handle
    std::unordered_map<const void *, lcbDevice *> owner;
};
Registry &registry() noexcept {
    static Registry r;
    return r;
}

// the host-layer device of the CUDA context that is current (the backend pushes its primary context around every call)
lcbDevice *current_device() noexcept {
    CUdevice ordinal = 0;
    LUISA_CHECK_CUDA(cuCtxGetDevice(&ordinal));
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    auto &d = r.devices[ordinal];
    if (d == nullptr) {
        d = lcbDeviceCreate(ordinal);
        LUISA_ASSERT(d != nullptr, "b200: lcbDeviceCreate({}) failed.", ordinal);
    }
    return d;
}
void remember(const void *object, lcbDevice *d, uint64_t h) noexcept {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    r.handles[object] = h, r.owner[object] = d;
}
uint64_t handle_of(const void *object) noexcept {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    auto it = r.handles.find(object);
    LUISA_ASSERT(it != r.handles.end(), "b200: unknown primitive / accel object.");
    return it->second;
}
lcbDevice *owner_of(const void *object) noexcept {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    return r.owner.at(object);
}
void forget(const void *object) noexcept {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    r.handles.erase(object), r.owner.erase(object);
}
lcbAccelOption to_option(const AccelOption &o) noexcept {
    return {o.hint == AccelOption::UsageHint::FAST_BUILD ? 1u : 0u, o.allow_compaction ? 1u : 0u, o.allow_update ? 1u : 0u};
}
#define B200_OK(expr)                                                                                                       \
    do {                                                                                                                    \
        if (int rc_ = (expr); rc_ != 0) {                                                                                   \
            LUISA_ERROR_WITH_LOCATION("b200: {} failed: {} ({})", #expr, lcrtGetErrorString(rc_), lcrtGetLastCudaError());  \
        }                                                                                                                   \
    } while (false)

}// namespace

// ---- CUDAPrimitiveBase / CUDAPrimitive
optix::TraversableHandle CUDAPrimitiveBase::handle() const noexcept {
    std::scoped_lock lock{_mutex};
    return _handle;
}
void CUDAPrimitiveBase::set_name(luisa::string &&name) noexcept {
    std::scoped_lock lock{_mutex};
    _name = std::move(name);
}
CUDAPrimitive::CUDAPrimitive(Tag tag, const AccelOption &option) noexcept : CUDAPrimitiveBase{tag}, _option{option} {
    auto d = current_device();
    auto o = to_option(option);
    switch (tag) {
        case Tag::MESH: remember(this, d, lcbCreateMesh(d, &o)); break;
        case Tag::PROCEDURAL: remember(this, d, lcbCreateProceduralPrimitive(d, &o)); break;
        default: LUISA_ERROR_WITH_LOCATION("b200: curves and motion instances are not supported by this backend.");
    }
}
CUDAPrimitive::~CUDAPrimitive() noexcept {
    auto d = owner_of(this);
    auto h = handle_of(this);
    if (tag() == Tag::MESH) { lcbDestroyMesh(d, h); } else { lcbDestroyProceduralPrimitive(d, h); }
    forget(this);
}
void CUDAPrimitive::_build(CUDACommandEncoder &) noexcept { LUISA_ERROR_WITH_LOCATION("b200: CUDAPrimitive::_build is not used."); }
void CUDAPrimitive::_update(CUDACommandEncoder &) noexcept { LUISA_ERROR_WITH_LOCATION("b200: CUDAPrimitive::_update is not used."); }
const CUdeviceptr *CUDAPrimitive::_motion_buffer_pointers(CUdeviceptr, size_t) const noexcept { return nullptr; }

// ---- CUDAMesh
CUDAMesh::CUDAMesh(const AccelOption &option) noexcept : CUDAPrimitive{Tag::MESH, option} {}
optix::BuildInput CUDAMesh::_make_build_input() const noexcept { return {}; }
void CUDAMesh::build(CUDACommandEncoder &encoder, MeshBuildCommand *command) noexcept {
    auto vertex_buffer = reinterpret_cast<const CUDABuffer *>(command->vertex_buffer());
    auto triangle_buffer = reinterpret_cast<const CUDABuffer *>(command->triangle_buffer());
    LUISA_ASSERT(command->vertex_buffer_offset() + command->vertex_buffer_size() <= vertex_buffer->size_bytes(), "Vertex buffer range exceeds the buffer.");
    LUISA_ASSERT(command->triangle_buffer_offset() + command->triangle_buffer_size() <= triangle_buffer->size_bytes(), "Triangle buffer range exceeds the buffer.");
    std::scoped_lock lock{_mutex};
    _vertex_buffer = vertex_buffer->device_address() + command->vertex_buffer_offset();
    _vertex_buffer_size = command->vertex_buffer_size(), _vertex_stride = command->vertex_stride();
    _triangle_buffer = triangle_buffer->device_address() + command->triangle_buffer_offset();
    _triangle_buffer_size = command->triangle_buffer_size();
    lcbMeshBuildCommand m{};
    m.handle = handle_of(this);
    m.request = command->request() == AccelBuildRequest::FORCE_BUILD ? LCB_BUILD_FORCE_BUILD : LCB_BUILD_PREFER_UPDATE;
    m.vertex_buffer = _vertex_buffer, m.vertex_buffer_offset = 0u, m.vertex_buffer_size = _vertex_buffer_size, m.vertex_stride = _vertex_stride;
    m.triangle_buffer = _triangle_buffer, m.triangle_buffer_offset = 0u, m.triangle_buffer_size = _triangle_buffer_size;
    auto d = owner_of(this);
    B200_OK(lcbMeshBuild(d, encoder.stream()->handle(), &m));
    _handle = lcbNativeHandle(d, m.handle);// the blob's device address: what a TLAS record points at
}

// ---- CUDAProceduralPrimitive
CUDAProceduralPrimitive::CUDAProceduralPrimitive(const AccelOption &option) noexcept : CUDAPrimitive{Tag::PROCEDURAL, option} {}
optix::BuildInput CUDAProceduralPrimitive::_make_build_input() const noexcept { return {}; }
void CUDAProceduralPrimitive::build(CUDACommandEncoder &encoder, ProceduralPrimitiveBuildCommand *command) noexcept {
    auto aabb_buffer = reinterpret_cast<const CUDABuffer *>(command->aabb_buffer());
    LUISA_ASSERT(command->aabb_buffer_offset() + command->aabb_buffer_size() <= aabb_buffer->size_bytes(), "AABB buffer range exceeds the buffer.");
    std::scoped_lock lock{_mutex};
    _aabb_buffer = aabb_buffer->device_address() + command->aabb_buffer_offset();
    _aabb_buffer_size = command->aabb_buffer_size();
    lcbProceduralPrimitiveBuildCommand p{};
    p.handle = handle_of(this);
    p.request = command->request() == AccelBuildRequest::FORCE_BUILD ? LCB_BUILD_FORCE_BUILD : LCB_BUILD_PREFER_UPDATE;
    p.aabb_buffer = _aabb_buffer, p.aabb_buffer_offset = 0u, p.aabb_buffer_size = _aabb_buffer_size;
    auto d = owner_of(this);
    B200_OK(lcbProceduralPrimitiveBuild(d, encoder.stream()->handle(), &p));
    _handle = lcbNativeHandle(d, p.handle);
}

// ---- CUDAAccel
CUDAAccel::CUDAAccel(const AccelOption &option) noexcept : _option{option} {
    auto d = current_device();
    auto o = to_option(option);
    remember(this, d, lcbCreateAccel(d, &o));
}
CUDAAccel::~CUDAAccel() noexcept {
    lcbDestroyAccel(owner_of(this), handle_of(this));
    forget(this);
}
void CUDAAccel::_build(CUDACommandEncoder &) noexcept {}
void CUDAAccel::_update(CUDACommandEncoder &) noexcept {}
void CUDAAccel::build(CUDACommandEncoder &encoder, AccelBuildCommand *command) noexcept {
    std::scoped_lock lock{_mutex};
    static_assert(sizeof(AccelBuildCommand::Modification) == 80u);
    auto mods = command->modifications();
    luisa::vector<AccelBuildCommand::Modification> translated{mods.begin(), mods.end()};
    for (auto &m : translated) {// `primitive` arrives as the backend object's address (cuda_device.cpp:909-951)
        if (m.flags & AccelBuildCommand::Modification::flag_primitive) { m.primitive = handle_of(reinterpret_cast<const void *>(m.primitive)); }
    }
    lcbAccelBuildCommand a{};
    a.handle = handle_of(this), a.instance_count = command->instance_count();
    a.request = command->request() == AccelBuildRequest::FORCE_BUILD ? LCB_BUILD_FORCE_BUILD : LCB_BUILD_PREFER_UPDATE;
    a.update_instance_buffer_only = command->update_instance_buffer_only() ? 1u : 0u;
    a.modification_count = static_cast<uint32_t>(translated.size()), a.modifications = translated.data();
    auto d = owner_of(this);
    B200_OK(lcbAccelBuild(d, encoder.stream()->handle(), &a));
    lcbAccelBinding b{};
    B200_OK(lcbAccelGetBinding(d, a.handle, &b));
    _handle = b.handle, _instance_buffer = b.instances;
}
optix::TraversableHandle CUDAAccel::handle() const noexcept {
    std::scoped_lock lock{_mutex};
    return _handle;
}
CUdeviceptr CUDAAccel::instance_buffer() const noexcept {
    std::scoped_lock lock{_mutex};
    return _instance_buffer;
}
CUDAAccel::Binding CUDAAccel::binding() const noexcept {
    std::scoped_lock lock{_mutex};
    return {_handle, _instance_buffer};
}
void CUDAAccel::set_name(luisa::string &&name) noexcept {
    std::scoped_lock lock{_mutex};
    _name = std::move(name);
}

}// namespace luisa::compute::cuda
