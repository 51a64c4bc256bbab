// lc_backend_b200rt.cpp — LuisaCompute backend plugin `lc-backend-b200rt` over liblcb200_rt.so.
//
// This is the drop-in boundary of SURVEY.md §8(b): the three `extern "C"` symbols the reference's Context loads from a
// backend module (`create`, `destroy`, `backend_device_names`; src/backends/cuda/cuda_device.cpp:1192-1228, typedefs
// include/luisa/runtime/device.h:105-106, discovery by file name src/runtime/context.cpp:117-139) and a DeviceInterface
// (include/luisa/runtime/rhi/device_interface.h:95-212) whose ray-tracing half is this repository:
//
//   create_mesh / create_procedural_primitive / create_accel / destroy_*   -> lcbCreate* / lcbDestroy*
//   MeshBuildCommand / ProceduralPrimitiveBuildCommand / AccelBuildCommand  -> lcbMeshBuild / lcbProceduralPrimitiveBuild / lcbAccelBuild
//   BufferUpload / BufferDownload / BufferCopy commands, streams, host callbacks, timeline events -> lcb plumbing calls
//   b200rt::TraceCommand (a CustomCommand, plugin/b200rt_ext.h)             -> lcbTraceClosest / lcbTraceAny / lcbQuery
//
// It is compiled against the reference's headers (C-named but C++-typed exports: Context&&, luisa::vector<luisa::string>)
// and contains no CUDA calls of its own.  Everything the path does not cover (textures, bindless arrays, swapchains,
// DSL shaders) reports "not supported by b200rt" through LUISA_ERROR_WITH_LOCATION, the reference's failure mode
// (src/backends/cuda/cuda_error.h:9-32): generated kernels reach this backend's traversal through the re-emitted device
// library (lcrtDeviceLibrarySource), which needs the CUDA backend's codegen and is the next step (f1).
#include <luisa/core/logging.h>
#include <luisa/core/stl/memory.h>
#include <luisa/ast/type.h>
#include <luisa/runtime/context.h>
#include <luisa/runtime/rhi/command.h>
#include <luisa/runtime/rhi/device_interface.h>
#include <luisa/runtime/command_list.h>

#include <lcb200_device.h>

#include "b200rt_ext.h"

namespace luisa::compute::b200rt {

#define LCB_OK(expr)                                                                                                      \
    do {                                                                                                                  \
        if (int rc_ = (expr); rc_ != 0) {                                                                                 \
            LUISA_ERROR_WITH_LOCATION("b200rt: {} failed: {} ({})", #expr, lcrtGetErrorString(rc_), lcrtGetLastCudaError()); \
        }                                                                                                                 \
    } while (false)

#define LCB_UNSUPPORTED(what) LUISA_ERROR_WITH_LOCATION("b200rt: {} is not supported by this backend (ray-tracing path only).", what)

class Device final : public DeviceInterface, public MutableCommandVisitor {

private:
    lcbDevice *_device{nullptr};
    lcrtStream _stream{nullptr};// stream of the dispatch in flight (dispatch() holds _dispatch_mutex)
    std::mutex _dispatch_mutex;

    using CallbackContainer = CommandList::CallbackContainer;
    static void _run_callbacks(void *user) noexcept {
        auto callbacks = static_cast<CallbackContainer *>(user);
        for (auto &&f : *callbacks) { f(); }
        luisa::delete_with_allocator(callbacks);
    }

public:
    Device(Context &&ctx, const DeviceConfig *config) noexcept
        : DeviceInterface{std::move(ctx)},
          _device{lcbDeviceCreate(config != nullptr && config->device_index != std::numeric_limits<size_t>::max() ? static_cast<int>(config->device_index) : 0)} {
        if (_device == nullptr) { LUISA_ERROR_WITH_LOCATION("b200rt: no usable CUDA device (this backend has no CPU fallback)."); }
    }
    ~Device() noexcept override { lcbDeviceDestroy(_device); }

    [[nodiscard]] void *native_handle() const noexcept override { return _device; }
    [[nodiscard]] uint compute_warp_size() const noexcept override { return 32u; }

    // ---- buffers: a handle is the device address, so MeshBuildCommand's (buffer, offset) pairs are plain pointer sums
    [[nodiscard]] BufferCreationInfo create_buffer(const Type *element, size_t elem_count, void *external_memory) noexcept override {
        BufferCreationInfo info{};
        info.element_stride = element == nullptr ? 1u : element->size();// ByteBuffer passes Type::of<void>() / nullptr
        if (info.element_stride == 0u) { info.element_stride = 1u; }
        info.total_size_bytes = info.element_stride * elem_count;
        auto address = external_memory != nullptr ? reinterpret_cast<uint64_t>(external_memory) : lcbBufferCreate(_device, info.total_size_bytes);
        if (address == ~0ull) { LUISA_ERROR_WITH_LOCATION("b200rt: failed to allocate {} bytes of device memory.", info.total_size_bytes); }
        info.handle = address;
        info.native_handle = reinterpret_cast<void *>(address);
        return info;
    }
    // the flat C API (src/api/runtime.cpp:341-353) creates every buffer through this overload; without the Rust IR library an
    // element type can not be named, so a null type means "bytes" (what a C caller passes), anything else is unsupported
    [[nodiscard]] BufferCreationInfo create_buffer(const ir::CArc<ir::Type> *element, size_t elem_count, void *external_memory) noexcept override {
        if (element != nullptr) { LCB_UNSUPPORTED("create_buffer from an IR type"); }
        return create_buffer(static_cast<const Type *>(nullptr), elem_count, external_memory);
    }
    void destroy_buffer(uint64_t handle) noexcept override { lcbBufferDestroy(_device, handle); }

    // ---- streams and events
    [[nodiscard]] ResourceCreationInfo create_stream(StreamTag) noexcept override {
        auto s = lcbStreamCreate(_device);
        if (s == nullptr) { LUISA_ERROR_WITH_LOCATION("b200rt: failed to create a stream."); }
        return {.handle = reinterpret_cast<uint64_t>(s), .native_handle = s};
    }
    void destroy_stream(uint64_t handle) noexcept override { lcbStreamDestroy(_device, reinterpret_cast<lcrtStream>(handle)); }
    void synchronize_stream(uint64_t handle) noexcept override { LCB_OK(lcbSynchronize(_device, reinterpret_cast<lcrtStream>(handle))); }
    void dispatch(uint64_t stream_handle, CommandList &&list) noexcept override {
        std::scoped_lock lock{_dispatch_mutex};
        _stream = reinterpret_cast<lcrtStream>(stream_handle);
        for (auto &&cmd : list.commands()) { cmd->accept(*this); }
        // host callbacks run after the GPU work queued so far, like cuda_stream.cpp:142-170
        if (auto callbacks = list.steal_callbacks(); !callbacks.empty()) {
            auto p = luisa::new_with_allocator<CallbackContainer>(std::move(callbacks));
            LCB_OK(lcbHostCallback(_device, _stream, _run_callbacks, p));
        }
        // commands own host memory a copy may still read (upload sources): keep them alive until the stream gets there
        auto keep = luisa::new_with_allocator<CommandList::CommandContainer>(list.steal_commands());
        LCB_OK(lcbHostCallback(_device, _stream, [](void *user) noexcept {
            luisa::delete_with_allocator(static_cast<CommandList::CommandContainer *>(user));
        }, keep));
        _stream = nullptr;
    }
    [[nodiscard]] ResourceCreationInfo create_event() noexcept override {
        auto e = lcbEventCreate(_device);
        return {.handle = reinterpret_cast<uint64_t>(e), .native_handle = e};
    }
    void destroy_event(uint64_t handle) noexcept override { lcbEventDestroy(_device, reinterpret_cast<lcbEvent *>(handle)); }
    void signal_event(uint64_t handle, uint64_t stream_handle, uint64_t value) noexcept override {
        LCB_OK(lcbEventSignal(_device, reinterpret_cast<lcbEvent *>(handle), reinterpret_cast<lcrtStream>(stream_handle), value));
    }
    void wait_event(uint64_t handle, uint64_t stream_handle, uint64_t value) noexcept override {
        LCB_OK(lcbEventWait(_device, reinterpret_cast<lcbEvent *>(handle), reinterpret_cast<lcrtStream>(stream_handle), value));
    }
    [[nodiscard]] bool is_event_completed(uint64_t handle, uint64_t value) const noexcept override {
        return lcbEventIsCompleted(_device, reinterpret_cast<lcbEvent *>(handle), value) != 0;
    }
    void synchronize_event(uint64_t handle, uint64_t value) noexcept override {
        LCB_OK(lcbEventSynchronize(_device, reinterpret_cast<lcbEvent *>(handle), value));
    }

    // ---- acceleration structures (device_interface.h:171-184 <-> cuda_device.cpp:909-981)
    [[nodiscard]] static lcbAccelOption _option(const AccelOption &o) noexcept {
        return {o.hint == AccelOption::UsageHint::FAST_BUILD ? 1u : 0u, o.allow_compaction ? 1u : 0u, o.allow_update ? 1u : 0u};
    }
    [[nodiscard]] ResourceCreationInfo create_mesh(const AccelOption &option) noexcept override {
        auto o = _option(option);
        auto h = lcbCreateMesh(_device, &o);
        return {.handle = h, .native_handle = reinterpret_cast<void *>(h)};
    }
    void destroy_mesh(uint64_t handle) noexcept override { lcbDestroyMesh(_device, handle); }
    [[nodiscard]] ResourceCreationInfo create_procedural_primitive(const AccelOption &option) noexcept override {
        auto o = _option(option);
        auto h = lcbCreateProceduralPrimitive(_device, &o);
        return {.handle = h, .native_handle = reinterpret_cast<void *>(h)};
    }
    void destroy_procedural_primitive(uint64_t handle) noexcept override { lcbDestroyProceduralPrimitive(_device, handle); }
    [[nodiscard]] ResourceCreationInfo create_accel(const AccelOption &option) noexcept override {
        auto o = _option(option);
        auto h = lcbCreateAccel(_device, &o);
        return {.handle = h, .native_handle = reinterpret_cast<void *>(h)};
    }
    void destroy_accel(uint64_t handle) noexcept override { lcbDestroyAccel(_device, handle); }

    // ---- commands (MutableCommandVisitor; command.h:160-233, 407-444, 489-510, 531-616)
    void visit(BufferUploadCommand *c) noexcept override {
        LCB_OK(lcbCopyAsync(_device, _stream, reinterpret_cast<void *>(c->handle() + c->offset()), c->data(), c->size()));
    }
    void visit(BufferDownloadCommand *c) noexcept override {
        LCB_OK(lcbCopyAsync(_device, _stream, c->data(), reinterpret_cast<const void *>(c->handle() + c->offset()), c->size()));
    }
    void visit(BufferCopyCommand *c) noexcept override {
        LCB_OK(lcbCopyAsync(_device, _stream, reinterpret_cast<void *>(c->dst_handle() + c->dst_offset()),
                            reinterpret_cast<const void *>(c->src_handle() + c->src_offset()), c->size()));
    }
    void visit(MeshBuildCommand *c) noexcept override {
        lcbMeshBuildCommand m{};
        m.handle = c->handle(), m.request = c->request() == AccelBuildRequest::FORCE_BUILD ? LCB_BUILD_FORCE_BUILD : LCB_BUILD_PREFER_UPDATE;
        m.vertex_buffer = c->vertex_buffer(), m.vertex_buffer_offset = c->vertex_buffer_offset();
        m.vertex_buffer_size = c->vertex_buffer_size(), m.vertex_stride = c->vertex_stride();
        m.triangle_buffer = c->triangle_buffer(), m.triangle_buffer_offset = c->triangle_buffer_offset();
        m.triangle_buffer_size = c->triangle_buffer_size();
        LCB_OK(lcbMeshBuild(_device, _stream, &m));
    }
    void visit(ProceduralPrimitiveBuildCommand *c) noexcept override {
        lcbProceduralPrimitiveBuildCommand p{};
        p.handle = c->handle(), p.request = c->request() == AccelBuildRequest::FORCE_BUILD ? LCB_BUILD_FORCE_BUILD : LCB_BUILD_PREFER_UPDATE;
        p.aabb_buffer = c->aabb_buffer(), p.aabb_buffer_offset = c->aabb_buffer_offset(), p.aabb_buffer_size = c->aabb_buffer_size();
        LCB_OK(lcbProceduralPrimitiveBuild(_device, _stream, &p));
    }
    void visit(AccelBuildCommand *c) noexcept override {
        static_assert(sizeof(AccelBuildCommand::Modification) == 80u, "the 80-byte Modification record is passed through verbatim");
        auto mods = c->modifications();
        lcbAccelBuildCommand a{};
        a.handle = c->handle(), a.instance_count = c->instance_count();
        a.request = c->request() == AccelBuildRequest::FORCE_BUILD ? LCB_BUILD_FORCE_BUILD : LCB_BUILD_PREFER_UPDATE;
        a.update_instance_buffer_only = c->update_instance_buffer_only() ? 1u : 0u;
        a.modification_count = static_cast<uint32_t>(mods.size()), a.modifications = mods.data();
        LCB_OK(lcbAccelBuild(_device, _stream, &a));
    }
    void visit(CustomCommand *c) noexcept override {
        if (c->uuid() != trace_command_uuid) { LCB_UNSUPPORTED(luisa::format("custom command {:#x}", c->uuid())); }
        auto t = static_cast<TraceCommand *>(c);
        auto rays = reinterpret_cast<const void *>(t->ray_buffer() + t->ray_offset());
        auto hits = reinterpret_cast<void *>(t->hit_buffer() + t->hit_offset());
        switch (t->mode()) {
            case TraceCommand::Mode::CLOSEST:
                LCB_OK(lcbTraceClosest(_device, _stream, t->accel(), rays, t->ray_count(), nullptr, t->visibility_mask(), hits, nullptr));
                break;
            case TraceCommand::Mode::ANY:
                LCB_OK(lcbTraceAny(_device, _stream, t->accel(), rays, t->ray_count(), nullptr, t->visibility_mask(), static_cast<uint32_t *>(hits), nullptr));
                break;
            case TraceCommand::Mode::QUERY_ALL:
            case TraceCommand::Mode::QUERY_ANY:
                LCB_OK(lcbQuery(_device, _stream, t->accel(), rays, t->ray_count(), nullptr, t->visibility_mask(),
                                t->mode() == TraceCommand::Mode::QUERY_ANY ? 1 : 0, t->triangle_program(), t->procedural_program(), hits));
                break;
        }
    }
    void visit(ShaderDispatchCommand *) noexcept override { LCB_UNSUPPORTED("ShaderDispatchCommand"); }
    void visit(BufferToTextureCopyCommand *) noexcept override { LCB_UNSUPPORTED("BufferToTextureCopyCommand"); }
    void visit(TextureUploadCommand *) noexcept override { LCB_UNSUPPORTED("TextureUploadCommand"); }
    void visit(TextureDownloadCommand *) noexcept override { LCB_UNSUPPORTED("TextureDownloadCommand"); }
    void visit(TextureCopyCommand *) noexcept override { LCB_UNSUPPORTED("TextureCopyCommand"); }
    void visit(TextureToBufferCopyCommand *) noexcept override { LCB_UNSUPPORTED("TextureToBufferCopyCommand"); }
    void visit(CurveBuildCommand *) noexcept override { LCB_UNSUPPORTED("CurveBuildCommand"); }
    void visit(MotionInstanceBuildCommand *) noexcept override { LCB_UNSUPPORTED("MotionInstanceBuildCommand"); }
    void visit(BindlessArrayUpdateCommand *) noexcept override { LCB_UNSUPPORTED("BindlessArrayUpdateCommand"); }

    [[nodiscard]] int trace_flat(uint64_t stream, uint32_t mode, uint64_t accel, uint64_t rays, uint64_t hits, uint64_t count, uint32_t mask) noexcept {
        auto s = reinterpret_cast<lcrtStream>(stream);
        if (mode == 0u) { return lcbTraceClosest(_device, s, accel, reinterpret_cast<const void *>(rays), count, nullptr, mask, reinterpret_cast<void *>(hits), nullptr); }
        if (mode == 1u) { return lcbTraceAny(_device, s, accel, reinterpret_cast<const void *>(rays), count, nullptr, mask, reinterpret_cast<uint32_t *>(hits), nullptr); }
        return LCRT_ERROR_INVALID_VALUE;
    }

    // ---- everything outside the ray-tracing path
    [[nodiscard]] ResourceCreationInfo create_texture(PixelFormat, uint, uint, uint, uint, uint, bool, bool) noexcept override { LCB_UNSUPPORTED("create_texture"); }
    void destroy_texture(uint64_t) noexcept override {}
    [[nodiscard]] ResourceCreationInfo create_bindless_array(size_t) noexcept override { LCB_UNSUPPORTED("create_bindless_array"); }
    void destroy_bindless_array(uint64_t) noexcept override {}
    [[nodiscard]] SwapchainCreationInfo create_swapchain(const SwapchainOption &, uint64_t) noexcept override { LCB_UNSUPPORTED("create_swapchain"); }
    void destroy_swap_chain(uint64_t) noexcept override {}
    void present_display_in_stream(uint64_t, uint64_t, uint64_t) noexcept override { LCB_UNSUPPORTED("present_display_in_stream"); }
    [[nodiscard]] ShaderCreationInfo create_shader(const ShaderOption &, Function) noexcept override { LCB_UNSUPPORTED("create_shader (DSL kernels)"); }
    [[nodiscard]] ShaderCreationInfo create_shader(const ShaderOption &, const ir::KernelModule *) noexcept override { LCB_UNSUPPORTED("create_shader (IR kernels)"); }
    [[nodiscard]] ShaderCreationInfo load_shader(luisa::string_view, luisa::span<const Type *const>) noexcept override { LCB_UNSUPPORTED("load_shader"); }
    Usage shader_argument_usage(uint64_t, size_t) noexcept override { LCB_UNSUPPORTED("shader_argument_usage"); }
    void destroy_shader(uint64_t) noexcept override {}
    void set_name(luisa::compute::Resource::Tag, uint64_t, luisa::string_view) noexcept override {}
    [[nodiscard]] luisa::string query(luisa::string_view property) noexcept override {
        if (property == "device_name") { return "b200rt"; }
        return {};
    }
};

}// namespace luisa::compute::b200rt

LUISA_EXPORT_API luisa::compute::DeviceInterface *create(luisa::compute::Context &&ctx, const luisa::compute::DeviceConfig *config) noexcept {
    return luisa::new_with_allocator<luisa::compute::b200rt::Device>(std::move(ctx), config);
}

LUISA_EXPORT_API void destroy(luisa::compute::DeviceInterface *device) noexcept {
    luisa::delete_with_allocator(static_cast<luisa::compute::b200rt::Device *>(device));
}

// Flat C entry for callers that can not build a CustomCommand (the reference's C API, src/api/runtime.cpp:487-568, has no
// custom-command variant): traces `count` rays of `ray_buffer` against `accel` on `stream`, stream-ordered like a
// ShaderDispatchCommand.  mode: 0 closest -> SurfaceHit[count], 1 any -> uint[count].  `device` is the DeviceInterface pointer the C
// API hands around as LCDevice._0; handles are the C API's.
LUISA_EXPORT_API int b200rt_trace(void *device, uint64_t stream, uint32_t mode, uint64_t accel, uint64_t ray_buffer, uint64_t hit_buffer,
                                  uint64_t count, uint32_t visibility_mask) noexcept {
    auto d = static_cast<luisa::compute::b200rt::Device *>(static_cast<luisa::compute::DeviceInterface *>(device));
    return d->trace_flat(stream, mode, accel, ray_buffer, hit_buffer, count, visibility_mask);
}

LUISA_EXPORT_API void backend_device_names(luisa::vector<luisa::string> &names) noexcept {
    names.clear();
    char name[256];
    for (int i = 0, n = lcbDeviceCount(); i < n; i++) {
        if (lcbDeviceName(i, name, sizeof(name)) == 0) { names.emplace_back(name); }
    }
}
