// b200_shader_rt.cpp — native launch path for kernels that trace rays.
//
// The reference launches every kernel with `kind == RAY_TRACING` through OptiX: CUDAShaderOptiX builds an OptiX module,
// program groups, a pipeline and a shader binding table around the generated `__raygen__main` and dispatches it with
// optixLaunch (src/backends/cuda/cuda_shader_optix.cpp:49-309, 478-483; selected at cuda_device.cpp:611-619).  With this
// repository's device library the traversal is ordinary device code, so such a kernel is an ordinary kernel:
//
//   * plugin/splice_device_library.py gives the RT branch of the device library the native launch-index helpers
//     (blockIdx / threadIdx, dispatch size from `params.ls_kid`) plus the out-of-range guard the code generator only emits
//     for compute kernels;
//   * this file implements the class the reference's CUDADevice instantiates for RT kernels — same name, same header
//     (cuda_shader_optix.h), linked INSTEAD OF cuda_shader_optix.cpp — over the CUDA driver API: the PTX is loaded with
//     cuModuleLoadData, the arguments are encoded exactly as the reference encodes them (the `Params` layout is a wire
//     contract with the code generator, cuda_codegen_ast.cpp:1346-1357) and copied into the module's `__constant__ Params
//     params`, and the kernel goes out with cuLaunchKernel, grid = ceil(dispatch_size / block_size) like
//     CUDAShaderNative (cuda_shader_native.cpp:150-186).
//
// Together with b200_optix_stub.cpp (an `optix::api()` that never opens libnvoptix) `lc-backend-b200` starts and runs on a
// machine without OptiX.
#include <algorithm>
#include <cstring>
#include <mutex>

#include <luisa/core/logging.h>
#include <luisa/runtime/rhi/command.h>

#include "cuda_error.h"
#include "cuda_device.h"
#include "cuda_stream.h"
#include "cuda_buffer.h"
#include "cuda_accel.h"
#include "cuda_texture.h"
#include "cuda_bindless_array.h"
#include "cuda_command_encoder.h"
#include "cuda_shader_optix.h"
#include "cuda_shader_printer.h"
#include "cuda_shader_metadata.h"

namespace luisa::compute::cuda {

namespace {

// Everything the native path needs; the class layout is the reference's (OptiX handles), so the state hangs off `_module`.
struct NativeRT {
    CUmodule module{};
    CUfunction function{};
    CUdeviceptr params{};     // the module's `__constant__ Params params`
    size_t params_size{};
    uint3 block_size{};
    // `params` is per module, not per launch: a launch from another stream waits for the previous launch of this shader
    std::mutex mutex;
    CUevent last_launch{};
    CUstream last_stream{};
    bool launched{false};
};

[[nodiscard]] NativeRT *state_of(optix::Module m) noexcept { return reinterpret_cast<NativeRT *>(m); }

}// namespace

CUDAShaderOptiX::CUDAShaderOptiX(optix::DeviceContext, luisa::vector<std::byte> ptx,
                                 const char *entry, const CUDAShaderMetadata &metadata,
                                 luisa::vector<ShaderDispatchCommand::Argument> bound_arguments) noexcept
    : CUDAShader{CUDAShaderPrinter::create(metadata.format_types), metadata.argument_usages},
      _bound_arguments{std::move(bound_arguments)} {

    // size of `Params`: every argument at 16-byte alignment, then the print buffer, then ls_kid (cuda_shader_optix.cpp:60-91)
    _argument_buffer_size = 0u;
    for (auto &&arg : metadata.argument_types) {
        auto type = Type::from(arg);
        switch (type->tag()) {
            case Type::Tag::BUFFER: _argument_buffer_size += sizeof(CUDABuffer::Binding); break;
            case Type::Tag::TEXTURE: _argument_buffer_size += sizeof(CUDATexture::Binding); break;
            case Type::Tag::BINDLESS_ARRAY: _argument_buffer_size += sizeof(CUDABindlessArray::Binding); break;
            case Type::Tag::ACCEL: _argument_buffer_size += sizeof(CUDAAccel::Binding); break;
            case Type::Tag::CUSTOM:
                LUISA_ERROR_WITH_LOCATION("Invalid custom type '{}' for a ray-tracing shader argument.", type->description());
            default: _argument_buffer_size += type->size(); break;
        }
        _argument_buffer_size = luisa::align(_argument_buffer_size, 16u);
    }
    if (metadata.requires_printing) {
        _argument_buffer_size += sizeof(CUDAShaderPrinter::Binding);
        _argument_buffer_size = luisa::align(_argument_buffer_size, 16u);
    }
    _argument_buffer_size += 16u;// dispatch size and kernel id

    auto s = luisa::new_with_allocator<NativeRT>();
    s->block_size = metadata.block_size;
    auto load = [&](const void *text, size_t size) noexcept {
        // Register cap: the reference hands kernels that trace rays to OptiX, which compiles them for at most 128 registers
        // per thread unless told otherwise (module_compile_options.maxRegisterCount, cuda_shader_optix.cpp:109-112); the plain
        // JIT has no such default and gives a path tracer 170+, i.e. one 256-thread block per SM (C5 through this backend:
        // 432 spp/s with OptiX's cap, 278 without; profiles/r2_native_launch.md).  Same rule here.
        CUlinkState link{};
        CUjit_option jit_options[] = {CU_JIT_MAX_REGISTERS};
        void *jit_values[] = {reinterpret_cast<void *>(static_cast<uintptr_t>(
            metadata.max_register_count == 0u ? 128u : std::clamp(metadata.max_register_count, 16u, 255u)))};
        LUISA_CHECK_CUDA(cuLinkCreate(1u, jit_options, jit_values, &link));
        auto ret = cuLinkAddData(link, CU_JIT_INPUT_PTX, const_cast<void *>(text), size, "rt_kernel.ptx", 0u, nullptr, nullptr);
        if (ret != CUDA_SUCCESS) {
            LUISA_CHECK_CUDA(cuLinkDestroy(link));
            return ret;
        }
        void *cubin = nullptr;
        size_t cubin_size = 0u;
        LUISA_CHECK_CUDA(cuLinkComplete(link, &cubin, &cubin_size));
        LUISA_CHECK_CUDA(cuModuleLoadData(&s->module, cubin));
        LUISA_CHECK_CUDA(cuLinkDestroy(link));
        return CUDA_SUCCESS;
    };
    auto ret = load(ptx.data(), ptx.size());
    if (ret == CUDA_ERROR_UNSUPPORTED_PTX_VERSION) {
        CUDAShader::_patch_ptx_version(ptx);
        ret = load(ptx.data(), ptx.size());
    }
    LUISA_CHECK_CUDA(ret);
    LUISA_CHECK_CUDA(cuModuleGetFunction(&s->function, s->module, entry));
    LUISA_CHECK_CUDA(cuModuleGetGlobal(&s->params, &s->params_size, s->module, "params"));
    LUISA_ASSERT(s->params_size >= _argument_buffer_size,
                 "b200: `params` of the kernel is {} bytes, the encoded arguments need {}.", s->params_size, _argument_buffer_size);
    LUISA_CHECK_CUDA(cuEventCreate(&s->last_launch, CU_EVENT_DISABLE_TIMING));
    _module = reinterpret_cast<optix::Module>(s);
    _pipeline = reinterpret_cast<optix::Pipeline>(s->function);// what handle() returns: the CUfunction
    LUISA_VERBOSE_WITH_LOCATION("b200: native ray-tracing shader {}: {} argument bytes, block {}x{}x{}.", entry, _argument_buffer_size,
                                s->block_size.x, s->block_size.y, s->block_size.z);
}

CUDAShaderOptiX::~CUDAShaderOptiX() noexcept {
    auto s = state_of(_module);
    if (s == nullptr) { return; }
    LUISA_CHECK_CUDA(cuEventDestroy(s->last_launch));
    LUISA_CHECK_CUDA(cuModuleUnload(s->module));
    luisa::delete_with_allocator(s);
}

// not used by the native path (declared by the reference's header)
optix::ShaderBindingTable CUDAShaderOptiX::_make_sbt() const noexcept { return {}; }

// `argument_buffer`: HOST pointer (pinned, from the stream's upload pool) to the encoded `Params` with ls_kid filled in; the
// reference passes a device address to OptiX here
void CUDAShaderOptiX::_do_launch(CUstream stream, CUdeviceptr argument_buffer, uint3 dispatch_size) const noexcept {
    auto s = state_of(_module);
    auto bs = s->block_size;
    auto blocks = (dispatch_size + bs - 1u) / bs;
    std::scoped_lock lock{s->mutex};
    if (s->launched && s->last_stream != stream) { LUISA_CHECK_CUDA(cuStreamWaitEvent(stream, s->last_launch, 0u)); }
    // stream-ordered: the copy into `params` runs after the previous launch of this stream has finished, the kernel after the copy
    LUISA_CHECK_CUDA(cuMemcpyHtoDAsync(s->params, reinterpret_cast<const void *>(argument_buffer), _argument_buffer_size, stream));
    LUISA_CHECK_CUDA(cuLaunchKernel(s->function, blocks.x, blocks.y, blocks.z, bs.x, bs.y, bs.z, 0u, stream, nullptr, nullptr));
    LUISA_CHECK_CUDA(cuEventRecord(s->last_launch, stream));
    s->last_stream = stream, s->launched = true;
}

// declared by the reference's header; indirect dispatches are expanded in _launch (one _do_launch per recorded dispatch)
void CUDAShaderOptiX::_do_launch_indirect(CUstream, CUdeviceptr, size_t, size_t, const IndirectParameters *, const IndirectParameters *) const noexcept {}

void CUDAShaderOptiX::_launch(CUDACommandEncoder &encoder, ShaderDispatchCommand *command) const noexcept {
    auto cuda_stream = encoder.stream()->handle();

    // indirect: read the dispatch buffer back first (cuda_shader_optix.cpp:329-349)
    luisa::vector<std::byte> indirect_host;
    if (command->is_indirect()) {
        auto indirect = command->indirect_dispatch();
        auto buffer = reinterpret_cast<CUDAIndirectDispatchBuffer *>(indirect.handle);
        indirect_host.resize(buffer->size_bytes());
        LUISA_CHECK_CUDA(cuMemcpyDtoHAsync(indirect_host.data(), buffer->device_address(), buffer->size_bytes(), cuda_stream));
    }

    // encode `Params` on the host: bound arguments, call arguments, print buffer, then room for ls_kid
    static thread_local std::array<std::byte, 65536u> argument_buffer;// the CUDA limit for __constant__ data
    LUISA_ASSERT(_argument_buffer_size <= argument_buffer.size(), "Too many arguments in ShaderDispatchCommand");
    auto offset = static_cast<size_t>(0u);
    auto allocate = [&](size_t bytes) noexcept {
        auto o = luisa::align(offset, 16u);
        LUISA_ASSERT(o + bytes <= _argument_buffer_size, "Too many arguments in ShaderDispatchCommand");
        offset = o + bytes;
        return argument_buffer.data() + o;
    };
    auto put = [&](const auto &binding) noexcept { std::memcpy(allocate(sizeof(binding)), &binding, sizeof(binding)); };
    auto encode = [&](const ShaderDispatchCommand::Argument &arg) noexcept {
        using Tag = ShaderDispatchCommand::Argument::Tag;
        switch (arg.tag) {
            case Tag::BUFFER: put(reinterpret_cast<const CUDABuffer *>(arg.buffer.handle)->binding(arg.buffer.offset, arg.buffer.size)); break;
            case Tag::TEXTURE: put(reinterpret_cast<const CUDATexture *>(arg.texture.handle)->binding(arg.texture.level)); break;
            case Tag::UNIFORM: {
                auto uniform = command->uniform(arg.uniform);
                std::memcpy(allocate(uniform.size_bytes()), uniform.data(), uniform.size_bytes());
                break;
            }
            case Tag::BINDLESS_ARRAY: put(reinterpret_cast<const CUDABindlessArray *>(arg.bindless_array.handle)->binding()); break;
            case Tag::ACCEL: put(reinterpret_cast<const CUDAAccel *>(arg.accel.handle)->binding()); break;
        }
    };
    for (auto &&arg : _bound_arguments) { encode(arg); }
    for (auto &&arg : command->arguments()) { encode(arg); }
    if (printer()) { put(printer()->encode(encoder)); }
    static_cast<void>(allocate(sizeof(uint4)));

    // one pinned copy of `Params` per dispatch (the copies are asynchronous: a dispatch's bytes must stay put until its copy ran);
    // the upload pool recycles the buffer when the stream has passed it (cuda_command_encoder.h:55-59)
    auto launch_one = [&](uint4 ls_kid) noexcept {
        encoder.with_upload_buffer(_argument_buffer_size, [&](CUDAHostBufferPool::View *view) noexcept {
            auto p = static_cast<std::byte *>(view->address());
            std::memcpy(p, argument_buffer.data(), _argument_buffer_size - sizeof(uint4));
            std::memcpy(p + _argument_buffer_size - sizeof(uint4), &ls_kid, sizeof(ls_kid));
            _do_launch(cuda_stream, reinterpret_cast<CUdeviceptr>(p), ls_kid.xyz());
        });
    };
    if (command->is_indirect()) {
        // as the reference does for RT kernels (cuda_shader_optix.cpp:485-503): the dispatch buffer has been read back, the stream is
        // drained once, then one launch per recorded dispatch with its own size and kernel id
        auto indirect = command->indirect_dispatch();
        LUISA_CHECK_CUDA(cuStreamSynchronize(cuda_stream));
        auto readback = reinterpret_cast<const IndirectParameters *>(indirect_host.data());
        auto n = std::min<size_t>(readback->header.size, indirect.max_dispatch_size);
        for (auto i = 0u; i < n; i++) {
            auto ls_kid = readback->dispatches[i + indirect.offset].dispatch_size_and_kernel_id;
            if (any(ls_kid.xyz() == make_uint3(0u))) { break; }
            launch_one(ls_kid);
        }
    } else if (command->is_multiple_dispatch()) {
        for (auto size : command->dispatch_sizes()) {
            if (any(size == make_uint3(0u))) { continue; }
            launch_one(make_uint4(size, 0u));
        }
    } else {
        launch_one(make_uint4(command->dispatch_size(), 0u));
    }
}

}// namespace luisa::compute::cuda
