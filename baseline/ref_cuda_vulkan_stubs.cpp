// ref_cuda_vulkan_stubs.cpp — link-time stand-ins for the two Vulkan-backed pieces of the REFERENCE's CUDA backend
// (timeline events: src/backends/cuda/cuda_event.cpp; swapchains: cuda_swapchain.cpp), which can not be compiled in an
// image without the Vulkan SDK and X11.  Every other translation unit of the backend — device, streams, buffers, the
// code generator, NVRTC compiler, OptiX shaders, meshes, accels — is the reference's own, unmodified, compiled by its own
// build system; this file only lets `liblc-backend-cuda.so` link.  The ray-tracing measurement
// (baseline/ref_optix_bench.cpp) creates neither events nor swapchains; anything that does lands in abort() below.
// Measurement tooling only; nothing here is part of the product.
#include <cstdio>
#include <cstdlib>

#include <cuda.h>

#include "cuda_event.h"
#include "cuda_swapchain.h"

#include <luisa/runtime/rhi/resource.h>

namespace {
[[noreturn]] void unavailable(const char *what) noexcept {
    fprintf(stderr, "reference CUDA backend built without the Vulkan SDK: %s is not available\n", what);
    abort();
}
}// namespace

namespace luisa::compute::cuda {

CUDAEvent::CUDAEvent(VkDevice device, VkSemaphore vk_semaphore, CUexternalSemaphore cuda_semaphore) noexcept
    : _device{device}, _vk_semaphore{vk_semaphore}, _cuda_semaphore{cuda_semaphore} {}
void CUDAEvent::notify(uint64_t) noexcept { unavailable("CUDAEvent::notify"); }
void CUDAEvent::signal(CUstream, uint64_t) noexcept { unavailable("CUDAEvent::signal"); }
void CUDAEvent::wait(CUstream, uint64_t) noexcept { unavailable("CUDAEvent::wait"); }
void CUDAEvent::synchronize(uint64_t) noexcept { unavailable("CUDAEvent::synchronize"); }
uint64_t CUDAEvent::signaled_value() noexcept { unavailable("CUDAEvent::signaled_value"); }
bool CUDAEvent::is_completed(uint64_t) noexcept { unavailable("CUDAEvent::is_completed"); }

CUDAEventManager::CUDAEventManager(const CUuuid &) noexcept { unavailable("CUDAEventManager (timeline events)"); }
CUDAEventManager::~CUDAEventManager() noexcept = default;
CUDAEvent *CUDAEventManager::create() noexcept { unavailable("CUDAEventManager::create"); }
void CUDAEventManager::destroy(CUDAEvent *) noexcept { unavailable("CUDAEventManager::destroy"); }

class CUDASwapchain::Impl {};
CUDASwapchain::CUDASwapchain(CUDADevice *, SwapchainOption) noexcept { unavailable("CUDASwapchain"); }
CUDASwapchain::~CUDASwapchain() noexcept = default;
VulkanSwapchain *CUDASwapchain::native_handle() noexcept { unavailable("CUDASwapchain::native_handle"); }
PixelStorage CUDASwapchain::pixel_storage() const noexcept { unavailable("CUDASwapchain::pixel_storage"); }
void CUDASwapchain::present(CUDAStream *, CUDATexture *) noexcept { unavailable("CUDASwapchain::present"); }
void CUDASwapchain::set_name(luisa::string &&) noexcept {}

}// namespace luisa::compute::cuda
