/* Minimal stand-in for <vulkan/vulkan_core.h>: just the handle types and constants that the reference's HEADERS
 * (src/backends/common/vulkan_instance.h, src/backends/cuda/cuda_event.h, include/luisa/backends/common/vulkan_swapchain.h)
 * mention, so that two translation units of its CUDA backend can be compiled in an image without the Vulkan SDK.
 * No Vulkan function is declared, none is called: see baseline/ref_cuda_vulkan_stubs.cpp.  Measurement tooling only. */
#pragma once
#include <stdint.h>
#define VK_DEFINE_HANDLE(object) typedef struct object##_T *object;
VK_DEFINE_HANDLE(VkInstance)
VK_DEFINE_HANDLE(VkPhysicalDevice)
VK_DEFINE_HANDLE(VkDevice)
VK_DEFINE_HANDLE(VkQueue)
VK_DEFINE_HANDLE(VkSemaphore)
VK_DEFINE_HANDLE(VkCommandPool)
VK_DEFINE_HANDLE(VkImageView)
VK_DEFINE_HANDLE(VkDebugUtilsMessengerEXT)
#define VK_UUID_SIZE 16U
#define VK_MAKE_API_VERSION(variant, major, minor, patch) ((((uint32_t)(variant)) << 29U) | (((uint32_t)(major)) << 22U) | (((uint32_t)(minor)) << 12U) | ((uint32_t)(patch)))
#define VK_API_VERSION_1_2 VK_MAKE_API_VERSION(0, 1, 2, 0)
typedef enum VkResult { VK_SUCCESS = 0, VK_ERROR_OUT_OF_DATE_KHR = -1000001004 } VkResult;
typedef enum VkImageLayout { VK_IMAGE_LAYOUT_UNDEFINED = 0 } VkImageLayout;
typedef enum VkFormat { VK_FORMAT_UNDEFINED = 0 } VkFormat;
typedef enum VkColorSpaceKHR { VK_COLOR_SPACE_SRGB_NONLINEAR_KHR = 0 } VkColorSpaceKHR;
typedef struct VkExtent2D { uint32_t width, height; } VkExtent2D;
typedef struct VkSurfaceFormatKHR { VkFormat format; VkColorSpaceKHR colorSpace; } VkSurfaceFormatKHR;
