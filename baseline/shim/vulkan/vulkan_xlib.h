/* stand-in, see vulkan_core.h */
#pragma once
