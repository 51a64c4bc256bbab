#pragma once
#include "vulkan_core.h"
