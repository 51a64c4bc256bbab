// b200_optix_stub.cpp — `luisa::compute::optix::api()` for lc-backend-b200, linked INSTEAD OF the reference's optix_api.cpp
// (which dlopen()s libnvoptix.so.1 and fills the table from optixQueryFunctionTable, src/backends/cuda/optix_api.cpp:44-179).
//
// With this repository's accel classes (b200_cuda_rt.cpp), device library and native shader launch (b200_shader_rt.cpp)
// nothing in the backend needs OptiX any more.  The reference's CUDADevice still asks for a device context when it
// creates a ray-tracing shader (cuda_device.cpp:611-615 -> Handle::optix_context(), :1117-1138) and destroys it at exit
// (:1103-1105): those two entries succeed with a placeholder.  Every other entry aborts with a message naming itself: it
// can only be reached by a feature this backend does not provide (curves, motion instances: SURVEY.md §8 f4).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "optix_api.h"

namespace luisa::compute::optix {

namespace {

[[noreturn]] void unsupported_entry() noexcept {
    std::fprintf(stderr, "[lc-backend-b200] an OptiX entry point was called: this backend does not use OptiX "
                         "(curves and motion instances are not supported; SURVEY.md 8 f4).\n");
    std::abort();
}

const char *error_name(Result) { return "OPTIX_NOT_USED"; }
const char *error_string(Result) { return "lc-backend-b200 does not use OptiX"; }

Result context_create(CUcontext, const DeviceContextOptions *, DeviceContext *context) {
    static int placeholder;
    *context = reinterpret_cast<DeviceContext>(&placeholder);
    return static_cast<Result>(0);// OPTIX_SUCCESS
}
Result context_destroy(DeviceContext) { return static_cast<Result>(0); }

FunctionTable make_table() noexcept {
    FunctionTable t;
    // a table of function pointers and nothing else: every slot traps unless set below
    static_assert(sizeof(FunctionTable) % sizeof(void *) == 0u);
    auto slots = reinterpret_cast<void **>(&t);
    for (size_t i = 0u; i < sizeof(FunctionTable) / sizeof(void *); i++) { slots[i] = reinterpret_cast<void *>(&unsupported_entry); }
    t.getErrorName = error_name;
    t.getErrorString = error_string;
    t.deviceContextCreate = context_create;
    t.deviceContextDestroy = context_destroy;
    return t;
}

}// namespace

const FunctionTable &api() noexcept {
    static auto table = make_table();
    return table;
}

}// namespace luisa::compute::optix
