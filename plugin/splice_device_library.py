"""Splices this repository's RT device library into the text of the reference's CUDA device library.

The reference's code generator pastes cuda_device_half.h + cuda_device_math.h + cuda_device_resource.h in front of every
generated kernel (src/backends/cuda/cuda_compiler.cpp:195-215).  This tool rewrites that text so that the ray-tracing names
the generated code calls are this repository's (lcrtDeviceLibrarySource(): lcrt_common.h + lcrt_traverse.h +
lcrt_device_library.cuh over the BVH8 format) instead of OptiX hit-object traversal:

  * `struct alignas(16) LCRay` ... `lc_accel_set_instance_user_id`   (cuda_device_resource.h:1279-1510)  -> our library
  * inside `#ifdef LUISA_ENABLE_OPTIX` (:1553-2387): everything is dropped.  A kernel that traces rays is ordinary device
    code here, so the branch gets the launch-index helpers of the reference's NATIVE branch (:2391-2416: blockIdx / threadIdx,
    `params.ls_kid`, lc_synchronize_block) — the generated `extern "C" __global__ void __raygen__main()` reading
    `__constant__ Params params` is then launched with cuLaunchKernel by plugin/hybrid/b200_shader_rt.cpp, no OptiX
    module / pipeline / SBT / optixLaunch (cuda_shader_optix.cpp:49-309, 478-483).  The code generator emits no
    out-of-range guard for RT kernels (cuda_codegen_ast.cpp:1470-1472: OptiX launches exactly dispatch_size threads), so the
    entry point's NAME becomes a function-like macro that wraps the generated body into an entry point with that guard, the
    same test the native branch's kernels carry.
Nothing of the reference is stored in this repository: the input is read from the reference tree (or from a kernel source the
reference dumped with LUISA_DUMP_SOURCE=1) at build / test time.
usage: python plugin/splice_device_library.py IN OUT [--library FILE]   (IN: cuda_device_resource.h or a dumped kernel_*.cu)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RT_BEGIN = "struct alignas(16) LCRay {"
RT_END_FN = "__device__ inline void lc_accel_set_instance_user_id(LCAccel accel, lc_uint index, lc_uint user_id) noexcept {"
OPTIX_BEGIN = "#ifdef LUISA_ENABLE_OPTIX\n\nenum LCPayloadTypeID"
GLUE_BEGIN = "[[nodiscard]] inline auto lc_dispatch_id() noexcept {"
GLUE_END_FN = "[[nodiscard]] inline auto lc_block_id() noexcept {"
NATIVE_BEGIN = "#else\n\n#define lc_dispatch_size() lc_make_uint3(params.ls_kid)"


def _end_of_function(text: str, start: int) -> int:
    """index just past the closing brace of the function whose header starts at `start`"""
    i = text.index("{", start)
    depth = 0
    while True:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def splice(text: str, library: str, native_launch: bool = True) -> str:
    a = text.index(RT_BEGIN)
    b = _end_of_function(text, text.index(RT_END_FN))
    o = text.index(OPTIX_BEGIN, b)
    g0 = text.index(GLUE_BEGIN, o)
    g1 = _end_of_function(text, text.index(GLUE_END_FN, g0))
    n = text.index(NATIVE_BEGIN, g1)
    # the native branch's helpers, taken from the reference text at build time: "#else" ... up to its "#endif"
    nat0 = n + len("#else\n")
    nat1 = text.index("\n#endif", text.index("lc_synchronize_block", nat0))
    native = text[nat0:nat1]
    assert "lc_dispatch_id" in native and "threadIdx" in native and "_optix_" not in native
    optix_branch = (
        "#ifdef LUISA_ENABLE_OPTIX\n\n"
        "// ray-tracing kernels are launched natively (cuLaunchKernel): launch-index helpers of the reference's native branch\n"
        + native + "\n\n"
        "// The code generator emits no out-of-range guard for ray-tracing kernels (OptiX launches exactly dispatch_size threads) and\n"
        "// writes the entry point as `extern \"C\" __global__ void __raygen__main() { body }` (cuda_codegen_ast.cpp:1391-1394).  As a\n"
        "// function-like macro the name turns that text into: the declaration of an unused kernel, the real entry point (still called\n"
        "// __raygen__main: a macro is not re-expanded inside its own expansion) that drops the threads outside the dispatch, and the\n"
        "// generated body as an inlined function.  Handlers outlined by the RayQuery lowering re-read the built-in variables and the\n"
        "// kernel arguments through the module-wide `__constant__ Params params` (:446-521), which is why the arguments stay there.\n"
        "#define __raygen__main() lc_rt_unused_entry(); \\\n"
        "    static __device__ __forceinline__ void lc_rt_kernel_body(); \\\n"
        "    extern \"C\" __global__ void __raygen__main() { \\\n"
        "        if (lc_any(lc_dispatch_id() >= lc_dispatch_size())) { return; } \\\n"
        "        lc_rt_kernel_body(); \\\n"
        "    } \\\n"
        "    static __device__ __forceinline__ void lc_rt_kernel_body()\n\n")
    if not native_launch:
        # developer variant (plugin/hybrid/build.sh HYBRID_OPTIX_LAUNCH=1 -> lc-backend-b200o): the kernel stays an OptiX raygen
        # program launched by the reference's CUDAShaderOptiX, only the traversal is ours — used to tell launch-path effects
        # from traversal effects
        optix_branch = (
            "#ifdef LUISA_ENABLE_OPTIX\n\n" + text[g0:g1] + "\n\n"
            "inline void lc_shader_execution_reorder(lc_uint hint, lc_uint hint_bits) noexcept {}\n\n"
            "#ifdef LUISA_ENABLE_OPTIX_RAY_QUERY\n"
            "extern \"C\" __global__ void __intersection__ray_query() {}\n"
            "extern \"C\" __global__ void __anyhit__ray_query() {}\n"
            "#endif\n\n")
    return text[:a] + "// ---- RT section: luisacompute_b200 device library (BVH8 traversal), spliced in by plugin/splice_device_library.py\n" + library + "\n" + \
        text[b:o] + optix_branch + text[n:]


def library_text() -> str:
    sys.path.insert(0, ROOT)
    import luisacompute_b200 as lc
    return lc.device_library_source()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("inp")
    ap.add_argument("out")
    ap.add_argument("--library", default=None)
    a = ap.parse_args()
    lib = open(a.library).read() if a.library else library_text()
    out = splice(open(a.inp).read(), lib)
    open(a.out, "w").write(out)
    print(f"{a.out}: {len(out)} bytes")
