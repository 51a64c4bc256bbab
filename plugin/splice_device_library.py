"""Splices this repository's RT device library into the text of the reference's CUDA device library.

The reference's code generator pastes cuda_device_half.h + cuda_device_math.h + cuda_device_resource.h in front of every
generated kernel (src/backends/cuda/cuda_compiler.cpp:195-215).  This tool rewrites that text so that the ray-tracing names
the generated code calls are this repository's (lcrtDeviceLibrarySource(): lcrt_common.h + lcrt_traverse.h +
lcrt_device_library.cuh over the BVH8 format) instead of OptiX hit-object traversal:

  * `struct alignas(16) LCRay` ... `lc_accel_set_instance_user_id`   (cuda_device_resource.h:1279-1510)  -> our library
  * inside `#ifdef LUISA_ENABLE_OPTIX` (:1553-2387): everything except the launch-index helpers (lc_dispatch_id,
    lc_dispatch_size, lc_kernel_id, lc_thread_id, lc_block_id, :1844-1880) is dropped; empty `__intersection__ray_query` /
    `__anyhit__ray_query` entry points keep the reference's OptiX pipeline creation (cuda_shader_optix.cpp:190-265) happy.
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


def splice(text: str, library: str) -> str:
    a = text.index(RT_BEGIN)
    b = _end_of_function(text, text.index(RT_END_FN))
    o = text.index(OPTIX_BEGIN, b)
    g0 = text.index(GLUE_BEGIN, o)
    g1 = _end_of_function(text, text.index(GLUE_END_FN, g0))
    n = text.index(NATIVE_BEGIN, g1)
    optix_branch = (
        "#ifdef LUISA_ENABLE_OPTIX\n\n"
        "// launch-index helpers of the reference (OptiX launches the kernel as a raygen program); the traversal is ordinary device code\n"
        + text[g0:g1] + "\n\n"
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
