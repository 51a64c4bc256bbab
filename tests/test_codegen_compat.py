"""Real generated code against the re-emitted device library.

tests/golden/codegen/*.cu are what the reference's own CUDA code generator emitted for the three kernels of
baseline/ref_optix_bench.cpp (a DSL kernel calling accel.intersect, one calling accel.intersect_any, one running a RayQuery
with a triangle and a procedural handler that reads accel.instance_transform) — dumped on a GPU box with
    LUISA_DUMP_SOURCE=1 baseline/_ref/bin/ref_optix_bench <dir> scene 0x5b
and stripped of the built-in device library text.  Here that library is taken from the reference tree, its ray-tracing
section is replaced by this repository's (plugin/splice_device_library.py), and the whole is compiled with NVRTC exactly as
the reference compiles kernels: every RT name the code generator emits must resolve against our library, with the
reference's own prelude types.  Needs the reference tree (device library text) and NVRTC; no GPU.
"""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = "/root/reference/src/backends/cuda/cuda_builtin"
sys.path.insert(0, os.path.join(ROOT, "plugin"))

nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_LIB, "cuda_device_resource.h")), reason="reference tree not mounted")


def _compile(src: bytes):
    err, prog = nvrtc.nvrtcCreateProgram(src, b"kernel.cu", 0, [], [])
    opts = [b"--gpu-architecture=compute_100", b"-std=c++17", b"-default-device", b"-restrict", b"-extra-device-vectorization", b"-dw", b"-w", b"-ewp",
            b"-DLC_NVRTC_VERSION=120900"]
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    ptx = b""
    if res[0] == nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, n = nvrtc.nvrtcGetPTXSize(prog)
        ptx = b" " * n
        nvrtc.nvrtcGetPTX(prog, ptx)
    return res[0], log.decode(errors="replace"), ptx.decode(errors="replace")


@pytest.mark.parametrize("kernel", ["closest", "any", "query", "query_inline"])
def test_generated_kernel_compiles_against_spliced_library(kernel):
    import splice_device_library as SP
    library = "".join(open(os.path.join(REF_LIB, h)).read() for h in ("cuda_device_half.h", "cuda_device_math.h", "cuda_device_resource.h"))
    rt_library = SP.library_text()
    if kernel == "query_inline":
        # emitted by THIS repository's RayQuery lowering (plugin/hybrid/b200_ray_query_lowering.inl, LUISA_B200_RAY_QUERY=inline): the
        # generator itself emits the lc_ray_query_inline_* section behind the library, which therefore comes without it
        sys.path.insert(0, os.path.join(ROOT, "plugin", "hybrid"))
        import make_device_resource as MDR
        rt_library, api = MDR.cut_inline_api(rt_library)
        assert "lc_ray_query_inline_proceed" in api and "lc_ray_query_inline_proceed" not in rt_library
    spliced = SP.splice(library, rt_library)
    assert "_optix_hitobject_traverse," in library and "_optix_hitobject_traverse," not in spliced     # the asm call is gone
    gen = open(os.path.join(ROOT, "tests", "golden", "codegen", kernel + ".cu")).read()
    src = gen.replace("/* built-in device library */", spliced)
    res, log, ptx = _compile(src.encode())
    assert res == nvrtc.nvrtcResult.NVRTC_SUCCESS, log[:4000]
    assert ".entry __raygen__main" in ptx and ".const .align 16 .b8 params" in ptx
    # no OptiX intrinsic is left: the kernel is ordinary device code that plugin/hybrid/b200_shader_rt.cpp launches with
    # cuLaunchKernel (launch index from %ctaid / %tid, dispatch size from params.ls_kid)
    assert not re.findall(r"_optix_[a-z_0-9]+", ptx)
    assert "%ctaid" in ptx and "%tid" in ptx
    assert "__intersection__ray_query" not in ptx and "__anyhit__ray_query" not in ptx
    if kernel == "query_inline":
        assert "lc_ray_query_inline_begin(lc_rq" in gen and "lc_ray_query_trace(" not in gen.split("/* built-in device library */")[1].split("inline RayQuery API end")[1]
