"""CPU-side checks of the drop-in boundary: the library loads (no GPU needed) and exports every symbol the headers declare."""
import ctypes
import os
import re

import numpy as np
import pytest

import luisacompute_b200 as lc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return re.findall(r"LCRT_API\s+[\w\s\*]+?\b(lc(?:rt|b)\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    L = lc.lib()
    names = _declared("lcb200_rt.h") + _declared("lcb200_device.h")
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), n
    assert sorted(set(names)) == sorted(set(lc.ABI_SYMBOLS))
    assert L.lcrtGetVersion() == 100


def test_memory_usage_and_layout_contract():
    for kind, stride in ((lc.KIND_TRIANGLES, 48), (lc.KIND_AABBS, 32), (lc.KIND_INSTANCES, 128)):
        for n in (0, 1, 7, 1000, 1_310_720):
            s = lc.compute_memory_usage(kind, n, 0)
            cap = (n * 4 + 6) // 7 + 2
            assert s.blob_bytes >= 128 + cap * 80 + max(n, 1) * stride
            assert s.blob_bytes % 256 == 0 and s.build_temp_bytes > 0 and s.update_temp_bytes == 0
            u = lc.compute_memory_usage(kind, n, 2)
            assert u.update_temp_bytes > 0 and u.blob_bytes >= s.blob_bytes
    # 50M triangles must fit comfortably in 180 GB (C4)
    s = lc.compute_memory_usage(lc.KIND_TRIANGLES, 50_000_000, 0)
    assert s.blob_bytes < 6e9 and s.build_temp_bytes < 12e9


def test_device_library_text_is_embedded():
    src = lc.device_library_source()
    assert "intersect_node8" in src and "struct alignas(16) Node8" in src
    assert "#include \"" not in src
    # every RT name the reference's codegen can emit (cuda_codegen_ast.cpp:983-1004, 1047-1050) is defined by the text
    for name in ("lc_accel_instance_transform", "lc_accel_instance_user_id", "lc_accel_instance_visibility_mask", "lc_accel_set_instance_transform",
                 "lc_accel_set_instance_visibility", "lc_accel_set_instance_opacity", "lc_accel_set_instance_user_id", "lc_accel_trace_closest",
                 "lc_accel_trace_any", "lc_accel_query_all", "lc_accel_query_any", "lc_accel_trace_closest_motion_blur",
                 "lc_accel_trace_any_motion_blur", "lc_accel_query_all_motion_blur", "lc_accel_query_any_motion_blur", "LC_RAY_QUERY_WORLD_RAY",
                 "LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT", "LC_RAY_QUERY_TRIANGLE_CANDIDATE_HIT", "lc_ray_query_committed_hit",
                 "LC_RAY_QUERY_COMMIT_TRIANGLE", "LC_RAY_QUERY_COMMIT_PROCEDURAL", "LC_RAY_QUERY_TERMINATE", "lc_ray_query_trace",
                 "LUISA_DECL_RAY_QUERY_TRIANGLE_IMPL", "LUISA_DECL_RAY_QUERY_PROCEDURAL_IMPL", "struct alignas(16u) LCAccel"):
        assert name in src, name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lc.LcrtError):
        lc.Device()
    # the raw C entry points must refuse to run without a device instead of computing on the host
    assert lc.lib().lcbDeviceCreate(0) is None


def test_wire_layouts():
    from luisacompute_b200 import scenes as S
    assert S.RAY_DTYPE.itemsize == 32 and S.HIT_DTYPE.itemsize == 24 and S.CHIT_DTYPE.itemsize == 24
    assert S.MOD_DTYPE.itemsize == 80 and S.INSTANCE_DTYPE.itemsize == 80
    assert S.MOD_DTYPE.fields["primitive"][1] == 64 and S.INSTANCE_DTYPE.fields["handle"][1] == 64
    assert ctypes.sizeof(lc._AccelBinding) == 16
