"""GPU tests of the re-emitted RT device library (luisacompute_b200/csrc/lcrt_device_library.cuh): kernels written
the way LuisaCompute's CUDA codegen emits them (tests/devlib/devlib_kernels.cu) against the CPU oracle.

Covers SURVEY.md 8a rows a17-a22: the LCAccel binding, lc_accel_trace_closest / _any, RayQuery with outlined triangle
and procedural candidate handlers, lc_ray_query_committed_hit and the in-kernel instance accessors.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import luisacompute_b200 as lc
from luisacompute_b200 import scenes as S
from tests.gpu_util import build_pair

pytestmark = pytest.mark.gpu
_HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dl():
    so = os.path.join(_HERE, "devlib", "libdevlibtest.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "devlib"), "-s"])
    L = C.CDLL(so)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.dl_trace.argtypes = [vp, u64, vp, vp, u32, u32, vp, vp]
    L.dl_query.argtypes = [vp, u64, vp, vp, u32, u32, vp, u32, vp, vp]
    L.dl_query_inline.argtypes = [vp, u64, vp, vp, u32, u32, vp, u32, vp, vp]
    L.dl_instance_ops.argtypes = [vp, u64, vp, u32, vp]
    return L


@pytest.fixture(scope="module")
def dev():
    d = lc.Device(0)
    yield d
    d.close()


def _dev_rays(rays):
    import torch
    return torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(-1)).cuda()


def test_inline_trace_closest_and_any(oracle, dev, dl):
    import torch
    verts, tris = S.triangle_soup(20000, seed=11, s=0.03)
    rng = np.random.default_rng(2)
    xf = [S.affine_identity()] + [S.affine_from_4x4(S.mat_translation(*rng.uniform(-0.5, 0.5, 3)) @ S.mat_rotation(rng.normal(size=3), rng.uniform(0, 6)) @ S.mat_scaling(rng.uniform(0.5, 1.5)))
                                  for _ in range(5)]
    masks = [1 << (i & 7) for i in range(6)]
    oa, da, _ = build_pair(oracle, dev, [(verts, tris)], xf, blas_index=[0] * 6, masks=masks)
    rays = S.uniform_rays(30000, seed=5)
    d_rays = _dev_rays(rays)
    bh, bi = da.binding()
    for mask in (0xFF, 0x15):
        hits = torch.empty((len(rays), 6), dtype=torch.int32, device="cuda")
        occ = torch.empty(len(rays), dtype=torch.int32, device="cuda")
        rc = dl.dl_trace(torch.cuda.current_stream().cuda_stream, bh, bi, d_rays.data_ptr(), len(rays), mask, hits.data_ptr(), occ.data_ptr())
        assert rc == 0
        torch.cuda.synchronize()
        got = lc.hits_to_numpy(hits)
        res = oracle.compare_closest(oa, rays, got, mask)
        assert res["id_mismatch"] == 0 and res["t_bad"] == 0 and res["bary_bad"] == 0, res
        assert res["tie"] + res["grazing"] <= 30, res
        # the inline entry point and the batch C-ABI run the same traversal: identical records
        ref = lc.hits_to_numpy(da.trace_closest(d_rays, mask))
        assert np.array_equal(ref["inst"], got["inst"]) and np.array_equal(ref["prim"], got["prim"])
        assert np.array_equal(ref["t"], got["t"]) and np.array_equal(ref["bary"], got["bary"])
        a, e = occ.cpu().numpy() != 0, oa.trace_any(rays, mask) != 0
        assert np.count_nonzero(a != e) <= 3


def test_ray_query_with_outlined_handlers(oracle, dev, dl):
    import torch
    aabbs = S.sphere_aabbs(64, radius=0.1)
    verts, tris, _ = S.test_rtx_scene()
    cv, ct = S.icosphere(2)
    xf = [S.affine_identity(), S.affine_from_4x4(S.mat_scaling(3.0)), S.affine_from_4x4(S.mat_translation(0, 0, -2) @ S.mat_scaling(0.7))]
    oa, da, _ = build_pair(oracle, dev, [aabbs, (verts, tris), (cv, ct)], xf, opaque=[False, False, True])
    rays = S.pinhole_rays(160, 160, pos=(0, 0, 4), vfov_deg=35)
    d_rays = _dev_rays(rays)
    d_aabbs = torch.from_numpy(np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1)).cuda()
    bh, bi = da.binding()
    n = len(rays)
    for tri_prog in (0, 1, 2, 3):
        h_all = torch.empty((n, 6), dtype=torch.int32, device="cuda")
        h_any = torch.empty((n, 6), dtype=torch.int32, device="cuda")
        rc = dl.dl_query(torch.cuda.current_stream().cuda_stream, bh, bi, d_rays.data_ptr(), n, 0xFF, d_aabbs.data_ptr(), tri_prog,
                         h_all.data_ptr(), h_any.data_ptr())
        assert rc == 0
        torch.cuda.synchronize()
        g = lc.chits_to_numpy(h_all)
        e = oa.query(rays, 0xFF, False, tri_prog, 0)
        if tri_prog == 3:
            # a terminating handler leaves *a* committed hit, not the nearest: hit / miss only
            assert np.array_equal(e["hit_type"] != 0, g["hit_type"] != 0)
        else:
            assert np.array_equal(e["hit_type"], g["hit_type"])
            assert np.array_equal(e["inst"], g["inst"])
            hit = e["hit_type"] != 0
            assert np.array_equal(e["prim"][hit], g["prim"][hit])
            np.testing.assert_allclose(g["t"][hit], e["t"][hit], rtol=1e-5)
            np.testing.assert_allclose(g["bary"][hit], e["bary"][hit], atol=1e-4)
            assert np.count_nonzero(e["hit_type"] == 2) > 100 and np.count_nonzero(e["hit_type"] == 1) > 100
        ga = lc.chits_to_numpy(h_any)
        ea = oa.query(rays, 0xFF, True, 0, 2)             # site #1: triangles commit, procedural candidates are ignored
        assert np.array_equal(ea["hit_type"] != 0, ga["hit_type"] != 0)
        # the same two sites lowered inline (`while (proceed) { handler }`, handlers in the kernel's scope): the traversal is the
        # same step by step, so query_all leaves identical records.  (A terminating handler and query_any keep the first hit that
        # commits, and which lane gets there first depends on the phase vote of the outlined form: hit / miss only.)
        i_all = torch.empty((n, 6), dtype=torch.int32, device="cuda")
        i_any = torch.empty((n, 6), dtype=torch.int32, device="cuda")
        rc = dl.dl_query_inline(torch.cuda.current_stream().cuda_stream, bh, bi, d_rays.data_ptr(), n, 0xFF, d_aabbs.data_ptr(), tri_prog,
                                i_all.data_ptr(), i_any.data_ptr())
        assert rc == 0
        torch.cuda.synchronize()
        gi, gia = lc.chits_to_numpy(i_all), lc.chits_to_numpy(i_any)
        if tri_prog == 3:
            assert np.array_equal(gi["hit_type"] != 0, g["hit_type"] != 0)
        else:
            assert gi.tobytes() == g.tobytes()
        assert np.array_equal(gia["hit_type"] != 0, ga["hit_type"] != 0)


def test_in_kernel_instance_accessors(oracle, dev, dl):
    """lc_accel_set_instance_* edit the 80-byte records; the next AccelBuildCommand makes traversal see them."""
    import torch
    verts, tris = S.icosphere(3)
    opt = lc.AccelOption(allow_update=True)
    mesh = dev.create_mesh(verts, tris, option=opt).build()
    acc = dev.create_accel(opt)
    xf = [S.affine_from_4x4(S.mat_translation(-3.0 + 3.0 * i, 0, 0)) for i in range(3)]
    for i in range(3):
        acc.emplace_back(mesh, xf[i], 0xFF, True, 7 + i)
    acc.build()
    bh, bi = acc.binding()
    out = torch.zeros(3, dtype=torch.int32, device="cuda")
    assert dl.dl_instance_ops(torch.cuda.current_stream().cuda_stream, bh, bi, 3, out.data_ptr()) == 0
    torch.cuda.synchronize()
    ib = acc.instance_buffer()
    assert list(out.cpu().numpy()) == [0, 1, 2] and list(ib["mask"]) == [0, 1, 2]          # masked to 8 bits
    assert list(ib["user_id"]) == [1007, 1008, 1009]
    np.testing.assert_allclose(ib["affine"][:, 3], [-2.0, 1.0, 4.0])
    assert [(f >> 2) & 3 for f in ib["flags"]] == [2, 1, 2]                                # ENFORCE_ANYHIT / DISABLE_ANYHIT
    # make every instance visible again from a kernel-edited state, rebuild, and trace against the oracle's view
    for i in range(3):
        acc.set_visibility_on_update(i, 0xFF)
    acc.update_instance_buffer()
    acc.build(lc.FORCE_BUILD)
    om = oracle.OracleMesh(verts, tris)
    oa = oracle.OracleAccel()
    oa._keep.append(om)
    moved = [S.affine_from_4x4(S.mat_translation(-2.0 + 3.0 * i, 0, 0)) for i in range(3)]
    assert oa.update(3, S.make_mods([om.handle] * 3, moved)) == 0
    rays = S.pinhole_rays(96, 96, pos=(1, 0, 9), vfov_deg=50)
    got = lc.hits_to_numpy(acc.trace_closest(rays))
    res = oracle.compare_closest(oa, rays, got)
    assert res["id_mismatch"] == 0 and res["t_bad"] == 0 and res["bary_bad"] == 0, res
    assert np.count_nonzero(got["inst"] != 0xFFFFFFFF) > 500
