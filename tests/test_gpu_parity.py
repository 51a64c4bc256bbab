"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json): inst / prim ids exact except inside the 1e-5 relative-t tie band; t within 1e-5 relative;
barycentrics within 1e-4 absolute.
"""
import numpy as np
import pytest

import luisacompute_b200 as lc
from luisacompute_b200 import scenes as S
from tests.gpu_util import build_pair, check_closest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = lc.Device(0)
    yield d
    d.close()


def test_radix_sort(dev):
    import torch
    L = lc.lib()
    for n, bits in ((1, 64), (33, 64), (4096, 64), (4097, 39), (1_000_003, 64), (300_000, 45)):
        g = torch.Generator(device="cuda").manual_seed(n)
        keys = torch.randint(0, 2 ** 62, (n,), dtype=torch.int64, device="cuda", generator=g)
        if bits < 64:
            keys &= (1 << bits) - 1
        keys[::7] = keys[0].clone()                        # duplicates: stability matters
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        ref_k, ref_i = torch.sort(keys, stable=True)
        temp = torch.empty(L.lcrtSortTempBytes(n), dtype=torch.uint8, device="cuda")
        k2, v2 = keys.clone(), vals.clone()
        rc = L.lcrtSortPairs(torch.cuda.current_stream().cuda_stream, k2.data_ptr(), v2.data_ptr(), n, bits, temp.data_ptr(), temp.numel())
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(k2, ref_k)
        assert torch.equal(v2.long(), ref_i)


def test_c1_test_rtx(oracle, dev):
    verts, tris, xf = S.test_rtx_scene()
    oa, da, dm = build_pair(oracle, dev, [(verts, tris)], xf, blas_index=[0, 0])
    info = dm[0].info()
    assert info.built == 1 and info.error == 0 and info.prim_count == 1 and info.node_count == 1
    rays = S.test_rtx_rays(256, 256)
    res, got = check_closest(oracle, oa, da, rays)
    assert res["exact"] == len(rays)
    occ = da.trace_any(rays).cpu().numpy()
    assert np.array_equal(occ != 0, got["inst"] != 0xFFFFFFFF)


@pytest.mark.parametrize("level", [0, 4, 6])
def test_icosphere_primary_and_bounce(oracle, dev, level):
    verts, tris = S.icosphere(level)
    oa, da, dm = build_pair(oracle, dev, [(verts, tris)], [S.affine_identity()])
    info = dm[0].info()
    assert info.built == 1 and info.error == 0 and info.prim_count == len(tris)
    w = h = 192
    prim = S.pinhole_rays(w, h)
    res, got = check_closest(oracle, oa, da, prim)
    assert res["exact"] == len(prim)
    bounce = S.bounce_rays(prim, got, verts, tris, w, h)
    res, gb = check_closest(oracle, oa, da, bounce)
    assert res["exact"] == len(bounce)
    occ = da.trace_any(bounce).cpu().numpy()
    assert np.array_equal(occ != 0, oa.trace_any(bounce) != 0)


def test_soup_instances_masks(oracle, dev):
    verts, tris = S.triangle_soup(20000, seed=3, s=0.03)
    rng = np.random.default_rng(5)
    xf = [S.affine_from_4x4(S.mat_translation(*rng.uniform(-0.5, 0.5, 3)) @ S.mat_rotation(rng.normal(size=3), rng.uniform(0, 6)) @ S.mat_scaling(rng.uniform(0.5, 1.5)))
          for _ in range(17)]
    masks = [1 << (i & 7) for i in range(17)]
    oa, da, _ = build_pair(oracle, dev, [(verts, tris)], xf, blas_index=[0] * 17, masks=masks, user_ids=list(range(100, 117)))
    rays = S.uniform_rays(40000, seed=77)
    for mask in (0xFF, 0x55, 0x01):
        check_closest(oracle, oa, da, rays, mask, allow_soft=40)
        a, b = da.trace_any(rays, mask).cpu().numpy() != 0, oa.trace_any(rays, mask) != 0
        assert np.count_nonzero(a != b) <= 4
    pm = (1 << (np.arange(len(rays)) % 8)).astype(np.uint32)
    check_closest(oracle, oa, da, rays, pm, allow_soft=40)
    ib = da.instance_buffer()
    assert list(ib["user_id"]) == list(range(100, 117)) and list(ib["mask"]) == masks


def test_degenerate_and_edge_cases(oracle, dev):
    verts = np.zeros((6, 4), dtype=np.float32)
    verts[:, :3] = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 0), (0, 0, 0), (2, 2, 2)]
    tris = np.array([[0, 1, 2], [0, 1, 2], [3, 4, 5], [3, 3, 3], [0, 1, 2]], dtype=np.uint32)
    oa, da, dm = build_pair(oracle, dev, [(verts, tris)], [S.affine_identity()])
    rays = np.zeros(3, dtype=S.RAY_DTYPE)
    rays["o"] = [(0.2, 0.2, 1), (0.9, 0.9, 1), (0.2, 0.2, -1)]
    rays["d"] = [(0, 0, -1), (0, 0, -1), (0, 0, 1)]
    rays["tmax"] = S.FLT_MAX
    got = lc.hits_to_numpy(da.trace_closest(rays))
    assert list(got["inst"]) == [0, 0xFFFFFFFF, 0] and list(got["prim"][[0, 2]]) == [0, 0]
    rays["tmin"], rays["tmax"] = 0.5, 0.9
    assert np.all(lc.hits_to_numpy(da.trace_closest(rays))["inst"] == 0xFFFFFFFF)
    rays["tmin"], rays["tmax"] = 0.5, 1.0
    assert list(lc.hits_to_numpy(da.trace_closest(rays))["inst"]) == [0, 0xFFFFFFFF, 0]
    rays["tmin"], rays["tmax"] = 2.0, 1.0
    assert np.all(lc.hits_to_numpy(da.trace_closest(rays))["inst"] == 0xFFFFFFFF)
    # zero rays, empty mesh next to a real one, instance_count == 0
    assert da.trace_closest(np.zeros(0, dtype=S.RAY_DTYPE)).shape[0] == 0
    em = dev.create_mesh(np.zeros((1, 4), np.float32), np.zeros((0, 3), np.uint32)).build()
    assert em.info().built == 1
    a2 = dev.create_accel()
    a2.emplace_back(em)
    a2.emplace_back(dm[0])
    a2.build()
    rays["tmin"], rays["tmax"] = 0.0, S.FLT_MAX
    assert list(lc.hits_to_numpy(a2.trace_closest(rays))["inst"]) == [1, 0xFFFFFFFF, 1]
    with pytest.raises(lc.LcrtError):
        dev.create_accel().build()


def test_refit_rebuild_decisions_and_dynamic_scene(oracle, dev):
    """test_rtx.cpp:122-139: per-frame vertex edit -> BLAS refit -> instance transform -> TLAS update."""
    verts, tris = S.icosphere(3)
    opt = lc.AccelOption(allow_update=True)
    mesh = dev.create_mesh(verts, tris, option=opt).build()
    assert mesh.last_build_kind == 1
    acc = dev.create_accel(opt)
    xf = [S.affine_identity(), S.affine_from_4x4(S.mat_translation(2.5, 0, 0))]
    acc.emplace_back(mesh, xf[0])
    acc.emplace_back(mesh, xf[1])
    acc.build()
    assert acc.last_build_kind == 1
    rays = S.pinhole_rays(96, 96, pos=(0, 0, 7), vfov_deg=50)
    import torch
    for frame in range(3):
        v2 = verts.copy()
        v2[:, :3] *= (1.0 + 0.2 * np.sin(5 * verts[:, 0:1] + frame)).astype(np.float32)
        mesh.vertices.copy_(torch.from_numpy(v2.view(np.uint8).reshape(-1)))       # same buffer, new contents
        mesh.build()
        assert mesh.last_build_kind == 2                                               # refit (cuda_mesh.cpp:42-68)
        t1 = S.affine_from_4x4(S.mat_translation(-2.5, 0.3 * frame, 0) @ S.mat_rotation((0, 1, 0), 0.7 * frame))
        acc.set_transform_on_update(1, t1)
        acc.build()
        assert acc.last_build_kind == 2                                                # update (cuda_accel.cpp:261-267)
        om = oracle.OracleMesh(v2, tris)
        oa = oracle.OracleAccel()
        oa._keep.append(om)
        assert oa.update(2, S.make_mods([om.handle] * 2, [xf[0], t1])) == 0
        check_closest(oracle, oa, acc, rays, allow_soft=8)
    # FORCE_BUILD and a changed instance count force rebuilds
    mesh.build(lc.FORCE_BUILD)
    assert mesh.last_build_kind == 1
    acc.build()
    assert acc.last_build_kind == 2
    acc.emplace_back(mesh, S.affine_from_4x4(S.mat_translation(0, 2.5, 0)))
    acc.build()
    assert acc.last_build_kind == 1
    # update_instance_buffer_only applies the modification without touching the BVH (cuda_accel.cpp:309)
    acc.set_visibility_on_update(0, 0x02)
    acc.update_instance_buffer()
    assert acc.instance_buffer()["mask"][0] == 0x02
    # a mesh built without allow_update always rebuilds
    m2 = dev.create_mesh(verts, tris).build()
    m2.build()
    assert m2.last_build_kind == 1


def test_ray_query(oracle, dev):
    aabbs = S.sphere_aabbs(64, radius=0.1)
    verts, tris, _ = S.test_rtx_scene()
    cv, ct = S.icosphere(2)
    xf = [S.affine_identity(), S.affine_from_4x4(S.mat_scaling(3.0)), S.affine_from_4x4(S.mat_translation(0, 0, -2) @ S.mat_scaling(0.7))]
    oa, da, _ = build_pair(oracle, dev, [aabbs, (verts, tris), (cv, ct)], xf, opaque=[False, False, True])
    rays = S.pinhole_rays(160, 160, pos=(0, 0, 4), vfov_deg=35)
    for tri_prog in (0, 1, 2):
        e = oa.query(rays, 0xFF, False, tri_prog, 0)
        g = lc.chits_to_numpy(da.query_all(rays, tri_program=tri_prog, proc_program=0))
        assert np.array_equal(e["hit_type"], g["hit_type"])
        assert np.array_equal(e["inst"], g["inst"])
        hit = e["hit_type"] != 0
        assert np.array_equal(e["prim"][hit], g["prim"][hit])
        np.testing.assert_allclose(g["t"][hit], e["t"][hit], rtol=1e-5)
        np.testing.assert_allclose(g["bary"][hit], e["bary"][hit], atol=1e-4)
        assert np.count_nonzero(e["hit_type"] == 2) > 100
    e = oa.query(rays, 0xFF, True, 0, 0)
    g = lc.chits_to_numpy(da.query_any(rays))
    assert np.array_equal(e["hit_type"] != 0, g["hit_type"] != 0)       # D5: hit / miss only
    got = lc.hits_to_numpy(da.trace_closest(rays))
    assert not np.any(got["inst"] == 0)                                    # intersect() does not see procedural prims
    assert oracle.compare_closest(oa, rays, got)["id_mismatch"] == 0


def test_host_buffer_path_matches_device_path(oracle, dev):
    import torch
    verts, tris = S.icosphere(5)
    oa, da, _ = build_pair(oracle, dev, [(verts, tris)], [S.affine_identity()])
    rays = S.pinhole_rays(300, 300)
    ref = da.trace_closest(rays).cpu()
    hr = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).pin_memory()
    hh = torch.empty((len(rays), 6), dtype=torch.int32).pin_memory()
    da.trace_closest_host(hr, hh, chunk_rays=20000)
    assert torch.equal(hh, ref)
