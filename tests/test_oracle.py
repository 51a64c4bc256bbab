"""Pins the CPU oracle: closed-form answers (C1 triangle, spheres) and its own brute-force mode.

The reference ships no golden vectors for this path (parity unpinned), so the oracle is pinned
against analytic ray/triangle and ray/sphere results and against O(rays x prims) evaluation.
"""
import numpy as np
import pytest

from luisacompute_b200 import scenes as S


def _single_mesh_accel(oracle, verts, tris, affines=None, **kw):
    mesh = oracle.OracleMesh(verts, tris)
    acc = oracle.OracleAccel()
    affines = [S.affine_identity()] if affines is None else affines
    mods = S.make_mods([mesh.handle] * len(affines), affines, **kw)
    assert acc.update(len(affines), mods) == 0
    acc._keep.append(mesh)
    return acc


def test_c1_closed_form(oracle):
    """test_rtx scene: orthographic -z rays against two transformed copies of one triangle, answer in closed form."""
    verts, tris, xf = S.test_rtx_scene()
    acc = _single_mesh_accel(oracle, verts, tris, xf)
    rays = S.test_rtx_rays(128, 128)
    hits, t2, tex = acc.trace_closest(rays)
    # closed form: transform the triangle to world space, 2D point-in-triangle in xy, z from barycentrics
    best_t = np.full(len(rays), np.inf)
    best_i = np.full(len(rays), 0xFFFFFFFF, dtype=np.uint32)
    best_uv = np.zeros((len(rays), 2))
    for ii, a in enumerate(xf):
        A = a.reshape(3, 4).astype(np.float64)
        w = verts[:, :3].astype(np.float64) @ A[:, :3].T + A[:, 3]
        p = rays["o"][:, :2].astype(np.float64)
        v0, v1, v2 = w
        d = (v1[1] - v2[1]) * (v0[0] - v2[0]) + (v2[0] - v1[0]) * (v0[1] - v2[1])
        l0 = ((v1[1] - v2[1]) * (p[:, 0] - v2[0]) + (v2[0] - v1[0]) * (p[:, 1] - v2[1])) / d
        l1 = ((v2[1] - v0[1]) * (p[:, 0] - v2[0]) + (v0[0] - v2[0]) * (p[:, 1] - v2[1])) / d
        l2 = 1 - l0 - l1
        inside = (l0 >= 0) & (l1 >= 0) & (l2 >= 0)
        z = l0 * v0[2] + l1 * v1[2] + l2 * v2[2]
        t = 1.0 - z
        upd = inside & (t < best_t)
        best_t[upd], best_i[upd] = t[upd], ii
        best_uv[upd] = np.stack([l1, l2], 1)[upd]
    assert np.array_equal(hits["inst"], best_i)
    h = best_i != 0xFFFFFFFF
    assert h.sum() > 1000
    assert np.all(hits["prim"][h] == 0)
    np.testing.assert_allclose(tex[h], best_t[h], rtol=1e-12)
    np.testing.assert_allclose(hits["bary"][h], best_uv[h], atol=1e-6)
    # miss record (accel.rs:501-507)
    m = ~h
    assert np.all(hits["prim"][m] == 0xFFFFFFFF) and np.all(hits["bary"][m] == 0) and np.all(hits["t"][m] == S.FLT_MAX)


def test_sphere_analytic(oracle):
    """Undisplaced icosphere: hit distance approaches the analytic ray/sphere distance from inside the chord error."""
    verts, tris = S.icosphere(5, displace=False)
    acc = _single_mesh_accel(oracle, verts, tris)
    rays = S.pinhole_rays(64, 64, vfov_deg=30)
    hits, _, tex = acc.trace_closest(rays)
    o = rays["o"].astype(np.float64)
    d = rays["d"].astype(np.float64)
    b = np.sum(o * d, 1)
    c = np.sum(o * o, 1) - 1.0
    disc = b * b - c
    ana = np.where(disc > 0, -b - np.sqrt(np.maximum(disc, 0)), np.inf)
    h = hits["inst"] != 0xFFFFFFFF
    assert h.sum() > 500
    # inscribed polyhedron: never closer than the sphere, and within the sagitta of a level-5 icosphere
    assert np.all(tex[h] >= ana[h] - 1e-9)
    inner = disc > 0.2
    assert np.all(np.abs(tex[h & inner] - ana[h & inner]) < 5e-3)
    assert not np.any(h & (disc < -1e-3))


@pytest.mark.parametrize("seed", [1, 2])
def test_bvh_equals_brute_force(oracle, seed):
    verts, tris = S.triangle_soup(3000, seed=seed, s=0.05)
    rng = np.random.default_rng(seed)
    xf = []
    for k in range(5):
        m = S.mat_translation(*rng.uniform(-0.5, 0.5, 3)) @ S.mat_rotation(rng.normal(size=3), rng.uniform(0, 6)) @ S.mat_scaling(rng.uniform(0.5, 1.5))
        xf.append(S.affine_from_4x4(m))
    masks = [0xFF, 0x01, 0x02, 0x80, 0x0F]
    acc = _single_mesh_accel(oracle, verts, tris, xf, masks=masks)
    rays = S.uniform_rays(4000, seed=seed + 10)
    for mask in (0xFF, 0x03, 0x80):
        a = acc.trace_closest(rays, mask)
        b = acc.trace_closest(rays, mask, brute=True)
        assert np.array_equal(a[0]["inst"], b[0]["inst"]) and np.array_equal(a[0]["prim"], b[0]["prim"])
        np.testing.assert_array_equal(a[2], b[2])
        np.testing.assert_array_equal(a[1], b[1])     # 2nd-nearest t too
        assert np.array_equal(acc.trace_any(rays, mask), acc.trace_any(rays, mask, brute=True))
        assert np.array_equal(acc.trace_any(rays, mask) != 0, a[0]["inst"] != 0xFFFFFFFF)


def test_masks_and_mods(oracle):
    verts, tris, xf = S.test_rtx_scene()
    mesh = oracle.OracleMesh(verts, tris)
    acc = oracle.OracleAccel()
    mods = S.make_mods([mesh.handle] * 2, xf, masks=[0x01, 0x02], user_ids=[7, 9])
    assert acc.update(2, mods) == 0
    ray = np.zeros(1, dtype=S.RAY_DTYPE)
    ray["o"], ray["d"], ray["tmax"] = (0.0, 0.0, 1.0), (0, 0, -1), S.FLT_MAX
    assert acc.trace_closest(ray, 0xFF)[0]["inst"][0] == 1        # instance 1 sits at z = 0.1, nearer
    assert acc.trace_closest(ray, 0x01)[0]["inst"][0] == 0
    assert acc.trace_closest(ray, 0x04)[0]["inst"][0] == 0xFFFFFFFF
    assert acc.trace_closest(ray, 0x102)[0]["inst"][0] == 1       # 8-bit masks (D4)
    # update_instance_buffer_only applies the mods but must not rebuild (accel.rs:428-430)
    m2 = np.zeros(1, dtype=S.MOD_DTYPE)
    m2["index"], m2["flags"], m2["vis_mask"] = 1, S.MOD_VISIBILITY, 0x04
    assert acc.update(2, m2, update_instance_buffer_only=True) == 0
    assert acc.trace_closest(ray, 0x04)[0]["inst"][0] == 1
    assert acc.update(0, m2) == -1                                  # instance_count == 0 is an error
    assert lib_user_id(oracle, acc, 1) == 9


def lib_user_id(oracle, acc, i):
    return oracle.lib().orc_accel_instance_user_id(acc.handle, i)


def test_ray_query_semantics(oracle):
    """Procedural spheres + one non-opaque triangle (test_procedural.cpp pattern), BVH vs brute force and analytic."""
    aabbs = S.sphere_aabbs(64, radius=0.1)
    proc = oracle.OracleProcedural(aabbs)
    verts, tris, _ = S.test_rtx_scene()
    mesh = oracle.OracleMesh(verts, tris)
    acc = oracle.OracleAccel()
    mods = S.make_mods([proc.handle, mesh.handle], [S.affine_identity(), S.affine_from_4x4(S.mat_scaling(3.0))], opaque=[False, False])
    assert acc.update(2, mods) == 0
    rays = S.pinhole_rays(96, 96, pos=(0, 0, 4), vfov_deg=35)
    for tri_prog in (0, 1, 2):
        a = acc.query(rays, 0xFF, False, tri_prog, 0)
        b = acc.query(rays, 0xFF, False, tri_prog, 0, brute=True)
        assert np.array_equal(a, b)
    q = acc.query(rays, 0xFF, False, 0, 0)
    proc_hits = q["hit_type"] == 2
    assert proc_hits.sum() > 50 and (q["hit_type"] == 1).sum() > 50
    # analytic sphere distance for procedural hits
    c = 0.5 * (aabbs[q["prim"][proc_hits], :3] + aabbs[q["prim"][proc_hits], 3:]).astype(np.float64)
    o = rays["o"][proc_hits].astype(np.float64)
    d = rays["d"][proc_hits].astype(np.float64)
    L = o - c
    hb = np.sum(L * d, 1)
    aa = np.sum(d * d, 1)
    bx = aabbs[q["prim"][proc_hits]].astype(np.float64)
    rad = 0.5 * (bx[:, 3] - bx[:, 0])
    disc = hb * hb - aa * (np.sum(L * L, 1) - rad ** 2)
    np.testing.assert_allclose(q["t"][proc_hits], (-hb - np.sqrt(disc)) / aa, rtol=2e-6)
    # intersect() must not see procedural primitives; with the triangle ignored everything else misses
    h = acc.trace_closest(rays)[0]
    assert np.all((h["inst"] == 1) | (h["inst"] == 0xFFFFFFFF))
    # query_any agrees with query_all on hit/miss (D5)
    qa = acc.query(rays, 0xFF, True, 0, 0)
    assert np.array_equal(qa["hit_type"] != 0, q["hit_type"] != 0)
    # TRI_IGNORE_ALL: only spheres remain
    q2 = acc.query(rays, 0xFF, False, 2, 0)
    assert np.all(q2["hit_type"] != 1)
