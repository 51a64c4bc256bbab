"""Helpers shared by the -m gpu parity tests: build the same scene in the oracle and on the device."""
import numpy as np

import luisacompute_b200 as lc
from luisacompute_b200 import scenes as S


def build_pair(oracle, dev, geoms, affines, blas_index=None, masks=None, opaque=None, user_ids=None, option=None):
    """geoms: list of (verts, tris) tuples or aabb arrays.  Returns (oracle_accel, device_accel, device_prims)."""
    n = len(affines)
    blas_index = list(range(n)) if blas_index is None else list(blas_index)
    masks = [0xFF] * n if masks is None else masks
    opaque = [True] * n if opaque is None else opaque
    user_ids = [0] * n if user_ids is None else user_ids
    om, dm = [], []
    for g in geoms:
        if isinstance(g, tuple):
            om.append(oracle.OracleMesh(*g))
            dm.append(dev.create_mesh(g[0], g[1], option=option).build())
        else:
            om.append(oracle.OracleProcedural(g))
            dm.append(dev.create_procedural_primitive(g, option=option).build())
    oa = oracle.OracleAccel()
    oa._keep = om
    assert oa.update(n, S.make_mods([om[b].handle for b in blas_index], affines, masks, opaque, user_ids)) == 0
    da = dev.create_accel(option)
    for i in range(n):
        da.emplace_back(dm[blas_index[i]], affines[i], int(masks[i]), bool(opaque[i]), int(user_ids[i]))
    da.build()
    return oa, da, dm


def check_closest(oracle, oa, da, rays, mask=0xFF, allow_soft=0):
    got = lc.hits_to_numpy(da.trace_closest(rays, mask))
    res = oracle.compare_closest(oa, rays, got, mask)
    assert res["id_mismatch"] == 0, res
    assert res["t_bad"] == 0 and res["bary_bad"] == 0, res
    assert res["tie"] + res["grazing"] <= allow_soft, res
    return res, got
