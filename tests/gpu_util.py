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


# what every closest-hit comparison of the run found inside the tolerance bands (the BASELINE.json rule allows ids to differ when the
# two nearest hits are within 1e-5 relative t, and a hit whose exact barycentric is within 1e-4 of an edge): tests/conftest.py prints
# the table at the end of the run, so the counts are on record and not just below a threshold
SOFT_LOG = []


def record_soft(what: str, res: dict, allowed) -> None:
    SOFT_LOG.append((what, int(res.get("n", 0)), int(res.get("exact", 0)), int(res.get("tie", 0)), int(res.get("grazing", 0)), allowed))


def check_closest(oracle, oa, da, rays, mask=0xFF, allow_soft=0):
    import inspect
    got = lc.hits_to_numpy(da.trace_closest(rays, mask))
    res = oracle.compare_closest(oa, rays, got, mask)
    record_soft(f"{inspect.stack()[1].function} mask " + (f"{int(mask):#x}" if isinstance(mask, (int, np.integer)) else "per ray"), res, allow_soft)
    assert res["id_mismatch"] == 0, res
    assert res["t_bad"] == 0 and res["bary_bad"] == 0, res
    assert res["tie"] + res["grazing"] <= allow_soft, res
    return res, got


def compare_query(oa, rays, e, g, t_rel=1e-5, bary_abs=1e-4):
    """RayQuery (CommittedHit) records of the device `g` against the oracle's `e` under the BASELINE.json rule: hit type,
    inst and prim exact, except surface hits whose claimed (inst, prim) the oracle evaluates to a hit within `t_rel` of its own
    nearest t (two coplanar walls meeting in an edge: the tie rule); t within t_rel, barycentrics within bary_abs."""
    same = (e["hit_type"] == g["hit_type"]) & (e["inst"] == g["inst"]) & ((e["prim"] == g["prim"]) | (e["hit_type"] == 0))
    res = {"n": len(rays), "exact": int(np.count_nonzero(same)), "tie": 0, "bad": 0, "bad_index": []}
    for i in np.nonzero(~same)[0]:
        ok = False
        if e["hit_type"][i] == 1 and g["hit_type"][i] == 1:
            valid, ev = oa.eval_hit(rays[i:i + 1], g["inst"][i], g["prim"][i])
            te = float(e["t"][i])
            ok = bool(valid) and ev[3] >= -bary_abs and abs(ev[0] - te) <= t_rel * abs(te) and abs(float(g["t"][i]) - ev[0]) <= t_rel * abs(ev[0])
        if ok:
            res["tie"] += 1
        else:
            res["bad"] += 1
            if len(res["bad_index"]) < 16:
                res["bad_index"].append(int(i))
    hit = same & (e["hit_type"] != 0)
    rel = np.abs(g["t"][hit].astype(np.float64) - e["t"][hit].astype(np.float64)) / np.maximum(np.abs(e["t"][hit].astype(np.float64)), 1e-30)
    res["t_bad"] = int(np.count_nonzero(rel > t_rel))
    surf = same & (e["hit_type"] == 1)
    db = np.abs(g["bary"][surf].astype(np.float64) - e["bary"][surf].astype(np.float64))
    res["bary_bad"] = int(np.count_nonzero(db.max(axis=1) > bary_abs)) if len(db) else 0
    return res
