"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_mesh_create.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_procedural_create.restype = C.c_void_p
        L.orc_procedural_create.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_geometry_destroy.argtypes = [C.c_void_p]
        L.orc_accel_create.restype = C.c_void_p
        L.orc_accel_destroy.argtypes = [C.c_void_p]
        L.orc_accel_update.restype = C.c_int
        L.orc_accel_update.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        L.orc_accel_set_instance.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.orc_accel_instance_user_id.restype = C.c_uint32
        L.orc_accel_instance_user_id.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_void_p]
        L.orc_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_query.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p,
                                C.c_int, C.c_int]
        L.orc_eval_hit.restype = C.c_int
        L.orc_eval_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


HIT_DTYPE = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", 2), ("t", "<f4"), ("pad", "<u4")])
CHIT_DTYPE = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", 2), ("hit_type", "<u4"), ("t", "<f4")])


class OracleMesh:
    def __init__(self, verts: np.ndarray, tris: np.ndarray):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        self.handle = lib().orc_mesh_create(_ptr(verts), verts.strides[0], len(verts), _ptr(tris), len(tris))

    def __del__(self):
        if getattr(self, "handle", None):
            lib().orc_geometry_destroy(self.handle)
            self.handle = None


class OracleProcedural:
    def __init__(self, aabbs: np.ndarray):
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        self.handle = lib().orc_procedural_create(_ptr(aabbs), len(aabbs))

    def __del__(self):
        if getattr(self, "handle", None):
            lib().orc_geometry_destroy(self.handle)
            self.handle = None


class OracleAccel:
    def __init__(self):
        self.handle = lib().orc_accel_create()
        self._keep = []

    def __del__(self):
        if getattr(self, "handle", None):
            lib().orc_accel_destroy(self.handle)
            self.handle = None

    def update(self, instance_count: int, mods: np.ndarray, update_instance_buffer_only: bool = False, cpu_flavour: bool = False) -> int:
        mods = np.ascontiguousarray(mods)
        assert mods.dtype.itemsize == 80
        return lib().orc_accel_update(self.handle, instance_count, _ptr(mods), len(mods), int(update_instance_buffer_only), int(cpu_flavour))

    def set_instance(self, index, affine=None, mask=None, opaque=None, user_id=None):
        a = None if affine is None else np.ascontiguousarray(affine, dtype=np.float32)
        lib().orc_accel_set_instance(self.handle, index, None if a is None else _ptr(a), mask is not None, mask or 0,
                                     opaque is not None, int(bool(opaque)), user_id is not None, user_id or 0)

    @staticmethod
    def _masks(mask, n):
        if np.isscalar(mask):
            return np.array([mask], dtype=np.uint32), 0
        m = np.ascontiguousarray(mask, dtype=np.uint32)
        assert len(m) == n
        return m, 1

    def trace_closest(self, rays: np.ndarray, mask=0xFF, threads: int = 0, brute: bool = False, want_counters: bool = False):
        rays = np.ascontiguousarray(rays)
        n = len(rays)
        m, ms = self._masks(mask, n)
        hits = np.zeros(n, dtype=HIT_DTYPE)
        t2 = np.zeros(n, dtype=np.float64)
        tex = np.zeros(n, dtype=np.float64)
        ctr = np.zeros(2, dtype=np.uint64)
        lib().orc_trace_closest(self.handle, _ptr(rays), n, _ptr(m), ms, _ptr(hits), _ptr(t2), _ptr(tex), threads, int(brute), _ptr(ctr))
        if want_counters:
            return hits, t2, tex, ctr
        return hits, t2, tex

    def trace_any(self, rays: np.ndarray, mask=0xFF, threads: int = 0, brute: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        n = len(rays)
        m, ms = self._masks(mask, n)
        out = np.zeros(n, dtype=np.uint8)
        lib().orc_trace_any(self.handle, _ptr(rays), n, _ptr(m), ms, _ptr(out), threads, int(brute))
        return out

    def query(self, rays: np.ndarray, mask=0xFF, terminate_on_first: bool = False, tri_prog: int = 0, proc_prog: int = 0,
              threads: int = 0, brute: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        n = len(rays)
        m, ms = self._masks(mask, n)
        hits = np.zeros(n, dtype=CHIT_DTYPE)
        lib().orc_query(self.handle, _ptr(rays), n, _ptr(m), ms, int(terminate_on_first), tri_prog, proc_prog, _ptr(hits), threads, int(brute))
        return hits

    def eval_hit(self, ray: np.ndarray, inst: int, prim: int):
        out = np.zeros(4, dtype=np.float64)
        r = np.ascontiguousarray(ray)
        ok = lib().orc_eval_hit(self.handle, _ptr(r), int(inst), int(prim), _ptr(out))
        return bool(ok), out


def compare_closest(oracle: OracleAccel, rays: np.ndarray, got: np.ndarray, mask=0xFF, threads: int = 0,
                    t_rel: float = 1e-5, bary_abs: float = 1e-4, ref=None) -> dict:
    """The BASELINE.json parity rule for closest-hit batches.

    inst/prim must match exactly, except where the claimed primitive — evaluated in double by the oracle —
    is itself a hit whose t lies within `t_rel` of the oracle's nearest t (the tie rule, generalised to
    N-way ties at shared vertices).  t within t_rel relative, barycentrics within bary_abs.
    Rays where one side reports a miss and the other a hit count as mismatches unless the hit grazes an
    edge within bary_abs (reported separately as `grazing`).
    """
    if ref is None:
        ref = oracle.trace_closest(rays, mask, threads)
    exp, t2, tex = ref
    n = len(rays)
    miss_e = exp["inst"] == 0xFFFFFFFF
    miss_g = got["inst"] == 0xFFFFFFFF
    same = (exp["inst"] == got["inst"]) & ((exp["prim"] == got["prim"]) | miss_e)
    res = {"n": n, "exact": int(np.count_nonzero(same)), "tie": 0, "grazing": 0, "id_mismatch": 0, "t_bad": 0, "bary_bad": 0,
           "bad_index": []}
    both = same & ~miss_e
    tg = got["t"][both].astype(np.float64)
    te = tex[both]
    rel = np.abs(tg - te) / np.maximum(np.abs(te), 1e-30)
    res["t_bad"] = int(np.count_nonzero(rel > t_rel))
    res["t_max_rel"] = float(rel.max()) if len(rel) else 0.0
    db = np.abs(got["bary"][both].astype(np.float64) - exp["bary"][both].astype(np.float64)).max(axis=1) if both.any() else np.zeros(0)
    res["bary_bad"] = int(np.count_nonzero(db > bary_abs))
    res["bary_max_abs"] = float(db.max()) if len(db) else 0.0
    for i in np.nonzero(~same)[0]:
        ok = False
        be = exp["bary"][i].astype(np.float64)
        graze_e = (not miss_e[i]) and min(be[0], be[1], 1 - be[0] - be[1]) <= bary_abs
        t_near = tex[i] if not miss_e[i] else np.inf
        if not miss_g[i]:
            valid, ev = oracle.eval_hit(rays[i:i + 1], got["inst"][i], got["prim"][i])
            t_claim, minb = ev[0], ev[3]
            valid = valid and minb >= -bary_abs and abs(float(got["t"][i]) - t_claim) <= t_rel * abs(t_claim)
            if valid and minb >= 0.0 and not miss_e[i] and abs(t_claim - t_near) <= t_rel * abs(t_near):
                res["tie"] += 1       # the tie rule (N-way ties included)
                ok = True
            elif valid and minb < 0.0 and t_claim <= t_near * (1 + t_rel):
                res["grazing"] += 1   # claimed hit grazes an edge within the barycentric tolerance
                ok = True
            elif valid and graze_e and t_claim > t_near:
                res["grazing"] += 1   # the oracle's nearer hit grazes an edge; the device skipped it
                ok = True
        elif graze_e:
            res["grazing"] += 1
            ok = True
        if not ok:
            res["id_mismatch"] += 1
            if len(res["bad_index"]) < 16:
                res["bad_index"].append(int(i))
    return res
