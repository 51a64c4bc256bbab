"""ctypes binding of tests/hostsim/libhostsim.so — TEST INFRASTRUCTURE ONLY (CPU execution of the device headers)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from luisacompute_b200 import scenes as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_TRIANGLES, KIND_AABBS, KIND_INSTANCES = 0, 1, 2
BUILD_ALLOW_UPDATE = 2
DEFAULT_CI = 0.5


_LIB_NAME = "libhostsim.so"


class variant:
    """with variant("libhostsim_recenter.so"): ...  — the same simulation built with other compile-time options (tests/hostsim/Makefile)"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        global _LIB, _LIB_NAME
        self.saved = (_LIB, _LIB_NAME)
        _LIB, _LIB_NAME = None, self.name
        return self

    def __exit__(self, *exc):
        global _LIB, _LIB_NAME
        _LIB, _LIB_NAME = self.saved
        return False


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(os.path.join(_HERE, _LIB_NAME))
        vp, u32, u64, f32, ci = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
        L.hs_blob_layout.argtypes = [u32, u32, u32, vp]
        L.hs_build_triangles.restype = ci
        L.hs_build_triangles.argtypes = [vp, u64, vp, u32, u32, f32, vp]
        L.hs_build_aabbs.restype = ci
        L.hs_build_aabbs.argtypes = [vp, u32, u32, f32, vp]
        L.hs_build_tlas.restype = ci
        L.hs_build_tlas.argtypes = [vp, u32, u32, f32, vp]
        L.hs_refit_triangles.restype = ci
        L.hs_refit_triangles.argtypes = [vp, u64, vp, u32, vp]
        L.hs_refit_aabbs.restype = ci
        L.hs_refit_aabbs.argtypes = [vp, u32, vp]
        L.hs_refit_tlas.restype = ci
        L.hs_refit_tlas.argtypes = [vp, u32, vp]
        L.hs_init_instances.argtypes = [vp, u32, u32]
        L.hs_apply_mods.argtypes = [vp, vp, u32]
        L.hs_validate.restype = ci
        L.hs_validate.argtypes = [vp, vp]
        L.hs_trace_closest.argtypes = [vp, vp, vp, u64, vp, ci, vp, vp]
        L.hs_trace_any.argtypes = [vp, vp, vp, u64, vp, ci, vp]
        L.hs_query.argtypes = [vp, vp, vp, u64, vp, ci, ci, u32, u32, vp]
        L.hs_query_suspending.argtypes = [vp, vp, vp, u64, vp, ci, ci, u32, u32, vp]
        L.hs_query_warp_schedule.restype = C.c_longlong
        L.hs_query_warp_schedule.argtypes = [vp, vp, vp, u64, u32, ci, u32, u32, vp, u32, u32, u32, u64]
        L.hs_set_defer.argtypes = [ci]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _blob(kind, n, flags):
    out = np.zeros(5, dtype=np.uint64)
    lib().hs_blob_layout(kind, n, flags, _p(out))
    # 256-byte aligned storage
    raw = np.zeros(int(out[0]) + 256, dtype=np.uint8)
    off = (-raw.ctypes.data) % 256
    return raw[off:off + int(out[0])], out


class SimMesh:
    def __init__(self, verts, tris, flags=0, ci=DEFAULT_CI):
        self.verts = np.ascontiguousarray(verts, dtype=np.float32)
        self.tris = np.ascontiguousarray(tris, dtype=np.uint32)
        self.flags = flags
        self.blob, self.layout = _blob(KIND_TRIANGLES, len(self.tris), flags)
        rc = lib().hs_build_triangles(_p(self.verts), self.verts.strides[0], _p(self.tris), len(self.tris), flags, ci, _p(self.blob))
        assert rc == 0, rc

    @property
    def handle(self):
        return self.blob.ctypes.data

    def refit(self, verts):
        self.verts = np.ascontiguousarray(verts, dtype=np.float32)
        rc = lib().hs_refit_triangles(_p(self.verts), self.verts.strides[0], _p(self.tris), len(self.tris), _p(self.blob))
        assert rc == 0, rc

    def validate(self):
        st = np.zeros(8, dtype=np.uint64)
        return lib().hs_validate(_p(self.blob), _p(st)), st


class SimProcedural:
    def __init__(self, aabbs, flags=0, ci=DEFAULT_CI):
        self.aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        self.blob, self.layout = _blob(KIND_AABBS, len(self.aabbs), flags)
        rc = lib().hs_build_aabbs(_p(self.aabbs), len(self.aabbs), flags, ci, _p(self.blob))
        assert rc == 0, rc

    @property
    def handle(self):
        return self.blob.ctypes.data

    def validate(self):
        st = np.zeros(8, dtype=np.uint64)
        return lib().hs_validate(_p(self.blob), _p(st)), st


class SimAccel:
    def __init__(self, flags=0, ci=DEFAULT_CI):
        self.instances = np.zeros(0, dtype=S.INSTANCE_DTYPE)
        self.flags, self.ci = flags, ci
        self.blob = None
        self._keep = []

    def update(self, instance_count, mods, update_instance_buffer_only=False, refit=False):
        if instance_count > len(self.instances):
            old = self.instances
            self.instances = np.zeros(instance_count, dtype=S.INSTANCE_DTYPE)
            self.instances[:len(old)] = old
            lib().hs_init_instances(_p(self.instances), len(old), instance_count - len(old))
        self.instances = self.instances[:instance_count]
        mods = np.ascontiguousarray(mods)
        if len(mods):
            lib().hs_apply_mods(_p(self.instances), _p(mods), len(mods))
        if update_instance_buffer_only:
            return
        if refit and self.blob is not None:
            rc = lib().hs_refit_tlas(_p(self.instances), instance_count, _p(self.blob))
        else:
            self.blob, self.layout = _blob(KIND_INSTANCES, instance_count, self.flags)
            rc = lib().hs_build_tlas(_p(self.instances), instance_count, self.flags, self.ci, _p(self.blob))
        assert rc == 0, rc

    def validate(self):
        st = np.zeros(8, dtype=np.uint64)
        return lib().hs_validate(_p(self.blob), _p(st)), st

    @staticmethod
    def _masks(mask, n):
        if np.isscalar(mask):
            return np.array([mask], dtype=np.uint32), 0
        return np.ascontiguousarray(mask, dtype=np.uint32), 1

    def trace_closest(self, rays, mask=0xFF, want_counters=False):
        rays = np.ascontiguousarray(rays)
        m, ms = self._masks(mask, len(rays))
        hits = np.zeros(len(rays), dtype=S.HIT_DTYPE)
        ctr = np.zeros(4, dtype=np.uint64)
        lib().hs_trace_closest(_p(self.blob), _p(self.instances), _p(rays), len(rays), _p(m), ms, _p(hits), _p(ctr))
        return (hits, ctr) if want_counters else hits

    def trace_any(self, rays, mask=0xFF):
        rays = np.ascontiguousarray(rays)
        m, ms = self._masks(mask, len(rays))
        out = np.zeros(len(rays), dtype=np.uint32)
        lib().hs_trace_any(_p(self.blob), _p(self.instances), _p(rays), len(rays), _p(m), ms, _p(out))
        return out

    def query_warp_schedule(self, rays, mask=0xFF, terminate_on_first=False, tri_prog=0, proc_prog=0, warps=8, vote=4, refill_min=12,
                            max_iterations=1 << 24):
        """The query through a lane-by-lane replay of the batch kernel's warp schedule.  Returns (hits, longest wait of a lane with
        a pending primitive in iterations; -1 if a warp exceeded max_iterations)."""
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), dtype=S.CHIT_DTYPE)
        worst = lib().hs_query_warp_schedule(_p(self.blob), _p(self.instances), _p(rays), len(rays), int(mask), int(terminate_on_first), tri_prog,
                                             proc_prog, _p(hits), warps, vote, refill_min, max_iterations)
        return hits, int(worst)

    def query(self, rays, mask=0xFF, terminate_on_first=False, tri_prog=0, proc_prog=0, suspending=False):
        rays = np.ascontiguousarray(rays)
        m, ms = self._masks(mask, len(rays))
        hits = np.zeros(len(rays), dtype=S.CHIT_DTYPE)
        (lib().hs_query_suspending if suspending else lib().hs_query)(_p(self.blob), _p(self.instances), _p(rays), len(rays), _p(m), ms, int(terminate_on_first), tri_prog, proc_prog, _p(hits))
        return hits
