"""luisacompute_b200 — B200-native ray-tracing acceleration backend for LuisaCompute's accel API.

Python front over the C-ABI shared library (luisacompute_b200/liblcb200_rt.so: sm_100a kernels + C++ host layer).
The class / method names follow the reference's runtime wrappers so that tests read like the reference's own
programs (paths relative to /root/reference):

    Device.create_mesh / create_procedural_primitive / create_accel   include/luisa/runtime/device.h, rtx/*.h
    Mesh.build(request)                                                src/runtime/rtx/mesh.cpp:6-12
    ProceduralPrimitive.build(request)                                 src/runtime/rtx/procedural_primitive.cpp:20-24
    Accel.emplace_back / set / pop_back / set_*_on_update / build / update_instance_buffer
                                                                       src/runtime/rtx/accel.cpp:68-257
    Accel.trace_closest / trace_any / query_all / query_any            DSL Expr<Accel>::intersect / intersect_any /
                                                                       query_all / query_any (src/dsl/rtx/accel.cpp:136-160)

PyTorch is used only for device memory and streams.  There is NO CPU fallback: importing works anywhere (so the
C-ABI surface can be checked), but creating a Device without a CUDA GPU or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import scenes  # noqa: F401  (host-side scene generators)

_HERE = os.path.dirname(os.path.abspath(__file__))
# LCB200_RT_LIB: developer override used by tools/trace_perf.py to time differently tuned builds of the same library
LIB_PATH = os.environ.get("LCB200_RT_LIB") or os.path.join(_HERE, "liblcb200_rt.so")
_LIB = None

KIND_TRIANGLES, KIND_AABBS, KIND_INSTANCES = 0, 1, 2
PREFER_UPDATE, FORCE_BUILD = 0, 1
TRI_COMMIT_ALL, TRI_CUTOUT_MOD3, TRI_IGNORE_ALL, TRI_TERMINATE_FIRST = 0, 1, 2, 3
PROC_SPHERE, PROC_BOX, PROC_IGNORE_ALL = 0, 1, 2


class LcrtError(RuntimeError):
    pass


class AccelOption(C.Structure):
    """include/luisa/runtime/rhi/resource.h:94-126"""
    _fields_ = [("hint", C.c_uint32), ("allow_compaction", C.c_uint32), ("allow_update", C.c_uint32)]

    def __init__(self, hint: int = 0, allow_compaction: bool = True, allow_update: bool = False):
        super().__init__(hint, int(allow_compaction), int(allow_update))


class _MeshBuildCommand(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("request", C.c_uint32), ("vertex_buffer", C.c_uint64), ("vertex_buffer_offset", C.c_uint64),
                ("vertex_buffer_size", C.c_uint64), ("vertex_stride", C.c_uint64), ("triangle_buffer", C.c_uint64),
                ("triangle_buffer_offset", C.c_uint64), ("triangle_buffer_size", C.c_uint64)]


class _ProcBuildCommand(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("request", C.c_uint32), ("aabb_buffer", C.c_uint64), ("aabb_buffer_offset", C.c_uint64),
                ("aabb_buffer_size", C.c_uint64)]


class _AccelBuildCommand(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("instance_count", C.c_uint32), ("request", C.c_uint32), ("update_instance_buffer_only", C.c_uint32),
                ("modification_count", C.c_uint32), ("modifications", C.c_void_p)]


class _AccelBinding(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("instances", C.c_uint64)]


class BlobInfo(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("prim_count", C.c_uint32), ("node_count", C.c_uint32), ("node_capacity", C.c_uint32),
                ("built", C.c_uint32), ("error", C.c_uint32), ("flags", C.c_uint32), ("prim_stride", C.c_uint32),
                ("lo", C.c_float * 3), ("hi", C.c_float * 3), ("nodes_offset", C.c_uint64), ("prims_offset", C.c_uint64),
                ("total_bytes", C.c_uint64), ("depth", C.c_uint32), ("max_stack", C.c_uint32), ("trace_error", C.c_uint32),
                ("reserved", C.c_uint32)]


class Sizes(C.Structure):
    _fields_ = [("blob_bytes", C.c_uint64), ("build_temp_bytes", C.c_uint64), ("update_temp_bytes", C.c_uint64)]


# every symbol include/lcb200_rt.h and include/lcb200_device.h declare (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "lcrtGetVersion", "lcrtGetErrorString", "lcrtGetLastCudaError", "lcrtAccelComputeMemoryUsage", "lcrtBuildTriangles",
    "lcrtRefitTriangles", "lcrtBuildAabbs", "lcrtRefitAabbs", "lcrtInstancesInit", "lcrtInstancesApplyMods",
    "lcrtInstancesApplyModsHost", "lcrtInstancesUpdateHandles", "lcrtBuildTlas", "lcrtRefitTlas", "lcrtTraceClosest", "lcrtTraceAny", "lcrtQuery",
    "lcrtTraceClosestTiled", "lcrtBlobGetInfo", "lcrtDeviceLibrarySource", "lcrtSortTempBytes", "lcrtSortPairs",
    "lcbDeviceCreate", "lcbDeviceDestroy", "lcbBackendName", "lcbCreateMesh", "lcbDestroyMesh", "lcbCreateProceduralPrimitive",
    "lcbDestroyProceduralPrimitive", "lcbCreateAccel", "lcbDestroyAccel", "lcbMeshBuild", "lcbProceduralPrimitiveBuild",
    "lcbAccelBuild", "lcbAccelGetBinding", "lcbLastBuildKind", "lcbNativeHandle", "lcbBlobInfo", "lcbTraceClosest", "lcbTraceAny",
    "lcbQuery", "lcbTraceClosestHost", "lcbSynchronize", "lcbDownload", "lcbUpload",
    "lcbPrimitiveExport", "lcbPrimitiveImport",
    "lcbTraceClosestTiled", "lcbFramebufferCreate", "lcbFramebufferOpen", "lcbFramebufferClose", "lcbFramebufferDestroy", "lcbFramebufferCopyTiles",
    # plumbing used by the backend plugin (plugin/lc_backend_b200rt.cpp)
    "lcbDeviceCount", "lcbDeviceName", "lcbBufferCreate", "lcbBufferDestroy", "lcbStreamCreate", "lcbStreamDestroy", "lcbCopyAsync",
    "lcbHostCallback", "lcbEventCreate", "lcbEventDestroy", "lcbEventSignal", "lcbEventWait", "lcbEventSynchronize", "lcbEventIsCompleted",
]


def lib():
    """Loads the CUDA library; raises if it has not been built (there is no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise LcrtError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"(make -C luisacompute_b200/csrc). luisacompute_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.lcrtGetVersion.restype = i32
    L.lcrtGetErrorString.restype = C.c_char_p
    L.lcrtGetErrorString.argtypes = [i32]
    L.lcrtGetLastCudaError.restype = C.c_char_p
    L.lcrtAccelComputeMemoryUsage.argtypes = [u32, u32, u32, C.POINTER(Sizes)]
    L.lcrtBuildTriangles.argtypes = [vp, vp, u64, vp, u32, u32, vp, u64, vp, u64]
    L.lcrtRefitTriangles.argtypes = [vp, vp, u64, vp, u32, vp, u64, vp]
    L.lcrtBuildAabbs.argtypes = [vp, vp, u32, u32, vp, u64, vp, u64]
    L.lcrtRefitAabbs.argtypes = [vp, vp, u32, vp, u64, vp]
    L.lcrtInstancesInit.argtypes = [vp, vp, u32, u32]
    L.lcrtInstancesApplyMods.argtypes = [vp, vp, vp, u32]
    L.lcrtInstancesApplyModsHost.argtypes = [vp, vp, vp, u32]
    L.lcrtInstancesUpdateHandles.argtypes = [vp, vp, vp, u32]
    L.lcrtBuildTlas.argtypes = [vp, vp, u32, u32, vp, u64, vp, u64]
    L.lcrtRefitTlas.argtypes = [vp, vp, u32, vp, u64, vp]
    L.lcrtTraceClosest.argtypes = [vp, vp, vp, vp, u64, vp, u32, vp, vp]
    L.lcrtTraceAny.argtypes = [vp, vp, vp, vp, u64, vp, u32, vp, vp]
    L.lcrtQuery.argtypes = [vp, vp, vp, vp, u64, vp, u32, i32, u32, u32, vp]
    L.lcrtBlobGetInfo.argtypes = [vp, vp, C.POINTER(BlobInfo)]
    L.lcrtDeviceLibrarySource.restype = C.c_char_p
    L.lcrtSortTempBytes.restype = u64
    L.lcrtSortTempBytes.argtypes = [u32]
    L.lcrtSortPairs.argtypes = [vp, vp, vp, u32, u32, vp, u64]
    L.lcbDeviceCreate.restype = vp
    L.lcbDeviceCreate.argtypes = [i32]
    L.lcbDeviceDestroy.argtypes = [vp]
    L.lcbBackendName.restype = C.c_char_p
    for name in ("lcbCreateMesh", "lcbCreateProceduralPrimitive", "lcbCreateAccel"):
        getattr(L, name).restype = u64
        getattr(L, name).argtypes = [vp, C.POINTER(AccelOption)]
    for name in ("lcbDestroyMesh", "lcbDestroyProceduralPrimitive", "lcbDestroyAccel"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [vp, u64]
    L.lcbMeshBuild.argtypes = [vp, vp, C.POINTER(_MeshBuildCommand)]
    L.lcbProceduralPrimitiveBuild.argtypes = [vp, vp, C.POINTER(_ProcBuildCommand)]
    L.lcbAccelBuild.argtypes = [vp, vp, C.POINTER(_AccelBuildCommand)]
    L.lcbAccelGetBinding.argtypes = [vp, u64, C.POINTER(_AccelBinding)]
    L.lcbLastBuildKind.argtypes = [vp, u64]
    L.lcbNativeHandle.restype = u64
    L.lcbNativeHandle.argtypes = [vp, u64]
    L.lcbBlobInfo.argtypes = [vp, vp, u64, C.POINTER(BlobInfo)]
    L.lcbTraceClosest.argtypes = [vp, vp, u64, vp, u64, vp, u32, vp, vp]
    L.lcbTraceAny.argtypes = [vp, vp, u64, vp, u64, vp, u32, vp, vp]
    L.lcbQuery.argtypes = [vp, vp, u64, vp, u64, vp, u32, i32, u32, u32, vp]
    L.lcrtTraceClosestTiled.argtypes = [vp, vp, vp, vp, u64, vp, u32, vp, vp, u64, u64, u32, vp]
    L.lcbTraceClosestTiled.argtypes = [vp, vp, u64, vp, u64, vp, u32, vp, vp, u64, u64, u32, vp]
    L.lcbFramebufferCreate.argtypes = [vp, u64, C.POINTER(u64), vp]
    L.lcbFramebufferOpen.argtypes = [vp, vp, C.POINTER(u64)]
    L.lcbFramebufferClose.argtypes = [vp, u64]
    L.lcbFramebufferDestroy.argtypes = [vp, u64]
    L.lcbFramebufferCopyTiles.argtypes = [vp, vp, vp, u64, u64, u64, u64]
    L.lcbTraceClosestHost.argtypes = [vp, u64, vp, u64, u32, vp, u64]
    L.lcbSynchronize.argtypes = [vp, vp]
    L.lcbDownload.argtypes = [vp, vp, u64, vp, u64]
    L.lcbPrimitiveExport.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64)]
    L.lcbPrimitiveImport.argtypes = [vp, vp, u64, vp, u64]
    L.lcbUpload.argtypes = [vp, vp, u64, vp, u64]
    # plumbing of the backend plugin (buffers, streams, copies, host callbacks, timeline events)
    L.lcbDeviceCount.restype = i32
    L.lcbDeviceName.argtypes = [i32, C.c_char_p, i32]
    L.lcbBufferCreate.restype = u64
    L.lcbBufferCreate.argtypes = [vp, u64]
    L.lcbBufferDestroy.restype = None
    L.lcbBufferDestroy.argtypes = [vp, u64]
    L.lcbStreamCreate.restype = vp
    L.lcbStreamCreate.argtypes = [vp]
    L.lcbStreamDestroy.restype = None
    L.lcbStreamDestroy.argtypes = [vp, vp]
    L.lcbCopyAsync.argtypes = [vp, vp, vp, vp, u64]
    L.lcbHostCallback.argtypes = [vp, vp, vp, vp]
    L.lcbEventCreate.restype = vp
    L.lcbEventCreate.argtypes = [vp]
    L.lcbEventDestroy.restype = None
    L.lcbEventDestroy.argtypes = [vp, vp]
    L.lcbEventSignal.argtypes = [vp, vp, vp, u64]
    L.lcbEventWait.argtypes = [vp, vp, vp, u64]
    L.lcbEventSynchronize.argtypes = [vp, vp, u64]
    L.lcbEventIsCompleted.argtypes = [vp, vp, u64]
    _LIB = L
    return L


_DEBUG_SYNC = os.environ.get("LCB200_DEBUG_SYNC") == "1"   # developer aid: name every call, then wait for the device


def _check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        raise LcrtError(f"{what} failed: {L.lcrtGetErrorString(rc).decode()} [{L.lcrtGetLastCudaError().decode()}]")
    if _DEBUG_SYNC:
        import sys
        import torch
        print(f"[lcb200] {what} ...", end="", file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        print(" done", file=sys.stderr, flush=True)


def _stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def _as_device_bytes(x, device):
    """numpy array / torch tensor -> contiguous CUDA tensor (kept alive by the caller)."""
    import torch
    if isinstance(x, np.ndarray):
        x = np.ascontiguousarray(x)
        t = torch.from_numpy(x.view(np.uint8).reshape(-1)).to(device, non_blocking=False)
        return t
    assert x.is_cuda, "device tensors only"
    return x.contiguous()


def device_library_source() -> str:
    return lib().lcrtDeviceLibrarySource().decode()


def compute_memory_usage(kind: int, prim_count: int, flags: int = 0) -> Sizes:
    s = Sizes()
    _check(lib().lcrtAccelComputeMemoryUsage(kind, prim_count, flags, C.byref(s)), "lcrtAccelComputeMemoryUsage")
    return s


class Device:
    """`Context::create_device("b200rt")` analogue (src/runtime/context.cpp:155-177 -> extern "C" create)."""

    def __init__(self, index: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise LcrtError("no CUDA device: luisacompute_b200 has no CPU fallback")
        self.index = torch.cuda.current_device() if index is None else index
        self.torch_device = torch.device("cuda", self.index)
        self._h = lib().lcbDeviceCreate(self.index)
        if not self._h:
            raise LcrtError("lcbDeviceCreate failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().lcbDeviceDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def backend_name(self) -> str:
        return lib().lcbBackendName().decode()

    def create_mesh(self, vertices=None, triangles=None, option: Optional[AccelOption] = None, vertex_stride: Optional[int] = None) -> "Mesh":
        return Mesh(self, vertices, triangles, option or AccelOption(), vertex_stride)

    def create_procedural_primitive(self, aabbs=None, option: Optional[AccelOption] = None) -> "ProceduralPrimitive":
        return ProceduralPrimitive(self, aabbs, option or AccelOption())

    def create_accel(self, option: Optional[AccelOption] = None) -> "Accel":
        return Accel(self, option or AccelOption())

    def synchronize(self):
        _check(lib().lcbSynchronize(self._h, _stream()), "lcbSynchronize")


class _Resource:
    _destroy = ""

    def __init__(self, device: Device, handle: int):
        self.device, self.handle = device, handle

    def destroy(self):
        if getattr(self, "handle", None) is not None and self.device._h:
            getattr(lib(), self._destroy)(self.device._h, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def native_handle(self) -> int:
        return lib().lcbNativeHandle(self.device._h, self.handle)

    @property
    def last_build_kind(self) -> int:
        return lib().lcbLastBuildKind(self.device._h, self.handle)

    def info(self) -> BlobInfo:
        b = BlobInfo()
        _check(lib().lcbBlobInfo(self.device._h, _stream(), self.handle, C.byref(b)), "lcbBlobInfo")
        return b


class _Primitive(_Resource):
    def export_blob(self):
        """The built BLAS as one position-independent uint8 device tensor (NCCL-broadcastable, serialisable)."""
        import torch
        n = C.c_uint64(0)
        _check(lib().lcbPrimitiveExport(self.device._h, _stream(), self.handle, None, 0, C.byref(n)), "lcbPrimitiveExport")
        t = torch.empty(n.value, dtype=torch.uint8, device=self.device.torch_device)
        _check(lib().lcbPrimitiveExport(self.device._h, _stream(), self.handle, t.data_ptr(), t.numel(), C.byref(n)), "lcbPrimitiveExport")
        return t

    def import_blob(self, blob):
        _check(lib().lcbPrimitiveImport(self.device._h, _stream(), self.handle, blob.data_ptr(), blob.numel()), "lcbPrimitiveImport")
        return self


class Mesh(_Primitive):
    """Triangle BLAS over borrowed vertex / triangle buffers (include/luisa/runtime/rtx/mesh.h)."""
    _destroy = "lcbDestroyMesh"

    def __init__(self, device: Device, vertices, triangles, option: AccelOption, vertex_stride: Optional[int]):
        super().__init__(device, lib().lcbCreateMesh(device._h, C.byref(option)))
        self.option = option
        self.vertices = self.triangles = None
        self.vertex_stride = vertex_stride
        if vertices is not None:
            self.set_buffers(vertices, triangles, vertex_stride)

    def set_buffers(self, vertices, triangles, vertex_stride: Optional[int] = None):
        if isinstance(vertices, np.ndarray):
            vertex_stride = vertex_stride or vertices.strides[0]
        else:
            vertex_stride = vertex_stride or vertices.stride(0) * vertices.element_size()
        self.vertex_stride = vertex_stride
        self.vertices = _as_device_bytes(vertices, self.device.torch_device)
        self.triangles = _as_device_bytes(triangles, self.device.torch_device)

    def build(self, request: int = PREFER_UPDATE):
        v, t = self.vertices, self.triangles
        cmd = _MeshBuildCommand(self.handle, request, v.data_ptr(), 0, v.numel() * v.element_size(), self.vertex_stride, t.data_ptr(), 0,
                                t.numel() * t.element_size())
        _check(lib().lcbMeshBuild(self.device._h, _stream(), C.byref(cmd)), "lcbMeshBuild")
        return self


class ProceduralPrimitive(_Primitive):
    _destroy = "lcbDestroyProceduralPrimitive"

    def __init__(self, device: Device, aabbs, option: AccelOption):
        super().__init__(device, lib().lcbCreateProceduralPrimitive(device._h, C.byref(option)))
        self.option = option
        self.aabbs = None
        if aabbs is not None:
            self.set_buffer(aabbs)

    def set_buffer(self, aabbs):
        self.aabbs = _as_device_bytes(aabbs, self.device.torch_device)

    def build(self, request: int = PREFER_UPDATE):
        a = self.aabbs
        cmd = _ProcBuildCommand(self.handle, request, a.data_ptr(), 0, a.numel() * a.element_size())
        _check(lib().lcbProceduralPrimitiveBuild(self.device._h, _stream(), C.byref(cmd)), "lcbProceduralPrimitiveBuild")
        return self


class Accel(_Resource):
    """TLAS + instance buffer; host-side modification bookkeeping as in src/runtime/rtx/accel.cpp."""
    _destroy = "lcbDestroyAccel"

    def __init__(self, device: Device, option: AccelOption):
        super().__init__(device, lib().lcbCreateAccel(device._h, C.byref(option)))
        self.option = option
        self._mods: dict[int, np.ndarray] = {}
        self._bulk: list = []
        self._instance_count = 0
        self._prims: list = []

    def __len__(self):
        return self._instance_count

    def _mod(self, index: int) -> np.ndarray:
        m = self._mods.get(index)
        if m is None:
            m = np.zeros(1, dtype=scenes.MOD_DTYPE)
            m["index"] = index
            self._mods[index] = m
        return m

    @staticmethod
    def _affine(transform) -> np.ndarray:
        a = np.asarray(transform, dtype=np.float32)
        if a.shape == (4, 4):
            a = a[:3, :]
        return a.reshape(12)

    # ---- src/runtime/rtx/accel.cpp:88-157
    def emplace_back(self, prim, transform=None, visibility_mask: int = 0xFF, opaque: bool = True, user_id: int = 0):
        index = self._instance_count
        self._instance_count += 1
        self._prims.append(prim)
        self.set(index, prim, transform, visibility_mask, opaque, user_id)

    def set(self, index: int, prim, transform=None, visibility_mask: int = 0xFF, opaque: bool = True, user_id: int = 0):
        assert index < self._instance_count
        if isinstance(prim, ProceduralPrimitive):
            opaque = False                     # src/runtime/rtx/accel.cpp:154-157
        m = self._mod(index)
        m["primitive"] = prim.handle
        m["affine"][0] = self._affine(scenes.affine_identity() if transform is None else transform)
        m["vis_mask"] = visibility_mask & 0xFF
        m["user_id"] = user_id
        m["flags"] = (scenes.MOD_PRIMITIVE | scenes.MOD_TRANSFORM | scenes.MOD_VISIBILITY | scenes.MOD_USER_ID |
                      (scenes.MOD_OPAQUE_ON if opaque else scenes.MOD_OPAQUE_OFF))
        self._prims[index] = prim

    def pop_back(self):
        if self._instance_count:
            self._instance_count -= 1
            self._mods.pop(self._instance_count, None)
            self._prims.pop()

    def set_transform_on_update(self, index: int, transform):
        m = self._mod(index)
        m["affine"][0] = self._affine(transform)
        m["flags"] |= scenes.MOD_TRANSFORM

    def set_visibility_on_update(self, index: int, mask: int):
        m = self._mod(index)
        m["vis_mask"] = mask & 0xFF
        m["flags"] |= scenes.MOD_VISIBILITY

    def set_opaque_on_update(self, index: int, opaque: bool):
        m = self._mod(index)
        m["flags"] = (int(m["flags"][0]) & ~(scenes.MOD_OPAQUE_ON | scenes.MOD_OPAQUE_OFF)) | (scenes.MOD_OPAQUE_ON if opaque else scenes.MOD_OPAQUE_OFF)

    def set_instance_user_id_on_update(self, index: int, user_id: int):
        m = self._mod(index)
        m["user_id"] = user_id
        m["flags"] |= scenes.MOD_USER_ID

    def add_mods(self, mods: np.ndarray, instance_count: Optional[int] = None, prims=None):
        """Bulk path: ready-made 80-byte Modification records (primitive = Mesh / ProceduralPrimitive handles).  The
        records must not repeat an instance index that already has a pending modification (the reference merges
        per-instance modifications on the host, src/runtime/rtx/accel.cpp:75-79; a bulk batch is taken as merged)."""
        assert mods.dtype.itemsize == 80
        self._bulk.append(np.ascontiguousarray(mods).copy())
        if instance_count is not None:
            self._instance_count = instance_count
        if prims is not None:
            self._prims = list(prims)

    # ---- src/runtime/rtx/accel.cpp:68-86, 255-257
    def build(self, request: int = PREFER_UPDATE, update_instance_buffer_only: bool = False):
        if self._instance_count == 0:
            raise LcrtError("Building acceleration structure without instances.")
        nb = sum(len(b) for b in self._bulk)
        mods = np.zeros(len(self._mods) + nb, dtype=scenes.MOD_DTYPE)     # (np.concatenate would repack the padded 80-byte dtype)
        for k, m in enumerate(self._mods.values()):
            mods[k] = m[0]
        k = len(self._mods)
        for b in self._bulk:
            mods[k:k + len(b)] = b
            k += len(b)
        assert mods.dtype.itemsize == 80 and mods.flags["C_CONTIGUOUS"]
        self._mods.clear()
        self._bulk.clear()
        cmd = _AccelBuildCommand(self.handle, self._instance_count, request, int(update_instance_buffer_only), len(mods),
                                 mods.ctypes.data_as(C.c_void_p) if len(mods) else None)
        _check(lib().lcbAccelBuild(self.device._h, _stream(), C.byref(cmd)), "lcbAccelBuild")
        return self

    def update_instance_buffer(self):
        return self.build(PREFER_UPDATE, update_instance_buffer_only=True)

    def binding(self):
        b = _AccelBinding()
        _check(lib().lcbAccelGetBinding(self.device._h, self.handle, C.byref(b)), "lcbAccelGetBinding")
        return b.handle, b.instances

    def instance_buffer(self) -> np.ndarray:
        """Debug read-back of the 80-byte instance records."""
        _, inst = self.binding()
        out = np.zeros(self._instance_count, dtype=scenes.INSTANCE_DTYPE)
        _check(lib().lcbDownload(self.device._h, _stream(), inst, out.ctypes.data_as(C.c_void_p), out.nbytes), "lcbDownload")
        return out

    # ---- batch tracing: rays / hits are device tensors (or numpy arrays, which are copied)
    def _rays(self, rays):
        import torch
        if isinstance(rays, np.ndarray):
            assert rays.dtype.itemsize == 32
            return torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(-1)).to(self.device.torch_device), len(rays)
        n = rays.numel() * rays.element_size() // 32
        return rays.contiguous(), n

    def _masks(self, mask):
        import torch
        if isinstance(mask, (int, np.integer)):
            return None, int(mask), None
        t = torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint32).view(np.int32) if isinstance(mask, np.ndarray) else mask)
        t = t.to(self.device.torch_device).contiguous()
        return t.data_ptr(), 0, t

    def trace_closest(self, rays, mask=0xFF, out=None, counters=None):
        """accel.intersect(ray, mask) over a batch.  Returns an int32 [n, 6] tensor of SurfaceHit records."""
        import torch
        r, n = self._rays(rays)
        mp, mv, keep = self._masks(mask)
        if out is None:
            out = torch.empty((n, 6), dtype=torch.int32, device=self.device.torch_device)
        _check(lib().lcbTraceClosest(self.device._h, _stream(), self.handle, r.data_ptr(), n, mp, mv, out.data_ptr(),
                                     counters.data_ptr() if counters is not None else None), "lcbTraceClosest")
        return out

    def trace_closest_tiled(self, rays, dst_address: int, tile_rays: int, tile_stride_bytes: int, mask=0xFF, scratch=None, counters=None,
                            store_each_record: bool = False):
        """trace_closest whose final records go to dst_address + (i // tile_rays) * tile_stride_bytes + (i % tile_rays) * 24
        (a framebuffer of tiles, possibly in a peer GPU's memory: luisacompute_b200.dist.PeerFramebuffer).  `scratch`: int32
        [n, 6] on this device where the records of a 128-ray batch wait until the batch is complete (allocated if None);
        store_each_record: every record goes out on its own when its ray ends (LCRT_TILED_STORE_EACH_RECORD)."""
        import torch
        r, n = self._rays(rays)
        mp, mv, keep = self._masks(mask)
        if scratch is None:
            scratch = torch.empty((n, 6), dtype=torch.int32, device=self.device.torch_device)
        _check(lib().lcbTraceClosestTiled(self.device._h, _stream(), self.handle, r.data_ptr(), n, mp, mv, scratch.data_ptr(), dst_address,
                                          tile_rays, tile_stride_bytes, 1 if store_each_record else 0,
                                          counters.data_ptr() if counters is not None else None),
               "lcbTraceClosestTiled")
        return scratch

    def trace_any(self, rays, mask=0xFF, out=None, counters=None):
        import torch
        r, n = self._rays(rays)
        mp, mv, keep = self._masks(mask)
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=self.device.torch_device)
        _check(lib().lcbTraceAny(self.device._h, _stream(), self.handle, r.data_ptr(), n, mp, mv, out.data_ptr(),
                                 counters.data_ptr() if counters is not None else None), "lcbTraceAny")
        return out

    def query(self, rays, mask=0xFF, terminate_on_first: bool = False, tri_program: int = 0, proc_program: int = 0, out=None):
        import torch
        r, n = self._rays(rays)
        mp, mv, keep = self._masks(mask)
        if out is None:
            out = torch.empty((n, 6), dtype=torch.int32, device=self.device.torch_device)
        _check(lib().lcbQuery(self.device._h, _stream(), self.handle, r.data_ptr(), n, mp, mv, int(terminate_on_first), tri_program,
                              proc_program, out.data_ptr()), "lcbQuery")
        return out

    def query_all(self, rays, mask=0xFF, **kw):
        return self.query(rays, mask, False, **kw)

    def query_any(self, rays, mask=0xFF, **kw):
        return self.query(rays, mask, True, **kw)

    def trace_closest_host(self, host_rays, host_hits, mask: int = 0xFF, chunk_rays: int = 1 << 20):
        """End-to-end path: pinned host rays -> (H2D, trace, D2H pipelined in chunks) -> pinned host hits."""
        n = host_rays.numel() * host_rays.element_size() // 32
        _check(lib().lcbTraceClosestHost(self.device._h, self.handle, host_rays.data_ptr(), n, mask, host_hits.data_ptr(), chunk_rays),
               "lcbTraceClosestHost")
        return host_hits


def hits_to_numpy(t) -> np.ndarray:
    """int32 [n, 6] device tensor -> structured SurfaceHit array."""
    return t.detach().cpu().numpy().view(scenes.HIT_DTYPE).reshape(-1)


def chits_to_numpy(t) -> np.ndarray:
    return t.detach().cpu().numpy().view(scenes.CHIT_DTYPE).reshape(-1)
