"""Drop-in boundary test: the `b200rt` backend plugin inside the reference's own (unmodified) host runtime.

`baseline/_ref/bin/test_plugin_rtx` (plugin/test_plugin_rtx.cpp) is a C++ program written against LuisaCompute's public
runtime API — Context / Device / Stream / Buffer / Mesh / ProceduralPrimitive / Accel — and linked against the
reference's liblc-runtime; the Context dlopens `liblc-backend-b200rt.so` (plugin/lc_backend_b200rt.cpp), whose
DeviceInterface forwards the accel path to liblcb200_rt.so.  The program dumps every case it ran (geometry, instances,
rays, results); here the results are checked against the oracle on the same inputs, to the BASELINE.json bar.

The binaries are built by `make -C plugin` when the reference tree and its built host libraries are present (this
container); they travel to the GPU box under baseline/_ref/ (git-ignored, not gpurun-ignored).  Absent binaries skip.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from luisacompute_b200 import scenes as S

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "bin")
EXE = os.path.join(BIN, "test_plugin_rtx")
needs_plugin = pytest.mark.skipif(not os.path.exists(EXE), reason="plugin integration binaries not built (need the reference tree; make -C plugin)")

INST_DTYPE = np.dtype([("blas", "<u4"), ("mask", "<u4"), ("opaque", "<u4"), ("user_id", "<u4"), ("affine", "<f4", 12)])


def _read_cases(path):
    raw = open(path, "rb").read()
    pos = 0

    def u32():
        nonlocal pos
        v = struct.unpack_from("<I", raw, pos)[0]
        pos += 4
        return v

    def arr(dtype, count):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=pos).copy()
        pos += a.nbytes
        return a

    cases = []
    while pos < len(raw):
        name = raw[pos:pos + 32].split(b"\0")[0].decode()
        pos += 32
        meshes = []
        for _ in range(u32()):
            v = arr("<f4", 4 * u32()).reshape(-1, 4)
            t = arr("<u4", 3 * u32()).reshape(-1, 3)
            meshes.append((v, t))
        boxes = [arr("<f4", 6 * u32()).reshape(-1, 6) for _ in range(u32())]
        inst = arr(INST_DTYPE, u32())
        n_rays, ray_mask = u32(), u32()
        rays = arr(S.RAY_DTYPE, n_rays)
        closest = arr("<u4", u32()).view(S.HIT_DTYPE)
        anyhit = arr("<u4", u32())
        query = arr("<u4", u32()).view(S.CHIT_DTYPE)
        cases.append(dict(name=name, meshes=meshes, boxes=boxes, inst=inst, rays=rays, ray_mask=ray_mask, closest=closest, any=anyhit, query=query))
    return cases


def _oracle_scene(oracle, case):
    geoms = [oracle.OracleMesh(v, t) for v, t in case["meshes"]]
    procs = [oracle.OracleProcedural(b) for b in case["boxes"]]
    acc = oracle.OracleAccel()
    acc._keep = geoms + procs
    inst = case["inst"]
    handles = [(procs[b & 0x7FFFFFFF] if b & 0x80000000 else geoms[b]).handle for b in inst["blas"]]
    mods = S.make_mods(handles, inst["affine"], inst["mask"], inst["opaque"] != 0, inst["user_id"])
    assert acc.update(len(inst), mods) == 0
    return acc


@needs_plugin
@pytest.mark.gpu
def test_plugin_in_reference_runtime(oracle, tmp_path):
    out = str(tmp_path / "plugin_cases.bin")
    r = subprocess.run([EXE, out, "192"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "b200rt device:" in r.stdout + r.stderr          # backend_device_names went through the plugin
    cases = _read_cases(out)
    assert [c["name"] for c in cases] == ["c1_static"] + [f"c1_frame{i}" for i in range(4)] + ["c1_three_masked", "c5_procedural"]
    for c in cases:
        acc = _oracle_scene(oracle, c)
        res = oracle.compare_closest(acc, c["rays"], c["closest"], int(c["ray_mask"]))
        assert res["id_mismatch"] == 0 and res["t_bad"] == 0 and res["bary_bad"] == 0, (c["name"], res)
        assert np.array_equal(c["any"] != 0, acc.trace_any(c["rays"], int(c["ray_mask"])) != 0), c["name"]
        if len(c["query"]):
            e = acc.query(c["rays"], int(c["ray_mask"]), False, 0, 0)
            g = c["query"]
            assert np.array_equal(e["hit_type"], g["hit_type"]), c["name"]
            hit = e["hit_type"] != 0
            assert np.array_equal(e["inst"][hit], g["inst"][hit]), c["name"]
            np.testing.assert_allclose(g["t"][hit], e["t"][hit], rtol=1e-5)
            # rays of the pixel grid that run exactly along the floor's diagonal hit both of its triangles at the same t:
            # a different primitive id is accepted only there (BASELINE.json tie rule), never for the spheres
            diff = hit & (e["prim"] != g["prim"])
            assert not np.any(diff & (e["hit_type"] != 1)) and diff.sum() <= 2 * int(np.sqrt(len(g))), (c["name"], int(diff.sum()))
            assert (e["hit_type"] == 2).sum() > 0 and (e["hit_type"] == 1).sum() > 0   # both spheres and the floor are seen
    # the frame loop really moved things: results differ between frames
    assert not np.array_equal(cases[1]["closest"]["t"], cases[4]["closest"]["t"])
    # the third instance (mask 0x0f, user id 7) is visible to ray mask 0x03 and nearest where it covers the others
    assert (cases[5]["closest"]["inst"] == 2).sum() > 0


@needs_plugin
def test_plugin_loads_and_fails_loudly_without_gpu():
    """CPU side of the boundary: the reference's Context discovers and dlopens the plugin and calls `create`; without a
    CUDA device that must be a logged error and an abort (the reference's failure mode), never a silent CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for GPU-less hosts")
    r = subprocess.run([EXE, "/dev/null", "8"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stdout + r.stderr
    assert "liblc-backend-b200rt.so" in r.stdout + r.stderr      # the stack trace goes through the plugin's create()


@pytest.mark.gpu
def test_plugin_plumbing_entry_points():
    """The C API the plugin builds DeviceInterface's buffer / stream / event methods from, driven directly: copies are
    stream ordered, host callbacks run after the work queued before them, events are timeline fences."""
    import ctypes as C

    import luisacompute_b200 as lc
    L = lc.lib()
    assert L.lcbDeviceCount() >= 1
    name = C.create_string_buffer(256)
    assert L.lcbDeviceName(0, name, 256) == 0 and len(name.value) > 0
    dev = lc.Device(0)
    h = dev._h
    n = 1 << 22
    src = np.arange(n, dtype=np.uint32)
    dst = np.zeros(n, dtype=np.uint32)
    a, b = L.lcbBufferCreate(h, src.nbytes), L.lcbBufferCreate(h, src.nbytes)
    assert a not in (0, 0xFFFFFFFFFFFFFFFF) and b not in (0, 0xFFFFFFFFFFFFFFFF)
    s0, s1 = L.lcbStreamCreate(h), L.lcbStreamCreate(h)
    assert s0 and s1
    ev = L.lcbEventCreate(h)
    assert L.lcbEventIsCompleted(h, ev, 0) == 1 and L.lcbEventIsCompleted(h, ev, 1) == 0
    order = []
    CB = C.CFUNCTYPE(None, C.c_void_p)
    cb = CB(lambda _: order.append(int(dst[-1])))
    # stream 0: upload, device-to-device copy, signal 1; stream 1: wait 1, download, callback (sees the data), signal 2
    assert L.lcbCopyAsync(h, s0, a, src.ctypes.data, src.nbytes) == 0
    assert L.lcbCopyAsync(h, s0, b, a, src.nbytes) == 0
    assert L.lcbEventSignal(h, ev, s0, 1) == 0
    assert L.lcbEventWait(h, ev, s1, 1) == 0
    assert L.lcbCopyAsync(h, s1, dst.ctypes.data, b, src.nbytes) == 0
    assert L.lcbHostCallback(h, s1, C.cast(cb, C.c_void_p), None) == 0
    assert L.lcbEventSignal(h, ev, s1, 2) == 0
    assert L.lcbEventWait(h, ev, s0, 5) != 0            # nobody has signalled 5: an error, not a hang
    assert L.lcbEventSynchronize(h, ev, 2) == 0
    assert L.lcbEventIsCompleted(h, ev, 2) == 1 and L.lcbEventIsCompleted(h, ev, 1) == 1
    assert order == [n - 1] and np.array_equal(dst, src)
    L.lcbEventDestroy(h, ev)
    L.lcbStreamDestroy(h, s0), L.lcbStreamDestroy(h, s1)
    L.lcbBufferDestroy(h, a), L.lcbBufferDestroy(h, b)
    dev.close()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "test_c_api_rtx")), reason="flat C API program not built (need the reference tree; make -C plugin)")
def test_flat_c_api_drives_the_plugin(oracle, tmp_path):
    """a plain-C program over the reference's flat C API (include/luisa/api/api.h; liblc-api.so is the reference's
    src/api/runtime.cpp + logging.cpp compiled where they lie) creates the mesh / accel, builds them through an LCCommandList
    (LCMeshBuildCommand, LCAccelBuildCommand with LCAccelBuildModification records) on the `b200rt` plugin and traces the
    test_rtx.cpp scene; its hit records must meet the oracle's on the same scene and rays."""
    out = str(tmp_path / "c_api_hits.bin")
    w = h = 320
    env = dict(os.environ, LD_LIBRARY_PATH=BIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([os.path.join(BIN, "test_c_api_rtx"), out, str(w), str(h)], capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, (r.returncode, (r.stdout + r.stderr)[-1500:])
    raw = open(out, "rb").read()
    assert struct.unpack_from("<II", raw, 0) == (w, h)
    n = w * h
    xf = np.frombuffer(raw, dtype="<f4", count=24, offset=8).reshape(2, 12).copy()
    rays = np.frombuffer(raw, dtype=S.RAY_DTYPE, count=n, offset=8 + 96).copy()
    got = np.frombuffer(raw, dtype=S.HIT_DTYPE, count=n, offset=8 + 96 + 32 * n).copy()
    occ = np.frombuffer(raw, dtype="<u4", count=n, offset=8 + 96 + 56 * n).copy()
    verts, tris, xf_ref = S.test_rtx_scene()
    np.testing.assert_allclose(xf, np.asarray(xf_ref), atol=1e-6)               # the program does build test_rtx.cpp's instances
    np.testing.assert_allclose(rays["o"], S.test_rtx_rays(w, h)["o"], atol=1e-6)
    om = oracle.OracleMesh(verts, tris)
    oa = oracle.OracleAccel()
    oa._keep = [om]
    assert oa.update(2, S.make_mods([om.handle, om.handle], [xf[0], xf[1]])) == 0
    res = oracle.compare_closest(oa, rays, got)
    assert res["id_mismatch"] == 0 and res["t_bad"] == 0 and res["bary_bad"] == 0, res
    assert res["exact"] + res["tie"] + res["grazing"] == n and res["tie"] + res["grazing"] <= 8, res
    assert np.array_equal(occ != 0, got["inst"] != 0xFFFFFFFF)
    assert 10_000 < np.count_nonzero(got["inst"] != 0xFFFFFFFF) < n
