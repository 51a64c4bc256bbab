"""Measure the BASELINE.json configurations that are not bench.py's headline (C1, C3, C4, C5) on one B200.

    python tools/bench_configs.py [--c4-tris 50000000] [--c4-rays 67108864] [--out gpurun_out/configs.json]

CUDA-event timings, 3 warm-ups, best of 5; inputs resident in HBM.  C4's geometry and rays are drawn on the device with
torch (same distributions as scenes.triangle_soup / uniform_rays, whose exact PCG32 streams are used by the parity tests
at smaller sizes): this tool measures throughput, tests/ prove parity.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import luisacompute_b200 as lc  # noqa: E402
from luisacompute_b200 import scenes as S  # noqa: E402


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def dev_rays(rays):
    return torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(-1)).cuda()


def counters(acc, d_rays, n, mask=0xFF):
    ctr = torch.zeros(4, dtype=torch.int64, device="cuda")
    acc.trace_closest(d_rays, mask, counters=ctr)
    torch.cuda.synchronize()
    c = ctr.cpu().numpy() / n
    return {"nodes": float(c[0]), "prims": float(c[1]), "instances": float(c[2]), "exact_evals": float(c[3]),
            "bytes_per_ray": float(32 + 24 + 80 * c[0] + 48 * c[1] + 80 * c[2])}


def c1(dev):
    verts, tris, xf = S.test_rtx_scene()
    opt = lc.AccelOption(allow_update=True)
    mesh = dev.create_mesh(verts, tris, option=opt).build()
    acc = dev.create_accel(opt)
    for a in xf:
        acc.emplace_back(mesh, a)
    acc.build()
    rays = S.test_rtx_rays(1024, 1024)
    d = dev_rays(rays)
    n = len(rays)
    hits = torch.empty((n, 6), dtype=torch.int32, device="cuda")
    t = timeit(lambda: acc.trace_closest(d, out=hits))
    # per-frame loop of test_rtx.cpp:122-139: vertex edit -> BLAS refit -> instance transform -> TLAS update
    frame = [0]

    def one_frame():
        frame[0] += 1
        mesh.vertices.view(torch.float32)[1] = 0.5 + 0.1 * np.sin(frame[0] * 0.1)      # same buffer, new contents
        mesh.build()
        acc.set_transform_on_update(1, S.affine_from_4x4(S.mat_translation(-0.25, 0.0, 0.1) @ S.mat_rotation((0, 0, 1), 0.5 + 0.01 * frame[0])))
        acc.build()
    tf = timeit(one_frame, reps=20, warm=5)
    return {"workload": "C1 test_rtx triangle, 2 instances, 1024x1024 orthographic closest-hit", "mrays_s": n / t / 1e3, "trace_ms": t,
            "dynamic_frame_us": tf * 1e3, "dynamic_frame": "vertex edit + BLAS refit + set_transform + TLAS update, host calls included",
            "hit_fraction": float((hits[:, 0] != -1).float().mean().item())}


def c3(dev, n_inst=100_000):
    blas = [S.icosphere(3, phase=float(k)) for k in range(16)]
    aff, bidx, masks, uids = S.instanced_grid(n_inst, 16, seed=42, grid=47)
    opt = lc.AccelOption(allow_update=True)
    meshes = [dev.create_mesh(v, t, option=opt).build() for v, t in blas]
    acc = dev.create_accel(opt)
    mods = S.make_mods([meshes[b].handle for b in bidx], aff, masks, None, uids)
    acc.add_mods(mods, n_inst, prims=[meshes[b] for b in bidx])
    t0 = time.perf_counter()
    acc.build()
    torch.cuda.synchronize()
    first_build_s = time.perf_counter() - t0
    tlas_ms = timeit(lambda: acc.build(lc.FORCE_BUILD), reps=5, warm=1)
    res = 2048
    cam = (47 * 0.9, 47 * 0.7, 47 * 1.6)
    prim = S.pinhole_rays(res, res, pos=cam, vfov_deg=45)
    n = len(prim)
    d_prim = dev_rays(prim)
    hits = torch.empty((n, 6), dtype=torch.int32, device="cuda")
    t_closest = timeit(lambda: acc.trace_closest(d_prim, 0xFF, out=hits))
    got = lc.hits_to_numpy(hits)
    shadow = S.shadow_rays(prim, got)
    d_sh = dev_rays(shadow)
    occ = torch.empty(n, dtype=torch.int32, device="cuda")
    t_any = timeit(lambda: acc.trace_any(d_sh, 0x55, out=occ))
    ctr = counters(acc, d_prim, n)
    # 10 % of the instances get a new transform, then one TLAS refit
    rng = np.random.default_rng(1)
    moved = rng.choice(n_inst, n_inst // 10, replace=False)

    def refit():
        m = np.zeros(len(moved), dtype=S.MOD_DTYPE)
        m["index"], m["flags"] = moved, S.MOD_TRANSFORM
        a = aff[moved].copy()
        a[:, [3, 7, 11]] += rng.uniform(-0.05, 0.05, (len(moved), 3)).astype(np.float32)
        m["affine"] = a
        acc.add_mods(m, n_inst)
        acc.build()
    t_refit = timeit(refit, reps=5, warm=2)
    assert acc.last_build_kind == 2
    return {"workload": f"C3 {n_inst} instances of 16 BLAS (icosphere L3), masks 1<<(i&7); {res}x{res} primary closest (0xff) + shadow any-hit (0x55)",
            "tlas_build_ms": tlas_ms, "tlas_build_minst_s": n_inst / tlas_ms / 1e3, "first_build_with_upload_s": first_build_s,
            "closest_mrays_s": n / t_closest / 1e3, "anyhit_shadow_mrays_s": n / t_any / 1e3, "primary_hit_fraction": float((got["inst"] != 0xFFFFFFFF).mean()),
            "shadow_occluded_fraction": float((occ != 0).float().mean().item()), "update_10pct_plus_refit_ms": t_refit, "per_primary_ray": ctr,
            "closest_achieved_gbs": ctr["bytes_per_ray"] * n / t_closest / 1e6}


def c4(dev, n_tris, n_rays):
    g = torch.Generator(device="cuda").manual_seed(1234)
    s = 0.004
    c = torch.rand((n_tris, 3), device="cuda", generator=g)
    e1 = (torch.rand((n_tris, 3), device="cuda", generator=g) * 2 - 1) * s
    e2 = (torch.rand((n_tris, 3), device="cuda", generator=g) * 2 - 1) * s
    verts = torch.zeros((n_tris, 3, 4), dtype=torch.float32, device="cuda")
    verts[:, 0, :3], verts[:, 1, :3], verts[:, 2, :3] = c, c + e1, c + e2
    del c, e1, e2
    verts = verts.view(-1, 4)
    tris = torch.arange(3 * n_tris, dtype=torch.int32, device="cuda").view(-1, 3)
    mesh = dev.create_mesh(verts, tris, vertex_stride=16)
    t0 = time.perf_counter()
    mesh.build()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    build_ms = timeit(lambda: mesh.build(lc.FORCE_BUILD), reps=3, warm=1)
    info = mesh.info()
    assert info.built == 1 and info.error == 0, (info.built, info.error)
    acc = dev.create_accel()
    acc.emplace_back(mesh)
    acc.build()
    g2 = torch.Generator(device="cuda").manual_seed(5678)
    rays = torch.zeros((n_rays, 8), dtype=torch.float32, device="cuda")
    rays[:, 0:3] = torch.rand((n_rays, 3), device="cuda", generator=g2)
    z = 1 - 2 * torch.rand(n_rays, device="cuda", generator=g2)
    ph = 2 * np.pi * torch.rand(n_rays, device="cuda", generator=g2)
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    rays[:, 4], rays[:, 5], rays[:, 6] = r * torch.cos(ph), r * torch.sin(ph), z
    rays[:, 7] = 3.4028234663852886e38
    del z, ph, r
    d_rays = rays.view(torch.uint8).reshape(-1)
    hits = torch.empty((n_rays, 6), dtype=torch.int32, device="cuda")
    t = timeit(lambda: acc.trace_closest(d_rays, out=hits), reps=3, warm=1)
    ctr = counters(acc, d_rays, n_rays)
    return {"workload": f"C4 {n_tris} random triangles (s=0.004), {n_rays} uniform incoherent closest-hit rays, 1 GPU",
            "build_ms": build_ms, "build_mtris_s": n_tris / build_ms / 1e3, "first_build_s": first, "bvh_nodes": int(info.node_count),
            "bvh_bytes": int(info.total_bytes), "trace_ms": t, "mrays_s": n_rays / t / 1e3, "hit_fraction": float((hits[:, 0] != -1).float().mean().item()),
            "per_ray": ctr, "achieved_gbs": ctr["bytes_per_ray"] * n_rays / t / 1e6}


def c5(dev, w=1920, h=1080, spp=16, so=None, reps=3, return_image=False):
    so = so or os.environ.get("LCRT_C5_LIB", os.path.join(ROOT, "tools", "libc5pt.so"))
    L = C.CDLL(so)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.c5_render.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, u32, u32, u32, u32, u32]
    meshes = S.cornell_room()
    aabbs = S.sphere_aabbs(64, radius=0.08)
    dm = [dev.create_mesh(v, t).build() for v, t in meshes]
    dp = dev.create_procedural_primitive(aabbs).build()
    acc = dev.create_accel()
    for m in dm:
        acc.emplace_back(m, S.affine_identity(), 0xFF, True, 0)
    acc.emplace_back(dp, S.affine_identity(), 0xFF, False, 0)
    acc.build()
    tri_first, tv = [], []
    for v, t in meshes:
        tri_first.append(len(tv) // 3)
        for tri in t:
            for k in tri:
                tv.append([*v[k, :3], 0.0])
    tri_first.append(len(tv) // 3)
    albedo = np.array([[0.725, 0.71, 0.68, 0], [0.725, 0.71, 0.68, 0], [0.725, 0.71, 0.68, 0], [0.63, 0.065, 0.05, 0], [0.14, 0.45, 0.091, 0],
                       [0.725, 0.71, 0.68, 0], [0.725, 0.71, 0.68, 0], [0, 0, 0, 0], [0.6, 0.6, 0.8, 0]], dtype=np.float32)
    d_tv = torch.from_numpy(np.asarray(tv, dtype=np.float32)).cuda()
    d_tf = torch.from_numpy(np.asarray(tri_first, dtype=np.uint32).view(np.int32)).cuda()
    d_al = torch.from_numpy(albedo).cuda()
    d_ab = torch.from_numpy(np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1)).cuda()
    image = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    seeds = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    rayc = torch.zeros(1, dtype=torch.int64, device="cuda")
    bh, bi = acc.binding()

    def render():
        rc = L.c5_render(torch.cuda.current_stream().cuda_stream, bh, bi, d_ab.data_ptr(), d_tv.data_ptr(), d_tf.data_ptr(), d_al.data_ptr(),
                         image.data_ptr(), seeds.data_ptr(), rayc.data_ptr(), w, h, spp, len(meshes) - 1, len(meshes))
        assert rc == 0
    if return_image:  # one dispatch from the initial sampler state: what two builds of the kernel must agree on bit for bit
        render()
        torch.cuda.synchronize()
        return image.cpu().numpy(), int(rayc.item())
    t = timeit(render, reps=reps, warm=1)
    rayc.zero_()
    render()
    torch.cuda.synchronize()
    rays = int(rayc.item())
    img = (image[..., :3] / image[..., 3:4]).cpu().numpy()
    return {"workload": f"C5 Cornell room (8 meshes, 36 triangles) + 64 procedural spheres via RayQuery handlers, {w}x{h}, {spp} spp per dispatch, depth <= 10",
            "dispatch_ms": t, "spp_per_s": spp / (t * 1e-3), "msamples_s": w * h * spp / t / 1e3, "mrays_s": rays / t / 1e3, "rays_per_sample": rays / (w * h * spp),
            "mean_radiance": [float(x) for x in img.reshape(-1, 3).mean(0)], "nan_pixels": int(np.isnan(img).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-tris", type=int, default=50_000_000)
    ap.add_argument("--c4-rays", type=int, default=67_108_864)
    ap.add_argument("--only", default="c1,c3,c4,c5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    a = ap.parse_args()
    dev = lc.Device(0)
    out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
           if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None}
    for name in a.only.split(","):
        t0 = time.perf_counter()
        if name == "c1":
            out["c1"] = c1(dev)
        elif name == "c3":
            out["c3"] = c3(dev)
        elif name == "c4":
            out["c4"] = c4(dev, a.c4_tris, a.c4_rays)
        elif name == "c5":
            out["c5"] = c5(dev)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        print(name, json.dumps(out[name]), f"[{time.perf_counter() - t0:.1f} s]", flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    dev.close()


if __name__ == "__main__":
    main()
