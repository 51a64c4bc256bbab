"""Developer tool (GPU box): BASELINE config C3 (100 000 instances of 16 BLAS, masks) through the reference's own CUDA/OptiX
backend and through this repository: TLAS build time, 2048 x 2048 primary closest-hit rays, record-by-record comparison.
usage: python tools/ref_optix_c3.py [--instances 100000] [--res 2048] [--out gpurun_out/ref_optix_c3.json]"""
import argparse
import json
import os
import struct
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import luisacompute_b200 as lc  # noqa: E402
from luisacompute_b200 import scenes as S  # noqa: E402
from tools.ref_optix_compare import timeit  # noqa: E402
from tools.ref_optix_scene import compare_closest  # noqa: E402

BIN = os.path.join(ROOT, "baseline", "_ref", "bin")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=100_000)
    ap.add_argument("--res", type=int, default=2048)
    ap.add_argument("--dir", default="/tmp/lc_ref_c3")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_optix_c3.json"))
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    blas = [S.icosphere(3, phase=float(k)) for k in range(16)]
    aff, bidx, masks, uids = S.instanced_grid(a.instances, 16, seed=42, grid=47)
    with open(os.path.join(a.dir, "scene.bin"), "wb") as f:
        f.write(struct.pack("<I", len(blas)))
        for v, t in blas:
            v4 = np.zeros((len(v), 4), dtype=np.float32)
            v4[:, :3] = np.asarray(v, dtype=np.float32)[:, :3]
            f.write(struct.pack("<I", len(v4))), f.write(v4.tobytes())
            f.write(struct.pack("<I", len(t))), f.write(np.ascontiguousarray(t, dtype=np.uint32).tobytes())
        f.write(struct.pack("<I", 0))
        f.write(struct.pack("<I", a.instances))
        rec = np.zeros(a.instances, dtype=np.dtype([("blas", "<u4"), ("mask", "<u4"), ("opaque", "<u4"), ("user_id", "<u4"), ("affine", "<f4", 12)]))
        rec["blas"], rec["mask"], rec["opaque"], rec["user_id"], rec["affine"] = bidx, masks, 1, uids, np.asarray(aff, dtype=np.float32).reshape(-1, 12)
        f.write(rec.tobytes())
    cam = (47 * 0.9, 47 * 0.7, 47 * 1.6)
    rays = S.pinhole_rays(a.res, a.res, pos=cam, vfov_deg=45)
    rays.tofile(os.path.join(a.dir, "rays_scene.bin"))
    dev = lc.Device(0)
    meshes = [dev.create_mesh(v, t).build() for v, t in blas]
    acc = dev.create_accel()
    acc.add_mods(S.make_mods([meshes[b].handle for b in bidx], aff, masks, None, uids), a.instances, prims=[meshes[b] for b in bidx])
    acc.build()
    tlas_ms = timeit(lambda: acc.build(lc.FORCE_BUILD), reps=5, warm=1)
    d = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    hits = torch.empty((len(rays), 6), dtype=torch.int32, device="cuda")
    tc = timeit(lambda: acc.trace_closest(d, 0xFF, out=hits))
    mine = lc.hits_to_numpy(hits)
    dev.close()
    res = {"workload": f"C3 {a.instances} instances of 16 BLAS, {a.res}x{a.res} primary closest-hit rays, mask 0xff",
           "mine": {"tlas_build_ms": tlas_ms, "closest_ms": tc, "closest_mrays_s": len(rays) / tc / 1e3}}
    env = dict(os.environ, LD_LIBRARY_PATH=BIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([os.path.join(BIN, "ref_optix_bench"), a.dir, "scene", "0xff"], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode != 0:
        res["reference"] = {"unavailable": (r.stdout + r.stderr)[-1200:]}
    else:
        res["reference"] = json.load(open(os.path.join(a.dir, "ref_scene_result.json")))
        ref = np.fromfile(os.path.join(a.dir, "ref_scene_hits.bin"), dtype=S.HIT_DTYPE)
        res["parity_vs_reference"] = compare_closest(mine, ref, 5e-4, 2e-2)
        res["speedup_closest"] = res["mine"]["closest_mrays_s"] / res["reference"]["closest_mrays_s"]
        res["speedup_tlas_build"] = res["reference"]["tlas_build_ms"] / tlas_ms
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
