"""Developer tool (GPU box): the C2 workload through this repository AND through the reference's own CUDA/OptiX backend
(baseline/_ref/bin/ref_optix_bench, built from baseline/ref_optix_bench.cpp against the unmodified reference), same files.

1. dumps the scene and both ray batches, 2. runs the reference program (times it, gets its hits), 3. traces the same
rays here, 4. compares the two result sets record by record under the BASELINE.json rule (ids exact unless the two
nearest hits are within 1e-5 relative t; t 1e-5 relative; barycentrics 1e-4) and prints one JSON object.
usage: python tools/ref_optix_compare.py [--level 8] [--res 2048] [--dir /tmp/lc_ref_case] [--out gpurun_out/ref_optix.json]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import luisacompute_b200 as lc  # noqa: E402
from luisacompute_b200 import scenes as S  # noqa: E402

EXE = os.path.join(ROOT, "baseline", "_ref", "bin", "ref_optix_bench")


def timeit(fn, reps=10, warm=3):
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


def compare(mine, ref, rays, oracle_acc=None):
    """mine / ref: HIT_DTYPE arrays.  Returns counts; disagreements on ids are re-judged by evaluating both candidates."""
    miss_m, miss_r = mine["inst"] == 0xFFFFFFFF, ref["inst"] == 0xFFFFFFFF
    both = ~miss_m & ~miss_r
    same_id = both & (mine["inst"] == ref["inst"]) & (mine["prim"] == ref["prim"])
    rel_t = np.abs(mine["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
    out = {"n": int(len(mine)), "hit_mine": int((~miss_m).sum()), "hit_ref": int((~miss_r).sum()), "hit_miss_disagree": int((miss_m != miss_r).sum()),
           "same_id": int(same_id.sum()), "diff_id": int((both & ~same_id).sum()),
           "diff_id_outside_tie_band": int((both & ~same_id & (rel_t > 1e-5)).sum()),
           "t_bad_same_id": int((same_id & (rel_t > 1e-5)).sum()),
           "bary_bad_same_id": int((same_id & (np.abs(mine["bary"] - ref["bary"]).max(axis=1) > 1e-4)).sum()),
           "max_rel_t_same_id": float(rel_t[same_id].max()) if same_id.any() else 0.0,
           "max_bary_diff_same_id": float(np.abs(mine["bary"] - ref["bary"])[same_id].max()) if same_id.any() else 0.0}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--res", type=int, default=2048)
    ap.add_argument("--dir", default="/tmp/lc_ref_case")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_optix.json"))
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    dev = lc.Device(0)
    verts, tris = S.icosphere(a.level)
    v4 = np.zeros((len(verts), 4), dtype=np.float32)
    v4[:, :3] = np.asarray(verts, dtype=np.float32)[:, :3]
    v4.tofile(os.path.join(a.dir, "verts.bin"))
    np.ascontiguousarray(tris, dtype=np.uint32).tofile(os.path.join(a.dir, "tris.bin"))
    mesh = dev.create_mesh(v4, np.ascontiguousarray(tris, dtype=np.uint32)).build()
    acc = dev.create_accel()
    acc.emplace_back(mesh, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    build_ms = timeit(lambda: mesh.build(lc.FORCE_BUILD), reps=5, warm=1)
    primary = S.pinhole_rays(a.res, a.res)
    hp = lc.hits_to_numpy(acc.trace_closest(primary))
    bounce = S.bounce_rays(primary, hp, verts, tris, a.res, a.res)
    primary.tofile(os.path.join(a.dir, "rays_primary.bin"))
    bounce.tofile(os.path.join(a.dir, "rays_bounce.bin"))
    result = {"workload": f"C2 icosphere level {a.level} ({len(tris)} triangles), {a.res}x{a.res} primary + bounce rays", "mine": {"mesh_build_ms": build_ms}}
    n = len(primary)
    tdev = torch.device("cuda:0")
    for name, rays in (("primary", primary), ("bounce", bounce)):
        d = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(tdev)
        hits = torch.empty((n, 6), dtype=torch.int32, device=tdev)
        occ = torch.empty(n, dtype=torch.int32, device=tdev)
        tc = timeit(lambda: acc.trace_closest(d, out=hits))
        ta = timeit(lambda: acc.trace_any(d, out=occ))
        result["mine"][name] = {"rays": n, "closest_ms": tc, "closest_mrays_s": n / tc / 1e3, "any_ms": ta, "any_mrays_s": n / ta / 1e3}
        lc.hits_to_numpy(hits).tofile(os.path.join(a.dir, f"mine_hits_{name}.bin"))
        occ.cpu().numpy().astype(np.uint32).tofile(os.path.join(a.dir, f"mine_any_{name}.bin"))
    dev.close()
    del acc, mesh
    torch.cuda.empty_cache()
    if not os.path.exists(EXE):
        result["reference"] = {"unavailable": "baseline/_ref/bin/ref_optix_bench not built"}
    else:
        t0 = time.time()
        r = subprocess.run([EXE, a.dir], capture_output=True, text=True, timeout=900)
        result["reference_wall_s"] = time.time() - t0
        if r.returncode != 0 or not os.path.exists(os.path.join(a.dir, "ref_result.json")):
            result["reference"] = {"unavailable": f"reference program failed (rc {r.returncode})", "log_tail": (r.stdout + r.stderr)[-1500:]}
        else:
            result["reference"] = json.load(open(os.path.join(a.dir, "ref_result.json")))
            result["parity_vs_reference"] = {}
            for name, rays in (("primary", primary), ("bounce", bounce)):
                mine = np.fromfile(os.path.join(a.dir, f"mine_hits_{name}.bin"), dtype=S.HIT_DTYPE)
                ref = np.fromfile(os.path.join(a.dir, f"ref_hits_{name}.bin"), dtype=S.HIT_DTYPE)
                c = compare(mine, ref, rays)
                ma = np.fromfile(os.path.join(a.dir, f"mine_any_{name}.bin"), dtype=np.uint32)
                ra = np.fromfile(os.path.join(a.dir, f"ref_any_{name}.bin"), dtype=np.uint32)
                c["any_disagree"] = int(((ma != 0) != (ra != 0)).sum())
                result["parity_vs_reference"][name] = c
            for name in ("primary", "bounce"):
                result[f"speedup_closest_{name}"] = result["mine"][name]["closest_mrays_s"] / result["reference"][name]["closest_mrays_s"]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(result, open(a.out, "w"), indent=1)
    print(json.dumps(result))


if __name__ == "__main__":
    main()
