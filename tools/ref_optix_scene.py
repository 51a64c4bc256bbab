"""Developer tool (GPU box): an INSTANCED scene with visibility masks, non-opaque instances and one set of procedural
AABB spheres through the reference's own CUDA/OptiX backend (baseline/_ref/bin/ref_optix_bench <dir> scene <mask>), through
this repository and through the oracle — intersect, intersect_any and a RayQuery whose triangle handler commits unless
prim % 3 == 0 (LCRT_TRI_CUTOUT_MOD3) and whose procedural handler intersects the inscribed sphere (LCRT_PROC_SPHERE).

The scene: 3 meshes (a level-3 icosphere, a 4000-triangle soup, the test_rtx triangle) and 40 spheres; 24 instances with
rotations / non-uniform placement, masks 1 << (i % 8), every third mesh instance non-opaque, user ids; ray mask 0x5b.
Prints one JSON object with the three-way comparison; with --golden PATH also packs the scene, the rays and the
REFERENCE's results into an .npz for tests/golden/.
usage: python tools/ref_optix_scene.py [--rays 65536] [--dir /tmp/lc_ref_scene] [--golden tests/golden/ref_optix_scene.npz]
"""
import argparse
import json
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import luisacompute_b200 as lc  # noqa: E402
from luisacompute_b200 import scenes as S  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests.gpu_util import build_pair  # noqa: E402

BIN = os.path.join(ROOT, "baseline", "_ref", "bin")
RAY_MASK = 0x5B


def make_scene():
    ico = S.icosphere(3)
    sv, st = S.triangle_soup(4000, seed=77, s=0.05)
    tv, tt, _ = S.test_rtx_scene()
    meshes = [(np.ascontiguousarray(ico[0]), np.ascontiguousarray(ico[1])), (np.ascontiguousarray(sv), np.ascontiguousarray(st)),
              (np.ascontiguousarray(tv), np.ascontiguousarray(tt))]
    boxes = S.sphere_aabbs(40, radius=0.12, lo=(-0.8, -0.8, -0.8), hi=(0.8, 0.8, 0.8), seed=5)
    rng = np.random.default_rng(11)
    aff, blas, masks, opaque, uid = [], [], [], [], []
    for i in range(24):
        axis = rng.normal(size=3)
        m = S.mat_translation(*rng.uniform(-2.0, 2.0, 3)) @ S.mat_rotation(axis / np.linalg.norm(axis), rng.uniform(0, 6.28)) @ S.mat_scaling(rng.uniform(0.4, 1.3))
        aff.append(S.affine_from_4x4(m))
        b = i % 4                                  # 0..2 meshes, 3 = the sphere set
        blas.append(b)
        masks.append(1 << (i % 8))
        opaque.append(False if b == 3 else (i % 3 != 0))
        uid.append(1000 + i)
    return meshes, boxes, np.asarray(aff, dtype=np.float32), blas, masks, opaque, uid


def make_rays(n):
    side = int(np.sqrt(n // 2))
    a = S.pinhole_rays(side, side, pos=(0.5, 0.3, 7.0), vfov_deg=50.0)
    b = S.uniform_rays(n - len(a), seed=99)
    b["o"] = (b["o"] - 0.5) * 5.0                  # origins in [-2.5, 2.5]^3, directions uniform
    return np.concatenate([a, b])


def write_scene(path, meshes, boxes, aff, blas, masks, opaque, uid):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(meshes)))
        for v, t in meshes:
            v4 = np.zeros((len(v), 4), dtype=np.float32)
            v4[:, :3] = np.asarray(v, dtype=np.float32)[:, :3]
            f.write(struct.pack("<I", len(v4))), f.write(v4.tobytes())
            f.write(struct.pack("<I", len(t))), f.write(np.ascontiguousarray(t, dtype=np.uint32).tobytes())
        f.write(struct.pack("<I", 1))
        f.write(struct.pack("<I", len(boxes))), f.write(np.ascontiguousarray(boxes, dtype=np.float32).tobytes())
        f.write(struct.pack("<I", len(blas)))
        for i in range(len(blas)):
            b = 0x80000000 if blas[i] == 3 else blas[i]
            f.write(struct.pack("<IIII", b, masks[i], int(opaque[i]), uid[i])), f.write(np.asarray(aff[i], dtype=np.float32).tobytes())


def compare_closest(a, b, tol_t, tol_bary):
    ma, mb = a["inst"] == 0xFFFFFFFF, b["inst"] == 0xFFFFFFFF
    both = ~ma & ~mb
    same = both & (a["inst"] == b["inst"]) & (a["prim"] == b["prim"])
    rel = np.abs(a["t"] - b["t"]) / np.maximum(np.abs(b["t"]), 1e-30)
    return {"hit_miss_disagree": int((ma != mb).sum()), "hits": int(both.sum()), "diff_id": int((both & ~same).sum()),
            "diff_id_outside_tie_band": int((both & ~same & (rel > 1e-5)).sum()), "t_bad": int((same & (rel > tol_t)).sum()),
            "bary_bad": int((same & (np.abs(a["bary"] - b["bary"]).max(axis=1) > tol_bary)).sum()), "max_rel_t": float(rel[same].max()) if same.any() else 0.0}


def compare_query(a, b, tol_t):
    """CommittedHit arrays: hit type, ids; t only (procedural hits carry no barycentrics)."""
    same_type = a["hit_type"] == b["hit_type"]
    hit = same_type & (a["hit_type"] != 0)
    same = hit & (a["inst"] == b["inst"]) & (a["prim"] == b["prim"])
    rel = np.abs(a["t"] - b["t"]) / np.maximum(np.abs(b["t"]), 1e-30)
    return {"type_disagree": int((~same_type).sum()), "surface": int((b["hit_type"] == 1).sum()), "procedural": int((b["hit_type"] == 2).sum()),
            "diff_id": int((hit & ~same).sum()), "diff_id_outside_tie_band": int((hit & ~same & (rel > 1e-5)).sum()),
            "t_bad": int((same & (rel > tol_t)).sum()), "max_rel_t": float(rel[same].max()) if same.any() else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--dir", default="/tmp/lc_ref_scene")
    ap.add_argument("--golden", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_optix_scene.json"))
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    meshes, boxes, aff, blas, masks, opaque, uid = make_scene()
    rays = make_rays(a.rays)
    write_scene(os.path.join(a.dir, "scene.bin"), meshes, boxes, aff, blas, masks, opaque, uid)
    rays.tofile(os.path.join(a.dir, "rays_scene.bin"))
    dev = lc.Device(0)
    oa, da, _ = build_pair(O, dev, meshes + [boxes], aff, blas_index=blas, masks=masks, opaque=opaque, user_ids=uid)
    mine = {"closest": lc.hits_to_numpy(da.trace_closest(rays, RAY_MASK)), "any": da.trace_any(rays, RAY_MASK).cpu().numpy() != 0,
            "query": lc.chits_to_numpy(da.query_all(rays, RAY_MASK, tri_program=1, proc_program=0))}
    orc = {"closest": oa.trace_closest(rays, RAY_MASK)[0], "any": oa.trace_any(rays, RAY_MASK) != 0,
           "query": oa.query(rays, RAY_MASK, False, 1, 0)}
    dev.close()
    env = dict(os.environ, LD_LIBRARY_PATH=BIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([os.path.join(BIN, "ref_optix_bench"), a.dir, "scene", hex(RAY_MASK)], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode != 0:
        print(json.dumps({"reference": {"unavailable": (r.stdout + r.stderr)[-1500:]}}))
        return 1
    ref = {"closest": np.fromfile(os.path.join(a.dir, "ref_scene_hits.bin"), dtype=S.HIT_DTYPE),
           "any": np.fromfile(os.path.join(a.dir, "ref_scene_any.bin"), dtype=np.uint32) != 0,
           "query": np.fromfile(os.path.join(a.dir, "ref_scene_query.bin"), dtype=S.CHIT_DTYPE)}
    res = {"scene": "3 meshes + 40 spheres, 24 instances, masks 1<<(i%8), ray mask 0x5b", "rays": int(len(rays)),
           "reference": json.load(open(os.path.join(a.dir, "ref_scene_result.json")))}
    for who, x in (("mine", mine), ("oracle", orc)):
        res[f"{who}_vs_reference"] = {"closest": compare_closest(x["closest"], ref["closest"], 5e-5, 2e-2), "any_disagree": int((x["any"] != ref["any"]).sum()),
                                      "query": compare_query(x["query"], ref["query"], 5e-5)}
    res["mine_vs_oracle"] = {"closest": compare_closest(mine["closest"], orc["closest"], 1e-5, 1e-4), "any_disagree": int((mine["any"] != orc["any"]).sum()),
                             "query": compare_query(mine["query"], orc["query"], 1e-5)}
    if a.golden:
        np.savez_compressed(a.golden, scene=np.fromfile(os.path.join(a.dir, "scene.bin"), dtype=np.uint8), rays=rays.view(np.uint8).reshape(-1, 32),
                            ray_mask=np.uint32(RAY_MASK), hits=ref["closest"].view(np.uint8).reshape(-1, 24), occluded=ref["any"].astype(np.uint8),
                            query=ref["query"].view(np.uint8).reshape(-1, 24))
        res["golden"] = {"path": a.golden, "bytes": os.path.getsize(a.golden)}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
