"""Developer tool (GPU box): BASELINE config C5 — Cornell room + 64 procedural spheres, path traced through RayQuery handlers,
1920 x 1080 — as a LuisaCompute DSL program (baseline/c5_dsl_path_tracer.cpp) on the reference's own `cuda` (OptiX) backend
and on the drop-in `b200` backend; spp/s, Mrays/s and image agreement.  The native counterpart over the device library
(tools/c5_path_tracer.cu) is in tools/bench_configs.py.
usage: python tools/ref_optix_c5.py [--width 1920 --height 1080 --spp 16] [--out gpurun_out/ref_optix_c5.json]"""
import argparse
import json
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luisacompute_b200 import scenes as S  # noqa: E402

BIN = os.path.join(ROOT, "baseline", "_ref", "bin")
ALBEDO = np.array([[0.725, 0.71, 0.68, 0], [0.725, 0.71, 0.68, 0], [0.725, 0.71, 0.68, 0], [0.63, 0.065, 0.05, 0], [0.14, 0.45, 0.091, 0],
                   [0.725, 0.71, 0.68, 0], [0.725, 0.71, 0.68, 0], [0, 0, 0, 0], [0.6, 0.6, 0.8, 0]], dtype=np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--dir", default="/tmp/lc_ref_c5")
    ap.add_argument("--flat", action="store_true", help="the program's flat (sample, bounce) loop instead of the nested loops")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_optix_c5.json"))
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    meshes = S.cornell_room()
    boxes = S.sphere_aabbs(64, radius=0.08)
    with open(os.path.join(a.dir, "c5_scene.bin"), "wb") as f:
        f.write(struct.pack("<I", len(meshes)))
        for v, t in meshes:
            f.write(struct.pack("<I", len(v))), f.write(np.ascontiguousarray(v, dtype=np.float32).tobytes())
            f.write(struct.pack("<I", len(t))), f.write(np.ascontiguousarray(t, dtype=np.uint32).tobytes())
        f.write(struct.pack("<I", 1)), f.write(struct.pack("<I", len(boxes))), f.write(np.ascontiguousarray(boxes, dtype=np.float32).tobytes())
        n_inst = len(meshes) + 1
        f.write(struct.pack("<I", n_inst))
        for i in range(n_inst):
            blas = 0x80000000 if i == len(meshes) else i
            f.write(struct.pack("<IIII", blas, 0xFF, 0 if i == len(meshes) else 1, i)), f.write(S.affine_identity().astype(np.float32).tobytes())
        f.write(ALBEDO.tobytes())
        f.write(struct.pack("<I", len(meshes) - 1))
    env = dict(os.environ, LD_LIBRARY_PATH=BIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    res, imgs = {"workload": f"C5 Cornell room + 64 procedural spheres via RayQuery handlers, {a.width}x{a.height}, {a.spp} spp per dispatch, depth <= 10, DSL program, " +
           ("one flat loop over (sample, bounce)" if a.flat else "nested sample / bounce loops")}, {}
    for backend in ("cuda", "b200"):
        r = subprocess.run([os.path.join(BIN, "c5_dsl_path_tracer"), a.dir, str(a.width), str(a.height), str(a.spp), "3"], capture_output=True, text=True,
                           timeout=900, env=dict(env, LC_BACKEND=backend, C5_FLAT_LOOP="1" if a.flat else "0"))
        line = [x for x in r.stdout.splitlines() if x.startswith("{")]
        if r.returncode != 0 or not line:
            res[backend] = {"unavailable": (r.stdout + r.stderr)[-1200:]}
            continue
        res[backend] = json.loads(line[-1])
        img = np.fromfile(os.path.join(a.dir, f"c5_{backend}.bin"), dtype=np.float32).reshape(a.height, a.width, 4)
        imgs[backend] = img[..., :3] / np.maximum(img[..., 3:4], 1.0)
    if len(imgs) == 2:
        d = np.abs(imgs["cuda"] - imgs["b200"])
        res["image_agreement"] = {"mean_abs_diff": float(d.mean()), "mean_radiance_cuda": float(imgs["cuda"].mean()), "mean_radiance_b200": float(imgs["b200"].mean()),
                                  "relative_mean_diff": float(d.mean() / max(imgs["cuda"].mean(), 1e-9))}
        res["speedup_spp_per_s"] = res["b200"]["spp_per_s"] / res["cuda"]["spp_per_s"]
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
