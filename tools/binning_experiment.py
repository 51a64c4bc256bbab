"""Developer tool (GPU box): what would ray reordering buy the batch kernel?  
Sorts the C2 batches by several keys WITH TORCH (cost not counted: this measures the upper bound of the benefit), traces the
reordered rays with the unchanged kernel and reports Mrays/s and nodes per ray next to pixel order.
usage: python tools/binning_experiment.py [--out gpurun_out/r2_binning.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import luisacompute_b200 as lc  # noqa: E402
from luisacompute_b200 import scenes as S  # noqa: E402


def timeit(fn, reps=5, warm=2):
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


def part1by2(x):
    x = x & 0x3ff
    x = (x | (x << 16)) & 0x30000ff
    x = (x | (x << 8)) & 0x300f00f
    x = (x | (x << 4)) & 0x30c30c3
    x = (x | (x << 2)) & 0x9249249
    return x


def morton3(p, lo, hi, bits=10):
    q = ((p - lo) / (hi - lo).clamp(min=1e-20) * ((1 << bits) - 1)).clamp(0, (1 << bits) - 1).to(torch.int64)
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = lc.Device(0)
    verts, tris = S.icosphere(8)
    mesh = dev.create_mesh(verts, tris).build()
    acc = dev.create_accel()
    acc.emplace_back(mesh, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    res = 2048
    primary = S.pinhole_rays(res, res)
    n = len(primary)
    d_p = torch.from_numpy(primary.view(np.uint8).reshape(-1)).cuda()
    hits = acc.trace_closest(d_p)
    bounce = S.bounce_rays(primary, lc.hits_to_numpy(hits), verts, tris, res, res)
    d_b = torch.from_numpy(bounce.view(np.uint8).reshape(-1)).cuda()
    out = {}

    def measure(name, rays_u8):
        h = torch.empty((n, 6), dtype=torch.int32, device="cuda")
        ms = timeit(lambda: acc.trace_closest(rays_u8, out=h))
        ctr = torch.zeros(4, dtype=torch.int64, device="cuda")
        acc.trace_closest(rays_u8, out=h, counters=ctr)
        torch.cuda.synchronize()
        c = ctr.cpu().numpy() / n
        out[name] = {"mrays_s": n / ms / 1e3, "ms": ms, "nodes_per_ray": float(c[0]), "prims_per_ray": float(c[1])}
        print(name, out[name], flush=True)

    for label, d in (("coherent", d_p), ("incoherent", d_b)):
        r = d.view(torch.float32).view(n, 8)
        o, dirs = r[:, 0:3], r[:, 4:7]
        measure(label + "/pixel_order", d)
        lo, hi = o.min(0).values, o.max(0).values
        octant = ((dirs[:, 0] < 0).long() | ((dirs[:, 1] < 0).long() << 1) | ((dirs[:, 2] < 0).long() << 2))
        keys = {
            "origin_morton30": morton3(o, lo, hi),
            "octant_then_origin_morton30": (octant << 30) | morton3(o, lo, hi),
            "origin_morton15_then_dir_morton15": (morton3(o, lo, hi, 5) << 15) | morton3(dirs, torch.full((3,), -1.0, device="cuda"), torch.full((3,), 1.0, device="cuda"), 5),
        }
        if label == "coherent":
            ys, xs = torch.div(torch.arange(n, device="cuda"), res, rounding_mode="floor"), torch.arange(n, device="cuda") % res
            keys["tiles_8x4"] = ((ys // 4) * (res // 8) + xs // 8) * 32 + (ys % 4) * 8 + xs % 8
        for kname, k in keys.items():
            perm = torch.argsort(k)
            measure(f"{label}/{kname}", r[perm].contiguous().view(torch.uint8).view(-1))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)
    dev.close()


if __name__ == "__main__":
    main()
