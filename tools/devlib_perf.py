"""Developer tool (GPU box): the inline entry points lc_accel_trace_closest / _any (what generated kernels call, one thread per
ray) on the C2 workload next to the batch kernel; developer variants of the test kernels in gpurun_variants/ are timed too.
usage: python tools/devlib_perf.py [--level 8] [--res 2048] [--block 128]"""
import argparse
import ctypes as C
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


def load(name):
    L = C.CDLL(name if os.path.isabs(name) else os.path.join(ROOT, "tests", "devlib", name))
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.dl_closest.argtypes = [vp, u64, vp, vp, u32, u32, vp, u32]
    L.dl_any.argtypes = [vp, u64, vp, vp, u32, u32, vp, u32]
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--res", type=int, default=2048)
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = lc.Device(0)
    verts, tris = S.icosphere(a.level)
    mesh = dev.create_mesh(verts, tris).build()
    acc = dev.create_accel()
    acc.emplace_back(mesh, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    primary = S.pinhole_rays(a.res, a.res)
    n = len(primary)
    d_p = torch.from_numpy(primary.view(np.uint8).reshape(-1)).cuda()
    ref_p = acc.trace_closest(d_p)
    bounce = S.bounce_rays(primary, lc.hits_to_numpy(ref_p), verts, tris, a.res, a.res)
    d_b = torch.from_numpy(bounce.view(np.uint8).reshape(-1)).cuda()
    ref_b = acc.trace_closest(d_b)
    ref_any = acc.trace_any(d_b)
    bh, bi = acc.binding()
    st = torch.cuda.current_stream().cuda_stream
    res = {"rays": n, "block": a.block,
           "batch_kernel_mrays_s": {"coherent": n / timeit(lambda: acc.trace_closest(d_p, out=ref_p)) / 1e3,
                                    "incoherent": n / timeit(lambda: acc.trace_closest(d_b, out=ref_b)) / 1e3,
                                    "incoherent_any": n / timeit(lambda: acc.trace_any(d_b, out=ref_any)) / 1e3}}
    hits = torch.empty((n, 6), dtype=torch.int32, device="cuda")
    occ = torch.empty(n, dtype=torch.int32, device="cuda")
    libs = [("inline_entry_points", "libdevlibtest.so")]
    vdir = os.path.join(ROOT, "gpurun_variants")   # developer variants (tools/_build_devlib_variants.sh)
    if os.path.isdir(vdir):
        libs += [(f[len("libdevlibtest_"):-3], os.path.join(vdir, f)) for f in sorted(os.listdir(vdir)) if f.startswith("libdevlibtest_")]
    for name, so in libs:
        L = load(so)
        r = {}
        for label, rays, ref in (("coherent", d_p, ref_p), ("incoherent", d_b, ref_b)):
            hits.zero_()
            assert L.dl_closest(st, bh, bi, rays.data_ptr(), n, 0xFF, hits.data_ptr(), a.block) == 0
            torch.cuda.synchronize()
            r[label + "_identical_to_batch"] = bool(torch.equal(hits, ref))
            r[label] = n / timeit(lambda: L.dl_closest(st, bh, bi, rays.data_ptr(), n, 0xFF, hits.data_ptr(), a.block)) / 1e3
        assert L.dl_any(st, bh, bi, d_b.data_ptr(), n, 0xFF, occ.data_ptr(), a.block) == 0
        torch.cuda.synchronize()
        r["any_identical_to_batch"] = bool(torch.equal(occ != 0, ref_any != 0))
        r["incoherent_any"] = n / timeit(lambda: L.dl_any(st, bh, bi, d_b.data_ptr(), n, 0xFF, occ.data_ptr(), a.block)) / 1e3
        res[name] = r
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    dev.close()


if __name__ == "__main__":
    main()
