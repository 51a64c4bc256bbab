"""Developer tool: time the C2 trace batches (coherent primary, incoherent bounce; closest and any) for one build of
the library.  Run once per variant: LCB200_RT_LIB=/path/to/variant.so python tools/trace_perf.py [--level 8] [--res 2048]
Prints one line; numbers are CUDA-event times (best of 5 after 3 warm-ups), never taken under a profiler.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--res", type=int, default=2048)
    ap.add_argument("--tag", default=os.environ.get("LCB200_RT_LIB", "default"))
    ap.add_argument("--cache", default="/tmp/trace_perf_cache.npz")
    a = ap.parse_args()
    dev = lc.Device(0)
    tdev = torch.device("cuda:0")
    verts, tris = S.icosphere(a.level)
    mesh = dev.create_mesh(verts, tris).build()
    acc = dev.create_accel()
    acc.emplace_back(mesh, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    build_ms = timeit(lambda: mesh.build(lc.FORCE_BUILD), reps=3, warm=1)
    n = a.res * a.res
    key = f"{a.level}_{a.res}"
    if os.path.exists(a.cache) and key in np.load(a.cache):
        z = np.load(a.cache)
        primary, bounce = z[key + "_p"].view(S.RAY_DTYPE), z[key].view(S.RAY_DTYPE)
        d_primary = torch.from_numpy(primary.view(np.uint8).reshape(-1)).to(tdev)
    else:
        primary = S.pinhole_rays(a.res, a.res)
        d_primary = torch.from_numpy(primary.view(np.uint8).reshape(-1)).to(tdev)
        hp = lc.hits_to_numpy(acc.trace_closest(d_primary))
        bounce = S.bounce_rays(primary, hp, verts, tris, a.res, a.res)
        np.savez(a.cache, **{key: bounce.view(np.uint8), key + "_p": primary.view(np.uint8)})
    d_bounce = torch.from_numpy(bounce.view(np.uint8).reshape(-1)).to(tdev)
    hits = torch.empty((n, 6), dtype=torch.int32, device=tdev)
    occ = torch.empty(n, dtype=torch.int32, device=tdev)
    ctr = torch.zeros(4, dtype=torch.int64, device=tdev)
    acc.trace_closest(d_bounce, out=hits, counters=ctr)
    torch.cuda.synchronize()
    c = ctr.cpu().numpy() / n
    t_p = timeit(lambda: acc.trace_closest(d_primary, out=hits))
    t_b = timeit(lambda: acc.trace_closest(d_bounce, out=hits))
    chk = int(hits[:, 0].to(torch.int64).sum().item())
    t_a = timeit(lambda: acc.trace_any(d_bounce, out=occ))
    info = mesh.info()
    print(f"{a.tag}: coherent {n / t_p / 1e3:8.1f}  incoherent {n / t_b / 1e3:8.1f}  any {n / t_a / 1e3:8.1f} Mrays/s | build {build_ms:.3f} ms "
          f"nodes {info.node_count} | per bounce ray: nodes {c[0]:.2f} prims {c[1]:.2f} inst {c[2]:.2f} exact {c[3]:.4f} | chk {chk}", flush=True)
    dev.close()


if __name__ == "__main__":
    main()
