"""Developer tool: BLAS refit (PREFER_UPDATE on an allow_update mesh) vs rebuild for icosphere level 8, and the traversal cost after the refit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


if __name__ == "__main__":
    dev = lc.Device(0)
    v, t = S.icosphere(8)
    dv = torch.from_numpy(np.ascontiguousarray(v)).cuda()
    dt = torch.from_numpy(np.ascontiguousarray(t)).cuda()
    m = dev.create_mesh(dv, dt, lc.AccelOption(allow_update=True)).build()
    acc = dev.create_accel(lc.AccelOption(allow_update=True))
    acc.emplace_back(m, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    rays = torch.from_numpy(S.pinhole_rays(1024, 1024).view(np.uint8).reshape(-1)).cuda()
    hits = torch.empty((1024 * 1024, 6), dtype=torch.int32, device="cuda")
    t_trace0 = timeit(lambda: acc.trace_closest(rays, out=hits))
    ms_build = timeit(lambda: m.build(lc.FORCE_BUILD))
    dv[:, :3] *= 1.01                                   # every vertex moves
    ms_refit = timeit(lambda: m.build(lc.PREFER_UPDATE))
    kind = m.last_build_kind
    acc.build()
    t_trace1 = timeit(lambda: acc.trace_closest(rays, out=hits))
    print(f"ico8 {len(t)} tris: rebuild {ms_build:.3f} ms ({len(t) / ms_build / 1e3:.0f} Mtris/s), refit {ms_refit:.3f} ms ({len(t) / ms_refit / 1e3:.0f} Mtris/s, kind {kind}); "
          f"1 Mi primary rays before {t_trace0:.3f} ms, after refit {t_trace1:.3f} ms", flush=True)
    dev.close()
