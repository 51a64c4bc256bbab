"""Developer tool: BLAS build time for icosphere level 8 (1.31 M) and a 5 M triangle soup, one line per library build."""
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


dev = lc.Device(0)
out = []
v, t = S.icosphere(8)
m = dev.create_mesh(v, t).build()
ms = timeit(lambda: m.build(lc.FORCE_BUILD))
i = m.info()
out.append(f"ico8 {len(t)} tris: {ms:.3f} ms = {len(t) / ms / 1e3:.0f} Mtris/s nodes {i.node_count} depth {i.depth} err {i.error}")
n = 5_000_000
g = torch.Generator(device="cuda").manual_seed(1)
c = torch.rand((n, 3), device="cuda", generator=g)
verts = torch.zeros((n, 3, 4), dtype=torch.float32, device="cuda")
verts[:, 0, :3] = c
verts[:, 1, :3] = c + (torch.rand((n, 3), device="cuda", generator=g) * 2 - 1) * 0.004
verts[:, 2, :3] = c + (torch.rand((n, 3), device="cuda", generator=g) * 2 - 1) * 0.004
tris = torch.arange(3 * n, dtype=torch.int32, device="cuda").view(-1, 3)
m2 = dev.create_mesh(verts.view(-1, 4), tris, vertex_stride=16).build()
ms = timeit(lambda: m2.build(lc.FORCE_BUILD), reps=3, warm=1)
i = m2.info()
out.append(f"soup {n} tris: {ms:.3f} ms = {n / ms / 1e3:.0f} Mtris/s nodes {i.node_count} depth {i.depth} err {i.error}")
print(os.environ.get("LCB200_RT_LIB", "default").split("/")[-1], "|", " | ".join(out), flush=True)
dev.close()
