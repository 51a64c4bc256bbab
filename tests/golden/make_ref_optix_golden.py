"""Packs reference-run golden vectors: outputs of the REFERENCE's own CUDA/OptiX backend (LuisaCompute `cuda` backend,
unmodified sources, built by baseline/build_reference_cuda_backend.sh) on a B200, for a level-4 displaced icosphere.

How the inputs were produced (GPU box, this repository's tree):
    bash baseline/build_reference_cuda_backend.sh && make -C plugin           # in the container that mounts /root/reference
    LD_LIBRARY_PATH=baseline/_ref/bin python tools/ref_optix_compare.py --level 4 --res 160 --dir gpurun_out/golden_case
baseline/ref_optix_bench.cpp (the program that ran) traces rays_*.bin with a DSL kernel calling accel.intersect /
accel.intersect_any and writes ref_hits_*.bin / ref_any_*.bin.  This script keeps every second ray.
usage: python tests/golden/make_ref_optix_golden.py gpurun_out/golden_case tests/golden/ref_optix_ico4.npz
"""
import os
import sys

import numpy as np

src, dst = sys.argv[1], sys.argv[2]
out = {"verts": np.fromfile(os.path.join(src, "verts.bin"), dtype=np.float32).reshape(-1, 4),
       "tris": np.fromfile(os.path.join(src, "tris.bin"), dtype=np.uint32).reshape(-1, 3)}
for name in ("primary", "bounce"):
    out[f"rays_{name}"] = np.fromfile(os.path.join(src, f"rays_{name}.bin"), dtype=np.uint8).reshape(-1, 32)[::2].copy()
    out[f"hits_{name}"] = np.fromfile(os.path.join(src, f"ref_hits_{name}.bin"), dtype=np.uint8).reshape(-1, 24)[::2].copy()
    out[f"any_{name}"] = np.fromfile(os.path.join(src, f"ref_any_{name}.bin"), dtype=np.uint32)[::2].astype(np.uint8)
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst), "bytes;", {k: v.shape for k, v in out.items()})
