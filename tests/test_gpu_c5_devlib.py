"""BASELINE config C5 through the re-emitted device library (tools/c5_path_tracer.cu: the integrator of
src/tests/test_path_tracing.cpp:165-256 with RayQuery handlers as src/tests/test_procedural.cpp:101-125): the kernel's two
loop forms (one flat loop over (sample, bounce); the reference program's nested loops) consume the same random sequence and add
radiance in the same order, so they must render the same image bit for bit although their warps schedule the shared,
phase-voted RayQuery traversal (lc_ray_query_trace -> Traversal::run_voting) completely differently."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_c5_flat_and_nested_loops_render_the_same_image():
    import luisacompute_b200 as lc
    import bench_configs as B
    flat, nested = (os.path.join(ROOT, "tools", n) for n in ("libc5pt.so", "libc5pt_nested.so"))
    assert os.path.exists(flat) and os.path.exists(nested), "run __graft_entry__.build() first"
    dev = lc.Device(0)
    w, h, spp = 320, 240, 8
    img_f, rays_f = B.c5(dev, w, h, spp, so=flat, return_image=True)
    img_n, rays_n = B.c5(dev, w, h, spp, so=nested, return_image=True)
    assert rays_f == rays_n
    assert 2.0 < rays_f / (w * h * spp) < 4.0  # a closed room: ~2.8 rays per sample
    assert np.array_equal(img_f.view(np.uint32), img_n.view(np.uint32))
    rgb = img_f[..., :3] / img_f[..., 3:4]
    assert not np.isnan(rgb).any()
    assert np.all(img_f[..., 3] == spp)
    # a lit room with one red and one green side wall (cornell_box.h)
    assert 0.02 < rgb.mean() < 0.3
    left, right = rgb[:, : w // 8].reshape(-1, 3).mean(0), rgb[:, -w // 8:].reshape(-1, 3).mean(0)
    assert (left[0] > 2.0 * left[1] and right[1] > 2.0 * right[0]) or (left[1] > 2.0 * left[0] and right[0] > 2.0 * right[1])
