"""The drop-in itself: LuisaCompute's CUDA backend with this repository's acceleration-structure classes and
device-library RT section linked in (`liblc-backend-b200.so`, plugin/hybrid/), driven by programs written against — or taken
unmodified from — the reference.

* a DSL program (baseline/ref_optix_bench.cpp: kernels calling accel.intersect / intersect_any, compiled by the reference's
  code generator and NVRTC against the spliced device library) must return, through `context.create_device("b200")`, exactly
  what this repository's batch C-ABI returns for the same scene and rays;
* the reference's own headless ray-tracing tests (src/tests/test_rtx.cpp, test_procedural.cpp, test_procedural_callable.cpp,
  test_indirect_rtx.cpp), compiled unmodified, must render with `b200` the image
  they render with the reference's `cuda` (OptiX) backend.
The binaries are built where the reference tree is mounted (__graft_entry__.build()) and travel under baseline/_ref/bin.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baseline", "_ref", "bin")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "liblc-backend-b200.so")), reason="hybrid backend not built (plugin/hybrid/build.sh)")]


def _env(**kw):
    return dict(os.environ, LD_LIBRARY_PATH=BIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""), **kw)


def test_dsl_kernels_through_b200_backend_match_the_batch_path(tmp_path):
    out = str(tmp_path / "hybrid.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_optix_compare.py"), "--level", "6", "--res", "384", "--dir", str(tmp_path / "case"),
                        "--out", out], capture_output=True, text=True, timeout=900, env=_env(LC_BACKEND="b200"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    if "unavailable" in res["reference"]:
        pytest.skip("OptiX launch path did not start on this box: " + res["reference"]["unavailable"][-300:])
    assert res["reference"]["backend"] == "b200"
    for name in ("primary", "bounce"):
        p = res["parity_vs_reference"][name]
        assert p["hit_miss_disagree"] == 0 and p["diff_id"] == 0 and p["any_disagree"] == 0, (name, p)
        assert p["max_rel_t_same_id"] == 0.0 and p["max_bary_diff_same_id"] == 0.0, (name, p)   # same code, same BVH: bit-identical
        assert p["same_id"] > 10_000


def test_reference_test_programs_render_the_same_image(tmp_path):
    if not os.path.exists(os.path.join(BIN, "test_rtx")):
        pytest.skip("reference test programs not built")
    out = str(tmp_path / "tests.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_tests.py"), "--out", out], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    for test, tol_pixels in (("test_rtx", 64), ("test_procedural", 400), ("test_procedural_callable", 400), ("test_indirect_rtx", 64)):
        e = res[test]
        if "unavailable" in e:
            continue
        if e.get("cuda", {}).get("rc") != 0:
            pytest.skip(f"{test} did not run with the reference's cuda backend on this box: {e.get('cuda')}")
        assert e["b200"]["rc"] == 0, e["b200"]
        assert e["image"]["nonblack_pixels"] > 100_000
        assert e["image"]["pixels_differing_by_more_than_2"] <= tol_pixels and e["image"]["mean_abs_diff"] < 0.05, (test, e["image"])
