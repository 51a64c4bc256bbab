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
    # test_procedural's sphere handler is numerically fragile by construction (d = sqrt(d_oc^2 - tc^2) goes NaN within rounding of a
    # sphere's centre, `d <= radius` flips within rounding of its silhouette): whether a mul and a sub of the USER'S handler are fused
    # into an FMA decides those pixels, and the driver's PTX compiler (native launch, `b200`) does not fuse exactly like OptiX's.  With
    # the kernels launched through OptiX (`b200o`, same traversal, same BVH) the images equal the reference's to 0-30 pixels (0 with the
    # traversal out of line as in round 1, 30 with it inlined into the single RayQuery site: inlining changes what the compiler fuses
    # in the handler too); natively 772 of 921 600 pixels differ (436 within 2 px of a projected sphere centre, the rest on
    # silhouettes; profiles/r2_native_launch.md).
    for test, tol_pixels, tol_optix in (("test_rtx", 64, 64), ("test_procedural", 1600, 100), ("test_procedural_callable", 1600, 100), ("test_indirect_rtx", 64, 64)):
        e = res[test]
        if "unavailable" in e:
            continue
        if e.get("cuda", {}).get("rc") != 0:
            pytest.skip(f"{test} did not run with the reference's cuda backend on this box: {e.get('cuda')}")
        assert e["b200"]["rc"] == 0, e["b200"]
        assert e["image"]["nonblack_pixels"] > 100_000
        assert e["image"]["pixels_differing_by_more_than_2"] <= tol_pixels and e["image"]["mean_abs_diff"] < 0.05, (test, e["image"])
        if "image_optix_launch" in e:
            assert e["image_optix_launch"]["pixels_differing_by_more_than_2"] <= tol_optix, (test, e["image_optix_launch"])
        if "procedural" in test:
            # RayQuery handlers lowered inline (`while (proceed) { handler }`, this repository's RayQueryLowering,
            # LUISA_B200_RAY_QUERY=inline) and the reference's outlined handlers run the same traversal and the same handler arithmetic...
            assert e["b200_inline"]["rc"] == 0, e["b200_inline"]
            # ...but not the same compiler schedule: the handler's `d <= radius` flips on FMA contraction (see above), so a few
            # silhouette pixels may differ between the two forms as they do between b200 and b200o
            assert e["image_inline_vs_outlined_lowering"]["pixels_differing"] <= 1600, e["image_inline_vs_outlined_lowering"]


def test_b200_backend_runs_without_optix(tmp_path):
    """kernels that trace rays are launched with cuLaunchKernel (plugin/hybrid/b200_shader_rt.cpp) and the
    backend's `optix::api()` is a stub (b200_optix_stub.cpp): the reference's unmodified test_rtx / test_procedural run on
    `b200` without libnvoptix ever being opened.  Evidence: the dynamic loader's own log (LD_DEBUG=files lists every object
    it maps, dlopen included) never names nvoptix or nvrtcore, and the backend binary does not contain the library's name."""
    so = os.path.join(BIN, "liblc-backend-b200.so")
    blob = open(so, "rb").read()
    assert b"libnvoptix" not in blob and b"optixQueryFunctionTable" not in blob
    for test in ("test_rtx", "test_procedural"):
        exe = os.path.join(BIN, test)
        if not os.path.exists(exe):
            pytest.skip("reference test programs not built")
        d = tmp_path / test
        d.mkdir()
        log = str(d / "ld_debug")
        r = subprocess.run([exe, "b200"], cwd=str(d), capture_output=True, text=True, timeout=900,
                           env=_env(LD_DEBUG="files", LD_DEBUG_OUTPUT=log))
        assert r.returncode == 0, (r.stdout + r.stderr)[-1500:]
        assert os.path.exists(str(d / (test + ".png")))
        loaded = ""
        for f in os.listdir(str(d)):
            if f.startswith("ld_debug"):
                loaded += open(str(d / f), errors="replace").read()
        assert "liblc-backend-b200" in loaded and "liblcb200_rt" in loaded      # the log does see what is loaded...
        assert "nvoptix" not in loaded and "nvrtcore" not in loaded             # ...and OptiX is not


def test_d6_shadow_rays_ending_on_the_light(tmp_path):
    """Divergence D6: a ray whose t_max is exactly the distance of a surface.  This repository accepts t <= t_max
    like the oracle and Embree do (hit iff tnear <= t <= tfar; accel.rs:449-535 hands the interval to rtcIntersect1 / rtcOccluded1
    unchanged), decided exactly; OptiX decides the same comparison in fp32 and mostly misses.  It shows in test_path_tracing-style
    next-event estimation, where the shadow ray ends ON the light quad: C5 renders 2-3 % darker here than on the reference's CUDA
    backend.  north_star names the Embree backend as the parity reference, so the inclusive rule stays; this test bounds the
    visible difference and pins its mechanism: with the shadow rays shortened by 0.1 % both backends render the same image."""
    exe = os.path.join(BIN, "c5_dsl_path_tracer")
    if not (os.path.exists(exe) and os.path.exists(os.path.join(BIN, "liblc-backend-cuda.so"))):
        pytest.skip("C5 DSL program / reference CUDA backend not built")
    res = {}
    for name, scale in (("as_written", None), ("shortened", "0.999")):
        out = str(tmp_path / f"c5_{name}.json")
        env = _env(**({"C5_SHADOW_TMAX_SCALE": scale} if scale else {}))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_optix_c5.py"), "--width", "640", "--height", "360", "--spp", "16",
                            "--dir", str(tmp_path / name), "--out", out], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res[name] = json.load(open(out))
        if "unavailable" in res[name].get("cuda", {}):
            pytest.skip("the reference's cuda backend did not start on this box: " + res[name]["cuda"]["unavailable"][-200:])
        assert "unavailable" not in res[name]["b200"], res[name]["b200"]
    a, s = res["as_written"]["image_agreement"], res["shortened"]["image_agreement"]
    rel = (a["mean_radiance_cuda"] - a["mean_radiance_b200"]) / a["mean_radiance_cuda"]
    print(f"D6: mean radiance cuda {a['mean_radiance_cuda']:.5f} b200 {a['mean_radiance_b200']:.5f} ({100 * rel:.2f} % darker); "
          f"with shadow rays x 0.999: {s['mean_radiance_cuda']:.5f} vs {s['mean_radiance_b200']:.5f}")
    assert 0.0 <= rel < 0.05, a                                    # darker, by a few per cent, never brighter
    assert abs(s["mean_radiance_cuda"] - s["mean_radiance_b200"]) < 2e-3 * s["mean_radiance_cuda"], s
