"""Developer tool (GPU box): the reference's OWN test programs, compiled unmodified from /root/reference/src/tests
(test_rtx.cpp: animated triangle, in-kernel set_instance_transform, BLAS refit + TLAS update per frame, 1024 frames;
test_procedural.cpp: AABB spheres + a mesh through RayQuery handlers), run once with the reference's `cuda` backend (OptiX)
and once with `b200` (the same backend with this repository's BVH builder, host layer, device-library traversal and native
kernel launch linked in, plugin/hybrid/) — plus `b200o` when built (the same, but kernels still launched through OptiX: the
developer variant of plugin/hybrid/build.sh) — and their output images compared.  usage: python tools/run_reference_tests.py [--out gpurun_out/reference_tests.json]"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baseline", "_ref", "bin")


def read_png(path):
    import zlib
    import struct
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, w, h, data = 8, 0, 0, b""
    while pos < len(raw):
        n, tag = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert depth == 8 and ctype == 6
        elif tag == b"IDAT":
            data += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(data), dtype=np.uint8).reshape(h, 1 + 4 * w).astype(np.int32)
    img = np.zeros((h, w * 4), dtype=np.int32)
    prev = np.zeros(w * 4, dtype=np.int32)
    for y in range(h):                      # undo the PNG row filters (stb writes 0..4)
        f, line = rows[y, 0], rows[y, 1:].copy()
        if f == 1:
            for x in range(4, w * 4):
                line[x] = (line[x] + line[x - 4]) & 255
        elif f == 2:
            line = (line + prev) & 255
        elif f == 3:
            for x in range(w * 4):
                left = line[x - 4] if x >= 4 else 0
                line[x] = (line[x] + ((left + prev[x]) >> 1)) & 255
        elif f == 4:
            for x in range(w * 4):
                a = line[x - 4] if x >= 4 else 0
                b, c = prev[x], (prev[x - 4] if x >= 4 else 0)
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                line[x] = (line[x] + pred) & 255
        img[y], prev = line, line
    return img.reshape(h, w, 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_tests.json"))
    a = ap.parse_args()
    env = dict(os.environ, LD_LIBRARY_PATH=BIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    res = {}
    for test in ("test_rtx", "test_procedural", "test_procedural_callable", "test_indirect_rtx"):
        exe = os.path.join(BIN, test)
        if not os.path.exists(exe):
            res[test] = {"unavailable": "not built"}
            continue
        imgs, entry = {}, {}
        # `b200_inline` is b200 told to lower RayQuery handlers inline (LUISA_B200_RAY_QUERY=inline, plugin/hybrid/b200_ray_query_lowering.inl)
        # instead of outlining them as the reference does: only programs with a RayQuery differ
        backends = ("cuda", "b200") + (("b200_inline",) if "procedural" in test else ()) + \
                   (("b200o",) if os.path.exists(os.path.join(BIN, "liblc-backend-b200o.so")) else ())
        for backend in backends:
            d = os.path.join("/tmp", f"lc_{test}_{backend}")
            os.makedirs(d, exist_ok=True)
            t0 = time.time()
            benv = dict(env, LUISA_B200_RAY_QUERY="inline") if backend == "b200_inline" else env
            r = subprocess.run([exe, backend.replace("_inline", "")], cwd=d, capture_output=True, text=True, timeout=900, env=benv)
            entry[backend] = {"rc": r.returncode, "wall_s": time.time() - t0}
            for line in (r.stdout + r.stderr).splitlines():
                if "Time:" in line or "time" in line.lower() and "ms" in line:
                    entry[backend]["log"] = line[-80:]
            png = os.path.join(d, test + ".png")
            if r.returncode == 0 and os.path.exists(png):
                imgs[backend] = read_png(png)
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                subprocess.run(["cp", png, os.path.join(ROOT, "gpurun_out", f"{test}_{backend}.png")])
            else:
                entry[backend]["log_tail"] = (r.stdout + r.stderr)[-600:]
        if "b200" in imgs and "b200_inline" in imgs:
            diff = np.abs(imgs["b200"] - imgs["b200_inline"])
            entry["image_inline_vs_outlined_lowering"] = {"max_abs_diff": int(diff.max()), "pixels_differing": int((diff.max(axis=2) > 0).sum())}
        for other, key in (("b200", "image"), ("b200o", "image_optix_launch")):
            if "cuda" in imgs and other in imgs:
                diff = np.abs(imgs["cuda"] - imgs[other])
                entry[key] = {"shape": list(imgs["cuda"].shape), "max_abs_diff": int(diff.max()), "mean_abs_diff": float(diff.mean()),
                              "pixels_differing_by_more_than_2": int((diff.max(axis=2) > 2).sum()),
                              "nonblack_pixels": int((imgs["cuda"][:, :, :3].sum(axis=2) > 0).sum())}
        res[test] = entry
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
