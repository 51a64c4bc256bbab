#!/usr/bin/env python
"""bench.py — the BASELINE.json headline benchmark of the ray-tracing acceleration path on B200.

Workload (BASELINE.json configs[1], "C2"): displaced icosphere level 8 (1 310 720 triangles, one instance), one step =
4 194 304 coherent primary closest-hit rays + 4 194 304 incoherent cosine-bounce closest-hit rays per GPU.
  value      Mrays/s, whole job (all ranks), rays and BVH resident in HBM when the timed region starts
  e2e        the same step through the public host-buffer call (Accel.trace_closest_host -> lcbTraceClosestHost): rays in
             pinned host memory, hits back in pinned host memory, copies inside the timed region
  roofline   dominant kernel = k_trace<closest> on the incoherent batch; algorithmic bytes per ray =
             32 (ray) + 24 (hit) + nodes*80 + prims*48 + instances*80 (the 80 bytes of a 256-byte instance record the traversal loads) with the per-ray averages COUNTED by the
             instrumented variant of the same kernel; peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle (CPU restatement of the reference's Embree-backed path) on all host cores, bounded sample
`--impl reference` times that CPU path alone on the same workload definition.

Multi-GPU (torchrun, one rank per GPU): replicated scene — rank 0 builds the BLAS, NCCL-broadcasts the position-
independent blob, every rank builds its own TLAS — and rays are sharded by 65 536-ray tiles round-robin.  The headline stays
the C2 step with the per-GPU work fixed ("weak": every rank traces 8 388 608 rays per step), and for N > 1 the timed region
INCLUDES getting every rank's 24-byte hit records into ray order on rank 0: the trace's refine pass stores them straight into
rank 0's framebuffer, mapped into every rank over NVLink (luisacompute_b200.dist.PeerFramebuffer, lcrtTraceClosestTiled); where
CUDA IPC is unavailable, or with LCB_GATHER=nccl, NCCL send/recv chunk by chunk on a side stream (dist.HitGatherer).  `detail.c4_strong` is BASELINE config 4
at every N (1 included): 50 M-triangle soup built on rank 0 and broadcast, the FIXED batch of 64 Mi rays tile-sharded over
the ranks (strong scaling) with the same in-region gather.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 65536


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--level", type=int, default=8, help="icosphere subdivision (8 = the BASELINE workload)")
    ap.add_argument("--res", type=int, default=2048, help="primary rays = res x res per GPU")
    ap.add_argument("--cpu-sample", type=int, default=1048576, help="rays per batch for the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the run of the reference's own CUDA/OptiX backend")
    ap.add_argument("--no-c4", action="store_true", help="skip detail.c4_strong (BASELINE config 4: 50 M triangles, 64 Mi rays sharded over the ranks)")
    ap.add_argument("--c4-tris", type=int, default=50_000_000)
    ap.add_argument("--c4-rays", type=int, default=64 << 20)
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])), smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_scene(level: int):
    from luisacompute_b200 import scenes as S
    t0 = time.time()
    verts, tris = S.icosphere(level)
    return verts, tris, time.time() - t0


def primary_for_rank(res: int, rank: int, world: int):
    """Global batch = `world` sub-pixel-shifted copies of the res x res camera; 65 536-ray tiles round-robin over ranks."""
    from luisacompute_b200 import scenes as S
    rays_per_sample = res * res
    tiles_per_sample = max(1, rays_per_sample // TILE)
    out = np.zeros(rays_per_sample, dtype=S.RAY_DTYPE)
    base = [S.pinhole_rays(res, res, pos=(0.0 + 1e-3 * s, 0.0, 3.0)) for s in range(world)] if world > 1 else [S.pinhole_rays(res, res)]
    if world == 1:
        return base[0]
    k = 0
    for j in range(tiles_per_sample * world):
        if j % world != rank:
            continue
        s, t = divmod(j, tiles_per_sample)
        out[k * TILE:(k + 1) * TILE] = base[s][t * TILE:(t + 1) * TILE]
        k += 1
    return out


def cpu_baseline(verts, tris, primary, bounce, sample: int, threads: int = 0) -> dict:
    """The oracle timed on the host cores over a bounded sample of the same two batches."""
    from luisacompute_b200 import scenes as S
    from oracle import oracle_py as O
    t0 = time.time()
    om = O.OracleMesh(verts, tris)
    oa = O.OracleAccel()
    assert oa.update(1, S.make_mods([om.handle], [S.affine_identity()])) == 0
    build_s = time.time() - t0
    n = min(sample, len(primary))
    # strided sample so that both batches keep their screen-space distribution
    stride = max(1, len(primary) // n)
    p, b = np.ascontiguousarray(primary[::stride][:n]), np.ascontiguousarray(bounce[::stride][:n])
    cores = O.lib().orc_hardware_threads() if threads == 0 else threads
    oa.trace_closest(p[:4096], 0xFF, threads)                      # warm the caches / thread pool
    t0 = time.time()
    oa.trace_closest(p, 0xFF, threads)
    tp = time.time() - t0
    t0 = time.time()
    oa.trace_closest(b, 0xFF, threads)
    tb = time.time() - t0
    return {"value": (len(p) + len(b)) / (tp + tb) / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{len(p)} primary + {len(b)} bounce closest-hit rays (every {stride}th ray of the GPU batches), same scene",
            "coherent_mrays_s": len(p) / tp / 1e6, "incoherent_mrays_s": len(b) / tb / 1e6, "bvh_build_s": build_s,
            "note": "scalar double-precision CPU restatement of the Embree-backed reference path; real Embree (AVX-512 BVH8) would be several x faster"}


def reference_gpu(verts, tris, primary, bounce, backend: str = "cuda") -> dict:
    """The REFERENCE's own GPU path on the same scene and rays, in the same run: LuisaCompute's `cuda` backend (OptiX,
    which is driver software traversal on B200) through baseline/_ref/bin/ref_optix_bench (baseline/ref_optix_bench.cpp: a
    DSL kernel calling accel.intersect over the ray buffer; unmodified reference sources, see
    baseline/build_reference_cuda_backend.sh).  Separate process, nothing of this repository loaded into it.  Returns
    {"unavailable": why} when the binaries did not travel with the tree or the backend does not start on this box."""
    import shutil
    import subprocess
    import tempfile
    bindir = os.path.join(ROOT, "baseline", "_ref", "bin")
    exe = os.path.join(bindir, "ref_optix_bench")
    if not (os.path.exists(exe) and os.path.exists(os.path.join(bindir, f"liblc-backend-{backend}.so"))):
        return {"unavailable": f"baseline/_ref/bin/ref_optix_bench or liblc-backend-{backend}.so not present"}
    d = tempfile.mkdtemp(prefix="lc_ref_")
    try:
        v4 = np.zeros((len(verts), 4), dtype=np.float32)
        v4[:, :3] = np.asarray(verts, dtype=np.float32)[:, :3]
        v4.tofile(os.path.join(d, "verts.bin"))
        np.ascontiguousarray(tris, dtype=np.uint32).tofile(os.path.join(d, "tris.bin"))
        primary.tofile(os.path.join(d, "rays_primary.bin"))
        bounce.tofile(os.path.join(d, "rays_bounce.bin"))
        env = dict(os.environ, LD_LIBRARY_PATH=bindir + ":" + os.environ.get("LD_LIBRARY_PATH", ""), LC_BACKEND=backend)
        r = subprocess.run([exe, d], capture_output=True, text=True, timeout=600, env=env)
        path = os.path.join(d, "ref_result.json")
        if r.returncode != 0 or not os.path.exists(path):
            return {"unavailable": f"reference program failed (rc {r.returncode}): " + (r.stdout + r.stderr)[-300:].replace("\n", " ")}
        res = json.load(open(path))
        n = res["primary"]["rays"] + res["bounce"]["rays"]
        ms = res["primary"]["closest_ms"] + res["bounce"]["closest_ms"]
        what = ("LuisaCompute cuda backend (OptiX) on this GPU" if backend == "cuda" else
                "lc-backend-b200 (the reference's CUDA backend with this repository's accel classes and device-library RT section, plugin/hybrid/)")
        return {"what": what + ", same scene and rays, DSL kernel accel.intersect per ray through the reference's runtime API; best of 10, host clock "
                        "around dispatch + synchronize", "value": n / ms / 1e3, "unit": "Mrays/s",
                "coherent_closest_mrays_s": res["primary"]["closest_mrays_s"], "incoherent_closest_mrays_s": res["bounce"]["closest_mrays_s"],
                "incoherent_any_mrays_s": res["bounce"]["any_mrays_s"], "mesh_build_ms": res["mesh_build_ms"],
                "mesh_build_mtris_s": res["triangles"] / res["mesh_build_ms"] / 1e3}
    except Exception as e:  # noqa: BLE001 - a measurement aid must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        shutil.rmtree(d, ignore_errors=True)



CHUNK_TILES = 8     # tiles per trace launch / gather in the multi-GPU step (512 Ki rays, 12.6 MB of hit records)


def make_peer_framebuffer(dev, n_local_tiles: int):
    """The gathering rank's framebuffer mapped into every rank (luisacompute_b200.dist.PeerFramebuffer), or None where CUDA IPC /
    peer access is not available (or LCB_GATHER=nccl): the callers then gather with NCCL send / recv (HitGatherer).  Collective."""
    import torch
    import torch.distributed as dist
    from luisacompute_b200.dist import PeerFramebuffer
    ok, fb = 1, None
    if gather_mode() == "nccl":
        ok = 0
    else:
        try:
            fb = PeerFramebuffer(dev, n_local_tiles)
        except Exception as e:   # noqa: BLE001 - reported, and every rank falls back together
            print(f"[bench] rank {dist.get_rank()}: peer framebuffer unavailable ({e}); NCCL gather instead", file=sys.stderr)
            ok = 0
    t = torch.tensor([ok], dtype=torch.int32, device=dev.torch_device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 0:
        if fb is not None:
            fb.address = 0   # not collective any more: leave the mapping to process exit
        return None
    return fb


def gather_mode() -> str:
    """stores (default): the trace kernel itself stores its finished records into rank 0's mapped framebuffer; dma: traced chunk
    by chunk, each chunk's records copied there by the copy engines while the next chunk is traced (the chunked launches cost
    more than the copies hide: profiles/r2_scaling.md); nccl: send / recv + reorder copy."""
    m = os.environ.get("LCB_GATHER", "stores")
    return m if m in ("dma", "stores", "nccl") else "stores"


DMA_CHUNK_TILES = int(os.environ.get("LCB_GATHER_CHUNK_TILES", "16"))


def gather_description(fb, gat) -> str:
    if fb is not None and gather_mode() == "dma":
        return (f"peer framebuffer, copy engines: rank 0's ray-ordered buffer is mapped into every rank (CUDA IPC); each {DMA_CHUNK_TILES}-tile chunk's "
                "24-byte records go there with one strided device-to-device copy over NVLink on a side stream while the next chunk is traced "
                "(lcbFramebufferCopyTiles); one completion all-reduce per step, no gather collective, no reorder copy")
    if fb is not None:
        how = ("one by one as the rays end" if fb.world <= 2 else "the records of each 128-ray batch together when the batch is complete (512-byte warp stores)")
        return ("peer framebuffer, stores: every rank's trace kernel stores its finished 24-byte records straight into rank 0's ray-ordered buffer over "
                f"NVLink, {how} (CUDA IPC mapping, lcrtTraceClosestTiled); one completion all-reduce per step, no gather collective, no reorder copy")
    if gat is not None:
        return "NCCL send / recv of 24-byte records, chunk by chunk on a side stream, one strided reorder copy per chunk (HitGatherer)"
    return "single GPU"


def c4_strong(args, lc, dev, world: int, rank: int, barrier) -> dict:
    """BASELINE config 4, strong scaling: 50 M random triangles (built on rank 0, blob broadcast over NCCL), the fixed batch of
    64 Mi uniform rays in 65 536-ray tiles dealt round-robin, hit records gathered into ray order on rank 0 inside the timed
    region.  Geometry and rays are drawn on the device with torch (same distributions as scenes.triangle_soup / uniform_rays,
    whose PCG32 streams the parity tests use at smaller sizes)."""
    import torch
    import torch.distributed as dist
    from luisacompute_b200 import scenes as S
    from luisacompute_b200.dist import HitGatherer
    tdev = dev.torch_device
    n_tris, n_rays = args.c4_tris, (args.c4_rays // (TILE * world)) * TILE * world
    out = {"workload": f"C4 {n_tris} random triangles (s=0.004), {n_rays} uniform incoherent closest-hit rays in {TILE}-ray tiles round-robin over "
                       f"{world} GPU(s), hits gathered to rank 0 in ray order inside the timed region", "scaling": "strong"}
    mesh = dev.create_mesh()
    if rank == 0:
        g = torch.Generator(device=tdev).manual_seed(1234)
        c = torch.rand((n_tris, 3), device=tdev, generator=g)
        verts = torch.zeros((n_tris, 3, 4), dtype=torch.float32, device=tdev)
        verts[:, 0, :3] = c
        verts[:, 1, :3] = c + (torch.rand((n_tris, 3), device=tdev, generator=g) * 2 - 1) * 0.004
        verts[:, 2, :3] = c + (torch.rand((n_tris, 3), device=tdev, generator=g) * 2 - 1) * 0.004
        del c
        tris = torch.arange(3 * n_tris, dtype=torch.int32, device=tdev).view(-1, 3)
        mesh.set_buffers(verts.view(-1, 4), tris, 16)
        mesh.build()
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mesh.build(lc.FORCE_BUILD)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        info = mesh.info()
        assert info.built == 1 and info.error == 0, (info.built, info.error)
        out.update(build_ms=min(times), build_mtris_s=n_tris / min(times) / 1e3, bvh_nodes=int(info.node_count), bvh_bytes=int(info.total_bytes))
    if world > 1:
        blob = mesh.export_blob() if rank == 0 else None
        nbytes = torch.tensor([blob.numel() if rank == 0 else 0], dtype=torch.int64, device=tdev)
        dist.broadcast(nbytes, 0)
        if rank != 0:
            blob = torch.empty(int(nbytes.item()), dtype=torch.uint8, device=tdev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(blob, 0)
        e1.record()
        torch.cuda.synchronize()
        out["bvh_broadcast_ms"] = e0.elapsed_time(e1)
        if rank != 0:
            mesh.import_blob(blob)
        torch.cuda.synchronize()
        del blob
    acc = dev.create_accel()
    acc.emplace_back(mesh)
    acc.build()
    # the same 64 Mi rays on every rank (same seed, same generator), of which it keeps its own tiles
    g2 = torch.Generator(device=tdev).manual_seed(5678)
    rays = torch.zeros((n_rays, 8), dtype=torch.float32, device=tdev)
    rays[:, 0:3] = torch.rand((n_rays, 3), device=tdev, generator=g2)
    z = 1 - 2 * torch.rand(n_rays, device=tdev, generator=g2)
    ph = 2 * np.pi * torch.rand(n_rays, device=tdev, generator=g2)
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    rays[:, 4], rays[:, 5], rays[:, 6] = r * torch.cos(ph), r * torch.sin(ph), z
    rays[:, 7] = 3.4028234663852886e38
    del z, ph, r
    n_tiles_local = n_rays // (TILE * world)
    local = rays.view(n_tiles_local, world, TILE * 8)[:, rank].contiguous().view(-1) if world > 1 else rays.view(-1)
    del rays
    n_local = n_tiles_local * TILE
    hits = torch.empty((n_local, 6), dtype=torch.int32, device=tdev)
    gat, fb = None, None
    if world > 1:
        fb = make_peer_framebuffer(dev, n_tiles_local)
        if fb is None:
            gat = HitGatherer(n_tiles_local, tdev, chunk_tiles=2 * CHUNK_TILES)
    out["gather"] = gather_description(fb, gat)
    lb = local.view(torch.uint8)
    ct = 2 * CHUNK_TILES

    def step():
        if fb is not None and gather_mode() == "stores":
            # one launch; the kernel stores every finished record into rank 0's ray-ordered framebuffer over NVLink
            fb.trace_closest(acc, lb, scratch=hits)
            fb.complete()
            return
        if fb is not None:
            for t0 in range(0, n_tiles_local, DMA_CHUNK_TILES):
                t1 = min(t0 + DMA_CHUNK_TILES, n_tiles_local)
                acc.trace_closest(lb[t0 * TILE * 32:t1 * TILE * 32], out=hits[t0 * TILE:t1 * TILE])
                ev = torch.cuda.Event()
                ev.record()
                fb.submit(hits[t0 * TILE:t1 * TILE], t0, ev)
            fb.complete()
            return
        if gat is None:
            acc.trace_closest(lb, out=hits)     # one GPU: one launch, records in place
            return
        for t0 in range(0, n_tiles_local, ct):
            t1 = min(t0 + ct, n_tiles_local)
            acc.trace_closest(lb[t0 * TILE * 32:t1 * TILE * 32], out=hits[t0 * TILE:t1 * TILE])
            ev = torch.cuda.Event()
            ev.record()
            gat.submit(hits[t0 * TILE:t1 * TILE], t0, t1, ev)
        gat.finish()
    step()
    barrier()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    # the trace alone (records stay on this rank)
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    tct = n_tiles_local if (gat is None and (fb is None or gather_mode() == "stores")) else (ct if gat is not None else DMA_CHUNK_TILES)
    for t0 in range(0, n_tiles_local, tct):
        t1 = min(t0 + tct, n_tiles_local)
        acc.trace_closest(lb[t0 * TILE * 32:t1 * TILE * 32], out=hits[t0 * TILE:t1 * TILE])
    t1e.record()
    torch.cuda.synchronize()
    trace_ms = t0e.elapsed_time(t1e)
    if world > 1:
        t = torch.tensor([ms, trace_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, trace_ms = (float(x) for x in t.cpu())
        # verification: every rank's records are where ray order puts them
        chk = torch.stack([hits[:, 0].long().sum(), hits[:, 1].long().sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)] if rank == 0 else None
        dist.gather(chk, allc, 0)
        if rank == 0:
            fbv = (fb if fb is not None else gat).full.view(n_tiles_local, world, TILE, 6)
            for rr in range(world):
                assert int(fbv[:, rr, :, 0].long().sum()) == int(allc[rr][0]) and int(fbv[:, rr, :, 1].long().sum()) == int(allc[rr][1]), f"gather mismatch, rank {rr}"
            assert torch.equal(fbv[:, 0].reshape(-1, 6), hits)
    out.update(ms_per_step=ms, mrays_s=n_rays / ms / 1e3, trace_only_ms=trace_ms, trace_only_mrays_s=n_rays / trace_ms / 1e3,
               gather_bytes_per_step=(world - 1) * n_local * 24 if world > 1 else 0, hit_fraction=float((hits[:, 0] != -1).float().mean().item()))
    if fb is not None:
        fb.close()
    del gat, fb, hits, local, lb, acc, mesh
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; Embree itself is not buildable here) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from luisacompute_b200 import scenes as S
    from oracle import oracle_py as O
    verts, tris, _ = make_scene(args.level)
    om = O.OracleMesh(verts, tris)
    oa = O.OracleAccel()
    assert oa.update(1, S.make_mods([om.handle], [S.affine_identity()])) == 0
    res = args.res
    primary = S.pinhole_rays(res, res)
    n = min(args.cpu_sample, len(primary))
    stride = max(1, len(primary) // n)
    # the bounce batch needs primary hits: the oracle supplies them for the sampled rays only
    ps = np.ascontiguousarray(primary[::stride][:n])
    hits = oa.trace_closest(ps)[0]
    ys, xs = np.divmod(np.arange(len(primary))[::stride][:n], res)
    # same generator as the GPU arm, evaluated on the sampled pixels
    full_hits = np.zeros(len(primary), dtype=S.HIT_DTYPE)
    full_hits["inst"] = 0xFFFFFFFF
    full_hits[::stride][:n] = hits
    bounce = S.bounce_rays(primary, full_hits, verts, tris, res, res)
    bs = np.ascontiguousarray(bounce[::stride][:n])
    cores = O.lib().orc_hardware_threads()
    for _ in range(args.warmup):
        oa.trace_closest(ps[:8192])
    t0 = time.time()
    for _ in range(args.steps):
        oa.trace_closest(ps)
        oa.trace_closest(bs)
    dt = time.time() - t0
    value = args.steps * 2 * n / dt / 1e6
    sample = f"each step = {n} primary + {n} bounce closest-hit rays (every {stride}th ray of the {res}x{res} batches), {cores} host threads"
    line = {"impl": "reference", "metric": "closest-hit ray throughput (coherent primary + incoherent bounce)", "value": value, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 in / f64 arithmetic", "data": "synthetic",
            "config": {"workload": f"C2 displaced icosphere level {args.level} ({len(tris)} triangles), {res}x{res} primary + bounce rays per GPU",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    # for information: the reference's own GPU path (CUDA backend / OptiX) on the same sampled rays, if its binaries are here
    if not args.no_reference_gpu:
        line["reference_gpu"] = reference_gpu(verts, tris, ps, bs)
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    import luisacompute_b200 as lc
    from luisacompute_b200 import scenes as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU: luisacompute_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from luisacompute_b200.dist import bind_to_gpu_numa
        numa = bind_to_gpu_numa(local_rank)   # pinned host buffers of the e2e path next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = lc.Device(local_rank)
    tdev = dev.torch_device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ scene: rank 0 builds, NCCL broadcasts the BLAS blob
    verts, tris, gen_s = make_scene(args.level)
    ntris = len(tris)
    mesh = dev.create_mesh(verts, tris)
    build_ms = None
    bcast_ms = None
    if rank == 0:
        mesh.build()                                              # warm-up build (allocations, module load)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(5):
            e0.record()
            mesh.build(lc.FORCE_BUILD)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        build_ms = min(times)
    if world > 1:
        if rank == 0:
            blob = mesh.export_blob()
            nbytes = torch.tensor([blob.numel()], dtype=torch.int64, device=tdev)
        else:
            nbytes = torch.zeros(1, dtype=torch.int64, device=tdev)
        dist.broadcast(nbytes, 0)
        if rank != 0:
            blob = torch.empty(int(nbytes.item()), dtype=torch.uint8, device=tdev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(blob, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            mesh.import_blob(blob)
    acc = dev.create_accel()
    acc.emplace_back(mesh, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    torch.cuda.synchronize()
    info = mesh.info()
    assert info.built == 1 and info.error == 0, (info.built, info.error)

    # ------------------------------------------------------------------ ray batches (resident in HBM before timing)
    res = args.res
    primary = primary_for_rank(res, rank, world)
    n = len(primary)
    d_primary = torch.from_numpy(primary.view(np.uint8).reshape(-1)).to(tdev)
    d_hits_p = torch.empty((n, 6), dtype=torch.int32, device=tdev)
    d_hits_b = torch.empty((n, 6), dtype=torch.int32, device=tdev)
    acc.trace_closest(d_primary, out=d_hits_p)
    hp = lc.hits_to_numpy(d_hits_p)
    bounce = S.bounce_rays(primary, hp, verts, tris, res, res)
    d_bounce = torch.from_numpy(bounce.view(np.uint8).reshape(-1)).to(tdev)
    d_any = torch.empty(n, dtype=torch.int32, device=tdev)

    # ------------------------------------------------------------------ counted pass (instrumented kernel, untimed)
    ctr_p = torch.zeros(4, dtype=torch.int64, device=tdev)
    ctr_b = torch.zeros(4, dtype=torch.int64, device=tdev)
    acc.trace_closest(d_primary, out=d_hits_p, counters=ctr_p)
    acc.trace_closest(d_bounce, out=d_hits_b, counters=ctr_b)
    torch.cuda.synchronize()
    cp, cb = ctr_p.cpu().numpy() / n, ctr_b.cpu().numpy() / n
    bytes_p = 32 + 24 + cp[0] * 80 + cp[1] * 48 + cp[2] * 80
    bytes_b = 32 + 24 + cb[0] * 80 + cb[1] * 48 + cb[2] * 80

    # ------------------------------------------------------------------ timed region
    # N = 1: one launch per batch.  N > 1: the same rays in chunks of CHUNK_TILES tiles, each chunk's hit records gathered
    # into ray order on rank 0 (side stream) while the next chunk is traced; the step ends when the gather has.
    K, W = args.steps, max(args.warmup, 3)
    gat_p = gat_b = fb_p = fb_b = None
    n_tiles = n // TILE
    if world > 1:
        from luisacompute_b200.dist import HitGatherer
        assert n % TILE == 0
        fb_p = make_peer_framebuffer(dev, n_tiles)
        fb_b = make_peer_framebuffer(dev, n_tiles) if fb_p is not None else None
        if fb_b is None:
            fb_p = None
        gat_p, gat_b = HitGatherer(n_tiles, tdev, chunk_tiles=CHUNK_TILES), HitGatherer(n_tiles, tdev, chunk_tiles=CHUNK_TILES)
    # scratch for the traversal's provisional records when the final ones go to the peer framebuffer
    d_scr = torch.empty((n, 6), dtype=torch.int32, device=tdev) if fb_p is not None else None

    def trace_batch(d_rays, d_hits, gat):
        if gat is None:
            acc.trace_closest(d_rays, out=d_hits)
            return
        if fb_p is not None and gather_mode() == "stores":
            (fb_p if gat is gat_p else fb_b).trace_closest(acc, d_rays, scratch=d_scr)
            return
        if fb_p is not None:
            fb = fb_p if gat is gat_p else fb_b
            for t0 in range(0, n_tiles, DMA_CHUNK_TILES):
                t1 = min(t0 + DMA_CHUNK_TILES, n_tiles)
                acc.trace_closest(d_rays[t0 * TILE * 32:t1 * TILE * 32], out=d_hits[t0 * TILE:t1 * TILE])
                e = torch.cuda.Event()
                e.record()
                fb.submit(d_hits[t0 * TILE:t1 * TILE], t0, e)
            return
        for t0 in range(0, n_tiles, CHUNK_TILES):
            t1 = min(t0 + CHUNK_TILES, n_tiles)
            acc.trace_closest(d_rays[t0 * TILE * 32:t1 * TILE * 32], out=d_hits[t0 * TILE:t1 * TILE])
            e = torch.cuda.Event()
            e.record()
            gat.submit(d_hits[t0 * TILE:t1 * TILE], t0, t1, e)

    def end_step():
        if fb_p is not None:
            if gather_mode() == "dma":
                fb_p.complete_local()
            fb_b.complete()     # one all-reduce orders rank 0's stream behind every rank's launches and copies of the step
        elif gat_p is not None:
            gat_p.finish(), gat_b.finish()

    for _ in range(W):
        trace_batch(d_primary, d_hits_p, gat_p)
        trace_batch(d_bounce, d_hits_b, gat_b)
        end_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    barrier()
    for k in range(K):
        ev[k][0].record()
        trace_batch(d_primary, d_hits_p, gat_p)
        ev[k][1].record()
        trace_batch(d_bounce, d_hits_b, gat_b)
        end_step()
        ev[k][2].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0][0].elapsed_time(ev[K - 1][2])
    tp_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]))
    tb_ms = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(K)]))
    gather_ms = None
    if world > 1:
        t = torch.tensor([total_ms, tp_ms, tb_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, tp_ms, tb_ms = (float(x) for x in t.cpu())
        # the gather by itself (nothing to hide behind): one batch, all chunks
        for t0 in range(0, n_tiles, CHUNK_TILES):   # (warm-up: NCCL sets its send / recv channels up on first use)
            gat_b.submit(d_hits_b[t0 * TILE:min(t0 + CHUNK_TILES, n_tiles) * TILE], t0, min(t0 + CHUNK_TILES, n_tiles), None)
        gat_b.finish()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for t0 in range(0, n_tiles, CHUNK_TILES):
            t1 = min(t0 + CHUNK_TILES, n_tiles)
            gat_b.submit(d_hits_b[t0 * TILE:t1 * TILE], t0, t1, None)
        gat_b.finish()
        g1.record()
        torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = float(t.item())
        # verification: rank 0's framebuffers hold every rank's records where ray order puts them
        chk = torch.stack([d_hits_p[:, 0].long().sum(), d_hits_p[:, 1].long().sum(), d_hits_b[:, 0].long().sum(), d_hits_b[:, 1].long().sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)] if rank == 0 else None
        dist.gather(chk, allc, 0)
        if rank == 0:
            for gat, base in ((fb_p if fb_p is not None else gat_p, 0), (fb_b if fb_b is not None else gat_b, 2)):
                fb = gat.full.view(n_tiles, world, TILE, 6)
                for rr in range(world):
                    assert int(fb[:, rr, :, 0].long().sum()) == int(allc[rr][base]) and int(fb[:, rr, :, 1].long().sum()) == int(allc[rr][base + 1]), \
                        f"gathered hit records of rank {rr} differ from what it traced"
            assert torch.equal((fb_p if fb_p is not None else gat_p).full.view(n_tiles, world, TILE, 6)[:, 0].reshape(-1, 6), d_hits_p)
    ms_per_step = total_ms / K
    value = world * 2 * n / (ms_per_step * 1e-3) / 1e6

    # any-hit (shadow-style) on the incoherent batch, reported as detail
    for _ in range(2):
        acc.trace_any(d_bounce, out=d_any)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(3):
        acc.trace_any(d_bounce, out=d_any)
    a1.record()
    torch.cuda.synchronize()
    any_ms = a0.elapsed_time(a1) / 3

    # ------------------------------------------------------------------ e2e: host buffers through the public call
    e2e = None
    e2e_chunk = int(os.environ.get("LCB_E2E_CHUNK", str(1 << 19)))
    if not args.no_e2e:
        h_rays = torch.empty(2 * n * 32, dtype=torch.uint8).pin_memory()
        h_rays[:n * 32] = torch.from_numpy(primary.view(np.uint8).reshape(-1))
        h_rays[n * 32:] = torch.from_numpy(bounce.view(np.uint8).reshape(-1))
        h_hits = torch.empty((2 * n, 6), dtype=torch.int32).pin_memory()
        for _ in range(2):
            acc.trace_closest_host(h_rays, h_hits, chunk_rays=e2e_chunk)
        barrier()
        t0 = time.perf_counter()
        reps = max(3, min(K, 10))
        for _ in range(reps):
            acc.trace_closest_host(h_rays, h_hits, chunk_rays=e2e_chunk)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * reps * 2 * n / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": 2 * n * 24,
               "ms_per_step": dt / reps * 1e3, "api": "Accel.trace_closest_host -> lcbTraceClosestHost (pinned host rays in, pinned host hits out, "
               f"{e2e_chunk >> 10}Ki-ray chunks, four in flight on four streams)"}
        # the ceiling of that path on THIS box: the same bytes (rays in, hits out) moved by plain copies on two streams, all ranks
        # at once, nothing traced — what host<->device bandwidth alone allows
        d_stage_in = torch.empty(2 * n * 32, dtype=torch.uint8, device=tdev)
        d_stage_out = torch.empty((2 * n, 6), dtype=torch.int32, device=tdev)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ck = (1 << 19) * 32
        best = 1e30
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for off in range(0, 2 * n * 32, ck):
                with torch.cuda.stream(s_in):
                    d_stage_in[off:off + ck].copy_(h_rays[off:off + ck], non_blocking=True)
                r0 = off // 32
                with torch.cuda.stream(s_out):
                    h_hits[r0:r0 + (1 << 19)].copy_(d_stage_out[r0:r0 + (1 << 19)], non_blocking=True)
            s_in.synchronize(), s_out.synchronize()
            barrier()
            best = min(best, time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([best], dtype=torch.float64, device=tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item())
        e2e["copy_only_ms_per_step"] = best * 1e3
        e2e["copy_only_mrays_s"] = world * 2 * n / best / 1e6
        e2e["copy_only_gbs_aggregate"] = world * 2 * n * 56 / best / 1e9
        e2e["fraction_of_copy_ceiling"] = e2e["value"] / e2e["copy_only_mrays_s"]
        del d_stage_in, d_stage_out
        # restore what the ceiling probe overwrote, through the real path
        acc.trace_closest_host(h_rays, h_hits, chunk_rays=e2e_chunk)
        # the host path must return exactly what the resident path returned
        assert torch.equal(h_hits[:n], d_hits_p.cpu()) and torch.equal(h_hits[n:], d_hits_b.cpu())

    # ------------------------------------------------------------------ BASELINE config 4, strong scaling (detail)
    c4 = None
    gather_text = gather_description(fb_p, gat_p)
    if not args.no_c4:
        if fb_p is not None:
            fb_p.close(), fb_b.close()
        fb_p = fb_b = gat_p = gat_b = d_scr = None
        torch.cuda.empty_cache()
        try:
            c4 = c4_strong(args, lc, dev, world, rank, barrier)
        except torch.cuda.OutOfMemoryError as e:  # a smaller GPU than the 180 GB this is sized for
            if world > 1:
                raise
            c4 = {"unavailable": f"out of memory: {e}"[:200]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        achieved = bytes_b * n / (tb_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, per launch, from the committed
        # `ncu --set full` capture (profiles/ncu_traffic.json names the report); null if the workload is not the profiled one
        traffic = None
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_trace_closest_incoherent"]
            if t["level"] == args.level and t["rays"] == n:
                traffic = t["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_trace<MODE_CLOSEST> on the incoherent bounce batch", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_ray": float(bytes_b), "per_ray": {"nodes": float(cb[0]), "prims": float(cb[1]), "instances": float(cb[2])},
                    "roofline_mrays_s": peak * 1e9 / bytes_b / 1e6, "coherent": {"bytes_per_ray": float(bytes_p), "achieved": bytes_p * n / (tp_ms * 1e-3) / 1e9,
                                                                                 "per_ray": {"nodes": float(cp[0]), "prims": float(cp[1]), "instances": float(cp[2])}}}
        # build roofline, SURVEY §8(d): bytes / primitive = 48 (3 vertices) + 12 (indices) + 2 * 8 * P (8-byte key read + written per
        # radix pass; the primitive index rides in the key's low bits, there is no payload array) + 16 (bounds / centroid scratch)
        # + 48 (triangle record out) + 80 * nodes / primitive (nodes out)
        def build_roofline(n_prims, n_nodes, ms):
            if not ms:
                return None
            lg = max(1, int(np.ceil(np.log2(max(n_prims, 2)))))
            bits = min(max((int(np.ceil(np.log2(max(n_prims, 1)))) + 2) // 3 + 6, 10), 21, (64 - lg) // 3)
            passes = (3 * bits + 7) // 8
            bpp = 48 + 12 + 16 * passes + 16 + 48 + 80.0 * n_nodes / n_prims
            ach = bpp * n_prims / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "bytes_per_prim": bpp, "radix_passes": passes, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "mprims_s": n_prims / ms / 1e3, "roofline_mprims_s": peak * 1e9 / bpp / 1e6, "ms": ms}
        roofline_build = build_roofline(ntris, info.node_count, build_ms)
        if c4 and c4.get("build_ms"):
            c4["roofline_build"] = build_roofline(args.c4_tris, c4["bvh_nodes"], c4["build_ms"])
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(verts, tris, primary, bounce, args.cpu_sample)
        ref_gpu = reference_gpu(verts, tris, primary, bounce) if (world == 1 and not args.no_reference_gpu) else None
        dropin = reference_gpu(verts, tris, primary, bounce, backend="b200") if (world == 1 and not args.no_reference_gpu) else None
        hit_frac_p = float(np.mean(hp["inst"] != 0xFFFFFFFF))
        line = {
            "metric": "closest-hit ray throughput (coherent primary + incoherent bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (box tests) + f64 (exact triangle test)", "data": "synthetic",
            "config": {"workload": f"C2 displaced icosphere level {args.level} ({ntris} triangles, 1 instance), {res}x{res} primary + {res}x{res} "
                                   f"cosine-bounce closest-hit rays per GPU",
                       "l2_policy": f"inputs larger than L2: {2 * n * 32 / 1e6:.0f} MB rays + {2 * n * 24 / 1e6:.0f} MB hits + "
                                    f"{info.total_bytes / 1e6:.0f} MB BVH per step vs 126 MB L2",
                       "sharding": ("single GPU" if world == 1 else f"replicated scene (BLAS built on rank 0, NCCL broadcast), 65536-ray tiles round-robin over {world} ranks; every "
                                    f"rank's hit records are gathered into ray order on rank 0 INSIDE the timed region (NCCL send/recv per {CHUNK_TILES}-tile chunk "
                                    f"on a side stream, overlapped with the next chunk's trace)")},
            "detail": {"coherent_closest_mrays_s": world * n / (tp_ms * 1e-3) / 1e6, "incoherent_closest_mrays_s": world * n / (tb_ms * 1e-3) / 1e6,
                       "incoherent_any_mrays_s": n / (any_ms * 1e-3) / 1e6, "primary_hit_fraction": hit_frac_p,
                       "blas_build_ms": build_ms, "blas_build_mtris_s": (ntris / (build_ms * 1e-3) / 1e6) if build_ms else None,
                       "numa_binding": list(numa) if numa else None, "bvh_nodes": info.node_count, "bvh_bytes": info.total_bytes, "bvh_broadcast_ms": bcast_ms, "hit_gather_ms": gather_ms,
                       "hit_gather": gather_text,
                       "hit_gather_note": None if world == 1 else (f"hit_gather_ms = NCCL send/recv gather of one batch ({(world - 1) * n * 24 / 1e6:.0f} MB into rank 0) by itself, for comparison"
                                                                   + ("; the timed step does not use it" if gather_text.startswith("peer") else "; in the step it hides behind the traces")),
                       "c4_strong": c4,
                       "scene_gen_s": gen_s},
            "roofline": roofline, "roofline_build": roofline_build, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "dropin_backend": dropin, "e2e": e2e, "clocks": clocks,
            # per step: k_trace + k_refine_hits per launch; one launch per batch at N = 1, one per CHUNK_TILES tiles otherwise
            "gpu_launches": 4 * K if (world == 1 or gather_text.startswith("peer framebuffer, stores")) else 4 * K * ((n_tiles + DMA_CHUNK_TILES - 1) // DMA_CHUNK_TILES) if gather_text.startswith("peer") else 4 * K * ((n_tiles + CHUNK_TILES - 1) // CHUNK_TILES),
        }
        print(json.dumps(line))
    if fb_p is not None:
        fb_p.close(), fb_b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
