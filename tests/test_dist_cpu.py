"""World-size-2 gloo test of the multi-GPU plumbing (tile sharding, blob broadcast, hit gather) on CPU tensors."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luisacompute_b200 import dist as D


def test_tiles_partition_all_rays_exactly_once():
    for n, world, tile in ((1000, 2, 64), (65536 * 5 + 17, 4, 65536), (10, 8, 4), (64 << 20, 8, 65536)):
        seen = np.zeros(n, dtype=np.int32) if n < (1 << 22) else None
        total = 0
        for r in range(world):
            idx = D.ray_indices_for_rank(n, r, world, tile)
            total += len(idx)
            if seen is not None:
                seen[idx] += 1
        assert total == n
        if seen is not None:
            assert np.all(seen == 1)
    # 64 Mi rays over 8 GPUs: every rank gets the same number of 65536-ray tiles (C4)
    assert len(D.tiles_for_rank(64 << 20, 3, 8)) == 128


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = torch.arange(12345, dtype=torch.int64).to(torch.uint8) if rank == 0 else None
        blob = D.broadcast_blob(blob, 0, device=torch.device("cpu"))
        ok_blob = bool(torch.equal(blob, torch.arange(12345, dtype=torch.int64).to(torch.uint8)))
        n = 1000
        idx = D.ray_indices_for_rank(n, rank, world, tile=64)
        local = torch.zeros((len(idx), 6), dtype=torch.int32)
        local[:, 0] = torch.from_numpy(idx).int()          # "hit" = the global ray index
        full = D.gather_hits(local, n, 0, tile=64)
        ok_gather = True if rank != 0 else bool(torch.equal(full[:, 0], torch.arange(n, dtype=torch.int32)))
        # the pipelined gatherer bench.py times: 6 local tiles of 16 rays per rank, chunks of 4 + 2 tiles
        tile, tl = 16, 6
        g = D.HitGatherer(tl, "cpu", chunk_tiles=4, tile=tile)
        gidx = (torch.arange(tl)[:, None] * world + rank) * tile + torch.arange(tile)[None, :]      # global ray index of every local ray
        local = torch.zeros((tl * tile, 6), dtype=torch.int32)
        local[:, 0] = gidx.reshape(-1).int()
        local[:, 5] = rank
        for t0, t1 in ((0, 4), (4, 6)):
            g.submit(local[t0 * tile:t1 * tile], t0, t1)
        g.finish()
        fb = g.framebuffer()
        ok_pipe = True
        if rank == 0:
            ok_pipe = bool(torch.equal(fb[:, 0], torch.arange(tl * tile * world, dtype=torch.int32)))
            ok_pipe = ok_pipe and bool(torch.equal(fb[:, 5].view(tl, world, tile)[:, 1], torch.ones((tl, tile), dtype=torch.int32)))
        ret[rank] = ok_blob and ok_gather and ok_pipe
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_peer_framebuffer_layout_is_ray_order():
    """The (base, stride) every rank stores with (dist.peer_tile_layout, used by PeerFramebuffer / lcrtTraceClosestTiled) puts
    record i of a tile-sharded batch where ray order wants it — the same order dist.gather_hits produces."""
    from luisacompute_b200 import dist as D
    tile, world, n_local_tiles, rec = 8, 3, 5, 24
    n = tile * world * n_local_tiles
    fb = np.full(n * rec, 0xFF, dtype=np.uint8)
    for r in range(world):
        idx = D.ray_indices_for_rank(n, r, world, tile)
        base, stride = D.peer_tile_layout(r, world, tile, rec)
        for i_local, ray in enumerate(idx):
            t, k = divmod(i_local, tile)
            off = base + t * stride + k * rec                      # the address computation of HitDest::at (lcrt_kernels.cu)
            fb[off:off + rec] = np.frombuffer(np.uint64(ray).tobytes() * 3, dtype=np.uint8)
    got = fb.view(np.uint64).reshape(n, 3)
    assert np.array_equal(got[:, 0], np.arange(n, dtype=np.uint64)) and np.array_equal(got[:, 2], np.arange(n, dtype=np.uint64))
