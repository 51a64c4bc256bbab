"""The tiled / peer-memory output path of closest-hit batches (lcrtTraceClosestTiled, dist.PeerFramebuffer): records
written through it are bit-identical to the in-place ones and land where ray order puts them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    import luisacompute_b200 as lc
    from luisacompute_b200 import scenes as S
    dev = lc.Device(0)
    v, t = S.icosphere(5)
    mesh = dev.create_mesh(v, t).build()
    acc = dev.create_accel()
    acc.emplace_back(mesh, S.affine_identity(), 0xFF, True, 0)
    acc.build()
    rays = S.pinhole_rays(512, 256)          # 131072 rays = 8 tiles of 16384; a quarter of them miss
    yield dev, acc, rays
    dev.close()


def test_framebuffer_single_rank_equals_in_place(scene):
    import torch
    from luisacompute_b200.dist import PeerFramebuffer
    dev, acc, rays = scene
    ref = acc.trace_closest(rays)
    fb = PeerFramebuffer(dev, n_local_tiles=8, tile=16384)
    for each in (False, True):      # records of a 128-ray batch stored together / every record on its own
        fb.full.fill_(0x55555555)
        fb.trace_closest(acc, rays, store_each_record=each)
        fb.complete()
        torch.cuda.synchronize()
        assert torch.equal(fb.full, ref), each
    assert int((ref[:, 0] == -1).sum()) > 1000 and int((ref[:, 0] != -1).sum()) > 1000   # misses and hits both went through
    fb.close()


def test_copy_engine_path_fills_ray_order(scene):
    """PeerFramebuffer.submit: finished local tiles go to their ray-order places with one strided copy per chunk."""
    import torch
    from luisacompute_b200.dist import PeerFramebuffer
    dev, acc, rays = scene
    ref = acc.trace_closest(rays)
    fb = PeerFramebuffer(dev, n_local_tiles=8, tile=16384)
    fb.full.fill_(0x55555555)
    for t0 in range(0, 8, 3):                       # ragged chunks: 3 + 3 + 2 tiles
        t1 = min(t0 + 3, 8)
        ev = torch.cuda.Event()
        ev.record()
        fb.submit(ref[t0 * 16384:t1 * 16384], t0, ev)
    fb.complete()
    torch.cuda.synchronize()
    assert torch.equal(fb.full, ref)
    # with a stride: the tiles of "rank 1 of 2" interleave with untouched ones
    fb.full.fill_(7)
    fb.world, fb.stride, fb.base = 2, 2 * fb.tile_bytes, fb.address + fb.tile_bytes
    fb.submit(ref[:4 * 16384], 0)
    torch.cuda.current_stream().wait_stream(fb._copy_stream)
    torch.cuda.synchronize()
    v = fb.full.view(4, 2, 16384, 6)
    assert torch.equal(v[:, 1].reshape(-1, 6), ref[:4 * 16384]) and bool((v[:, 0] == 7).all())
    fb.world = 1
    fb.close()


def test_two_interleaved_writers_fill_ray_order(scene):
    """Two 'ranks' in one process: writer r traces the tiles j = r mod 2 and stores them with base r * tile, stride 2 * tile."""
    import ctypes as C

    import torch
    import luisacompute_b200 as lc
    dev, acc, rays = scene
    tile, n_tiles = 16384, 8
    ref = acc.trace_closest(rays)
    addr, hb = C.c_uint64(0), (C.c_ubyte * 64)()
    lc._check(lc.lib().lcbFramebufferCreate(dev._h, n_tiles * tile * 24, C.byref(addr), hb), "lcbFramebufferCreate")
    try:
        r8 = torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(n_tiles // 2, 2, tile * 32)).to(dev.torch_device)
        for r in range(2):
            local = r8[:, r].contiguous().view(-1)
            acc.trace_closest_tiled(local, addr.value + r * tile * 24, tile, 2 * tile * 24)
        torch.cuda.synchronize()
        got = np.zeros(n_tiles * tile * 6, dtype=np.int32)
        lc._check(lc.lib().lcbDownload(dev._h, None, addr.value, got.ctypes.data_as(C.c_void_p), got.nbytes), "lcbDownload")
        assert np.array_equal(got.reshape(-1, 6), ref.cpu().numpy())
    finally:
        lc._check(lc.lib().lcbFramebufferDestroy(dev._h, addr.value), "lcbFramebufferDestroy")


def test_tiled_arguments_are_validated(scene):
    import torch
    import luisacompute_b200 as lc
    dev, acc, rays = scene
    scratch = torch.empty((len(rays), 6), dtype=torch.int32, device=dev.torch_device)
    with pytest.raises(lc.LcrtError):
        acc.trace_closest_tiled(rays, scratch.data_ptr(), 16384, 16384 * 24 - 8, scratch=scratch)   # tiles would overlap
    with pytest.raises(lc.LcrtError):
        acc.trace_closest_tiled(rays, 0, 16384, 16384 * 24, scratch=scratch)                        # no destination
