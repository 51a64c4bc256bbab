"""Multi-GPU plumbing: replicated scene, rays sharded by tile.

Nothing like this exists in the reference (no NCCL / MPI anywhere).  One process per GPU; NCCL is used only
to broadcast the built BLAS blob from the rank that built it and to gather hit records / checksums; there is no
collective during traversal.  The functions work with any torch.distributed backend (the CPU tests use gloo).
"""
from __future__ import annotations

import numpy as np

TILE = 65536


def tiles_for_rank(n_rays: int, rank: int, world: int, tile: int = TILE) -> np.ndarray:
    """Indices of the tiles (of `tile` consecutive rays) that `rank` traces: tile j -> rank j mod world."""
    n_tiles = (n_rays + tile - 1) // tile
    return np.arange(rank, n_tiles, world, dtype=np.int64)


def ray_indices_for_rank(n_rays: int, rank: int, world: int, tile: int = TILE) -> np.ndarray:
    t = tiles_for_rank(n_rays, rank, world, tile)
    idx = (t[:, None] * tile + np.arange(tile, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n_rays]


def broadcast_blob(blob, src: int = 0, device=None):
    """Broadcast a uint8 tensor whose size only `src` knows.  Returns the tensor on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = blob.device if blob is not None else device
    n = torch.tensor([blob.numel() if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if rank != src:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(blob, src)
    return blob


def gather_hits(local_hits, n_rays: int, dst: int = 0, tile: int = TILE):
    """Gather per-rank hit records ([n_local, 6] int32, tile-sharded) back into ray order on `dst`."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = [len(ray_indices_for_rank(n_rays, r, world, tile)) for r in range(world)]
    pad = max(counts)
    buf = torch.zeros((pad, local_hits.shape[1]), dtype=local_hits.dtype, device=local_hits.device)
    buf[:local_hits.shape[0]] = local_hits
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst)
    if rank != dst:
        return None
    full = torch.zeros((n_rays, local_hits.shape[1]), dtype=local_hits.dtype, device=local_hits.device)
    for r in range(world):
        idx = torch.from_numpy(ray_indices_for_rank(n_rays, r, world, tile)).to(full.device)
        full[idx] = out[r][:counts[r]]
    return full


def bind_to_gpu_numa(device_index: int):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are allocated:
    with one rank per GPU on a two-socket box, pinned ray / hit buffers that all live on socket 0 make the ranks of the
    other socket cross the inter-socket link for every host<->device copy (measured on 8 x B200: end-to-end throughput
    fell from 4 to 8 ranks).  Returns (node, cpu count) or None if the topology cannot be read."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except Exception:
        return None



class HitGatherer:
    """Gather of tile-sharded 24-byte hit records into RAY ORDER on `dst`, chunk by chunk, overlapped with tracing.

    The local batch of every rank is `n_local_tiles` tiles of `tile` rays; local tile t of rank r is global tile
    t * world + r (tiles dealt round-robin).  `submit(chunk, t0, t1, after)` queues the gather of the local
    tiles [t0, t1) on a high-priority side stream behind the CUDA event `after` (recorded when the chunk's trace has been
    enqueued), so the NCCL transfer and the reorder copy of chunk c run while chunk c + 1 is being traced.  `dst` receives
    into a double-buffered staging area [world, chunk] and one strided copy per chunk puts the records into the
    framebuffer `full` ([n_local_tiles, world, tile * 6] == ray order).  `finish()` makes the caller's stream wait for the
    side stream.  Works with any backend (CPU tensors / gloo run inline)."""

    def __init__(self, n_local_tiles: int, device, chunk_tiles: int = 8, tile: int = TILE, dst: int = 0, words: int = 6):
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.dst = dist.get_rank(), dist.get_world_size(), dst
        self.tile, self.words, self.chunk_tiles, self.n_local_tiles = tile, words, chunk_tiles, n_local_tiles
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(self.device, priority=-1) if self.cuda else None
        self.full = None
        self.stage = None
        if self.rank == dst:
            self.full = torch.empty((n_local_tiles, self.world, tile * words), dtype=torch.int32, device=self.device)
            self.stage = [torch.empty((self.world, chunk_tiles * tile, words), dtype=torch.int32, device=self.device) for _ in range(2)]
        self._k = 0

    def _ctx(self):
        import contextlib
        import torch
        return torch.cuda.stream(self.comm) if self.cuda else contextlib.nullcontext()

    def submit(self, chunk, t0: int, t1: int, after=None):
        import torch.distributed as dist
        n = (t1 - t0) * self.tile
        assert chunk.shape[0] == n and t1 - t0 <= self.chunk_tiles
        with self._ctx():
            if self.cuda and after is not None:
                self.comm.wait_event(after)
            if self.rank == self.dst:
                st = self.stage[self._k & 1]
                dist.gather(chunk, [st[r, :n] for r in range(self.world)], self.dst)
                self.full[t0:t1].copy_(st[:, :n].reshape(self.world, t1 - t0, self.tile * self.words).transpose(0, 1))
            else:
                dist.gather(chunk, None, self.dst)
        self._k += 1

    def finish(self):
        import torch
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)

    def framebuffer(self):
        """[n_rays, words] in ray order (dst only)."""
        return None if self.full is None else self.full.view(-1, self.words)


def peer_tile_layout(rank: int, world: int, tile: int = TILE, record_bytes: int = 24):
    """(byte offset of this rank's local tile 0, byte stride between its consecutive local tiles) in a ray-ordered framebuffer of
    `record_bytes` records when tiles of `tile` rays are dealt round-robin: local tile t of rank r is global tile t * world + r."""
    tile_bytes = tile * record_bytes
    return rank * tile_bytes, world * tile_bytes


class _RawCudaBuffer:
    """A device address as something torch.as_tensor understands (int32 view of `words` elements)."""

    def __init__(self, address: int, words: int):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i4", "data": (address, False), "version": 3, "strides": None}


class PeerFramebuffer:
    """The gathering rank's ray-ordered framebuffer of 24-byte hit records, mapped into EVERY rank of the node so that each
    rank's trace writes its tiles where ray order puts them (Accel.trace_closest_tiled -> lcrtTraceClosestTiled): the hit
    gather of SURVEY.md 8e without a gather — the stores travel over NVLink while the rank is still tracing, rank `dst`
    receives nothing it has to reorder, and the only collective left in a step is the completion barrier.

    Layout: global tile j (`tile` consecutive rays) at byte j * tile * 24; local tile t of rank r is global tile
    t * world + r, i.e. rank r writes with base = r * tile * 24 and stride = world * tile * 24.
    `dst` allocates (lcbFramebufferCreate: cudaMalloc + CUDA IPC handle), the 64-byte handle is broadcast with
    torch.distributed, the others map it (lcbFramebufferOpen, peer access enabled on first use).  Single-process use
    (world 1) works the same without the exchange."""

    def __init__(self, device, n_local_tiles: int, tile: int = TILE, dst: int = 0, words: int = 6):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _check, lib
        self.device, self.tile, self.words, self.dst = device, tile, words, dst
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_local_tiles = n_local_tiles
        self.bytes = n_local_tiles * self.world * tile * words * 4
        self.owner = self.rank == dst
        tdev = device.torch_device
        handle = torch.zeros(64, dtype=torch.uint8, device=tdev)
        addr = C.c_uint64(0)
        if self.owner:
            hb = (C.c_ubyte * 64)()
            _check(lib().lcbFramebufferCreate(device._h, self.bytes, C.byref(addr), hb), "lcbFramebufferCreate")
            handle.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
        if self.world > 1:
            dist.broadcast(handle, dst)
        if not self.owner:
            hb = (C.c_ubyte * 64).from_buffer_copy(bytes(handle.cpu().numpy().tobytes()))
            _check(lib().lcbFramebufferOpen(device._h, hb, C.byref(addr)), "lcbFramebufferOpen")
        self.address = int(addr.value)
        self.tile_bytes = tile * words * 4
        offset, self.stride = peer_tile_layout(self.rank, self.world, tile, words * 4)
        self.base = self.address + offset                               # where this rank's local tile 0 goes
        # the owner's view of the whole buffer, ray order
        self.full = torch.as_tensor(_RawCudaBuffer(self.address, self.bytes // 4), device=tdev).view(-1, words) if self.owner else None

    def trace_closest(self, accel, rays, mask=0xFF, scratch=None, first_local_tile: int = 0, store_each_record=None):
        """Traces this rank's rays (local tiles first_local_tile ...) and stores the records into the framebuffer FROM THE TRACE
        KERNEL: no second pass, no copy, the stores leave over NVLink while the kernel is still tracing.  By default the
        records of a 128-ray batch go out together when the batch is complete (512-byte warp stores) when more than two
        ranks write into the owner, and one by one as the rays end otherwise: single 24-byte stores cost the kernel less, but
        eight GPUs sending them into one use its NVLink ingress badly (profiles/r2_scaling.md)."""
        if store_each_record is None:
            import os
            forced = os.environ.get("LCB200_PEER_STORE")     # developer knob: each | batch
            store_each_record = forced == "each" if forced in ("each", "batch") else self.world <= 2
        return accel.trace_closest_tiled(rays, self.base + first_local_tile * self.stride, self.tile, self.stride, mask=mask, scratch=scratch,
                                         store_each_record=store_each_record)

    def submit(self, hits, first_local_tile: int, after=None):
        """Queues the copy of finished local tiles (`hits`: int32 [k * tile, words], contiguous, final records) to where ray
        order puts them in the framebuffer: ONE strided device-to-device copy on a side stream, executed by the copy engines
        over NVLink, behind the CUDA event `after` (recorded when the chunk's trace was enqueued) — it runs while the SMs
        trace the next chunk."""
        import torch

        from . import _check, lib
        k = hits.shape[0] // self.tile
        assert hits.is_contiguous() and hits.shape[0] == k * self.tile
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device.torch_device, priority=-1)
        if after is not None:
            self._copy_stream.wait_event(after)
        else:
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device.torch_device))
        _check(lib().lcbFramebufferCopyTiles(self.device._h, self._copy_stream.cuda_stream, hits.data_ptr(), self.base + first_local_tile * self.stride,
                                             self.tile_bytes, self.stride, k), "lcbFramebufferCopyTiles")

    def complete_local(self):
        """The caller's stream waits for this framebuffer's queued copies (no collective)."""
        import torch
        if hasattr(self, "_copy_stream"):
            torch.cuda.current_stream(self.device.torch_device).wait_stream(self._copy_stream)

    def complete(self):
        """Every rank's records of the step are in the owner's memory once this returns on the owner's stream: kernel / copy
        completion makes them visible system-wide, the (tiny) all-reduce orders the owner's stream behind every rank's work."""
        import torch
        import torch.distributed as dist
        if hasattr(self, "_copy_stream"):
            torch.cuda.current_stream(self.device.torch_device).wait_stream(self._copy_stream)
        if self.world > 1:
            if not hasattr(self, "_token"):
                self._token = torch.zeros(1, dtype=torch.int32, device=self.device.torch_device)
            dist.all_reduce(self._token)

    def close(self):
        """Collective: the peers unmap, then the owner frees."""
        import torch

        from . import _check, lib
        if not self.address:
            return
        torch.cuda.synchronize()
        multi = self.world > 1
        if multi:
            import torch.distributed as dist
            dist.barrier()          # nobody unmaps or frees while a peer may still be storing
        self.full = None
        if not self.owner:
            _check(lib().lcbFramebufferClose(self.device._h, self.address), "lcbFramebufferClose")
        if multi:
            dist.barrier()          # every peer has unmapped
        if self.owner:
            _check(lib().lcbFramebufferDestroy(self.device._h, self.address), "lcbFramebufferDestroy")
        self.address = 0
