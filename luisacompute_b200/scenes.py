"""Deterministic synthetic scenes and ray batches for the BASELINE.json configs.

Host-side numpy only; nothing here is timed.  Scene sources in the reference that these follow:
  C1  src/tests/test_rtx.cpp:24-28,67-73,101-105   (triangle, two instance transforms, orthographic rays)
  C2  icosphere level 8 + analytic displacement, pinhole primaries, cosine-hemisphere bounce with the
      LCG / tea generators of src/tests/test_path_tracing.cpp:112-134 and offset_ray_origin of
      src/dsl/rtx/ray.cpp:16-23
  C3  16 BLAS x 100 000 instances, masks 1 << (i & 7), PCG32(seed 42, stream i)
  C4  random triangle soup, PCG32(1234) / rays PCG32(5678)
  C5  Cornell-box style room (5 walls, 2 boxes, 1 light: 8 meshes / 36 triangles) + AABB spheres
      with the lcg of src/tests/test_procedural.cpp:29-40
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])          # 32 B
HIT_DTYPE = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", 2), ("t", "<f4"), ("pad", "<u4")])  # 24 B
CHIT_DTYPE = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", 2), ("hit_type", "<u4"), ("t", "<f4")])
MOD_DTYPE = np.dtype({"names": ["index", "user_id", "flags", "vis_mask", "affine", "primitive"],
                      "formats": ["<u4", "<u4", "<u4", "<u4", ("<f4", 12), "<u8"],
                      "offsets": [0, 4, 8, 12, 16, 64], "itemsize": 80})                              # 80 B, align 16
INSTANCE_DTYPE = np.dtype([("affine", "<f4", 12), ("user_id", "<u4"), ("sbt_offset", "<u4"), ("mask", "<u4"),
                           ("flags", "<u4"), ("handle", "<u8"), ("pad", "<u8")])                      # 80 B
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 24 and CHIT_DTYPE.itemsize == 24
assert MOD_DTYPE.itemsize == 80 and INSTANCE_DTYPE.itemsize == 80

MOD_PRIMITIVE, MOD_TRANSFORM, MOD_OPAQUE_ON, MOD_OPAQUE_OFF, MOD_VISIBILITY, MOD_USER_ID = 1, 2, 4, 8, 16, 32


# ----------------------------------------------------------------------------- affine helpers (row-major 3x4)
def affine_identity() -> np.ndarray:
    return np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)


def affine_from_4x4(m: np.ndarray) -> np.ndarray:
    """m is a conventional 4x4 (row r, col c) matrix acting on column vectors."""
    return np.asarray(m, dtype=np.float32)[:3, :].reshape(12).copy()


def mat_translation(x, y, z) -> np.ndarray:
    m = np.eye(4, dtype=np.float64)
    m[:3, 3] = (x, y, z)
    return m


def mat_scaling(s) -> np.ndarray:
    m = np.eye(4, dtype=np.float64)
    m[0, 0] = m[1, 1] = m[2, 2] = s
    return m


def mat_rotation(axis, angle) -> np.ndarray:
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    c, s = np.cos(angle), np.sin(angle)
    x, y, z = axis
    r = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = r
    return m


def make_mods(prims, affines, masks=None, opaque=None, user_ids=None, first_index=0) -> np.ndarray:
    """Modification records as Accel::emplace_back would produce (src/runtime/rtx/accel.cpp:88-100)."""
    n = len(prims)
    mods = np.zeros(n, dtype=MOD_DTYPE)
    mods["index"] = np.arange(first_index, first_index + n, dtype=np.uint32)
    mods["primitive"] = np.asarray(prims, dtype=np.uint64)
    mods["affine"] = np.asarray(affines, dtype=np.float32).reshape(n, 12)
    mods["vis_mask"] = 0xFF if masks is None else np.asarray(masks, dtype=np.uint32)
    mods["user_id"] = 0 if user_ids is None else np.asarray(user_ids, dtype=np.uint32)
    op = np.ones(n, dtype=bool) if opaque is None else np.asarray(opaque, dtype=bool)
    mods["flags"] = (MOD_PRIMITIVE | MOD_TRANSFORM | MOD_VISIBILITY | MOD_USER_ID) | np.where(op, MOD_OPAQUE_ON, MOD_OPAQUE_OFF)
    return mods


# ----------------------------------------------------------------------------- RNGs
def lcg_next(state: np.ndarray):
    """test_path_tracing.cpp:128-134: state = a*state + c; float = (state & 0xffffff) / 2^24."""
    state = (state.astype(np.uint64) * np.uint64(1664525) + np.uint64(1013904223)).astype(np.uint32)
    return state, (state & np.uint32(0x00FFFFFF)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def tea(v0: np.ndarray, v1: np.ndarray) -> np.ndarray:
    """test_path_tracing.cpp:112-120 (4 rounds)."""
    v0 = v0.astype(np.uint64)
    v1 = v1.astype(np.uint64)
    s0 = np.uint64(0)
    M = np.uint64(0xFFFFFFFF)
    for _ in range(4):
        s0 = (s0 + np.uint64(0x9E3779B9)) & M
        v0 = (v0 + ((((v1 << np.uint64(4)) & M) + np.uint64(0xA341316C)) ^ ((v1 + s0) & M) ^ ((v1 >> np.uint64(5)) + np.uint64(0xC8013EA4)))) & M
        v1 = (v1 + ((((v0 << np.uint64(4)) & M) + np.uint64(0xAD90777D)) ^ ((v0 + s0) & M) ^ ((v0 >> np.uint64(5)) + np.uint64(0x7E95761E)))) & M
    return v0.astype(np.uint32)


class PCG32:
    """Vectorised PCG32 (XSH-RR 64/32): one independent stream per element."""

    MULT = np.uint64(6364136223846793005)

    def __init__(self, seed: int, streams: np.ndarray | int):
        streams = np.atleast_1d(np.asarray(streams, dtype=np.uint64))
        self.inc = (streams << np.uint64(1)) | np.uint64(1)
        self.state = np.zeros_like(self.inc)
        self.next_u32()
        with np.errstate(over="ignore"):
            self.state = self.state + np.uint64(seed)
        self.next_u32()

    def next_u32(self) -> np.ndarray:
        old = self.state
        with np.errstate(over="ignore"):
            self.state = old * self.MULT + self.inc
        xorshifted = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
        rot = (old >> np.uint64(59)).astype(np.uint32)
        return (xorshifted >> rot) | (xorshifted << ((np.uint32(32) - rot) & np.uint32(31)))

    def next_f32(self) -> np.ndarray:
        return (self.next_u32() >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def pcg_block(seed: int, n: int, k: int) -> np.ndarray:
    """n x k floats in [0,1): element i uses PCG32(seed, stream=i)."""
    g = PCG32(seed, np.arange(n, dtype=np.uint64))
    return np.stack([g.next_f32() for _ in range(k)], axis=1)


# ----------------------------------------------------------------------------- meshes
def icosphere(level: int, displace: bool = True, phase: float = 0.0, radius: float = 1.0):
    """Returns (verts[N,4] float32 — stride 16, tris[M,3] uint32). Level 8 = 1 310 720 triangles."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        key = e[:, 0] * (len(v) + 1) + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a = uniq // (len(v) + 1)
        b = uniq % (len(v) + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        nf = len(f)
        m01, m12, m20 = base + inv[:nf], base + inv[nf:2 * nf], base + inv[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    if displace:
        n = v.copy()
        d = 0.05 * np.sin(13 * v[:, 0] + phase) * np.sin(17 * v[:, 1] + phase) * np.sin(19 * v[:, 2] + phase)
        v = v + n * d[:, None]
    v = v * radius
    verts = np.zeros((len(v), 4), dtype=np.float32)
    verts[:, :3] = v.astype(np.float32)
    return verts, f.astype(np.uint32)


def test_rtx_scene():
    """C1: src/tests/test_rtx.cpp:24-28 (vertices), :101-105 (instance transforms)."""
    verts = np.zeros((3, 4), dtype=np.float32)
    verts[:, :3] = [(-0.5, -0.5, 0.0), (0.5, -0.5, 0.0), (0.0, 0.5, 0.0)]
    tris = np.array([[0, 1, 2]], dtype=np.uint32)
    t0 = affine_from_4x4(mat_scaling(1.5))
    t1 = affine_from_4x4(mat_translation(-0.25, 0.0, 0.1) @ mat_rotation((0, 0, 1), 0.5))
    return verts, tris, [t0, t1]


def test_rtx_rays(w: int = 1024, h: int = 1024) -> np.ndarray:
    """C1 rays: test_rtx.cpp:67-73 with the jitter replaced by +0.5."""
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    rays = np.zeros(w * h, dtype=RAY_DTYPE)
    px = ((x.reshape(-1) + 0.5) / w * 2.0 - 1.0).astype(np.float32)
    py = (-((y.reshape(-1) + 0.5) / h * 2.0 - 1.0)).astype(np.float32)
    rays["o"][:, 0], rays["o"][:, 1], rays["o"][:, 2] = px, py, 1.0
    rays["d"][:, 2] = -1.0
    rays["tmin"], rays["tmax"] = 0.0, FLT_MAX
    return rays


def pinhole_rays(w: int, h: int, pos=(0.0, 0.0, 3.0), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), vfov_deg: float = 40.0) -> np.ndarray:
    pos, look_at, up = (np.asarray(a, dtype=np.float64) for a in (pos, look_at, up))
    fwd = look_at - pos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    th = np.tan(np.radians(vfov_deg) / 2.0)
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    sx = ((x.reshape(-1) + 0.5) / w * 2.0 - 1.0) * th * (w / h)
    sy = -((y.reshape(-1) + 0.5) / h * 2.0 - 1.0) * th
    d = fwd[None, :] + sx[:, None] * right[None, :] + sy[:, None] * upv[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(w * h, dtype=RAY_DTYPE)
    rays["o"] = pos.astype(np.float32)[None, :]
    rays["d"] = d.astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.0, FLT_MAX
    return rays


def offset_ray_origin(p: np.ndarray, n: np.ndarray) -> np.ndarray:
    """src/dsl/rtx/ray.cpp:16-23 (Wächter & Binder, 'A fast and robust method for avoiding self-intersection')."""
    origin, float_scale, int_scale = np.float32(1.0 / 32.0), np.float32(1.0 / 65536.0), np.float32(256.0)
    p = p.astype(np.float32)
    n = n.astype(np.float32)
    of_i = (int_scale * n).astype(np.int32)
    p_i_bits = p.view(np.int32) + np.where(p < 0, -of_i, of_i)
    p_i = p_i_bits.view(np.float32)
    return np.where(np.abs(p) < origin, p + float_scale * n, p_i).astype(np.float32)


def bounce_rays(primary: np.ndarray, hits: np.ndarray, verts: np.ndarray, tris: np.ndarray, w: int, h: int) -> np.ndarray:
    """C2 incoherent batch: one cosine-hemisphere bounce per primary hit (pixel order, no sorting); primaries that
    miss are replaced by rays from a uniform point on the radius-2 sphere toward a uniform point in the unit ball."""
    n = len(primary)
    y, x = np.meshgrid(np.arange(h, dtype=np.uint32), np.arange(w, dtype=np.uint32), indexing="ij")
    state = tea(x.reshape(-1), y.reshape(-1))
    state, u1 = lcg_next(state)
    state, u2 = lcg_next(state)
    state, u3 = lcg_next(state)
    state, u4 = lcg_next(state)
    state, u5 = lcg_next(state)
    hit = hits["inst"] != 0xFFFFFFFF
    prim = np.where(hit, hits["prim"], 0).astype(np.int64)
    v0 = verts[tris[prim, 0], :3].astype(np.float64)
    v1 = verts[tris[prim, 1], :3].astype(np.float64)
    v2 = verts[tris[prim, 2], :3].astype(np.float64)
    bu = hits["bary"][:, 0].astype(np.float64)[:, None]
    bv = hits["bary"][:, 1].astype(np.float64)[:, None]
    p = (1 - bu - bv) * v0 + bu * v1 + bv * v2
    ng = np.cross(v1 - v0, v2 - v0)
    ng /= np.maximum(np.linalg.norm(ng, axis=1, keepdims=True), 1e-30)
    wo = -primary["d"].astype(np.float64)
    ng = np.where((np.sum(ng * wo, axis=1) < 0)[:, None], -ng, ng)
    # cosine-weighted hemisphere about ng
    r = np.sqrt(u1.astype(np.float64))
    phi = 2 * np.pi * u2.astype(np.float64)
    lx, ly = r * np.cos(phi), r * np.sin(phi)
    lz = np.sqrt(np.maximum(0.0, 1.0 - lx * lx - ly * ly))
    sgn = np.where(ng[:, 2] >= 0, 1.0, -1.0)
    a = -1.0 / (sgn + ng[:, 2])
    b = ng[:, 0] * ng[:, 1] * a
    tx = np.stack([1 + sgn * ng[:, 0] ** 2 * a, sgn * b, -sgn * ng[:, 0]], 1)
    ty = np.stack([b, sgn + ng[:, 1] ** 2 * a, -ng[:, 1]], 1)
    d = lx[:, None] * tx + ly[:, None] * ty + lz[:, None] * ng
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = offset_ray_origin(p.astype(np.float32), ng.astype(np.float32))
    # replacement rays for misses
    z = 1 - 2 * u1.astype(np.float64)
    rr = np.sqrt(np.maximum(0.0, 1 - z * z))
    so = 2.0 * np.stack([rr * np.cos(phi), rr * np.sin(phi), z], 1)
    zz = 1 - 2 * u3.astype(np.float64)
    r2 = np.sqrt(np.maximum(0.0, 1 - zz * zz))
    ph2 = 2 * np.pi * u4.astype(np.float64)
    tgt = np.cbrt(u5.astype(np.float64))[:, None] * np.stack([r2 * np.cos(ph2), r2 * np.sin(ph2), zz], 1)
    sd = tgt - so
    sd /= np.linalg.norm(sd, axis=1, keepdims=True)
    out = np.zeros(n, dtype=RAY_DTYPE)
    out["o"] = np.where(hit[:, None], o, so.astype(np.float32))
    out["d"] = np.where(hit[:, None], d.astype(np.float32), sd.astype(np.float32))
    out["tmin"], out["tmax"] = 0.0, FLT_MAX
    return out


def triangle_soup(n: int, seed: int = 1234, s: float = 0.004):
    """C4: centre ~ U[0,1]^3, two edge vectors ~ U[-s,s]^3; unindexed, stride-16 vertices."""
    r = pcg_block(seed, n, 9)
    c = r[:, 0:3]
    e1 = (r[:, 3:6] * 2 - 1) * np.float32(s)
    e2 = (r[:, 6:9] * 2 - 1) * np.float32(s)
    verts = np.zeros((3 * n, 4), dtype=np.float32)
    verts[0::3, :3] = c
    verts[1::3, :3] = c + e1
    verts[2::3, :3] = c + e2
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    return verts, tris


def uniform_rays(n: int, seed: int = 5678, stream_offset: int = 0) -> np.ndarray:
    """C4 rays: origin ~ U[0,1]^3, direction uniform on S^2."""
    g = PCG32(seed, np.arange(stream_offset, stream_offset + n, dtype=np.uint64))
    o = np.stack([g.next_f32() for _ in range(3)], 1)
    z = 1 - 2 * g.next_f32().astype(np.float64)
    ph = 2 * np.pi * g.next_f32().astype(np.float64)
    r = np.sqrt(np.maximum(0.0, 1 - z * z))
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = o
    rays["d"] = np.stack([r * np.cos(ph), r * np.sin(ph), z], 1).astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.0, FLT_MAX
    return rays


def instanced_grid(n_inst: int = 100_000, n_blas: int = 16, seed: int = 42, grid: int = 47):
    """C3: returns (affines[n,12], blas_index[n], masks[n], user_ids[n])."""
    r = pcg_block(seed, n_inst, 7)
    i = np.arange(n_inst)
    cell = np.stack([i % grid, (i // grid) % grid, i // (grid * grid)], 1).astype(np.float64)
    pos = cell - (grid - 1) / 2.0 + (r[:, 0:3].astype(np.float64) * 2 - 1) * 0.25
    # uniform unit quaternion (Shoemake) from three uniforms
    u1, u2, u3 = (r[:, k].astype(np.float64) for k in (3, 4, 5))
    q = np.stack([np.sqrt(1 - u1) * np.sin(2 * np.pi * u2), np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
                  np.sqrt(u1) * np.sin(2 * np.pi * u3), np.sqrt(u1) * np.cos(2 * np.pi * u3)], 1)
    x, y, z, w = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)
    A = np.zeros((n_inst, 3, 4), dtype=np.float64)
    A[:, :, :3] = R * 0.3
    A[:, :, 3] = pos
    return (A.reshape(n_inst, 12).astype(np.float32), (i % n_blas).astype(np.int64),
            (1 << (i & 7)).astype(np.uint32), i.astype(np.uint32))


def shadow_rays(primary: np.ndarray, hits: np.ndarray, light=(60.0, 80.0, 100.0)) -> np.ndarray:
    """C3 any-hit batch: from each primary hit point toward a point light, tmax = distance (direction normalised)."""
    hit = hits["inst"] != 0xFFFFFFFF
    t = np.where(hit, hits["t"], 1.0).astype(np.float64)
    p = primary["o"].astype(np.float64) + primary["d"].astype(np.float64) * t[:, None]
    L = np.asarray(light, dtype=np.float64)[None, :] - p
    dist = np.linalg.norm(L, axis=1)
    out = np.zeros(len(primary), dtype=RAY_DTYPE)
    out["o"] = p.astype(np.float32)
    out["d"] = (L / dist[:, None]).astype(np.float32)
    out["tmin"] = np.float32(1e-3)
    out["tmax"] = np.where(hit, dist, 0.0).astype(np.float32)
    return out


# ----------------------------------------------------------------------------- Cornell-style room (C5)
def _quad(a, b, c, d):
    return [a, b, c, d], [(0, 1, 2), (0, 2, 3)]


def _box(lo, hi, rot_y=0.0):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    c = 0.5 * (lo + hi)
    corners = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                        [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
    R = mat_rotation((0, 1, 0), rot_y)[:3, :3]
    corners = (corners - c) @ R.T + c
    faces = [(0, 1, 2, 3), (5, 4, 7, 6), (4, 0, 3, 7), (1, 5, 6, 2), (3, 2, 6, 7), (4, 5, 1, 0)]
    tris = []
    for f in faces:
        tris += [(f[0], f[1], f[2]), (f[0], f[2], f[3])]
    return corners.tolist(), tris


def cornell_room():
    """Eight meshes / 36 triangles in the spirit of the classic Cornell box (unit scale, room spans [-1,1]^3):
    floor, ceiling, back wall, left wall, right wall, short box, tall box, area light.
    Returns a list of (verts[N,4] float32, tris[M,3] uint32)."""
    meshes = []

    def add(vs, ts):
        v = np.zeros((len(vs), 4), dtype=np.float32)
        v[:, :3] = np.asarray(vs, dtype=np.float32)
        meshes.append((v, np.asarray(ts, dtype=np.uint32)))

    add(*_quad((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)))       # floor
    add(*_quad((-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1)))           # ceiling
    add(*_quad((-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1)))       # back wall
    add(*_quad((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1)))       # left wall
    add(*_quad((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1)))           # right wall
    add(*_box((0.1, -1.0, 0.0), (0.7, -0.4, 0.6), rot_y=-0.3))            # short box
    add(*_box((-0.7, -1.0, -0.7), (-0.1, 0.2, -0.1), rot_y=0.35))         # tall box
    add(*_quad((-0.25, 0.995, -0.25), (-0.25, 0.995, 0.25), (0.25, 0.995, 0.25), (0.25, 0.995, -0.25)))  # light
    assert sum(len(t) for _, t in meshes) == 36
    return meshes


def lcg_f32_scalar(state: int):
    """test_procedural.cpp:9-15 style scalar lcg (same constants as lcg_next)."""
    state = (1664525 * state + 1013904223) & 0xFFFFFFFF
    return state, np.float32((state & 0x00FFFFFF) / 16777216.0)


def sphere_aabbs(n: int = 64, radius: float = 0.08, lo=(-0.85, -0.85, -0.85), hi=(0.85, 0.85, 0.85), seed: int = 0) -> np.ndarray:
    """C5 procedural primitive: n AABBs (24 B each) of spheres with centres from the scalar lcg (seed 0)."""
    state = seed
    out = np.zeros((n, 6), dtype=np.float32)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    for i in range(n):
        c = np.zeros(3, dtype=np.float32)
        for a in range(3):
            state, f = lcg_f32_scalar(state)
            c[a] = lo[a] + (hi[a] - lo[a]) * f
        out[i, :3] = c - np.float32(radius)
        out[i, 3:] = c + np.float32(radius)
    return out
