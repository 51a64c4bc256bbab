// lcrt_build.h — per-thread builder functions (host+device portable): primitive bounds, Morton codes, agglomerative
// LBVH (Apetrei 2014) with SAH leaf marking, collapse to the compressed BVH8 layout, refit.
//
// Replaces optixAccelBuild / optixAccelCompact as called by the reference
// (src/backends/cuda/cuda_primitive.cpp:20-132, src/backends/cuda/cuda_accel.cpp:26-142).
// The __global__ wrappers live in lcrt_kernels.cu; tests/hostsim runs the same functions serially on the CPU.
#pragma once
#include "lcrt_common.h"

namespace lcrt {

#if defined(__CUDA_ARCH__)
#define LCRT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define LCRT_FENCE() __threadfence()
// atomic add that releases this thread's earlier writes and acquires those released by earlier arrivals at the same
// counter (one instruction instead of fence.sc / atom / fence.sc)
__device__ __forceinline__ uint32_t lcrt_atomic_add_acq_rel(uint32_t *p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
#define LCRT_ATOMIC_ADD_ACQ_REL(p, v) lcrt_atomic_add_acq_rel((p), (v))
// atomicAdd on a counter that every thread of the grid hits: the lanes that arrive together add up first and send ONE atomic
// (same-address atomics are served one at a time by a single L2 slice)
__device__ __forceinline__ uint32_t lcrt_atomic_add_aggregated(uint32_t *p, uint32_t v) {
    const uint32_t m = __activemask(), lane = threadIdx.x & 31u;
    uint32_t before = 0u, total = 0u;
    for (uint32_t mm = m; mm; mm &= mm - 1u) {
        const int b = __ffs((int)mm) - 1;
        const uint32_t x = __shfl_sync(m, v, b);
        before += (uint32_t)b < lane ? x : 0u;
        total += x;
    }
    const int leader = __ffs((int)m) - 1;
    uint32_t base = 0u;
    if ((int)lane == leader) base = atomicAdd(p, total);
    return __shfl_sync(m, base, leader) + before;
}
#define LCRT_ATOMIC_ADD_AGG(p, v) lcrt_atomic_add_aggregated((p), (v))
#define LCRT_VOLATILE_LOAD_U32(p) (*reinterpret_cast<const volatile uint32_t *>(p))
// data produced by other SMs inside the same kernel must not be served from this SM's (non-coherent) L1
#define LCRT_LOAD_CG_F4(p) __ldcg(reinterpret_cast<const float4 *>(p))
#else
inline uint32_t host_atomic_add(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p += v;
    return o;
}
#define LCRT_ATOMIC_ADD(p, v) ::lcrt::host_atomic_add((p), (v))
#define LCRT_ATOMIC_ADD_ACQ_REL(p, v) ::lcrt::host_atomic_add((p), (v))
#define LCRT_ATOMIC_ADD_AGG(p, v) ::lcrt::host_atomic_add((p), (v))
#define LCRT_FENCE() ((void)0)
#define LCRT_VOLATILE_LOAD_U32(p) (*(p))
#define LCRT_LOAD_CG_F4(p) (*(p))
#endif

constexpr uint32_t INVALID_U32 = 0xffffffffu;
constexpr uint32_t LEAF_MARK = 0x80000000u;

// ------------------------------------------------------------------------------------------ build inputs
struct TriInput {
    const char *verts;
    uint64_t stride;
    const uint32_t *tris;
    uint32_t n;
    static constexpr uint32_t kind = KIND_TRIANGLES;
    static constexpr uint32_t rec_bytes = sizeof(TriRecord);
    static constexpr uint32_t max_leaf = 3;
    LCRT_HD void fetch(uint32_t i, float v[9]) const {
        const uint32_t i0 = tris[3 * (uint64_t)i], i1 = tris[3 * (uint64_t)i + 1], i2 = tris[3 * (uint64_t)i + 2];
        const float *a = reinterpret_cast<const float *>(verts + i0 * stride);
        const float *b = reinterpret_cast<const float *>(verts + i1 * stride);
        const float *c = reinterpret_cast<const float *>(verts + i2 * stride);
        v[0] = a[0], v[1] = a[1], v[2] = a[2], v[3] = b[0], v[4] = b[1], v[5] = b[2], v[6] = c[0], v[7] = c[1], v[8] = c[2];
    }
    LCRT_HD Box box(uint32_t i) const {
        float v[9];
        fetch(i, v);
        Box b = box_empty();
        box_grow(b, v[0], v[1], v[2]), box_grow(b, v[3], v[4], v[5]), box_grow(b, v[6], v[7], v[8]);
        return b;
    }
    LCRT_HD void write(char *dst, uint32_t i) const {
        float v[9];
        fetch(i, v);
        f4 *d = reinterpret_cast<f4 *>(dst);
        d[0] = make_f4(v[0], v[1], v[2], as_float(i));
        d[1] = make_f4(v[3], v[4], v[5], 0.f);
        d[2] = make_f4(v[6], v[7], v[8], 0.f);
    }
    LCRT_HD static uint32_t record_prim(const char *rec) { return as_uint(reinterpret_cast<const f4 *>(rec)->w); }
};

struct AabbInput {
    const float *aabbs;// 24-byte records (include/luisa/runtime/rtx/aabb.h:7-10)
    uint32_t n;
    static constexpr uint32_t kind = KIND_AABBS;
    static constexpr uint32_t rec_bytes = sizeof(BoxRecord);
    static constexpr uint32_t max_leaf = 2;
    LCRT_HD Box box(uint32_t i) const {
        const float *p = aabbs + 6 * (uint64_t)i;
        Box b;
        for (int a = 0; a < 3; a++) b.lo[a] = p[a], b.hi[a] = p[3 + a];
        return b;
    }
    LCRT_HD void write(char *dst, uint32_t i) const {
        const float *p = aabbs + 6 * (uint64_t)i;
        f4 *d = reinterpret_cast<f4 *>(dst);
        d[0] = make_f4(p[0], p[1], p[2], as_float(i));
        d[1] = make_f4(p[3], p[4], p[5], 0.f);
    }
    LCRT_HD static uint32_t record_prim(const char *rec) { return as_uint(reinterpret_cast<const f4 *>(rec)->w); }
};

LCRT_HD bool invert_affine(const float *m, double *inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0 || det != det) return false;
    const double id = 1.0 / det;
    inv[0] = (e * i - f * h) * id, inv[1] = (c * h - b * i) * id, inv[2] = (b * f - c * e) * id;
    inv[4] = (f * g - d * i) * id, inv[5] = (a * i - c * g) * id, inv[6] = (c * d - a * f) * id;
    inv[8] = (d * h - e * g) * id, inv[9] = (b * g - a * h) * id, inv[10] = (a * e - b * d) * id;
    inv[3] = -(inv[0] * m[3] + inv[1] * m[7] + inv[2] * m[11]);
    inv[7] = -(inv[4] * m[3] + inv[5] * m[7] + inv[6] * m[11]);
    inv[11] = -(inv[8] * m[3] + inv[9] * m[7] + inv[10] * m[11]);
    return true;
}

struct InstInput {
    const Instance *inst;
    uint32_t n;
    static constexpr uint32_t kind = KIND_INSTANCES;
    static constexpr uint32_t rec_bytes = sizeof(InstanceRecord);
    static constexpr uint32_t max_leaf = 1;
    LCRT_HD static const BlobHeader *blas_of(const Instance &in) {
        const BlobHeader *h = reinterpret_cast<const BlobHeader *>(in.handle);
        if (h == nullptr || h->magic != BLOB_MAGIC || !h->built || h->kind == KIND_INSTANCES) return nullptr;
        return h;
    }
    LCRT_HD Box box(uint32_t i) const {
        const Instance in = inst[i];
        const BlobHeader *h = blas_of(in);
        Box b = box_empty();
        if (h == nullptr || !(h->lo[0] <= h->hi[0])) return b;
        double tmp[12];
        if (!invert_affine(in.affine, tmp)) return b;// singular transform: the instance can not be hit
        const float *m = in.affine;
        for (int c = 0; c < 8; c++) {
            const double x = (c & 1) ? h->hi[0] : h->lo[0], y = (c & 2) ? h->hi[1] : h->lo[1], z = (c & 4) ? h->hi[2] : h->lo[2];
            const double wx = m[0] * x + m[1] * y + m[2] * z + m[3], wy = m[4] * x + m[5] * y + m[6] * z + m[7],
                         wz = m[8] * x + m[9] * y + m[10] * z + m[11];
            box_grow(b, (float)wx, (float)wy, (float)wz);
        }
        // pad: the ray is moved to object space with an fp32 inverse, so the world box must cover that rounding
        for (int a = 0; a < 3; a++) {
            const float e = 2e-6f * (fmaxf(fabsf(b.lo[a]), fabsf(b.hi[a])) + (b.hi[a] - b.lo[a])) + 1e-30f;
            b.lo[a] -= e, b.hi[a] += e;
        }
        return b;
    }
    LCRT_HD void write(char *dst, uint32_t i) const {
        const Instance in = inst[i];
        const BlobHeader *h = blas_of(in);
        InstanceRecord r;
        memset(&r, 0, sizeof(r));
        if (!invert_affine(in.affine, r.inv_exact)) {
            for (int k = 0; k < 12; k++) r.inv_exact[k] = 0.0;
        }
        for (int k = 0; k < 12; k++) r.inv[k] = (float)r.inv_exact[k];
#if LCRT_RECENTER
        for (int a = 0; a < 3; a++) memcpy(&r.spare[a], &in.affine[4 * a + 3], 4);
#endif
        bool identity = true;
        for (int k = 0; k < 12; k++) identity = identity && in.affine[k] == ((k == 0 || k == 5 || k == 10) ? 1.f : 0.f);
        r.nodes = h ? (uint64_t)(reinterpret_cast<const char *>(h) + h->nodes_offset) : 0ull;
        r.prims = h ? (uint64_t)(reinterpret_cast<const char *>(h) + h->prims_offset) : 0ull;
        r.inst = i;
        r.mask = h ? (in.mask & 0xffu) : 0u;// an unbound instance can never be entered
        const bool opaque = !(in.flags & INSTANCE_FLAG_ENFORCE_ANYHIT);
        r.flags = (h ? h->kind : KIND_TRIANGLES) | (opaque ? IREC_OPAQUE : 0u) | (identity ? IREC_IDENTITY : 0u);
        copy_vec16(dst, &r, 16);
    }
    LCRT_HD static uint32_t record_prim(const char *rec) { return reinterpret_cast<const InstanceRecord *>(rec)->inst; }
};

// ------------------------------------------------------------------------------------------ scene bounds / Morton
// order-preserving float <-> uint map so that atomicMin / atomicMax on uints reduce floats
LCRT_HD uint32_t float_to_ordered(float f) {
    const uint32_t u = as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
LCRT_HD float ordered_to_float(uint32_t u) { return as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

LCRT_HD uint64_t expand_bits_21(uint64_t v) {// spread the low 21 bits to every third bit
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// Morton code (3 x `bits`, bits <= 21) of the box centre relative to the centroid bounds [clo, chi]
LCRT_HD uint64_t morton63(const Box &b, const float clo[3], const float inv_ext[3], uint32_t bits) {
    uint64_t code = 0;
    const float scale = (float)(1u << bits);
    const uint32_t qmax = (1u << bits) - 1u;
    for (int a = 0; a < 3; a++) {
        const float c = 0.5f * (b.lo[a] + b.hi[a]);
        float f = (c - clo[a]) * inv_ext[a];// [0, 1]
        if (!(f >= 0.f)) f = 0.f;           // also catches NaN (empty boxes)
        if (f > 1.f) f = 1.f;
        uint32_t q = (uint32_t)(f * scale);
        if (q > qmax) q = qmax;
        code |= expand_bits_21(q) << a;
    }
    return code;
}

// Sort key = (Morton code << index_bits) | primitive index: ONE 64-bit word per primitive goes through the radix sort (no
// payload array), keys are unique by construction (duplicate Morton codes split by index, which is what the LBVH tie-break
// did explicitly before) and the sorted order IS the key's low bits.  Only the Morton bits are sorted: the keys start in
// index order and the sort is stable.
LCRT_HD uint32_t index_bits_for(uint32_t n) {
    uint32_t lg = 1;
    while ((1ull << lg) < (uint64_t)n) lg++;
    return lg;
}
// Morton bits per axis: enough to resolve 64 cells per primitive (per axis) of a uniform distribution, at most 21 and at
// most what the key has room for next to the index.  The number of 8-bit radix passes follows: ceil(3 * bits / 8) —
// 5 passes at 1.3 M primitives (13 bits per axis), 5 at 50 M (12 bits per axis).
#ifndef LCRT_MORTON_EXTRA_BITS
#define LCRT_MORTON_EXTRA_BITS 6u
#endif
LCRT_HD uint32_t morton_bits_for(uint32_t n) {
    uint32_t lg = 0;
    while ((1ull << lg) < (uint64_t)n) lg++;
    uint32_t b = (lg + 2u) / 3u + LCRT_MORTON_EXTRA_BITS;
    if (b < 10u) b = 10u;
    if (b > 21u) b = 21u;
    const uint32_t room = (64u - index_bits_for(n)) / 3u;
    if (b > room) b = room;
    return b;
}

// ------------------------------------------------------------------------------------------ binary hierarchy
// Node ids: internal j in [0, n-1), leaf i -> (n-1) + i.  Everything known about a node lives in its PARENT's record, one
// 48-byte half per child: opening a child during the collapse is ONE dependent load (both grandchildren with their
// boxes, leaf marks, leaf-slot counts and ranges arrive together), and a leaf-marked subtree's primitives are the
// sorted positions first .. first + count - 1, so nothing below it is ever traversed.
struct alignas(16) Bvh2Half {
    float lo[3];
    float cost;          // SAH cost of the subtree
    float hi[3];
    uint32_t count_mark; // primitives below | LEAF_MARK if the subtree is collapsed into one leaf slot
    uint32_t id;         // node id of this child
    uint32_t slots;      // leaf slots below it (1 for a leaf-marked subtree)
    uint32_t first, last;// range of sorted positions below it
};
static_assert(sizeof(Bvh2Half) == 48, "Bvh2Half must be 48 bytes");
struct alignas(16) Bvh2Node {
    Bvh2Half l, r;
};
static_assert(sizeof(Bvh2Node) == 96, "Bvh2Node must be 96 bytes");

struct Bvh2 {
    uint32_t n;
    uint32_t index_bits;   // low bits of a key = input primitive
    const uint64_t *keys;  // sorted keys
    Bvh2Node *node;        // [n-1]
    uint32_t *flag;        // [n-1]  zero-initialised arrival counters
    Bvh2Half *root;        // [1]    the root described like a child
    float ci;              // cost of one binary split relative to one primitive test
    uint32_t max_leaf;
    LCRT_HD uint32_t order(uint32_t pos) const { return (uint32_t)(keys[pos] & ((1ull << index_bits) - 1ull)); }
};

LCRT_HD bool delta_less(const uint64_t *keys, uint32_t a, uint32_t b) {// delta(a) < delta(b); delta(j) between j and j+1
    const uint64_t xa = keys[a] ^ keys[a + 1], xb = keys[b] ^ keys[b + 1];// keys are unique: no tie-break needed
    return xa < xb;
}

LCRT_HD float area_of(const float *lo, const float *hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}

LCRT_HD void store_half(Bvh2Half *dst, const Bvh2Half &h) { copy_vec16(dst, &h, 3); }
LCRT_HD Bvh2Half load_half_cg(const Bvh2Half *src) {// written by another SM inside the same kernel: not from this SM's L1
    Bvh2Half h;
#if defined(__CUDA_ARCH__)
    const float4 *s = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(&h);
    d[0] = __ldcg(s), d[1] = __ldcg(s + 1), d[2] = __ldcg(s + 2);
#else
    h = *src;
#endif
    return h;
}

// merges the sibling half `s` into `h` (which becomes the description of their parent `parent`)
LCRT_HD void merge_halves(const Bvh2 &t, Bvh2Half &h, const Bvh2Half &s, uint32_t parent, uint32_t left, uint32_t right) {
    for (int a = 0; a < 3; a++) h.lo[a] = fminf(h.lo[a], s.lo[a]), h.hi[a] = fmaxf(h.hi[a], s.hi[a]);
    const uint32_t count = (h.count_mark & ~LEAF_MARK) + (s.count_mark & ~LEAF_MARK);
    const float area = area_of(h.lo, h.hi);
    const float c_split = t.ci * area + h.cost + s.cost;
    const float c_leaf = area * (float)count;
    if (count <= t.max_leaf && c_leaf <= c_split) h.cost = c_leaf, h.count_mark = count | LEAF_MARK, h.slots = 1u;
    else h.cost = c_split, h.count_mark = count, h.slots = h.slots + s.slots;
    h.id = parent, h.first = left, h.last = right;
}

LCRT_HD Bvh2Half leaf_half(uint32_t n, uint32_t i, const Box &leaf_box) {
    Bvh2Half h;
    for (int a = 0; a < 3; a++) h.lo[a] = leaf_box.lo[a], h.hi[a] = leaf_box.hi[a];
    h.cost = area_of(h.lo, h.hi), h.count_mark = 1u | LEAF_MARK, h.id = (n - 1u) + i, h.slots = 1u, h.first = i, h.last = i;
    return h;
}

// Apetrei 2014: a thread that holds the finished subtree [left, right] (described by `h`) climbs while it is the second child
// to arrive at the parent; the halves and the arrival counters live in global memory.
LCRT_HD void lbvh_climb_from(const Bvh2 &t, uint32_t left, uint32_t right, Bvh2Half h) {
    const uint32_t n = t.n;
    for (;;) {
        if (left == 0u && right == n - 1u) {
            store_half(t.root, h);
            return;
        }
        uint32_t parent;
        bool from_left;
        if (left == 0u || (right != n - 1u && delta_less(t.keys, right, left - 1u))) parent = right, from_left = true;
        else parent = left - 1u, from_left = false;
        store_half(from_left ? &t.node[parent].l : &t.node[parent].r, h);
        if (LCRT_ATOMIC_ADD_ACQ_REL(&t.flag[parent], 1u) == 0u) return;// the sibling will finish this node
        // second arrival: merge with the sibling
        const Bvh2Half s = load_half_cg(from_left ? &t.node[parent].r : &t.node[parent].l);
        if (from_left) right = s.last;
        else left = s.first;
        merge_halves(t, h, s, parent, left, right);
    }
}

// one thread per leaf
LCRT_HD void lbvh_climb(const Bvh2 &t, uint32_t i, const Box &leaf_box) { lbvh_climb_from(t, i, i, leaf_half(t.n, i, leaf_box)); }

// ------------------------------------------------------------------------------------------ BVH8 emission
struct Bvh8Out {
    BlobHeader *header;
    Node8 *nodes;
    char *prims;
    uint32_t *parent;     // optional (allow_update): (parent << 3) | slot
    u2 *task;             // [node_capacity] { binary node id, depth of the wide node } per wide node, x == INVALID until scheduled
    uint32_t *rec_src;    // [prim_count] sorted position whose primitive goes into record j (records are written by a
                          // separate, fully parallel pass: emit_record)
    uint32_t *counters;   // nodes allocated, prims allocated, prims finished — at [0], [COUNTER_STRIDE], [2 * COUNTER_STRIDE]
    uint32_t node_capacity;
};
constexpr uint32_t COUNTER_STRIDE = 32;// words: every counter on its own 128-byte line

// Quantisation of child boxes onto the node's 8-bit grid: origin p = node lo, per-axis step 2^(e-127) with e the
// smallest exponent for which 255 steps span the node, child lo rounded down and hi rounded up.  The host evaluates
// the differences exactly in double; the device uses fp32 with directed rounding — floor(fl_down(lo - p) * 2^-k) equals
// floor((lo - p) * 2^-k) because every grid line n * 2^k is itself a float (DESIGN.md "quantisation"), so both produce
// the same bytes.
LCRT_HD uint32_t quant_exponent(float lo, float hi) {// biased exponent e, 0 for an empty / degenerate extent
#if defined(__CUDA_ARCH__)
    const float ext = __fsub_ru(hi, lo);
#else
    const double ext = (double)hi - (double)lo;
#endif
    if (!(ext > 0)) return 0u;
    int ex;
#if defined(__CUDA_ARCH__)
    const float m = frexpf(ext, &ex);// ext = m * 2^ex, m in [0.5, 1)
    int ee = (m <= 0.99609375f ? ex - 8 : ex - 7) + 127;// 255 * 2^(e-127) >= ext  <=>  m * 2^(ex-k) <= 255/256 * 2^8
#else
    const double m = frexp(ext, &ex);
    int ee = (m <= 0.99609375 ? ex - 8 : ex - 7) + 127;
#endif
    if (ee < 1) ee = 1;
    if (ee > 254) ee = 254;
    return (uint32_t)ee;
}
LCRT_HD uint32_t quant_down(float x, float p, uint32_t e) {
    if (e == 0u) return 0u;
#if defined(__CUDA_ARCH__)
    float q = floorf(__fsub_rd(x, p) * as_float((254u - e) << 23));// 2^-(e-127), exact scaling
#else
    double q = floor(((double)x - (double)p) * ldexp(1.0, 127 - (int)e));
#endif
    if (!(q >= 0)) q = 0;
    if (q > 255) q = 255;
    return (uint32_t)q;
}
LCRT_HD uint32_t quant_up(float x, float p, uint32_t e) {
    if (e == 0u) return 0u;
#if defined(__CUDA_ARCH__)
    float q = ceilf(__fsub_ru(x, p) * as_float((254u - e) << 23));
#else
    double q = ceil(((double)x - (double)p) * ldexp(1.0, 127 - (int)e));
#endif
    if (!(q >= 0)) q = 0;
    if (q > 255) q = 255;
    return (uint32_t)q;
}

LCRT_HD void quantize_node(Node8 &nd, const Box &nb, const Box *cb, const bool *valid) {
    float p[3];
    uint32_t eb[3];
    for (int a = 0; a < 3; a++) {
        const bool ok = nb.lo[a] <= nb.hi[a];// false for empty / NaN bounds
        p[a] = ok ? nb.lo[a] : 0.f;
        eb[a] = ok ? quant_exponent(nb.lo[a], nb.hi[a]) : 0u;
    }
    nd.px = p[0], nd.py = p[1], nd.pz = p[2];
    nd.ex = (uint8_t)eb[0], nd.ey = (uint8_t)eb[1], nd.ez = (uint8_t)eb[2];
    uint8_t *qlo[3] = {nd.qlo_x, nd.qlo_y, nd.qlo_z};
    uint8_t *qhi[3] = {nd.qhi_x, nd.qhi_y, nd.qhi_z};
    for (int s = 0; s < 8; s++) {
        for (int a = 0; a < 3; a++) {
            uint32_t ql = 255u, qh = 0u;// inverted: never hit
            if (valid[s] && cb[s].lo[a] <= cb[s].hi[a]) ql = quant_down(cb[s].lo[a], p[a], eb[a]), qh = quant_up(cb[s].hi[a], p[a], eb[a]);
            qlo[a][s] = (uint8_t)ql, qhi[a][s] = (uint8_t)qh;
        }
    }
}

// Builds wide node `widx` (at depth `depth`, root = 1) from the binary node `bid`; for widx == 0 the root comes described as
// a half (Bvh2::root).  Returns false on capacity overflow.  This is the serial statement of the algorithm (tests/hostsim
// runs it); the device runs the same decisions with eight lanes per node (collapse_octet, lcrt_kernels.cu).
template<class Input>
LCRT_HD_NOINLINE bool collapse_node(const Bvh2 &t, const Bvh8Out &o, const Input &in, uint32_t widx, uint32_t bid, uint32_t depth) {
    Bvh2Half c[8];
    float carea[8];
    bool cleaf[8];
    int nc = 0;
    auto set_child = [&](int k, const Bvh2Half &h) {
        c[k] = h;
        carea[k] = area_of(h.lo, h.hi);
        cleaf[k] = (h.count_mark & LEAF_MARK) != 0u;
    };
    if (widx == 0u) {
        const Bvh2Half r = *t.root;
        if (r.count_mark & LEAF_MARK) {
            set_child(0, r), nc = 1;
        } else {
            bid = r.id;
        }
    }
    if (nc == 0) {
        const Bvh2Node nd = t.node[bid];
        set_child(0, nd.l), set_child(1, nd.r), nc = 2;
    }
    // Opening policy.  Every wide node below the top levels should be as full as possible, so the binary subtree is
    // opened by leaf-slot counts (lbvh_climb leaves them in the halves):
    //   * a subtree of <= 8 leaf slots becomes ONE wide node whose children are all leaves;
    //   * otherwise children with more than 8 leaf slots are opened first, largest surface area first (the classic
    //     greedy rule, which is what the upper levels see);
    //   * then, smallest first, children small enough to be dissolved into leaf slots of this node entirely are
    //     dissolved; what is left (<= 8 leaf slots each) becomes full-ish all-leaf nodes one level down.
    auto open_child = [&](int k) {
        const Bvh2Node nd = t.node[c[k].id];
        set_child(k, nd.l), set_child(nc, nd.r);
        nc++;
    };
    uint32_t total_slots = 0;
    for (int k = 0; k < nc; k++) total_slots += c[k].slots;
    if (total_slots <= 8u) {
        // the whole subtree fits into this node (six wide nodes out of seven are of this kind): dissolve it level by level,
        // every internal child of a round at once — left halves stay in place, right halves are appended in child order
        for (;;) {
            uint32_t opening = 0u;
            for (int k = 0; k < nc; k++)
                if (!cleaf[k]) opening |= 1u << k;
            if (!opening) break;
            const int base = nc;
            int rank = 0;
            for (int k = 0; k < base; k++) {
                if (!(opening & (1u << k))) continue;
                const Bvh2Node nd = t.node[c[k].id];
                set_child(k, nd.l), set_child(base + rank, nd.r);
                rank++;
            }
            nc = base + rank;
        }
    } else {
        while (nc < 8) {
            int best = -1;
            float ba = -1.f;
            for (int k = 0; k < nc; k++)
                if (!cleaf[k] && c[k].slots > 8u && carea[k] > ba) ba = carea[k], best = k;
            if (best < 0) break;
            open_child(best);
        }
        for (;;) {
            int best = -1;
            uint32_t bs = 0xffffffffu;
            for (int k = 0; k < nc; k++)
                if (!cleaf[k] && (uint32_t)nc - 1u + c[k].slots <= 8u && c[k].slots < bs) bs = c[k].slots, best = k;
            if (best < 0) break;
            // dissolve the whole subtree of that child: every opening yields pieces that still fit
            uint32_t pending = 1u << best;
            while (pending) {
                const int k = bfind(pending);
                pending &= ~(1u << k);
                if (cleaf[k]) continue;
                open_child(k);
                pending |= (1u << k) | (1u << (nc - 1));
            }
        }
    }
    Box cbox[8];
    Box nb = box_empty();
    for (int k = 0; k < nc; k++) {
        for (int a = 0; a < 3; a++) cbox[k].lo[a] = c[k].lo[a], cbox[k].hi[a] = c[k].hi[a];
        box_grow(nb, cbox[k]);
    }
    // ---- octant-ordered slot assignment (greedy maximum of (centroid - centre) . octant sign)
    float cx = 0.5f * (nb.lo[0] + nb.hi[0]), cy = 0.5f * (nb.lo[1] + nb.hi[1]), cz = 0.5f * (nb.lo[2] + nb.hi[2]);
    if (!(cx == cx)) cx = 0.f;
    if (!(cy == cy)) cy = 0.f;
    if (!(cz == cz)) cz = 0.f;
    int slot_of[8], child_at[8];
    for (int k = 0; k < 8; k++) slot_of[k] = -1, child_at[k] = -1;
    for (int round = 0; round < nc; round++) {
        float bestv = -INFINITY;
        int bk = -1, bs = -1;
        for (int k = 0; k < nc; k++) {
            if (slot_of[k] >= 0) continue;
            float ddx = 0.5f * (cbox[k].lo[0] + cbox[k].hi[0]) - cx, ddy = 0.5f * (cbox[k].lo[1] + cbox[k].hi[1]) - cy,
                  ddz = 0.5f * (cbox[k].lo[2] + cbox[k].hi[2]) - cz;
            if (!(ddx == ddx)) ddx = 0.f;
            if (!(ddy == ddy)) ddy = 0.f;
            if (!(ddz == ddz)) ddz = 0.f;
            for (int s = 0; s < 8; s++) {
                if (child_at[s] >= 0) continue;
                const float v = ((s & 1) ? ddx : -ddx) + ((s & 2) ? ddy : -ddy) + ((s & 4) ? ddz : -ddz);
                if (v > bestv || bk < 0) bestv = v, bk = k, bs = s;
            }
        }
        slot_of[bk] = bs, child_at[bs] = bk;
    }
    // ---- allocate children / primitive records
    uint32_t n_int = 0, n_prims = 0;
    for (int k = 0; k < nc; k++) {
        if (cleaf[k]) n_prims += c[k].count_mark & ~LEAF_MARK;
        else n_int++;
    }
    const uint32_t child_base = n_int ? LCRT_ATOMIC_ADD(&o.counters[0 * COUNTER_STRIDE], n_int) : 0u;
    const uint32_t prim_base = n_prims ? LCRT_ATOMIC_ADD_AGG(&o.counters[1 * COUNTER_STRIDE], n_prims) : 0u;
    if (n_int && child_base + n_int > o.node_capacity) {
        o.header->error = 1u;
        return false;
    }
    Node8 nd;
    nd.imask = 0;
    nd.child_base = child_base, nd.prim_base = prim_base;
    Box sbox[8];
    bool svalid[8];
    uint32_t ci = 0, po = 0;
    for (int s = 0; s < 8; s++) {
        const int k = child_at[s];
        svalid[s] = k >= 0;
        nd.meta[s] = 0;
        if (k < 0) continue;
        sbox[s] = cbox[k];
        if (!cleaf[k]) {
            const uint32_t w = child_base + ci++;
            nd.imask |= (uint8_t)(1u << s);
            nd.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            if (o.parent) o.parent[w] = (widx << 3) | (uint32_t)s;
            o.task[w] = make_u2(c[k].id, depth + 1u);
        } else {
            // a leaf slot holds the primitives at the sorted positions first .. first + count - 1 of its subtree
            const uint32_t cnt = c[k].count_mark & ~LEAF_MARK;
            for (uint32_t q = 0; q < cnt; q++) o.rec_src[prim_base + po + q] = c[k].first + q;
            const uint32_t unary = cnt == 1u ? 1u : cnt == 2u ? 3u : 7u;
            nd.meta[s] = (uint8_t)((unary << 5) | po);
            po += cnt;
        }
    }
    quantize_node(nd, nb, sbox, svalid);
    copy_vec16(&o.nodes[widx], &nd, 5);
    if (widx == 0u) {
        for (int a = 0; a < 3; a++) o.header->lo[a] = nb.lo[a], o.header->hi[a] = nb.hi[a];
        if (o.parent) o.parent[0] = INVALID_U32;
    }
    if (n_prims) LCRT_ATOMIC_ADD_AGG(&o.counters[2 * COUNTER_STRIDE], n_prims);
    return true;
}

// record j of the blob <- primitive order[rec_src[j]] (one thread per record)
template<class Input>
LCRT_HD void emit_record(const Bvh2 &t, const Bvh8Out &o, const Input &in, uint32_t j) {
    in.write(o.prims + (uint64_t)j * Input::rec_bytes, t.order(o.rec_src[j]));
}

// ------------------------------------------------------------------------------------------ refit
struct RefitCtx {
    BlobHeader *header;
    Node8 *nodes;
    char *prims;
    const uint32_t *parent;
    f4 *box_lo, *box_hi;  // [node_count] exact boxes (temp)
    uint32_t *arrived;    // [node_count] zero-initialised (temp)
};

LCRT_HD Box record_box_tri(const char *rec) {
    const f4 *r = reinterpret_cast<const f4 *>(rec);
    Box b = box_empty();
    box_grow(b, r[0].x, r[0].y, r[0].z), box_grow(b, r[1].x, r[1].y, r[1].z), box_grow(b, r[2].x, r[2].y, r[2].z);
    return b;
}

// Recompute one node from its (already rewritten) primitive records and its internal children's exact boxes.
template<class Input>
LCRT_HD void refit_node(const RefitCtx &c, const Input &in, uint32_t widx) {
    Node8 nd = c.nodes[widx];
    Box sbox[8];
    bool svalid[8];
    Box nb = box_empty();
    uint32_t ci = 0;
    for (int s = 0; s < 8; s++) {
        const uint32_t meta = nd.meta[s];
        svalid[s] = meta != 0u;
        if (!meta) continue;
        Box b = box_empty();
        if (nd.imask & (1u << s)) {
            const uint32_t w = nd.child_base + ci++;
            const f4 lo = LCRT_LOAD_CG_F4(&c.box_lo[w]), hi = LCRT_LOAD_CG_F4(&c.box_hi[w]);
            b.lo[0] = lo.x, b.lo[1] = lo.y, b.lo[2] = lo.z, b.hi[0] = hi.x, b.hi[1] = hi.y, b.hi[2] = hi.z;
        } else {
            const uint32_t cnt = popc(meta >> 5), off = meta & 31u;
            for (uint32_t q = 0; q < cnt; q++) {
                const char *rec = c.prims + (uint64_t)(nd.prim_base + off + q) * Input::rec_bytes;
                box_grow(b, in.box(Input::record_prim(rec)));
            }
        }
        sbox[s] = b;
        box_grow(nb, b);
    }
    quantize_node(nd, nb, sbox, svalid);
    copy_vec16(&c.nodes[widx], &nd, 5);
    c.box_lo[widx] = make_f4(nb.lo[0], nb.lo[1], nb.lo[2], 0.f);
    c.box_hi[widx] = make_f4(nb.hi[0], nb.hi[1], nb.hi[2], 0.f);
    if (widx == 0u)
        for (int a = 0; a < 3; a++) c.header->lo[a] = nb.lo[a], c.header->hi[a] = nb.hi[a];
}

// one thread per wide node: nodes without internal children start, the last child to arrive refits the parent
template<class Input>
LCRT_HD void refit_climb(const RefitCtx &c, const Input &in, uint32_t widx) {
    if (c.nodes[widx].imask != 0) return;
    uint32_t cur = widx;
    for (;;) {
        refit_node(c, in, cur);
        if (cur == 0u) return;
        const uint32_t p = c.parent[cur] >> 3;
        const uint32_t need = (uint32_t)popc(c.nodes[p].imask);
        if (LCRT_ATOMIC_ADD_ACQ_REL(&c.arrived[p], 1u) + 1u < need) return;
        cur = p;
    }
}

}// namespace lcrt
