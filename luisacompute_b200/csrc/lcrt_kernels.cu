// lcrt_kernels.cu — sm_100a kernels and the C-ABI (include/lcb200_rt.h) of liblcb200_rt.so.
//
// Build pipeline (all stream-ordered, zero host synchronisation — the reference stalls the host once per compacted
// build, src/backends/cuda/cuda_primitive.cpp:74, cuda_accel.cpp:90):
//   k_build_init -> k_centroid_bounds -> k_morton -> radix sort (lcrt_sort.cuh) -> k_climb (agglomerative LBVH + SAH
//   leaf marking) -> k_collapse (persistent work-queue kernel: BVH2 -> compressed BVH8 + primitive records)
//   -> k_build_finalize.
// Trace: persistent-thread kernels (one resident grid, warp-cooperative ray fetch with __ballot_sync / __shfl_sync
// compaction of idle lanes, short stack in shared memory with local-memory overflow).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/lcb200_rt.h"
#include "lcrt_accel.h"
#include "lcrt_build.h"
#include "lcrt_layout.h"
#include "lcrt_query_programs.h"
#include "lcrt_sort.cuh"
#include "lcrt_traverse.h"

namespace lcrt {

static_assert(SORT_TILE == SORT_TILE_KEYS && SORT_CONTROL_PASSES == SORT_MAX_PASSES, "lcrt_layout.h and lcrt_sort.cuh disagree on the sort tile");

// =====================================================================================================================
// build kernels
// =====================================================================================================================
enum SmallSlots : uint32_t {
    SM_CMIN = 0, SM_CMAX = 3, SM_LEVELS = 15 /* depth of the wide hierarchy */, SM_CHILD_DEPTH = 16 /* TLAS: deepest instanced BLAS */,
    SM_ROOT_HALF = 32 /* Bvh2Half of the root (12 words, 16-byte aligned) */,
    // the words every wide node touches, each on its own 128-byte line (they are served by one L2 slice each; sharing a
    // line with the polled flags made the counters' atomics queue behind the polls)
    SM_COUNTERS = 64 /* nodes, prims, done prims at COUNTER_STRIDE words */, SM_DONE = 160 /* every primitive has its record slot */, SM_ABORT = 192
};
static_assert(SM_COUNTERS + 2 * COUNTER_STRIDE < SM_DONE && SM_ABORT < SMALL_WORDS, "scratch words");
static_assert(SM_ROOT_HALF * 4 % 16 == 0 && SM_ROOT_HALF + sizeof(Bvh2Half) / 4 <= SMALL_WORDS, "scratch words");

__global__ void k_build_init(BlobHeader *h, uint32_t kind, uint32_t n, uint32_t flags, BlobLayout L, uint32_t *small) {
    for (uint32_t k = threadIdx.x; k < SMALL_WORDS; k += blockDim.x) small[k] = 0u;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        BlobHeader hd;
        memset(&hd, 0, sizeof(hd));
        hd.magic = BLOB_MAGIC, hd.kind = kind, hd.prim_count = n, hd.node_capacity = L.node_capacity;
        hd.flags = flags, hd.prim_stride = L.prim_stride;
        hd.plant_magic = PLANT_MAGIC;
        hd.nodes_offset = L.nodes_offset, hd.prims_offset = L.prims_offset, hd.parent_offset = L.parent_offset, hd.total_bytes = L.total_bytes;
        for (int a = 0; a < 3; a++) hd.lo[a] = INFINITY, hd.hi[a] = -INFINITY;
        *h = hd;
        small[SM_CMIN + 0] = small[SM_CMIN + 1] = small[SM_CMIN + 2] = 0xffffffffu;
        small[SM_COUNTERS + 0 * COUNTER_STRIDE] = 1u;// the root
        if (n == 0u) {
            // an empty structure: one node with no children; nothing can be hit
            Node8 nd;
            memset(&nd, 0, sizeof(nd));
            Box nb = box_empty();
            Box sb[8];
            bool sv[8] = {false, false, false, false, false, false, false, false};
            for (int s = 0; s < 8; s++) sb[s] = nb;
            quantize_node(nd, nb, sb, sv);
            Node8 *nodes = reinterpret_cast<Node8 *>(reinterpret_cast<char *>(h) + L.nodes_offset);
            copy_vec16(&nodes[0], &nd, 5);
            if (L.parent_offset) *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(h) + L.parent_offset) = INVALID_U32;
            h->node_count = 1u, h->built = 1u, h->depth = 1u, h->max_stack = 1u;
        }
    }
}

// primitive boxes (kept, in input order: the Morton pass streams them, the climb gathers 32 bytes per leaf instead of
// re-fetching indices and vertices) and the bounds of their centres
template<class Input>
__global__ void __launch_bounds__(256) k_prim_bounds(Input in, f4 *__restrict__ pbox, uint32_t *small) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 2
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += gridDim.x * blockDim.x) {
        const Box b = in.box(i);
        pbox[2ull * i] = make_f4(b.lo[0], b.lo[1], b.lo[2], 0.f), pbox[2ull * i + 1] = make_f4(b.hi[0], b.hi[1], b.hi[2], 0.f);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float c = 0.5f * (b.lo[a] + b.hi[a]);
            if (c == c) lo[a] = fminf(lo[a], c), hi[a] = fmaxf(hi[a], c);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    // one set of atomics per block (the six words are a hot spot)
    __shared__ float s_lo[3][8], s_hi[3][8];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (lane == 0u) {
#pragma unroll
        for (int a = 0; a < 3; a++) s_lo[a][warp] = lo[a], s_hi[a][warp] = hi[a];
    }
    __syncthreads();
    if (threadIdx.x < 3u) {
        const int a = (int)threadIdx.x;
        float l = INFINITY, h = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; w++) l = fminf(l, s_lo[a][w]), h = fmaxf(h, s_hi[a][w]);
        if (l <= h) {
            atomicMin(&small[SM_CMIN + a], float_to_ordered(l));
            atomicMax(&small[SM_CMAX + a], float_to_ordered(h));
        }
    }
}

LCRT_HD Box box_from_pbox(const f4 lo, const f4 hi) {
    Box b;
    b.lo[0] = lo.x, b.lo[1] = lo.y, b.lo[2] = lo.z, b.hi[0] = hi.x, b.hi[1] = hi.y, b.hi[2] = hi.z;
    return b;
}

__global__ void __launch_bounds__(256) k_morton(const f4 *__restrict__ pbox, uint32_t n, const uint32_t *__restrict__ small, uint64_t *__restrict__ keys,
                                                uint32_t bits, uint32_t index_bits) {
    float clo[3], inv_ext[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const uint32_t ulo = small[SM_CMIN + a], uhi = small[SM_CMAX + a];
        clo[a] = ordered_to_float(ulo);
        const float ext = ordered_to_float(uhi) - clo[a];
        inv_ext[a] = ext > 0.f ? 1.0f / ext : 0.f;
        if (ulo == 0xffffffffu) clo[a] = 0.f, inv_ext[a] = 0.f;
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (morton63(box_from_pbox(pbox[2ull * i], pbox[2ull * i + 1]), clo, inv_ext, bits) << index_bits) | (uint64_t)i;
}

// One thread per sorted leaf.  (Measured and rejected, profiles/r2_build_experiments.md: a block-local climb that keeps the nodes
// whose range lies inside a 256-leaf chunk in shared memory and writes each finished record once: 151 -> 169 us.)
__global__ void __launch_bounds__(64) k_climb(Bvh2 t, const f4 *__restrict__ pbox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t.n) {
        const uint32_t p = t.order(i);
        lbvh_climb(t, i, box_from_pbox(pbox[2ull * p], pbox[2ull * p + 1]));
    }
}

#ifndef LCRT_COLLAPSE_MIN_BLOCKS
#define LCRT_COLLAPSE_MIN_BLOCKS 8
#endif
#ifndef LCRT_COLLAPSE_SLEEP_AFTER
#define LCRT_COLLAPSE_SLEEP_AFTER 16u
#define LCRT_COLLAPSE_SLEEP_SHORT 32u
#define LCRT_COLLAPSE_SLEEP_LONG 200u
#endif
#ifndef LCRT_COLLAPSE_PREFETCH
#define LCRT_COLLAPSE_PREFETCH 0
#endif
constexpr int COLLAPSE_BLOCK = 128;

// ---- collapse_node() (lcrt_build.h, the serial statement the host simulation runs) done by EIGHT lanes per wide node:
// lane k of an "octet" owns child k (its Bvh2Half) in registers, so nothing is indexed dynamically in local memory; the
// choices (which child to open, which (child, slot) pair to assign next) are arg-max reductions over the octet with three
// __shfl_xor steps, and the per-child work (leaf records, quantisation) runs in parallel.  Opening a child is ONE
// dependent load: the binary node record carries both grandchildren completely.  Every decision is the serial one
// (same tie-breaks), so both produce the same tree.  All shuffles are executed by the full warp: control flow is
// warp-uniform, per-octet activity is a predicate.
struct OctChild {
    uint32_t id;
    float lox, loy, loz, hix, hiy, hiz, area;
    bool valid, leaf;
    uint32_t slots, count, first;// leaf slots below the child; primitives in it and their first sorted position if it is a leaf
};
// NC: the binary hierarchy was written by an EARLIER kernel (read-only path); the single-block small build reads what it
// wrote itself a moment ago and uses ordinary loads
template<bool NC>
__device__ __forceinline__ void oct_set(OctChild &c, const Bvh2Half *src) {
    const float4 *s = reinterpret_cast<const float4 *>(src);
    const float4 a = NC ? __ldg(s) : s[0], b = NC ? __ldg(s + 1) : s[1];
    const uint4 m = NC ? __ldg(reinterpret_cast<const uint4 *>(s + 2)) : *reinterpret_cast<const uint4 *>(s + 2);
    c.valid = true;
    c.lox = a.x, c.loy = a.y, c.loz = a.z, c.hix = b.x, c.hiy = b.y, c.hiz = b.z;
    const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    c.area = (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) ? 0.f : dx * dy + dy * dz + dz * dx;// area_of
    const uint32_t cm = __float_as_uint(b.w);
    c.leaf = (cm & LEAF_MARK) != 0u;
    c.count = cm & ~LEAF_MARK;
    c.id = m.x, c.slots = m.y, c.first = m.z;
}
// the record an opening of this child would read: asked for early, it is in L1 when the (sequential) opening loop gets there
__device__ __forceinline__ void oct_prefetch(const Bvh2 &t, const OctChild &c) {
    if (LCRT_COLLAPSE_PREFETCH && c.valid && !c.leaf) {
        const char *p = reinterpret_cast<const char *>(&t.node[c.id]);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 64));
    }
}
// opens child `k` of every octet with `go`: its left child stays in lane k, its right child goes to lane nc
template<bool NC>
__device__ __forceinline__ void oct_open(const Bvh2 &t, OctChild &c, uint32_t l, bool go, uint32_t k, uint32_t &nc) {
    const uint32_t id = __shfl_sync(0xffffffffu, c.id, (int)k, 8);
    if (go && l == k) oct_set<NC>(c, &t.node[id].l);
    if (go && l == nc) oct_set<NC>(c, &t.node[id].r);
    if (go && (l == k || l == nc)) oct_prefetch(t, c);
    if (go) nc++;
}
// arg-max over the octet of an integer key (0 = "not a candidate"); ties go to the lower lane; returns 8 if there is no
// candidate.  Two REDUX instructions on the octet's own lane mask instead of a three-step shuffle tree.
__device__ __forceinline__ uint32_t oct_argmax(uint32_t key, uint32_t l, uint32_t omask) {
    const uint32_t m = __reduce_max_sync(omask, key);
    return __reduce_min_sync(omask, (key == m && m != 0u) ? l : 8u);
}
__device__ __forceinline__ uint32_t oct_or(uint32_t v, uint32_t omask) { return __reduce_or_sync(omask, v); }

// COUNTED: the persistent kernel; the octet that hands out the last primitive slot raises `done`.  (Measured and rejected,
// profiles/r2_build_experiments.md: leaving the six nodes out of seven whose whole subtree fits into one node to a second kernel
// with one thread per node, 414 -> 337 + 195 us; counting pending nodes instead of primitives, two more fences per node.)
template<class Input, bool NC, bool COUNTED>
__device__ bool collapse_octet(const Bvh2 &t, const Bvh8Out &o, const Input &in, uint32_t widx, uint32_t bid, uint32_t depth, bool act,
                               unsigned char *node_smem, uint32_t *done_flag) {
    const uint32_t l = threadIdx.x & 7u;
    OctChild c;
    c.id = INVALID_U32, c.valid = false, c.leaf = true, c.slots = 0u, c.count = 0u, c.first = 0u, c.area = 0.f;
    c.lox = c.loy = c.loz = INFINITY, c.hix = c.hiy = c.hiz = -INFINITY;
    uint32_t nc = 0u;
    if (act && widx == 0u) {
        // the root arrives described as a half: a leaf-marked root (<= 3 primitives in all) is the node's only child
        const uint32_t cm = t.root->count_mark;
        if (cm & LEAF_MARK) {
            nc = 1u;
            if (l == 0u) oct_set<NC>(c, t.root);
        } else {
            bid = t.root->id;
        }
    }
    if (act && nc == 0u) {
        nc = 2u;
        if (l == 0u) oct_set<NC>(c, &t.node[bid].l);
        if (l == 1u) oct_set<NC>(c, &t.node[bid].r);
        if (l < 2u) oct_prefetch(t, c);
    }
    const uint32_t omask = 0xffu << (threadIdx.x & 24u), oshift = threadIdx.x & 24u;
    // a subtree that fits into one node entirely (six wide nodes out of seven) is dissolved level by level, every
    // internal child of a round at once: one dependent load per LEVEL instead of one per opening
    const bool fits = act && __reduce_add_sync(omask, c.valid ? c.slots : 0u) <= 8u;
    if (__any_sync(0xffffffffu, fits)) {
        for (;;) {
            const bool opening = fits && c.valid && !c.leaf;
            const uint32_t om = (__ballot_sync(0xffffffffu, opening) >> oshift) & 0xffu;
            if (!__any_sync(0xffffffffu, om != 0u)) break;
            // lane k (opening) keeps the left half; lane nc + r takes the right half of the r-th opening lane
            const int src = (l >= nc && l - nc < (uint32_t)__popc(om)) ? (int)__fns(om, 0u, (int)(l - nc) + 1) : (int)l;
            const uint32_t id = __shfl_sync(0xffffffffu, c.id, src, 8);
            if (opening) oct_set<NC>(c, &t.node[id].l);
            else if (om && l >= nc && l - nc < (uint32_t)__popc(om)) oct_set<NC>(c, &t.node[id].r);
            nc += (uint32_t)__popc(om);
        }
    }
    if (__any_sync(0xffffffffu, act && !fits)) {
        const bool gen = act && !fits;
        // ---- open the largest child with more than 8 leaf slots until the node is full
        for (;;) {
            const bool cand = gen && nc < 8u && c.valid && !c.leaf && c.slots > 8u;
            const uint32_t best = oct_argmax(cand ? __float_as_uint(c.area) + 1u : 0u, l, omask);
            if (!__any_sync(0xffffffffu, best < 8u)) break;
            oct_open<NC>(t, c, l, best < 8u, best & 7u, nc);
        }
        // ---- dissolve, smallest first, children whose whole subtree fits into the free slots
        uint32_t pending = 0u;
        bool more = gen;
        for (;;) {
            const bool cand = more && pending == 0u && c.valid && !c.leaf && nc - 1u + c.slots <= 8u;
            const uint32_t best = oct_argmax(cand ? 0xffffffffu - c.slots : 0u, l, omask);
            if (more && pending == 0u) {
                if (best < 8u) pending = 1u << best;
                else more = false;
            }
            if (!__any_sync(0xffffffffu, more)) break;
            const uint32_t k = pending ? (uint32_t)bfind(pending) : 0u;
            pending &= ~(1u << k);
            const bool kleaf = __shfl_sync(0xffffffffu, (int)c.leaf, (int)k, 8) != 0;
            const bool go = more && !kleaf;
            oct_open<NC>(t, c, l, go, k, nc);
            if (go) pending |= (1u << k) | (1u << (nc - 1u));
        }
    }
    // ---- node box and centre
    float nlx = c.lox, nly = c.loy, nlz = c.loz, nhx = c.hix, nhy = c.hiy, nhz = c.hiz;
    if (!c.valid) nlx = nly = nlz = INFINITY, nhx = nhy = nhz = -INFINITY;
    {
        // box_grow semantics: NaN coordinates never enter (fminf / fmaxf return the other operand)
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
            nlx = fminf(nlx, __shfl_xor_sync(0xffffffffu, nlx, m, 8)), nly = fminf(nly, __shfl_xor_sync(0xffffffffu, nly, m, 8));
            nlz = fminf(nlz, __shfl_xor_sync(0xffffffffu, nlz, m, 8)), nhx = fmaxf(nhx, __shfl_xor_sync(0xffffffffu, nhx, m, 8));
            nhy = fmaxf(nhy, __shfl_xor_sync(0xffffffffu, nhy, m, 8)), nhz = fmaxf(nhz, __shfl_xor_sync(0xffffffffu, nhz, m, 8));
        }
    }
    float cx = 0.5f * (nlx + nhx), cy = 0.5f * (nly + nhy), cz = 0.5f * (nlz + nhz);
    if (!(cx == cx)) cx = 0.f;
    if (!(cy == cy)) cy = 0.f;
    if (!(cz == cz)) cz = 0.f;
    // ---- octant-ordered slot assignment: nc rounds of "best (child, slot) pair among the unassigned"
    int slot = -1;
    {
        float ddx = 0.5f * (c.lox + c.hix) - cx, ddy = 0.5f * (c.loy + c.hiy) - cy, ddz = 0.5f * (c.loz + c.hiz) - cz;
        if (!(ddx == ddx)) ddx = 0.f;
        if (!(ddy == ddy)) ddy = 0.f;
        if (!(ddz == ddz)) ddz = 0.f;
        float v[8];// this child's affinity to each slot (octant), constant over the rounds
#pragma unroll
        for (int s = 0; s < 8; s++) v[s] = ((s & 1) ? ddx : -ddx) + ((s & 2) ? ddy : -ddy) + ((s & 4) ? ddz : -ddz);
        uint32_t free_slots = 0xffu;
        const uint32_t rounds = __reduce_max_sync(0xffffffffu, act ? nc : 0u);
        for (uint32_t round = 0u; round < rounds; round++) {
            float bv = -INFINITY;
            int bs = -1;
#pragma unroll
            for (int s = 0; s < 8; s++)
                if (((free_slots >> s) & 1u) && (v[s] > bv || bs < 0)) bv = v[s], bs = s;
            const bool cand = act && round < nc && c.valid && slot < 0;
            const uint32_t who = oct_argmax(cand ? float_to_ordered(bv) : 0u, l, omask);
            const int ws = __shfl_sync(0xffffffffu, bs, (int)(who & 7u), 8);
            if (who < 8u) {
                if (l == who) slot = ws;
                free_slots &= ~(1u << ws);
            }
        }
    }
    // ---- allocate internal children and primitive records
    const bool is_int = act && c.valid && !c.leaf, is_leaf = act && c.valid && c.leaf;
    const uint32_t int_by_slot = oct_or(is_int ? (1u << slot) : 0u, omask);
    const uint32_t cnt_by_slot = oct_or(is_leaf ? (c.count << (2 * slot)) : 0u, omask);// counts <= 3: two bits per slot
    const uint32_t n_int = (uint32_t)__popc(int_by_slot);
    const uint32_t n_prims = (uint32_t)__popc(cnt_by_slot & 0x5555u) + 2u * (uint32_t)__popc(cnt_by_slot & 0xaaaau);
    // one atomic per WARP and counter: every wide node of the build goes through these two words, and same-address atomics are
    // served one at a time by one L2 slice (the four octets of a warp add up first)
    uint32_t child_base, prim_base, warp_prims;
    {
        const uint32_t oct = (threadIdx.x >> 3) & 3u, lane = threadIdx.x & 31u;
        const uint32_t ai = act ? n_int : 0u, ap = act ? n_prims : 0u;
        const uint32_t i0 = __shfl_sync(0xffffffffu, ai, 0), i1 = __shfl_sync(0xffffffffu, ai, 8), i2 = __shfl_sync(0xffffffffu, ai, 16),
                       i3 = __shfl_sync(0xffffffffu, ai, 24);
        const uint32_t p0 = __shfl_sync(0xffffffffu, ap, 0), p1 = __shfl_sync(0xffffffffu, ap, 8), p2 = __shfl_sync(0xffffffffu, ap, 16),
                       p3 = __shfl_sync(0xffffffffu, ap, 24);
        const uint32_t ti = i0 + i1 + i2 + i3, tpr = p0 + p1 + p2 + p3;
        uint32_t bi = 0u, bp = 0u;
        if (lane == 0u) {
            if (ti) bi = atomicAdd(&o.counters[0 * COUNTER_STRIDE], ti);
            if (tpr) bp = atomicAdd(&o.counters[1 * COUNTER_STRIDE], tpr);
        }
        bi = __shfl_sync(0xffffffffu, bi, 0), bp = __shfl_sync(0xffffffffu, bp, 0);
        child_base = bi + (oct > 0u ? i0 : 0u) + (oct > 1u ? i1 : 0u) + (oct > 2u ? i2 : 0u);
        prim_base = bp + (oct > 0u ? p0 : 0u) + (oct > 1u ? p1 : 0u) + (oct > 2u ? p2 : 0u);
        warp_prims = tpr;
    }
    bool ok = true;
    if (act && n_int && child_base + n_int > o.node_capacity) {
        if (l == 0u) o.header->error = 1u;
        ok = false, act = false;
    }
    // ---- per child: schedule the internal ones, note the leaves' records, quantise the box
    Node8 *stage = reinterpret_cast<Node8 *>(node_smem) + (threadIdx.x >> 3);
    uint32_t *sw = reinterpret_cast<uint32_t *>(stage);
    const bool okx = nlx <= nhx, oky = nly <= nhy, okz = nlz <= nhz;
    const float px = okx ? nlx : 0.f, py = oky ? nly : 0.f, pz = okz ? nlz : 0.f;
    const uint32_t ex = okx ? quant_exponent(nlx, nhx) : 0u, ey = oky ? quant_exponent(nly, nhy) : 0u, ez = okz ? quant_exponent(nlz, nhz) : 0u;
    // defaults: empty slots have meta 0 and inverted boxes (lo 255, hi 0)
    if (l == 0u) sw[0] = as_uint(px), sw[1] = as_uint(py), sw[2] = as_uint(pz), sw[3] = ex | (ey << 8) | (ez << 16) | (int_by_slot << 24);
    if (l == 1u) sw[4] = child_base, sw[5] = prim_base, sw[6] = 0u, sw[7] = 0u;
    if (l == 2u) sw[8] = sw[9] = sw[10] = sw[11] = 0xffffffffu;            // qlo_x, qlo_y
    if (l == 3u) sw[12] = sw[13] = 0xffffffffu, sw[14] = sw[15] = 0u;       // qlo_z | qhi_x
    if (l == 4u) sw[16] = sw[17] = sw[18] = sw[19] = 0u;                     // qhi_y, qhi_z
    __syncwarp();
    if (act && c.valid) {
        uint8_t meta;
        if (!c.leaf) {
            const uint32_t w = child_base + (uint32_t)__popc(int_by_slot & ((1u << slot) - 1u));
            meta = (uint8_t)((1u << 5) | (24u + (uint32_t)slot));
            if (o.parent) o.parent[w] = (widx << 3) | (uint32_t)slot;
            // one 64-bit store: a consumer that sees the node id also sees the depth
            __stcg(reinterpret_cast<unsigned long long *>(&o.task[w]), ((unsigned long long)(depth + 1u) << 32) | c.id);
        } else {
            const uint32_t below = cnt_by_slot & ((1u << (2 * slot)) - 1u);
            const uint32_t po = (uint32_t)__popc(below & 0x5555u) + 2u * (uint32_t)__popc(below & 0xaaaau);
            for (uint32_t q = 0; q < c.count; q++) o.rec_src[prim_base + po + q] = c.first + q;
            const uint32_t unary = c.count == 1u ? 1u : c.count == 2u ? 3u : 7u;
            meta = (uint8_t)((unary << 5) | po);
        }
        stage->meta[slot] = meta;
        uint32_t ql, qh;
        ql = 255u, qh = 0u;
        if (c.lox <= c.hix) ql = quant_down(c.lox, px, ex), qh = quant_up(c.hix, px, ex);
        stage->qlo_x[slot] = (uint8_t)ql, stage->qhi_x[slot] = (uint8_t)qh;
        ql = 255u, qh = 0u;
        if (c.loy <= c.hiy) ql = quant_down(c.loy, py, ey), qh = quant_up(c.hiy, py, ey);
        stage->qlo_y[slot] = (uint8_t)ql, stage->qhi_y[slot] = (uint8_t)qh;
        ql = 255u, qh = 0u;
        if (c.loz <= c.hiz) ql = quant_down(c.loz, pz, ez), qh = quant_up(c.hiz, pz, ez);
        stage->qlo_z[slot] = (uint8_t)ql, stage->qhi_z[slot] = (uint8_t)qh;
    }
    __syncwarp();
    if (act && l < 5u) reinterpret_cast<uint4 *>(&o.nodes[widx])[l] = reinterpret_cast<const uint4 *>(stage)[l];
    if (act && l == 0u) {
        if (widx == 0u) {
            o.header->lo[0] = nlx, o.header->lo[1] = nly, o.header->lo[2] = nlz;
            o.header->hi[0] = nhx, o.header->hi[1] = nhy, o.header->hi[2] = nhz;
            if (o.parent) o.parent[0] = INVALID_U32;
        }
    }
    __syncwarp();// the warp's node, task and parent stores precede the release below
    if ((threadIdx.x & 31u) == 0u && warp_prims) {
        const uint32_t before = LCRT_ATOMIC_ADD_ACQ_REL(&o.counters[2 * COUNTER_STRIDE], warp_prims);
        if (COUNTED && before + warp_prims == t.n) {
            // the last primitive has its slot: every wide node has been processed
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(done_flag) = 1u;
        }
    }
    __syncwarp();
    return ok;
}

// BVH2 -> BVH8 in ONE persistent kernel without levels or grid barriers: wide node w is built by the octet (eight lanes)
// that owns index w (round-robin over the grid), as soon as its parent has scheduled it (task[w], a 64-bit word written
// with one store).  Indices are handed out by an atomic counter in the order nodes are opened, so a node's parent always
// has a smaller index: the lowest unprocessed index is always ready, and with the whole grid co-resident (cooperative
// launch) nothing can wait forever.  The last primitive to get its record slot raises `done`; octets whose next index was
// never scheduled leave then.  (Round 1 ran level by level with a grid barrier per level: 0.58 ms at 1.3 M triangles.)
template<class Input>
__global__ void __launch_bounds__(COLLAPSE_BLOCK, LCRT_COLLAPSE_MIN_BLOCKS) k_collapse(Bvh2 t, Bvh8Out o, Input in, uint32_t *abort_flag, uint32_t *done_flag,
                                                                                        uint32_t *levels_out) {
    __shared__ __align__(16) unsigned char node_smem[(COLLAPSE_BLOCK / 8) * sizeof(Node8)];
    __shared__ uint32_t s_depth;
    if (threadIdx.x == 0) s_depth = 0u;
    __syncthreads();
    const uint32_t octets = COLLAPSE_BLOCK / 8u, stride = gridDim.x * octets, l = threadIdx.x & 7u;
    uint32_t w = blockIdx.x * octets + (threadIdx.x >> 3);
    bool finished = w >= o.node_capacity;
    for (;;) {
        // ---- wait until a task of this warp's four octets is there (or nothing will ever come): a tight loop, the polls of
        // idle warps compete for issue slots with the warps that build nodes
        unsigned long long word = 0xffffffffull;
        uint32_t idle = 0u;
        bool act = false;
        for (;;) {
            if (!finished) {
                if (w == 0u) word = 1ull << 32;// the root: described by t.root, depth 1
                else if (l == 0u) word = __ldcg(reinterpret_cast<const unsigned long long *>(&o.task[w]));
            }
            const uint32_t ready = __ballot_sync(0xffffffffu, !finished && (uint32_t)word != INVALID_U32);
            if (ready) {
                act = ((ready >> (threadIdx.x & 24u)) & 1u) != 0u;
                break;
            }
            if ((++idle & 15u) == 0u) {
                // nothing for a while: done / abort are raised after the last task has been written, so an index that is
                // still empty behind them was never going to be scheduled (one more look at it, behind a fence)
                const uint32_t stop = LCRT_VOLATILE_LOAD_U32(done_flag) | LCRT_VOLATILE_LOAD_U32(abort_flag);
                if (stop) {
                    __threadfence();
                    if (!finished && l == 0u) word = __ldcg(reinterpret_cast<const unsigned long long *>(&o.task[w]));
                    const uint32_t late = __ballot_sync(0xffffffffu, !finished && (uint32_t)word != INVALID_U32 && !LCRT_VOLATILE_LOAD_U32(abort_flag));
                    if (late) {
                        act = ((late >> (threadIdx.x & 24u)) & 1u) != 0u;
                        finished = finished || !act;
                    } else {
                        finished = true;
                    }
                    break;
                }
                if (idle > (1u << 22)) {// safety net (seconds): never hang the device, flag the build instead
                    atomicExch(abort_flag, 2u);
                    finished = true;
                    break;
                }
            }
            __nanosleep(idle < LCRT_COLLAPSE_SLEEP_AFTER ? LCRT_COLLAPSE_SLEEP_SHORT : LCRT_COLLAPSE_SLEEP_LONG);// back off: idle polls cost the busy warps issue slots
        }
        if (__all_sync(0xffffffffu, finished)) break;
        word = __shfl_sync(0xffffffffu, word, (int)(threadIdx.x & 24u));
        const uint32_t bid = (uint32_t)word, depth = (uint32_t)(word >> 32);
        if (!collapse_octet<Input, true, true>(t, o, in, w, bid, depth, act, node_smem, done_flag)) atomicExch(abort_flag, 1u);
        if (act) {
            if (l == 0u) atomicMax(&s_depth, depth);
            w += stride;
            if (w >= o.node_capacity) finished = true;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_depth) atomicMax(levels_out, s_depth);
}

// =====================================================================================================================
// Small structures (n <= SMALL_BUILD_MAX): the whole build in ONE launch of one block.  test_rtx.cpp:122-139 rebuilds a
// one-triangle mesh and a three-instance TLAS every frame; the general pipeline costs sixteen stream operations there
// (three memsets, thirteen kernels), i.e. the launch latency sixteen times over.  Same algorithm, same functions, same
// tree: boxes and keys in shared memory, the keys ranked against each other (n^2 / 256 comparisons per thread), the climb
// and the collapse on the global scratch arrays with block barriers between the stages.
// =====================================================================================================================
constexpr uint32_t SMALL_BUILD_MAX = 512;
constexpr int SMALL_BLOCK = 256;

template<class Input>
__global__ void __launch_bounds__(SMALL_BLOCK) k_build_small(Input in, BlobHeader *h, uint32_t flags, BlobLayout L, uint32_t *small, uint32_t *flag, u2 *task,
                                                             uint64_t *keys, uint32_t *rec_src, f4 *pbox, Bvh2Node *node, float ci) {
    __shared__ uint64_t skey[SMALL_BUILD_MAX];
    __shared__ __align__(16) unsigned char node_smem[(SMALL_BLOCK / 8) * sizeof(Node8)];
    __shared__ uint32_t s_cb[6], s_depth, s_end;
    const uint32_t n = in.n, tid = threadIdx.x;
    // ---- header, counters, scratch
    for (uint32_t k = tid; k < SMALL_WORDS; k += SMALL_BLOCK) small[k] = 0u;
    for (uint32_t k = tid; k < n; k += SMALL_BLOCK) flag[k] = 0u;
    for (uint32_t k = tid; k < L.node_capacity; k += SMALL_BLOCK) task[k] = make_u2(INVALID_U32, 0u);
    if (tid < 3u) s_cb[tid] = 0xffffffffu, s_cb[3 + tid] = 0u;
    if (tid == 0u) {
        BlobHeader hd;
        memset(&hd, 0, sizeof(hd));
        hd.magic = BLOB_MAGIC, hd.kind = Input::kind, hd.prim_count = n, hd.node_capacity = L.node_capacity;
        hd.flags = flags, hd.prim_stride = L.prim_stride, hd.plant_magic = PLANT_MAGIC;
        hd.nodes_offset = L.nodes_offset, hd.prims_offset = L.prims_offset, hd.parent_offset = L.parent_offset, hd.total_bytes = L.total_bytes;
        for (int a = 0; a < 3; a++) hd.lo[a] = INFINITY, hd.hi[a] = -INFINITY;
        *h = hd;
        s_depth = 0u;
    }
    __syncthreads();
    if (tid == 0u) small[SM_COUNTERS] = 1u;// the root
    // ---- boxes and centre bounds
    for (uint32_t i = tid; i < n; i += SMALL_BLOCK) {
        const Box b = in.box(i);
        pbox[2u * i] = make_f4(b.lo[0], b.lo[1], b.lo[2], 0.f), pbox[2u * i + 1u] = make_f4(b.hi[0], b.hi[1], b.hi[2], 0.f);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float c = 0.5f * (b.lo[a] + b.hi[a]);
            if (c == c) atomicMin(&s_cb[a], float_to_ordered(c)), atomicMax(&s_cb[3 + a], float_to_ordered(c));
        }
    }
    __syncthreads();
    // ---- keys, ranked in shared memory (keys are unique: the primitive index sits in the low bits)
    const uint32_t bits = morton_bits_for(n), index_bits = index_bits_for(n);
    {
        float clo[3], inv_ext[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            clo[a] = ordered_to_float(s_cb[a]);
            const float ext = ordered_to_float(s_cb[3 + a]) - clo[a];
            inv_ext[a] = ext > 0.f ? 1.0f / ext : 0.f;
            if (s_cb[a] == 0xffffffffu) clo[a] = 0.f, inv_ext[a] = 0.f;
        }
        for (uint32_t i = tid; i < n; i += SMALL_BLOCK)
            skey[i] = (morton63(box_from_pbox(pbox[2u * i], pbox[2u * i + 1u]), clo, inv_ext, bits) << index_bits) | (uint64_t)i;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SMALL_BLOCK) {
        const uint64_t k = skey[i];
        uint32_t rank = 0u;
        for (uint32_t j = 0; j < n; j++) rank += skey[j] < k ? 1u : 0u;
        keys[rank] = k;
    }
    __syncthreads();
    // ---- binary hierarchy
    Bvh2 t;
    t.n = n, t.index_bits = index_bits, t.keys = keys, t.node = node, t.flag = flag, t.root = reinterpret_cast<Bvh2Half *>(small + SM_ROOT_HALF);
    t.ci = ci, t.max_leaf = Input::max_leaf;
    for (uint32_t i = tid; i < n; i += SMALL_BLOCK) {
        const uint32_t p = t.order(i);
        lbvh_climb(t, i, box_from_pbox(pbox[2u * p], pbox[2u * p + 1u]));
    }
    __threadfence_block();
    __syncthreads();
    // ---- wide hierarchy, level by level (the nodes of a level are the index range allocated while the previous one ran)
    Bvh8Out o;
    o.header = h, o.nodes = reinterpret_cast<Node8 *>(reinterpret_cast<char *>(h) + L.nodes_offset), o.prims = reinterpret_cast<char *>(h) + L.prims_offset;
    o.parent = L.parent_offset ? reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(h) + L.parent_offset) : nullptr;
    o.task = task, o.counters = small + SM_COUNTERS, o.node_capacity = L.node_capacity, o.rec_src = rec_src;
    const uint32_t octets = SMALL_BLOCK / 8u, oct = tid >> 3;
    bool failed = false;
    for (uint32_t begin = 0u, end = 1u, depth = 1u; begin < end; depth++) {
        for (uint32_t base = begin; base < end; base += octets) {// block-uniform trip count
            const uint32_t w = base + oct;
            const bool act = w < end;
            uint32_t bid = INVALID_U32;
            if (act && w != 0u) bid = *reinterpret_cast<volatile uint32_t *>(&task[w].x);
            if (!collapse_octet<Input, false, false>(t, o, in, w, bid, depth, act, node_smem, small + SM_DONE)) failed = true;
        }
        __threadfence_block();
        __syncthreads();
        if (tid == 0u) {
            const uint32_t count = *reinterpret_cast<volatile uint32_t *>(&small[SM_COUNTERS]);
            s_end = count < L.node_capacity ? count : L.node_capacity;
            s_depth = depth;
        }
        __syncthreads();
        begin = end, end = s_end;
    }
    if (failed) small[SM_ABORT] = 1u;
    // ---- records, header
    uint32_t child_depth = 0u;
    for (uint32_t j = tid; j < n; j += SMALL_BLOCK) {
        emit_record(t, o, in, j);
        if constexpr (Input::kind == KIND_INSTANCES) {
            const BlobHeader *bh = InstInput::blas_of(in.inst[t.order(rec_src[j])]);
            if (bh) child_depth = child_depth > bh->depth ? child_depth : bh->depth;
        }
    }
    if constexpr (Input::kind == KIND_INSTANCES) {
        child_depth = __reduce_max_sync(0xffffffffu, child_depth);
        if ((tid & 31u) == 0u && child_depth) atomicMax(&small[SM_CHILD_DEPTH], child_depth);
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0u) {
        const uint32_t levels = s_depth;
        h->node_count = small[SM_COUNTERS], h->prim_written = small[SM_COUNTERS + COUNTER_STRIDE];
        h->depth = levels;
        h->max_stack = Input::kind == KIND_INSTANCES ? levels + 2u + *reinterpret_cast<volatile uint32_t *>(&small[SM_CHILD_DEPTH]) : levels;
        bool ok = small[SM_ABORT] == 0u && *reinterpret_cast<volatile uint32_t *>(&small[SM_COUNTERS + 2 * COUNTER_STRIDE]) == n &&
                  small[SM_COUNTERS] <= L.node_capacity;
        uint32_t err = ok ? 0u : (small[SM_ABORT] ? small[SM_ABORT] : 3u);
        if (ok && h->max_stack > (uint32_t)LCRT_STACK_SIZE) ok = false, err = 5u;
        h->error = err, h->built = ok ? 1u : 0u;
    }
}

template<class Input>
__global__ void __launch_bounds__(256) k_emit_records(Bvh2 t, Bvh8Out o, Input in, uint32_t *child_depth) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < t.n) emit_record(t, o, in, j);
    if constexpr (Input::kind == KIND_INSTANCES) {
        // deepest instanced BLAS (one atomic per warp): the TLAS header records the stack a ray can need
        uint32_t d = 0u;
        if (j < t.n) {
            const BlobHeader *h = InstInput::blas_of(in.inst[t.order(o.rec_src[j])]);
            if (h) d = h->depth;
        }
        d = __reduce_max_sync(0xffffffffu, d);
        if ((threadIdx.x & 31u) == 0u && d) atomicMax(child_depth, d);
    }
}

// error codes in BlobHeader::error: 1 node capacity, 2 grid barrier gave up, 3 primitive count mismatch, 4 refit of a blob
// that is not updatable, 5 the hierarchy is deeper than the traversal stack (LCRT_STACK_SIZE)
__global__ void k_build_finalize(BlobHeader *h, const uint32_t *small, uint32_t n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        h->node_count = small[SM_COUNTERS + 0 * COUNTER_STRIDE];
        h->prim_written = small[SM_COUNTERS + 1 * COUNTER_STRIDE];
        h->depth = small[SM_LEVELS];
        h->max_stack = h->kind == KIND_INSTANCES ? small[SM_LEVELS] + 2u + small[SM_CHILD_DEPTH] : small[SM_LEVELS];
        bool ok = small[SM_ABORT] == 0u && small[SM_COUNTERS + 2 * COUNTER_STRIDE] == n && small[SM_COUNTERS + 0 * COUNTER_STRIDE] <= h->node_capacity;
        uint32_t err = ok ? 0u : (small[SM_ABORT] ? small[SM_ABORT] : 3u);
        if (ok && h->max_stack > (uint32_t)LCRT_STACK_SIZE) ok = false, err = 5u;
        h->error = err;
        h->built = ok ? 1u : 0u;
    }
}

// ---- refit
// A blob that was not built with ALLOW_UPDATE (no parent array), is unbuilt, or has another primitive count is left
// untouched and flagged (error = 4) instead of being corrupted.
__device__ __forceinline__ bool refit_allowed(const BlobHeader *h, uint32_t n) {
    return h->magic == BLOB_MAGIC && h->built && h->parent_offset != 0ull && h->prim_count == n;
}

template<class Input>
__global__ void __launch_bounds__(256) k_refit_records(Input in, BlobHeader *h, char *prims) {
    if (!refit_allowed(h, in.n)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) h->error = 4u;
        return;
    }
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < in.n) {
        char *rec = prims + (uint64_t)j * Input::rec_bytes;
        in.write(rec, Input::record_prim(rec));
    }
}

template<class Input>
__global__ void __launch_bounds__(128) k_refit_climb(RefitCtx c, Input in) {
    if (!refit_allowed(c.header, in.n)) return;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < c.header->node_count) refit_climb(c, in, w);
}

// small structures: the arrival counters, the records and the climb in one launch of one block
template<class Input>
__global__ void __launch_bounds__(SMALL_BLOCK) k_refit_small(RefitCtx c, Input in) {
    if (!refit_allowed(c.header, in.n)) {
        if (threadIdx.x == 0) c.header->error = 4u;
        return;
    }
    const uint32_t nodes = c.header->node_count;
    for (uint32_t w = threadIdx.x; w < nodes; w += SMALL_BLOCK) c.arrived[w] = 0u;
    for (uint32_t j = threadIdx.x; j < in.n; j += SMALL_BLOCK) {
        char *rec = c.prims + (uint64_t)j * Input::rec_bytes;
        in.write(rec, Input::record_prim(rec));
    }
    __threadfence_block();
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < nodes; w += SMALL_BLOCK) refit_climb(c, in, w);
}

// ---- instance buffer
__global__ void k_instances_init(Instance *instances, uint32_t first, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        Instance in;
        init_instance(in);
        copy_vec16(&instances[first + i], &in, 5);
    }
}
__global__ void k_apply_mods(Instance *instances, const Modification *__restrict__ mods, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        Modification m;
        copy_vec16(&m, &mods[i], 5);
        apply_modification(instances, m);
    }
}
// a handful of modifications travel in the launch parameters (no staging copy): the per-frame case of test_rtx.cpp
constexpr uint32_t INLINE_MODS_MAX = 16;
struct InlineMods {
    Modification m[INLINE_MODS_MAX];
};
__global__ void k_apply_mods_inline(Instance *instances, const __grid_constant__ InlineMods mods, uint32_t n) {
    const uint32_t i = threadIdx.x;
    if (i < n) {
        Modification m = mods.m[i];
        apply_modification(instances, m);
    }
}
__global__ void k_update_handles(Instance *instances, const InstanceHandleUpdate *__restrict__ ups, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) apply_handle_update(instances, ups[i]);
}

// =====================================================================================================================
// trace kernels
// =====================================================================================================================
#ifndef LCRT_TRACE_BLOCK
#define LCRT_TRACE_BLOCK 128
#endif
constexpr int TRACE_BLOCK = LCRT_TRACE_BLOCK;
constexpr int TRACE_SMEM_STACK = 8;   // stack entries per thread kept in shared memory
#ifndef LCRT_TRACE_BATCH
#define LCRT_TRACE_BATCH 128
#endif
constexpr uint32_t TRACE_BATCH = LCRT_TRACE_BATCH; // rays a warp reserves per atomic
#ifndef LCRT_TRACE_MIN_BLOCKS
#define LCRT_TRACE_MIN_BLOCKS 5     // resident blocks per SM the register allocation is bounded for
#endif
#ifndef LCRT_TRI_VOTE_PATIENCE
#define LCRT_TRI_VOTE_PATIENCE 128
#endif
// The vote makes a lane with a pending primitive wait for company.  In RayQuery batches such a lane was seen to wait FOREVER
// (BASELINE C5, the 2 Mi bounce rays of tests/test_gpu_fullsize.py::test_c5_fullsize: the launch never ended in about half of
// the runs, never without the vote, never in the host simulation of the same schedule): RayQuery launches therefore bound the
// wait (a warp-uniform counter: after LCRT_TRI_VOTE_PATIENCE refusals in a row the primitive phase runs regardless).
// Closest / any-hit launches have never shown it and pay 4 % for the counter (2392 -> 2296 Mrays/s coherent on C2), so there it
// is compiled in only with LCRT_TRI_VOTE_VALVE_ALL_MODES=1.  profiles/r2_trace_experiments.md.
#ifndef LCRT_TRI_VOTE_VALVE_ALL_MODES
#define LCRT_TRI_VOTE_VALVE_ALL_MODES 0
#endif
#ifndef LCRT_TRI_VOTE
#define LCRT_TRI_VOTE 4               // lanes with a pending primitive that make the warp run the primitive phase
#endif
#ifndef LCRT_STREAM_HITS
#define LCRT_STREAM_HITS 1
#endif
#ifndef LCRT_PREFETCH_RAYS
#define LCRT_PREFETCH_RAYS 1
#endif
#ifndef LCRT_REFILL_MIN
#define LCRT_REFILL_MIN 12            // idle lanes that trigger a warp-cooperative refill
#endif

// Per-thread storage outside the register file (see LocalStore in lcrt_traverse.h): every array is laid out
// [entry][thread] so that a warp's accesses to one entry are consecutive words / vectors (conflict free).
struct TraceSmem {
    u2 stack[TRACE_SMEM_STACK * TRACE_BLOCK];
    f4 rows[2 * TRACE_BLOCK];
    uint32_t words[CW_COUNT * TRACE_BLOCK];
    uint32_t slots[(TRACE_BLOCK / 32) * 4];// FUSED_BATCH, per warp: batch index of slot 0, of slot 1 (~0 = free), rays still missing (16 bits each)
};
__device__ __forceinline__ TraceSmem *trace_smem() {
    // dynamic shared memory (sizeof(TraceSmem) at launch): referencing the array directly keeps every access an
    // LDS / STS (a pointer member would decay to generic loads and stores)
    extern __shared__ __align__(16) unsigned char lcrt_trace_smem[];
    return reinterpret_cast<TraceSmem *>(lcrt_trace_smem);
}
// FUSED kernels: batches (128 consecutive rays reserved by one warp) whose finished records are collected before they are
// stored to the destination: a record stored on its own is a 24-byte write, and eight GPUs sending those into one use a
// quarter of its NVLink ingress (profiles/r2_scaling.md).  The records wait in the caller's local scratch array (L2), not in
// shared memory: 3 KB per slot and warp there would leave the SM without L1 for the hierarchy (measured: 3.72 -> 5.51 ms per
// C2 step at two GPUs); only the bookkeeping is per warp, in registers.
#ifndef LCRT_STAGE_SLOTS
#define LCRT_STAGE_SLOTS 2
#endif
constexpr uint32_t STAGE_SLOTS = LCRT_STAGE_SLOTS;

struct HybridStore {
    static constexpr bool world_in_global = true;
    u2 local[LCRT_STACK_SIZE - TRACE_SMEM_STACK];
    __device__ __forceinline__ void store(int i, u2 v) {
        if (i < TRACE_SMEM_STACK) trace_smem()->stack[i * TRACE_BLOCK + threadIdx.x] = v;
        else local[i - TRACE_SMEM_STACK] = v;
    }
    __device__ __forceinline__ u2 load(int i) const {
        return i < TRACE_SMEM_STACK ? trace_smem()->stack[i * TRACE_BLOCK + threadIdx.x] : local[i - TRACE_SMEM_STACK];
    }
    __device__ __forceinline__ void st4(int k, f4 v) { trace_smem()->rows[k * TRACE_BLOCK + threadIdx.x] = v; }
    __device__ __forceinline__ f4 ld4(int k) const { return trace_smem()->rows[k * TRACE_BLOCK + threadIdx.x]; }
    __device__ __forceinline__ void stw(int k, uint32_t v) { trace_smem()->words[k * TRACE_BLOCK + threadIdx.x] = v; }
    __device__ __forceinline__ uint32_t ldw(int k) const { return trace_smem()->words[k * TRACE_BLOCK + threadIdx.x]; }
};

// Where the final record of ray i goes: in place (dst == nullptr), or into a framebuffer of tiles — `tile` consecutive rays are
// contiguous, tile j starts at dst + j * stride.  The framebuffer may be PEER memory (another GPU's HBM mapped over NVLink):
// tile-sharded multi-GPU tracing writes its records straight into the gathering rank's ray-ordered buffer from this pass, so
// there is no gather step and no reorder copy behind the trace (luisacompute_b200.dist.PeerFramebuffer).
struct HitDest {
    char *dst;
    uint64_t tile, stride;
    __device__ __forceinline__ u2 *at(void *in_place, uint64_t i) const {
        if (dst == nullptr) return reinterpret_cast<u2 *>(static_cast<char *>(in_place) + i * 24u);
        return reinterpret_cast<u2 *>(dst + (i / tile) * stride + (i % tile) * 24u);
    }
};
struct TraceParams {
    const BlobHeader *tlas;
    const Instance *instances;
    const Ray *rays;
    uint64_t n;
    const uint32_t *masks;
    uint32_t mask;
    void *out;
    unsigned long long *next;    // ray cursor (zeroed before the launch)
    unsigned long long *counters;// optional
    uint32_t tri_prog, proc_prog;
    int terminate_on_first;
    uint32_t magic;              // PLANT_MAGIC, see plant_byte()
    HitDest dest;                // FUSED kernels: where the final records go
};

// Persistent threads: every warp reserves TRACE_BATCH consecutive rays per atomic and hands them to its idle lanes
// (ballot + popc rank) whenever at least LCRT_REFILL_MIN lanes are idle; between refills every lane advances its
// ray by one step() per iteration.  Surface hits are written in a provisional form (u = TLAS record index, prim =
// primitive record index, t = fp32 distance) that k_refine_hits turns into the final record.
// FUSED (closest-hit only): the retiring lanes evaluate the exact record themselves; the records of a 128-ray batch are collected
// (in p.out, the local scratch array) and stored where HitDest says when the batch is complete — a framebuffer that may live in a peer GPU's memory.
// The stores then leave over NVLink while the kernel is still tracing, spread over its whole duration (not in one burst from a
// second pass: eight ranks storing into one at the same moment queue at its NVLink ingress), in 512-byte warp stores.
// FUSED: 0 provisional records + k_refine_hits; 1 every final record stored as its ray ends; 2 final records stored per complete batch
enum : int { FUSED_NONE = 0, FUSED_EACH = 1, FUSED_BATCH = 2 };
template<int MODE, bool COUNT, class Callbacks, int FUSED = FUSED_NONE>
__global__ void __launch_bounds__(TRACE_BLOCK, LCRT_TRACE_MIN_BLOCKS) k_trace(TraceParams p) {
    HybridStore store;
    Callbacks cb;
    if constexpr (MODE == MODE_QUERY) cb = Callbacks{p.tri_prog, p.proc_prog};
    Traversal<MODE, COUNT, HybridStore, Callbacks> tr;
    tr.done = true;
    tr.magic = p.magic;
    const u4 *tlas_nodes = reinterpret_cast<const u4 *>(reinterpret_cast<const char *>(p.tlas) + p.tlas->nodes_offset);
    const InstanceRecord *tlas_recs = reinterpret_cast<const InstanceRecord *>(reinterpret_cast<const char *>(p.tlas) + p.tlas->prims_offset);
    const bool single = p.tlas->prim_count == 1u;
    if (p.tlas->max_stack > (uint32_t)LCRT_STACK_SIZE) {
        // a hierarchy deeper than the traversal stack (never one this library built, k_build_finalize): nothing is traced,
        // the error is raised; the caller sees it at its next synchronisation (lcbSynchronize / lcrtBlobInfo.trace_error)
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(const_cast<uint32_t *>(&p.tlas->trace_error), 1u);
        return;
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    bool active = false, unsaved = false;// unsaved: this lane's finished ray has not been written yet
    uint64_t warp_next = 0, warp_end = 0;// this warp's reserved range of ray indices
    unsigned long long c_nodes = 0, c_prims = 0, c_insts = 0, c_exact = 0;
    bool exhausted = false;
    uint32_t denied = 0u;// consecutive iterations in which lanes with a pending primitive were made to wait (warp-uniform)
    // FUSED_BATCH: the batches whose records are being collected (warp-uniform): index of the batch (first ray / TRACE_BATCH, ~0 =
    // slot free) and, 16 bits per slot, the rays still missing
    // (kept in shared memory, three words per warp, read at refills only: as registers they spilled traversal state in the hot loop)
    static_assert(STAGE_SLOTS == 2u && TRACE_BATCH < 65536u, "slot bookkeeping");
    volatile uint32_t *slot = trace_smem()->slots + (threadIdx.x >> 5) * 4u;
    if constexpr (FUSED == FUSED_BATCH) {
        if (lane == 0u) slot[0] = ~0u, slot[1] = ~0u, slot[2] = 0u;
        __syncwarp();
    }
    for (;;) {
        // The ballot below is meant to see the whole warp.  nvcc 12.9 lowers __ballot_sync(0xffffffff, ..) here to a VOTE / VOTEU
        // with NO preceding WARPSYNC in some instantiations — it takes the BSYNC.RECONVERGENT in front of the back edge for
        // convergence, turns an explicit __syncwarp() into a NOP and ignores a run-time mask — and the RayQuery instantiation of
        // round 1, one of those, did not finish in half of its runs on C5's bounce rays (profiles/r2_trace_experiments.md §3):
        // a part of the warp that votes on its own can never find "every lane idle" and spins for ever once the rays are
        // exhausted.  Every build with a WARPSYNC.ALL on the back edge has always finished.  The instantiations that do not get one
        // from the compiler are therefore written so that a PART of a warp is also correct on its own: "everybody" is whoever
        // takes part in the vote, and the leader is its lowest lane.  (With the whole warp present: the same values as before.)
        // Closest-hit in place and FUSED_EACH carry the WARPSYNC.ALL and keep the constants: their SASS is unchanged.
        constexpr bool converge = MODE != MODE_CLOSEST || FUSED == FUSED_BATCH || COUNT;
        const uint32_t here = converge ? __ballot_sync(0xffffffffu, true) : 0xffffffffu;
        const int leader = converge ? __ffs((int)here) - 1 : 0;
        // ---- warp-cooperative retire + refill of idle lanes
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle && ((!exhausted && (uint32_t)__popc(idle) >= (uint32_t)LCRT_REFILL_MIN) || idle == here)) {
            if constexpr (MODE == MODE_CLOSEST && FUSED != FUSED_NONE) {
                // the retiring lanes finish their records themselves (exact t, u, v)
                uint64_t ray_idx = ~0ull;
                uint2 r0 = make_uint2(0u, 0u), r1 = r0, r2 = r0;
                if (unsaved) {
                    ray_idx = (uint64_t)(tr.world - p.rays);
                    uint32_t prim = ~0u;
                    float t = tr.hit_t(), u = 0.f, v = 0.f;
                    if (tr.has_hit()) refine_surface_hit(tr.world, tlas_recs, tr.hit_rec(), tr.hit_pi(), prim, t, u, v);
                    r0 = make_uint2(tr.hit_inst(), prim), r1 = make_uint2(as_uint(u), as_uint(v)), r2 = make_uint2(as_uint(t), 0u);
                }
                bool staged = false;
                if constexpr (FUSED == FUSED_BATCH) {
                    // records go to their batch's place in the local scratch array; a batch without a slot (both were busy when it
                    // was reserved) is stored record by record
                    const uint32_t my_batch = unsaved ? (uint32_t)(ray_idx / TRACE_BATCH) : ~0u - 1u;
#pragma unroll
                    for (uint32_t sl = 0; sl < STAGE_SLOTS; sl++) {
                        const uint32_t sb = slot[sl];
                        const bool mine = my_batch == sb;
                        const uint32_t m = __ballot_sync(0xffffffffu, mine);
                        if (m == 0u) continue;
                        const uint64_t base = (uint64_t)sb * TRACE_BATCH;
                        uint2 *st = reinterpret_cast<uint2 *>(static_cast<char *>(p.out) + base * 24u);
                        if (mine) {
                            uint2 *e = st + (uint32_t)(ray_idx - base) * 3u;
                            e[0] = r0, e[1] = r1, e[2] = r2;
                            staged = true;
                        }
                        const uint32_t left = slot[2] - ((uint32_t)__popc(m) << (16u * sl));
                        __syncwarp();
                        if (lane == 0u) slot[2] = left;
                        if (((left >> (16u * sl)) & 0xffffu) == 0u) {
                            // the batch is complete: its records are consecutive in the destination (a batch never straddles a tile)
                            __syncwarp();
                            char *dst = reinterpret_cast<char *>(p.dest.at(p.out, base));
                            const uint64_t left_in_launch = p.n - base;
                            const uint32_t bytes = (uint32_t)(left_in_launch < TRACE_BATCH ? left_in_launch : TRACE_BATCH) * 24u;
                            if (dst != reinterpret_cast<char *>(st)) {
                                if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(st) | bytes) & 15u) == 0u) {
                                    const uint4 *src = reinterpret_cast<const uint4 *>(st);
                                    for (uint32_t k = lane; k < bytes / 16u; k += 32u) __stcs(reinterpret_cast<uint4 *>(dst) + k, __ldcg(src + k));
                                } else {
                                    for (uint32_t k = lane; k < bytes / 8u; k += 32u) __stcs(reinterpret_cast<uint2 *>(dst) + k, __ldcg(st + k));
                                }
                            }
                            if (lane == 0u) slot[sl] = ~0u;
                        }
                        __syncwarp();
                    }
                }
                if (unsaved && !staged) {
                    uint2 *o = reinterpret_cast<uint2 *>(p.dest.at(p.out, ray_idx));
                    __stcs(o, r0), __stcs(o + 1, r1), __stcs(o + 2, r2);
                }
                unsaved = false;
            }
            if (unsaved) {
                // results are written here, together by all lanes that finished since the last refill
                const uint64_t ray_idx = (uint64_t)(tr.world - p.rays);
                const bool hit = tr.has_hit();
                if constexpr (MODE == MODE_CLOSEST) {
                    u2 *o = reinterpret_cast<u2 *>(static_cast<SurfaceHit *>(p.out) + ray_idx);
#if LCRT_STREAM_HITS
                    __stcs(reinterpret_cast<uint2 *>(o), make_uint2(tr.hit_inst(), hit ? tr.hit_pi() : ~0u));
                    __stcs(reinterpret_cast<uint2 *>(o) + 1, make_uint2(hit ? tr.hit_rec() : 0u, 0u));
                    __stcs(reinterpret_cast<uint2 *>(o) + 2, make_uint2(as_uint(tr.hit_t()), 0u));
#else
                    o[0] = make_u2(tr.hit_inst(), hit ? tr.hit_pi() : ~0u), o[1] = make_u2(hit ? tr.hit_rec() : 0u, 0u), o[2] = make_u2(as_uint(tr.hit_t()), 0u);
#endif
                } else if constexpr (MODE == MODE_ANY) {
                    static_cast<uint32_t *>(p.out)[ray_idx] = hit ? 1u : 0u;
                } else {
                    const uint32_t type = tr.hit_type();
                    u2 *o = reinterpret_cast<u2 *>(static_cast<CommittedHit *>(p.out) + ray_idx);
                    o[0] = make_u2(tr.hit_inst(), type == HIT_SURFACE ? tr.hit_pi() : tr.hit_prim());
                    o[1] = make_u2(type == HIT_SURFACE ? tr.hit_rec() : 0u, 0u), o[2] = make_u2(type, as_uint(tr.hit_t()));
                }
                if constexpr (COUNT) c_nodes += tr.ctr.nodes, c_prims += tr.ctr.prims, c_insts += tr.ctr.insts, c_exact += tr.ctr.exact;
                unsaved = false;
            }
            if (exhausted) break;// every lane idle and nothing left to fetch
            const uint32_t want = (uint32_t)__popc(idle);
            if (warp_next == warp_end) {
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(p.next, (unsigned long long)TRACE_BATCH);
                base = __shfl_sync(0xffffffffu, base, leader);
                warp_next = base < p.n ? base : p.n;
                warp_end = base + TRACE_BATCH < p.n ? base + TRACE_BATCH : p.n;
                if constexpr (MODE == MODE_CLOSEST && FUSED == FUSED_BATCH) {
                    if (warp_next != warp_end) {
                        const uint32_t cnt = (uint32_t)(warp_end - warp_next), s0 = slot[0], s1 = slot[1], left = slot[2];
                        __syncwarp();
                        if (lane == 0u) {
                            if (s0 == ~0u) slot[0] = (uint32_t)(warp_next / TRACE_BATCH), slot[2] = (left & 0xffff0000u) | cnt;
                            else if (s1 == ~0u) slot[1] = (uint32_t)(warp_next / TRACE_BATCH), slot[2] = (left & 0x0000ffffu) | (cnt << 16);
                        }
                        __syncwarp();
                    }
                }
#if LCRT_PREFETCH_RAYS
                // the reserved rays (TRACE_BATCH x 32 B = one 128-byte line per lane) stream from HBM exactly once
                if (warp_next + lane * 4u < warp_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.rays + warp_next + lane * 4u));
#endif
                if (warp_next == warp_end) exhausted = true;
            }
            const uint64_t avail = warp_end - warp_next;
            const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
            if (!active && rank < avail) {
                const uint64_t ray_idx = warp_next + rank;
                tr.init(tlas_nodes, tlas_recs, single, p.instances, p.rays + ray_idx, p.masks ? p.masks[ray_idx] : p.mask, &store, &cb);
                if constexpr (MODE == MODE_QUERY) tr.terminate_on_first = p.terminate_on_first != 0;
                active = true;
            }
            warp_next += want < avail ? want : avail;
        }
#ifdef LCRT_TRACE_WATCHDOG
        // developer aid: a ray that is still being traversed after this many iterations is reported and dropped
        {
            static_assert(sizeof(unsigned long long) == 8, "");
            if (active && ++c_nodes > (unsigned long long)LCRT_TRACE_WATCHDOG) {
                printf("[lcrt watchdog] ray %llu block %u thread %u: sp %d blas_sp %d G (%u, %08x) Gt (%u, %08x) tp %u inst %u flags %08x octinv %08x tmin %g tmax %g done %d\n",
                       (unsigned long long)(tr.world - p.rays), blockIdx.x, threadIdx.x, tr.sp, tr.blas_sp, tr.G.x, tr.G.y, tr.Gt.x, tr.Gt.y, tr.tp, tr.cur_inst,
                       tr.cur_flags, tr.octinv, (double)tr.tmin, (double)tr.tmax, (int)tr.done);
                tr.done = true;
            }
            if (!active) c_nodes = 0;
        }
#endif
        // ---- one traversal step per lane: a node, then (if the warp votes for it) a primitive
        bool alive = active && tr.step_node();
#if LCRT_TRI_VOTE > 0
        {
            // the primitive phase runs at a few lanes per warp when every lane enters it as soon as it has a candidate;
            // lanes with a pending primitive wait until LCRT_TRI_VOTE of them (or every live lane) can go together
            const uint32_t am = __ballot_sync(0xffffffffu, alive);
            const uint32_t pm = __ballot_sync(0xffffffffu, alive && tr.prim_pending());
            // (a lane is never made to wait more than LCRT_TRI_VOTE_PATIENCE iterations: liveness does not rest on the other
            // lanes' traversals)
#if LCRT_TRI_VOTE_VALVE_ALL_MODES
            constexpr bool valve = true;
#else
            constexpr bool valve = MODE == MODE_QUERY;
#endif
            bool go = (uint32_t)__popc(pm) >= (uint32_t)LCRT_TRI_VOTE || pm == am;
            if constexpr (valve) go = go || denied >= (uint32_t)LCRT_TRI_VOTE_PATIENCE;
            if (go) {
                alive = alive && tr.step_prim();
                if constexpr (valve) denied = 0u;
            } else if constexpr (valve) {
                if (pm != 0u) denied++;
            }
        }
#else
        alive = alive && tr.step_prim();
#endif
        if (active && !alive) active = false, unsaved = true;
    }
    if constexpr (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_nodes += __shfl_xor_sync(0xffffffffu, c_nodes, o);
            c_prims += __shfl_xor_sync(0xffffffffu, c_prims, o);
            c_insts += __shfl_xor_sync(0xffffffffu, c_insts, o);
            c_exact += __shfl_xor_sync(0xffffffffu, c_exact, o);
        }
        if (lane == 0 && p.counters) {
            atomicAdd(&p.counters[0], c_nodes);
            atomicAdd(&p.counters[1], c_prims);
            atomicAdd(&p.counters[2], c_insts);
            atomicAdd(&p.counters[3], c_exact);
        }
    }
}

// Second pass of closest-hit / RayQuery batches: the reported (t, u, v) of every surface hit are recomputed in
// double with the oracle's formula (refine_surface_hit), one thread per ray, fully convergent and coalesced.
template<int MODE>
__global__ void __launch_bounds__(256) k_refine_hits(const BlobHeader *tlas, const Ray *__restrict__ rays, void *out, uint64_t n, HitDest d) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const InstanceRecord *tlas_recs = reinterpret_cast<const InstanceRecord *>(reinterpret_cast<const char *>(tlas) + tlas->prims_offset);
    const u2 *in = reinterpret_cast<const u2 *>(static_cast<const char *>(out) + i * 24u);
    u2 *o = d.at(out, i);
    const u2 h0 = in[0], h1 = in[1], h2 = in[2];
    if constexpr (MODE == MODE_CLOSEST) {
        if (h0.x == ~0u) {
            if (d.dst) o[0] = h0, o[1] = h1, o[2] = h2;// a miss is final as the trace wrote it
            return;
        }
        uint32_t prim = 0;
        float t = as_float(h2.x), u = 0.f, v = 0.f;
        refine_surface_hit(rays + i, tlas_recs, h1.x, h0.y, prim, t, u, v);
        o[0] = make_u2(h0.x, prim), o[1] = make_u2(as_uint(u), as_uint(v)), o[2] = make_u2(as_uint(t), 0u);
    } else {
        if (h2.x != HIT_SURFACE) {
            if (d.dst) o[0] = h0, o[1] = h1, o[2] = h2;
            return;
        }
        uint32_t prim = 0;
        float t = as_float(h2.y), u = 0.f, v = 0.f;
        refine_surface_hit(rays + i, tlas_recs, h1.x, h0.y, prim, t, u, v);
        o[0] = make_u2(h0.x, prim), o[1] = make_u2(as_uint(u), as_uint(v)), o[2] = make_u2(HIT_SURFACE, as_uint(t));
    }
}

// =====================================================================================================================
// host side
// =====================================================================================================================
static thread_local char g_last_cuda_error[256] = "";

static int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
    return LCRT_ERROR_CUDA;
}
#define LCRT_CUDA(call)                                     \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)
#define LCRT_LAUNCH_CHECK(name)                            \
    do {                                                   \
        cudaError_t e_ = cudaGetLastError();               \
        if (e_ != cudaSuccess) return cuda_fail(e_, name); \
    } while (0)

struct DeviceInfo {
    int sm_count = 0;
    unsigned long long *cursors = nullptr;// ring of ray cursors
    int collapse_blocks_per_sm[3] = {0, 0, 0};
    int trace_blocks_per_sm = 0;
};
constexpr int CURSOR_RING = 4096;
static std::mutex g_dev_mutex;
static DeviceInfo g_dev[64];
static std::atomic<uint32_t> g_cursor_next{0};

static int device_info(DeviceInfo **out) {
    int dev = 0;
    LCRT_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return LCRT_ERROR_INVALID_VALUE;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceInfo &d = g_dev[dev];
    if (d.sm_count == 0) {
        int sms = 0;
        LCRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        LCRT_CUDA(cudaMalloc(&d.cursors, CURSOR_RING * sizeof(unsigned long long)));
        d.sm_count = sms;
    }
    *out = &d;
    return LCRT_SUCCESS;
}

static float sah_ci(uint32_t flags) {
    static const char *env = getenv("LCRT_SAH_CI");
    if (env) return (float)atof(env);
    return (flags & BUILD_FAST_BUILD) ? 1.0f : 0.5f;
}

template<class Input>
static int build_generic(cudaStream_t stream, const Input &in, uint32_t flags, void *temp, uint64_t temp_bytes, void *blob, uint64_t blob_bytes) {
    const uint32_t n = in.n;
    const BlobLayout L = blob_layout(Input::kind, n, flags);
    const TempLayout T = temp_layout(n, L.node_capacity);
    if (blob == nullptr || blob_bytes < L.total_bytes) return LCRT_ERROR_BUFFER_TOO_SMALL;
    if (temp == nullptr || temp_bytes < T.total_bytes) return LCRT_ERROR_BUFFER_TOO_SMALL;
    if ((reinterpret_cast<uintptr_t>(blob) & 127u) || (reinterpret_cast<uintptr_t>(temp) & 127u)) return LCRT_ERROR_MISALIGNED;
    if (n > SORT_MAX_KEYS) return LCRT_ERROR_INVALID_VALUE;
    DeviceInfo *dev;
    if (int rc = device_info(&dev)) return rc;
    char *tp = static_cast<char *>(temp);
    char *bp = static_cast<char *>(blob);
    uint32_t *small = reinterpret_cast<uint32_t *>(tp + T.small);
    BlobHeader *h = reinterpret_cast<BlobHeader *>(bp);
    uint32_t *flag = reinterpret_cast<uint32_t *>(tp + T.flag);
    u2 *task = reinterpret_cast<u2 *>(tp + T.task);
    if (n != 0u && n <= SMALL_BUILD_MAX) {
        // one launch of one block does it all (k_build_small)
        k_build_small<Input><<<1, SMALL_BLOCK, 0, stream>>>(in, h, flags, L, small, flag, task, reinterpret_cast<uint64_t *>(tp + T.keys0),
                                                            reinterpret_cast<uint32_t *>(tp + T.keys1), reinterpret_cast<f4 *>(tp + T.pbox),
                                                            reinterpret_cast<Bvh2Node *>(tp + T.node), sah_ci(flags));
        LCRT_LAUNCH_CHECK("k_build_small");
        return LCRT_SUCCESS;
    }
    k_build_init<<<1, 32, 0, stream>>>(h, Input::kind, n, flags, L, small);
    LCRT_LAUNCH_CHECK("k_build_init");
    if (n == 0) return LCRT_SUCCESS;
    // flag and task are adjacent in the scratch layout: arrival counters 0, tasks INVALID
    LCRT_CUDA(cudaMemsetAsync(flag, 0, (uint64_t)n * 4, stream));
    LCRT_CUDA(cudaMemsetAsync(task, 0xff, (uint64_t)L.node_capacity * 8, stream));
    const uint32_t blocks256 = (n + 255u) / 256u;
    f4 *pbox = reinterpret_cast<f4 *>(tp + T.pbox);
    {
        const uint32_t g = blocks256 < (uint32_t)dev->sm_count * 16u ? blocks256 : (uint32_t)dev->sm_count * 16u;
        k_prim_bounds<Input><<<g, 256, 0, stream>>>(in, pbox, small);
        LCRT_LAUNCH_CHECK("k_prim_bounds");
    }
    SortBuffers sb;
    sb.keys[0] = reinterpret_cast<uint64_t *>(tp + T.keys0), sb.keys[1] = reinterpret_cast<uint64_t *>(tp + T.keys1);
    sb.vals[0] = sb.vals[1] = nullptr;
    sb.control = reinterpret_cast<uint32_t *>(tp + T.control);
    const uint32_t bits = morton_bits_for(n), index_bits = index_bits_for(n);
    k_morton<<<blocks256, 256, 0, stream>>>(pbox, n, small, sb.keys[0], bits, index_bits);
    LCRT_LAUNCH_CHECK("k_morton");
    const int cur = radix_sort(stream, sb, n, index_bits, 3u * bits, dev->sm_count);
    if (cur < 0) return LCRT_ERROR_INVALID_VALUE;
    LCRT_LAUNCH_CHECK("radix_sort");
    Bvh2 t;
    t.n = n, t.index_bits = index_bits, t.keys = sb.keys[cur];
    t.node = reinterpret_cast<Bvh2Node *>(tp + T.node);
    t.flag = flag, t.root = reinterpret_cast<Bvh2Half *>(small + SM_ROOT_HALF), t.ci = sah_ci(flags), t.max_leaf = Input::max_leaf;
    // 64-thread blocks: a block lives as long as its longest climber (most threads leave after one or two levels)
    k_climb<<<(n + 63u) / 64u, 64, 0, stream>>>(t, pbox);
    LCRT_LAUNCH_CHECK("k_climb");
    Bvh8Out o;
    o.header = h, o.nodes = reinterpret_cast<Node8 *>(bp + L.nodes_offset), o.prims = bp + L.prims_offset;
    o.parent = L.parent_offset ? reinterpret_cast<uint32_t *>(bp + L.parent_offset) : nullptr;
    o.task = task, o.counters = small + SM_COUNTERS, o.node_capacity = L.node_capacity;
    o.rec_src = reinterpret_cast<uint32_t *>(sb.keys[cur ^ 1]);// the idle half of the sort's ping-pong buffers
    {
        int &bps = dev->collapse_blocks_per_sm[Input::kind];
        if (bps == 0) {
            LCRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_collapse<Input>, COLLAPSE_BLOCK, 0));
            if (bps < 1) bps = 1;
            static const char *cap = getenv("LCRT_COLLAPSE_BLOCKS_PER_SM");// developer knob (tools/build_perf.py sweeps it)
            if (cap && atoi(cap) > 0 && atoi(cap) < bps) bps = atoi(cap);
        }
        uint32_t g = (uint32_t)(dev->sm_count * bps);
        const uint32_t need = (L.node_capacity + COLLAPSE_BLOCK / 8u - 1u) / (COLLAPSE_BLOCK / 8u);
        if (g > need) g = need;
        // cooperative launch: the runtime guarantees that all blocks are co-resident (octets wait for their parents' tasks)
        uint32_t *abort_flag = small + SM_ABORT, *done_flag = small + SM_DONE, *levels = small + SM_LEVELS;
        Input in_arg = in;
        void *args[] = {&t, &o, &in_arg, &abort_flag, &done_flag, &levels};
        LCRT_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(&k_collapse<Input>), dim3(g), dim3(COLLAPSE_BLOCK), args, 0, stream));
    }
    k_emit_records<Input><<<blocks256, 256, 0, stream>>>(t, o, in, small + SM_CHILD_DEPTH);
    LCRT_LAUNCH_CHECK("k_emit_records");
    k_build_finalize<<<1, 32, 0, stream>>>(h, small, n);
    LCRT_LAUNCH_CHECK("k_build_finalize");
    return LCRT_SUCCESS;
}

template<class Input>
static int refit_generic(cudaStream_t stream, const Input &in, void *temp, uint64_t temp_bytes, void *blob) {
    // the header lives on the device; sizes are derived from the primitive count the caller states (validated on
    // the device: a mismatching blob makes the kernels no-ops through node_count / prim_count checks)
    const uint32_t n = in.n;
    const BlobLayout L = blob_layout(Input::kind, n, BUILD_ALLOW_UPDATE);
    if (blob == nullptr) return LCRT_ERROR_INVALID_VALUE;
    if (temp == nullptr || temp_bytes < refit_temp_bytes(L.node_capacity)) return LCRT_ERROR_BUFFER_TOO_SMALL;
    if (n == 0) return LCRT_SUCCESS;
    char *bp = static_cast<char *>(blob);
    char *tp = static_cast<char *>(temp);
    RefitCtx c;
    c.header = reinterpret_cast<BlobHeader *>(bp);
    c.nodes = reinterpret_cast<Node8 *>(bp + L.nodes_offset), c.prims = bp + L.prims_offset;
    c.parent = reinterpret_cast<const uint32_t *>(bp + L.parent_offset);
    c.box_lo = reinterpret_cast<f4 *>(tp), c.box_hi = c.box_lo + L.node_capacity;
    c.arrived = reinterpret_cast<uint32_t *>(c.box_hi + L.node_capacity);
    if (n <= SMALL_BUILD_MAX) {
        k_refit_small<Input><<<1, SMALL_BLOCK, 0, stream>>>(c, in);
        LCRT_LAUNCH_CHECK("k_refit_small");
        return LCRT_SUCCESS;
    }
    LCRT_CUDA(cudaMemsetAsync(c.arrived, 0, (uint64_t)L.node_capacity * 4, stream));
    k_refit_records<Input><<<(n + 255u) / 256u, 256, 0, stream>>>(in, c.header, c.prims);
    LCRT_LAUNCH_CHECK("k_refit_records");
    k_refit_climb<Input><<<(L.node_capacity + 127u) / 128u, 128, 0, stream>>>(c, in);
    LCRT_LAUNCH_CHECK("k_refit_climb");
    return LCRT_SUCCESS;
}

template<int MODE, class Callbacks>
static int trace_generic(cudaStream_t stream, const void *tlas, const void *instances, const void *rays, uint64_t n, const uint32_t *masks,
                         uint32_t mask, void *out, uint64_t *counters, uint32_t tri_prog, uint32_t proc_prog, int terminate_on_first,
                         HitDest dest = HitDest{nullptr, 1u, 0u}, bool stage = true) {
    if (n == 0) return LCRT_SUCCESS;
    if (dest.dst != nullptr && (MODE == MODE_ANY || dest.tile == 0u || dest.tile % TRACE_BATCH != 0u || (reinterpret_cast<uintptr_t>(dest.dst) & 7u) || (dest.stride & 7u) ||
                                dest.stride < dest.tile * 24u))
        return LCRT_ERROR_INVALID_VALUE;
    if (tlas == nullptr || rays == nullptr || out == nullptr) return LCRT_ERROR_INVALID_VALUE;
    if (reinterpret_cast<uintptr_t>(rays) & 15u) return LCRT_ERROR_MISALIGNED;
    if (reinterpret_cast<uintptr_t>(out) & 7u) return LCRT_ERROR_MISALIGNED;
    DeviceInfo *dev;
    if (int rc = device_info(&dev)) return rc;
    TraceParams p;
    p.tlas = static_cast<const BlobHeader *>(tlas), p.instances = static_cast<const Instance *>(instances);
    p.rays = static_cast<const Ray *>(rays), p.n = n, p.masks = masks, p.mask = mask, p.out = out;
    p.next = dev->cursors + (g_cursor_next.fetch_add(1) % CURSOR_RING);
    p.counters = reinterpret_cast<unsigned long long *>(counters);
    p.tri_prog = tri_prog, p.proc_prog = proc_prog, p.terminate_on_first = terminate_on_first, p.magic = PLANT_MAGIC;
    LCRT_CUDA(cudaMemsetAsync(p.next, 0, sizeof(unsigned long long), stream));
    // closest-hit records that go to a tiled (possibly peer) framebuffer are finished inside the traversal kernel
    static const bool force_fused = getenv("LCRT_FUSED_REFINE") != nullptr && atoi(getenv("LCRT_FUSED_REFINE")) != 0;
    const bool fused = MODE == MODE_CLOSEST && counters == nullptr && (dest.dst != nullptr || force_fused);
    p.dest = dest;
    auto launch = [&](auto kernel, size_t smem) -> int {
        int bps = 0;
        LCRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, TRACE_BLOCK, smem));
        if (bps < 1) bps = 1;
        static const bool verbose = getenv("LCRT_VERBOSE") != nullptr;
        if (verbose) fprintf(stderr, "[lcrt] k_trace mode %d: %zu B shared memory per block, %d blocks per SM\n", MODE, smem, bps);
        uint64_t g = (uint64_t)dev->sm_count * (uint64_t)bps;
        const uint64_t need = (n + TRACE_BLOCK - 1) / TRACE_BLOCK;
        if (g > need) g = need;
        kernel<<<(uint32_t)g, TRACE_BLOCK, smem, stream>>>(p);
        LCRT_LAUNCH_CHECK("k_trace");
        return LCRT_SUCCESS;
    };
    int rc;
    if (counters) rc = launch(k_trace<MODE, true, Callbacks, FUSED_NONE>, sizeof(TraceSmem));
    else if (fused && stage && n / TRACE_BATCH < 0xfffffff0ull) rc = launch(k_trace<MODE, false, Callbacks, MODE == MODE_CLOSEST ? FUSED_BATCH : FUSED_NONE>, sizeof(TraceSmem));
    else if (fused) rc = launch(k_trace<MODE, false, Callbacks, MODE == MODE_CLOSEST ? FUSED_EACH : FUSED_NONE>, sizeof(TraceSmem));
    else rc = launch(k_trace<MODE, false, Callbacks, FUSED_NONE>, sizeof(TraceSmem));
    if (rc != LCRT_SUCCESS) return rc;
    if (MODE != MODE_ANY && !fused) {
        k_refine_hits<MODE><<<(uint32_t)((n + 255u) / 256u), 256, 0, stream>>>(p.tlas, p.rays, out, n, dest);
        LCRT_LAUNCH_CHECK("k_refine_hits");
    }
    return LCRT_SUCCESS;
}

}// namespace lcrt

// =====================================================================================================================
// C-ABI
// =====================================================================================================================
using namespace lcrt;

extern "C" {

int lcrtGetVersion(void) { return 100; }

const char *lcrtGetErrorString(int code) {
    switch (code) {
        case LCRT_SUCCESS: return "success";
        case LCRT_ERROR_INVALID_VALUE: return "invalid value";
        case LCRT_ERROR_BUFFER_TOO_SMALL: return "buffer too small";
        case LCRT_ERROR_CUDA: return "CUDA error (see lcrtGetLastCudaError)";
        case LCRT_ERROR_NOT_UPDATABLE: return "acceleration structure was not built with LCRT_BUILD_ALLOW_UPDATE";
        case LCRT_ERROR_MISALIGNED: return "misaligned pointer";
        case LCRT_ERROR_BUILD_FAILED: return "device-side build / refit / traversal error (see lcrtBlobInfo.error, .trace_error)";
        default: return "unknown error";
    }
}

const char *lcrtGetLastCudaError(void) { return g_last_cuda_error; }

int lcrtAccelComputeMemoryUsage(uint32_t kind, uint32_t prim_count, uint32_t flags, lcrtSizes *out) {
    if (out == nullptr || kind > KIND_INSTANCES) return LCRT_ERROR_INVALID_VALUE;
    const BlobLayout L = blob_layout(kind, prim_count, flags);
    const TempLayout T = temp_layout(prim_count, L.node_capacity);
    out->blob_bytes = L.total_bytes;
    out->build_temp_bytes = T.total_bytes;
    out->update_temp_bytes = (flags & BUILD_ALLOW_UPDATE) ? refit_temp_bytes(L.node_capacity) : 0;
    return LCRT_SUCCESS;
}

int lcrtBuildTriangles(lcrtStream stream, const void *vertices, uint64_t vertex_stride, const void *triangles, uint32_t triangle_count,
                       uint32_t flags, void *temp, uint64_t temp_bytes, void *blob, uint64_t blob_bytes) {
    if (triangle_count && (vertices == nullptr || triangles == nullptr)) return LCRT_ERROR_INVALID_VALUE;
    if (vertex_stride < 12 || (vertex_stride & 3u)) return LCRT_ERROR_INVALID_VALUE;
    TriInput in{static_cast<const char *>(vertices), vertex_stride, static_cast<const uint32_t *>(triangles), triangle_count};
    return build_generic(static_cast<cudaStream_t>(stream), in, flags, temp, temp_bytes, blob, blob_bytes);
}

int lcrtRefitTriangles(lcrtStream stream, const void *vertices, uint64_t vertex_stride, const void *triangles, uint32_t triangle_count,
                       void *temp, uint64_t temp_bytes, void *blob) {
    if (triangle_count && (vertices == nullptr || triangles == nullptr)) return LCRT_ERROR_INVALID_VALUE;
    if (vertex_stride < 12 || (vertex_stride & 3u)) return LCRT_ERROR_INVALID_VALUE;
    TriInput in{static_cast<const char *>(vertices), vertex_stride, static_cast<const uint32_t *>(triangles), triangle_count};
    return refit_generic(static_cast<cudaStream_t>(stream), in, temp, temp_bytes, blob);
}

int lcrtBuildAabbs(lcrtStream stream, const void *aabbs, uint32_t aabb_count, uint32_t flags, void *temp, uint64_t temp_bytes, void *blob,
                   uint64_t blob_bytes) {
    if (aabb_count && aabbs == nullptr) return LCRT_ERROR_INVALID_VALUE;
    AabbInput in{static_cast<const float *>(aabbs), aabb_count};
    return build_generic(static_cast<cudaStream_t>(stream), in, flags, temp, temp_bytes, blob, blob_bytes);
}

int lcrtRefitAabbs(lcrtStream stream, const void *aabbs, uint32_t aabb_count, void *temp, uint64_t temp_bytes, void *blob) {
    if (aabb_count && aabbs == nullptr) return LCRT_ERROR_INVALID_VALUE;
    AabbInput in{static_cast<const float *>(aabbs), aabb_count};
    return refit_generic(static_cast<cudaStream_t>(stream), in, temp, temp_bytes, blob);
}

int lcrtInstancesInit(lcrtStream stream, void *instances, uint32_t first, uint32_t count) {
    if (count == 0) return LCRT_SUCCESS;
    if (instances == nullptr) return LCRT_ERROR_INVALID_VALUE;
    k_instances_init<<<(count + 255u) / 256u, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<Instance *>(instances), first, count);
    LCRT_LAUNCH_CHECK("k_instances_init");
    return LCRT_SUCCESS;
}

int lcrtInstancesApplyMods(lcrtStream stream, void *instances, const void *mods_device, uint32_t mod_count) {
    if (mod_count == 0) return LCRT_SUCCESS;
    if (instances == nullptr || mods_device == nullptr) return LCRT_ERROR_INVALID_VALUE;
    // same launch shape as the reference: 256 threads per block, one thread per modification (cuda_accel.cpp:234-238)
    k_apply_mods<<<(mod_count + 255u) / 256u, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<Instance *>(instances),
                                                                                             static_cast<const Modification *>(mods_device), mod_count);
    LCRT_LAUNCH_CHECK("k_apply_mods");
    return LCRT_SUCCESS;
}

int lcrtInstancesApplyModsHost(lcrtStream stream, void *instances, const void *mods_host, uint32_t mod_count) {
    if (mod_count == 0) return LCRT_SUCCESS;
    if (instances == nullptr || mods_host == nullptr || mod_count > INLINE_MODS_MAX) return LCRT_ERROR_INVALID_VALUE;
    InlineMods p;
    memcpy(p.m, mods_host, (size_t)mod_count * sizeof(Modification));
    k_apply_mods_inline<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<Instance *>(instances), p, mod_count);
    LCRT_LAUNCH_CHECK("k_apply_mods_inline");
    return LCRT_SUCCESS;
}

int lcrtInstancesUpdateHandles(lcrtStream stream, void *instances, const void *updates_device, uint32_t count) {
    if (count == 0) return LCRT_SUCCESS;
    if (instances == nullptr || updates_device == nullptr) return LCRT_ERROR_INVALID_VALUE;
    k_update_handles<<<(count + 255u) / 256u, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<Instance *>(instances), static_cast<const InstanceHandleUpdate *>(updates_device), count);
    LCRT_LAUNCH_CHECK("k_update_handles");
    return LCRT_SUCCESS;
}

int lcrtBuildTlas(lcrtStream stream, const void *instances, uint32_t instance_count, uint32_t flags, void *temp, uint64_t temp_bytes,
                  void *blob, uint64_t blob_bytes) {
    if (instance_count == 0 || instances == nullptr) return LCRT_ERROR_INVALID_VALUE;// cuda_accel.cpp:159
    InstInput in{static_cast<const Instance *>(instances), instance_count};
    return build_generic(static_cast<cudaStream_t>(stream), in, flags, temp, temp_bytes, blob, blob_bytes);
}

int lcrtRefitTlas(lcrtStream stream, const void *instances, uint32_t instance_count, void *temp, uint64_t temp_bytes, void *blob) {
    if (instance_count == 0 || instances == nullptr) return LCRT_ERROR_INVALID_VALUE;
    InstInput in{static_cast<const Instance *>(instances), instance_count};
    return refit_generic(static_cast<cudaStream_t>(stream), in, temp, temp_bytes, blob);
}

int lcrtTraceClosest(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count, const uint32_t *masks,
                     uint32_t mask, void *hits, uint64_t *counters) {
    return trace_generic<MODE_CLOSEST, NoCallbacks>(static_cast<cudaStream_t>(stream), tlas, instances, rays, ray_count, masks, mask, hits, counters,
                                                    0, 0, 0);
}

int lcrtTraceClosestTiled(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count, const uint32_t *masks,
                          uint32_t mask, void *scratch_hits, void *dst, uint64_t tile_rays, uint64_t tile_stride_bytes, uint32_t flags,
                          uint64_t *counters) {
    if (dst == nullptr) return LCRT_ERROR_INVALID_VALUE;
    return trace_generic<MODE_CLOSEST, NoCallbacks>(static_cast<cudaStream_t>(stream), tlas, instances, rays, ray_count, masks, mask, scratch_hits,
                                                    counters, 0, 0, 0, HitDest{static_cast<char *>(dst), tile_rays, tile_stride_bytes},
                                                    (flags & LCRT_TILED_STORE_EACH_RECORD) == 0u);
}

int lcrtTraceAny(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count, const uint32_t *masks,
                 uint32_t mask, uint32_t *occluded, uint64_t *counters) {
    return trace_generic<MODE_ANY, NoCallbacks>(static_cast<cudaStream_t>(stream), tlas, instances, rays, ray_count, masks, mask, occluded, counters, 0,
                                                0, 0);
}

int lcrtQuery(lcrtStream stream, const void *tlas, const void *instances, const void *rays, uint64_t ray_count, const uint32_t *masks,
              uint32_t mask, int terminate_on_first, uint32_t tri_program, uint32_t proc_program, void *hits) {
    if (tri_program > TRI_TERMINATE_FIRST || proc_program > PROC_IGNORE_ALL) return LCRT_ERROR_INVALID_VALUE;
    if (instances == nullptr) return LCRT_ERROR_INVALID_VALUE;
    return trace_generic<MODE_QUERY, BuiltinPrograms>(static_cast<cudaStream_t>(stream), tlas, instances, rays, ray_count, masks, mask, hits, nullptr,
                                                      tri_program, proc_program, terminate_on_first);
}

int lcrtBlobGetInfo(lcrtStream stream, const void *blob, lcrtBlobInfo *out) {
    if (blob == nullptr || out == nullptr) return LCRT_ERROR_INVALID_VALUE;
    BlobHeader h;
    LCRT_CUDA(cudaMemcpyAsync(&h, blob, sizeof(h), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    LCRT_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (h.magic != BLOB_MAGIC) return LCRT_ERROR_INVALID_VALUE;
    out->kind = h.kind, out->prim_count = h.prim_count, out->node_count = h.node_count, out->node_capacity = h.node_capacity;
    out->built = h.built, out->error = h.error, out->flags = h.flags, out->prim_stride = h.prim_stride;
    for (int a = 0; a < 3; a++) out->lo[a] = h.lo[a], out->hi[a] = h.hi[a];
    out->nodes_offset = h.nodes_offset, out->prims_offset = h.prims_offset, out->total_bytes = h.total_bytes;
    out->depth = h.depth, out->max_stack = h.max_stack, out->trace_error = h.trace_error, out->reserved = 0u;
    return LCRT_SUCCESS;
}

uint64_t lcrtSortTempBytes(uint32_t n) {
    const uint64_t nn = n ? n : 1;
    return align_up(nn * 8, 256) + align_up(nn * 4, 256) + align_up(sort_control_bytes(n), 256);
}

int lcrtSortPairs(lcrtStream stream_, uint64_t *keys, uint32_t *values, uint32_t n, uint32_t key_bits, void *temp, uint64_t temp_bytes) {
    if (n == 0) return LCRT_SUCCESS;
    if (keys == nullptr || values == nullptr || temp == nullptr || key_bits == 0 || key_bits > 64 || n > SORT_MAX_KEYS) return LCRT_ERROR_INVALID_VALUE;
    if (temp_bytes < lcrtSortTempBytes(n)) return LCRT_ERROR_BUFFER_TOO_SMALL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DeviceInfo *dev;
    if (int rc = device_info(&dev)) return rc;
    char *tp = static_cast<char *>(temp);
    SortBuffers sb;
    sb.keys[0] = keys, sb.vals[0] = values;
    sb.keys[1] = reinterpret_cast<uint64_t *>(tp);
    sb.vals[1] = reinterpret_cast<uint32_t *>(tp + align_up((uint64_t)n * 8, 256));
    sb.control = reinterpret_cast<uint32_t *>(tp + align_up((uint64_t)n * 8, 256) + align_up((uint64_t)n * 4, 256));
    const int cur = radix_sort(stream, sb, n, 0u, key_bits, dev->sm_count);
    if (cur < 0) return LCRT_ERROR_INVALID_VALUE;
    LCRT_LAUNCH_CHECK("radix_sort");
    if (cur == 1) {
        LCRT_CUDA(cudaMemcpyAsync(keys, sb.keys[1], (uint64_t)n * 8, cudaMemcpyDeviceToDevice, stream));
        LCRT_CUDA(cudaMemcpyAsync(values, sb.vals[1], (uint64_t)n * 4, cudaMemcpyDeviceToDevice, stream));
    }
    return LCRT_SUCCESS;
}

extern const char lcrt_device_library_text[];
const char *lcrtDeviceLibrarySource(void) { return lcrt_device_library_text; }

}// extern "C"
