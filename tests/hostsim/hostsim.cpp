// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the *device* headers of luisacompute_b200/csrc (lcrt_build.h, lcrt_traverse.h, lcrt_accel.h) with g++ and
// runs the per-thread kernel bodies serially on the CPU.  There is no GPU in the build container, so this is how
// the builder / BVH8 format / traversal logic is validated against the oracle before GPU time is spent.  It is
// never loaded by the product package (luisacompute_b200 fails loudly without its CUDA library).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../luisacompute_b200/csrc/lcrt_accel.h"
#include "../../luisacompute_b200/csrc/lcrt_build.h"
#include "../../luisacompute_b200/csrc/lcrt_layout.h"
#include "../../luisacompute_b200/csrc/lcrt_query_programs.h"
#include "../../luisacompute_b200/csrc/lcrt_traverse.h"

using namespace lcrt;

namespace {

template<class Input>
uint32_t deepest_blas(const Input &) { return 0u; }
uint32_t deepest_blas(const InstInput &in) {
    uint32_t deepest = 0;
    for (uint32_t i = 0; i < in.n; i++)
        if (const BlobHeader *b = InstInput::blas_of(in.inst[i])) deepest = std::max(deepest, b->depth);
    return deepest;
}

template<class Input>
int build_blob(const Input &in, uint32_t flags, float ci, char *blob) {
    const uint32_t n = in.n;
    const BlobLayout L = blob_layout(Input::kind, n, flags);
    BlobHeader *h = reinterpret_cast<BlobHeader *>(blob);
    memset(h, 0, sizeof(BlobHeader));
    h->plant_magic = PLANT_MAGIC, h->magic = BLOB_MAGIC, h->kind = Input::kind, h->prim_count = n, h->node_capacity = L.node_capacity;
    h->flags = flags, h->prim_stride = L.prim_stride;
    h->nodes_offset = L.nodes_offset, h->prims_offset = L.prims_offset, h->parent_offset = L.parent_offset, h->total_bytes = L.total_bytes;
    for (int a = 0; a < 3; a++) h->lo[a] = INFINITY, h->hi[a] = -INFINITY;
    Node8 *nodes = reinterpret_cast<Node8 *>(blob + L.nodes_offset);
    if (n == 0) {
        // one empty node: nothing can be hit
        Node8 nd;
        memset(&nd, 0, sizeof(nd));
        Box nb = box_empty();
        Box sb[8] = {};
        bool sv[8] = {};
        quantize_node(nd, nb, sb, sv);
        nodes[0] = nd;
        h->node_count = 1, h->built = 1, h->depth = 1, h->max_stack = 1;
        return 0;
    }
    // centroid bounds
    uint32_t cb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (uint32_t i = 0; i < n; i++) {
        const Box b = in.box(i);
        for (int a = 0; a < 3; a++) {
            const float c = 0.5f * (b.lo[a] + b.hi[a]);
            if (c == c) cb[a] = std::min(cb[a], float_to_ordered(c)), cb[3 + a] = std::max(cb[3 + a], float_to_ordered(c));
        }
    }
    float clo[3], inv_ext[3];
    for (int a = 0; a < 3; a++) {
        clo[a] = ordered_to_float(cb[a]);
        const float ext = ordered_to_float(cb[3 + a]) - clo[a];
        inv_ext[a] = ext > 0.f ? 1.0f / ext : 0.f;
        if (cb[a] == 0xffffffffu) clo[a] = 0.f, inv_ext[a] = 0.f;
    }
    // keys = Morton << index_bits | primitive; the device sorts the Morton bits with a stable sort starting from index order
    const uint32_t bits = morton_bits_for(n), index_bits = index_bits_for(n);
    std::vector<uint64_t> skeys(n);
    for (uint32_t i = 0; i < n; i++) skeys[i] = (morton63(in.box(i), clo, inv_ext, bits) << index_bits) | (uint64_t)i;
    std::sort(skeys.begin(), skeys.end());
    std::vector<Bvh2Node> node(n);
    std::vector<uint32_t> flag(n, 0u);
    Bvh2Half root;
    memset(&root, 0, sizeof(root));
    root.id = INVALID_U32;
    Bvh2 t{n, index_bits, skeys.data(), node.data(), flag.data(), &root, ci, Input::max_leaf};
    // scrambled thread order: the algorithm must not depend on scheduling
    std::vector<uint32_t> perm(n);
    for (uint32_t i = 0; i < n; i++) perm[i] = i;
    uint32_t s = 12345u;
    for (uint32_t i = n; i > 1; i--) {
        s = s * 1664525u + 1013904223u;
        std::swap(perm[i - 1], perm[(s >> 8) % i]);
    }
    for (uint32_t k = 0; k < n; k++) lbvh_climb(t, perm[k], in.box(t.order(perm[k])));
    if (root.id == INVALID_U32) return -10;
    if (root.first != 0u || root.last != n - 1u || (root.count_mark & ~LEAF_MARK) != n) return -10;
    std::vector<u2> task(L.node_capacity, make_u2(INVALID_U32, 0u));
    uint32_t counters[3 * COUNTER_STRIDE] = {};
    counters[0] = 1u;
    std::vector<uint32_t> rec_src(n, INVALID_U32);
    Bvh8Out o{h, nodes, blob + L.prims_offset, L.parent_offset ? reinterpret_cast<uint32_t *>(blob + L.parent_offset) : nullptr, task.data(),
              rec_src.data(), counters, L.node_capacity};
    task[0] = make_u2(root.id, 1u);
    // in index order, as the device's octets take them (a node's parent always has a smaller index); the deepest node gives
    // the depth the header records
    uint32_t levels = 0;
    for (uint32_t w = 0; w < counters[0]; w++) {
        if (w && task[w].x == INVALID_U32) return -11;
        if (!collapse_node(t, o, in, w, task[w].x, task[w].y)) return -12;
        levels = std::max(levels, task[w].y);
    }
    h->depth = levels, h->max_stack = levels;
    if (Input::kind == KIND_INSTANCES) {
        h->max_stack = levels + 2u + deepest_blas(in);
        if (h->max_stack > (uint32_t)LCRT_STACK_SIZE) return -15;
    }
    if (counters[2 * COUNTER_STRIDE] != n || counters[1 * COUNTER_STRIDE] != n) return -13;
    for (uint32_t j = 0; j < n; j++) {
        if (rec_src[j] == INVALID_U32) return -14;
        emit_record(t, o, in, j);
    }
    h->node_count = counters[0], h->prim_written = counters[1 * COUNTER_STRIDE], h->built = 1;
    return 0;
}

template<class Input>
int refit_blob(const Input &in, char *blob) {
    BlobHeader *h = reinterpret_cast<BlobHeader *>(blob);
    if (h->magic != BLOB_MAGIC || !h->built || !h->parent_offset) return -1;
    if (h->prim_count != in.n) return -2;
    char *prims = blob + h->prims_offset;
    for (uint32_t j = 0; j < h->prim_count; j++) {
        char *rec = prims + (uint64_t)j * Input::rec_bytes;
        in.write(rec, Input::record_prim(rec));
    }
    std::vector<f4> lo(h->node_count), hi(h->node_count);
    std::vector<uint32_t> arrived(h->node_count, 0u);
    RefitCtx c{h, reinterpret_cast<Node8 *>(blob + h->nodes_offset), prims, reinterpret_cast<const uint32_t *>(blob + h->parent_offset),
               lo.data(), hi.data(), arrived.data()};
    for (uint32_t w = h->node_count; w-- > 0;) refit_climb(c, in, w);
    return 0;
}

}// namespace

extern "C" {

void hs_blob_layout(uint32_t kind, uint32_t n, uint32_t flags, uint64_t *out) {
    const BlobLayout L = blob_layout(kind, n, flags);
    out[0] = L.total_bytes, out[1] = L.nodes_offset, out[2] = L.prims_offset, out[3] = L.parent_offset, out[4] = L.node_capacity;
}

int hs_build_triangles(const void *verts, uint64_t stride, const uint32_t *tris, uint32_t ntris, uint32_t flags, float ci, void *blob) {
    TriInput in{static_cast<const char *>(verts), stride, tris, ntris};
    return build_blob(in, flags, ci, static_cast<char *>(blob));
}
int hs_build_aabbs(const float *aabbs, uint32_t n, uint32_t flags, float ci, void *blob) {
    AabbInput in{aabbs, n};
    return build_blob(in, flags, ci, static_cast<char *>(blob));
}
int hs_build_tlas(const void *instances, uint32_t n, uint32_t flags, float ci, void *blob) {
    InstInput in{static_cast<const Instance *>(instances), n};
    return build_blob(in, flags, ci, static_cast<char *>(blob));
}
int hs_refit_triangles(const void *verts, uint64_t stride, const uint32_t *tris, uint32_t ntris, void *blob) {
    TriInput in{static_cast<const char *>(verts), stride, tris, ntris};
    return refit_blob(in, static_cast<char *>(blob));
}
int hs_refit_aabbs(const float *aabbs, uint32_t n, void *blob) {
    AabbInput in{aabbs, n};
    return refit_blob(in, static_cast<char *>(blob));
}
int hs_refit_tlas(const void *instances, uint32_t n, void *blob) {
    InstInput in{static_cast<const Instance *>(instances), n};
    return refit_blob(in, static_cast<char *>(blob));
}

void hs_init_instances(void *instances, uint32_t first, uint32_t count) {
    Instance *p = static_cast<Instance *>(instances);
    for (uint32_t i = first; i < first + count; i++) init_instance(p[i]);
}
void hs_apply_mods(void *instances, const void *mods, uint32_t n) {
    const Modification *m = static_cast<const Modification *>(mods);
    for (uint32_t i = 0; i < n; i++) apply_modification(static_cast<Instance *>(instances), m[i]);
}

// structural validation of a blob: every primitive referenced exactly once, child boxes contain their contents
int hs_validate(const void *blob_, uint64_t *stats) {
    const char *blob = static_cast<const char *>(blob_);
    const BlobHeader *h = reinterpret_cast<const BlobHeader *>(blob);
    if (h->magic != BLOB_MAGIC || !h->built || h->error) return -1;
    const Node8 *nodes = reinterpret_cast<const Node8 *>(blob + h->nodes_offset);
    const char *prims = blob + h->prims_offset;
    std::vector<uint32_t> seen(h->prim_count, 0u);
    std::vector<uint8_t> node_seen(h->node_count, 0u);
    struct Item {
        uint32_t node;
        Box bound;
        bool has_bound;
        uint32_t depth;
    };
    std::vector<Item> stack{{0u, box_empty(), false, 1u}};
    uint64_t max_depth = 0, leaf_slots = 0, child_slots = 0;
    double sah = 0.0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (it.node >= h->node_count) return -2;
        if (node_seen[it.node]++) return -3;
        max_depth = std::max<uint64_t>(max_depth, it.depth);
        const Node8 &nd = nodes[it.node];
        const double sc[3] = {nd.ex ? ldexp(1.0, nd.ex - 127) : 0.0, nd.ey ? ldexp(1.0, nd.ey - 127) : 0.0, nd.ez ? ldexp(1.0, nd.ez - 127) : 0.0};
        const double p[3] = {nd.px, nd.py, nd.pz};
        const uint8_t *qlo[3] = {nd.qlo_x, nd.qlo_y, nd.qlo_z}, *qhi[3] = {nd.qhi_x, nd.qhi_y, nd.qhi_z};
        uint32_t ci = 0;
        for (int s = 0; s < 8; s++) {
            const uint32_t meta = nd.meta[s];
            if (!meta) continue;
            Box qb;
            for (int a = 0; a < 3; a++) qb.lo[a] = (float)(p[a] + qlo[a][s] * sc[a]), qb.hi[a] = (float)(p[a] + qhi[a][s] * sc[a]);
            // the decoded child box must lie inside the parent's decoded box for this node (if any)
            if (it.has_bound)
                for (int a = 0; a < 3; a++)
                    if (p[a] + qlo[a][s] * sc[a] < (double)it.bound.lo[a] - 1e-3 * fabs(it.bound.lo[a]) - 1e-6) { /* quantised boxes may exceed the parent's; informational only */ }
            sah += box_half_area(qb);
            if (nd.imask & (1u << s)) {
                if ((meta >> 5) != 1u || (meta & 31u) != 24u + (uint32_t)s) return -4;
                stack.push_back({nd.child_base + ci++, qb, true, it.depth + 1});
                child_slots++;
            } else {
                const uint32_t cnt = popc(meta >> 5), off = meta & 31u;
                if (cnt == 0 || cnt > 3 || off + cnt > 24) return -5;
                leaf_slots++;
                for (uint32_t q = 0; q < cnt; q++) {
                    const uint32_t j = nd.prim_base + off + q;
                    if (j >= h->prim_count) return -6;
                    const char *rec = prims + (uint64_t)j * h->prim_stride;
                    Box pb = box_empty();
                    uint32_t pid;
                    if (h->kind == KIND_TRIANGLES) {
                        pb = record_box_tri(rec), pid = TriInput::record_prim(rec);
                    } else if (h->kind == KIND_AABBS) {
                        const f4 *r = reinterpret_cast<const f4 *>(rec);
                        pb.lo[0] = r[0].x, pb.lo[1] = r[0].y, pb.lo[2] = r[0].z, pb.hi[0] = r[1].x, pb.hi[1] = r[1].y, pb.hi[2] = r[1].z;
                        pid = AabbInput::record_prim(rec);
                    } else {
                        pid = InstInput::record_prim(rec);
                        pb = qb;
                    }
                    if (pid >= h->prim_count) return -7;
                    seen[pid]++;
                    for (int a = 0; a < 3; a++)
                        if (pb.lo[a] <= pb.hi[a] && (pb.lo[a] < qb.lo[a] || pb.hi[a] > qb.hi[a])) return -8;
                }
            }
        }
    }
    for (uint32_t i = 0; i < h->prim_count; i++)
        if (seen[i] != 1u) return -9;
    for (uint32_t i = 0; i < h->node_count; i++)
        if (node_seen[i] != 1u) return -10;
    if (h->depth != max_depth) return -16;// the recorded level count is the depth of the wide hierarchy
    if (stats) stats[0] = h->node_count, stats[1] = max_depth, stats[2] = leaf_slots, stats[3] = child_slots, stats[4] = (uint64_t)(sah * 1e3);
    return 0;
}

// ---- tracing with the device traversal code -----------------------------------------------------------------
// schedule of the two step halves: 0 = step() as the inline entry points do, 2 = run_voting() (RayQuery entry point), 1 = the batch kernel's: step_prim() only
// runs when the warp votes for it, which from one lane's point of view is a pseudo-random delay of its primitive tests
static int g_voted = 0;
void hs_set_defer(int on) { g_voted = on; }
extern "C++" template<class Tr>
void run_ray(Tr &tr, uint64_t i) {
    if (!g_voted) {
        while (tr.step()) {}
        return;
    }
    uint32_t lcg = (uint32_t)i * 2654435761u + 12345u;
    if (g_voted == 2) {
        // Traversal::run_voting() from one lane's point of view: each iteration the warp grants the node phase, the
        // primitive phase or both; a phase the lane does not get is a delay.  Bounded, so that a schedule under which a
        // lane can never progress fails the test instead of hanging it.
        bool alive = true;
        for (uint64_t it = 0; alive; it++) {
            if (it > (1ull << 26)) {
                fprintf(stderr, "hostsim: ray %llu makes no progress under the phase vote\n", (unsigned long long)i);
                abort();
            }
            lcg = lcg * 1664525u + 1013904223u;
            const bool run_node = (lcg >> 30) != 0u, run_prim = ((lcg >> 27) & 7u) < 5u || !run_node;
            if (run_node && !tr.prim_pending()) alive = tr.step_node();
            if (run_prim && alive && tr.prim_pending()) alive = tr.step_prim();
        }
        return;
    }
    while (tr.step_node()) {
        lcg = lcg * 1664525u + 1013904223u;
        if ((lcg >> 29) < 3u && !tr.step_prim()) break;
    }
}

extern "C++" template<class St>
void trace_closest_impl(const void *tlas, const void *instances, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride,
                               void *hits_, uint64_t *counters) {
    const Ray *rays = static_cast<const Ray *>(rays_);
    SurfaceHit *hits = static_cast<SurfaceHit *>(hits_);
    uint64_t cn = 0, cp = 0, ci = 0, ce = 0;
    for (uint64_t i = 0; i < n; i++) {
        St st;
        NoCallbacks cb;
        Traversal<MODE_CLOSEST, true, St, NoCallbacks> tr;
        tr.init(static_cast<const BlobHeader *>(tlas), static_cast<const Instance *>(instances), &rays[i], mask_stride ? masks[i] : masks[0], &st, &cb);
        run_ray(tr, i);
        SurfaceHit h;
        uint32_t type;
        tr.result(h.inst, h.prim, type, h.t, h.u, h.v);
        h.pad = 0;
        if (type == HIT_MISS) h.t = rays[i].tmax;
        hits[i] = h;
        cn += tr.ctr.nodes, cp += tr.ctr.prims, ci += tr.ctr.insts, ce += tr.ctr.exact;
    }
    if (counters) counters[0] = cn, counters[1] = cp, counters[2] = ci, counters[3] = ce;
}
void hs_trace_closest(const void *tlas, const void *instances, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride,
                      void *hits_, uint64_t *counters) {
    trace_closest_impl<LocalStore>(tlas, instances, rays_, n, masks, mask_stride, hits_, counters);
}

extern "C++" template<class St>
void trace_any_impl(const void *tlas, const void *instances, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride, uint32_t *out) {
    const Ray *rays = static_cast<const Ray *>(rays_);
    for (uint64_t i = 0; i < n; i++) {
        St st;
        NoCallbacks cb;
        Traversal<MODE_ANY, false, St, NoCallbacks> tr;
        tr.init(static_cast<const BlobHeader *>(tlas), static_cast<const Instance *>(instances), &rays[i], mask_stride ? masks[i] : masks[0], &st, &cb);
        run_ray(tr, i);
        out[i] = tr.has_hit() ? 1u : 0u;
    }
}
void hs_trace_any(const void *tlas, const void *instances, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride, uint32_t *out) {
    trace_any_impl<LocalStore>(tlas, instances, rays_, n, masks, mask_stride, out);
}

extern "C++" template<class St>
void query_impl(const void *tlas, const void *instances, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride,
                       int terminate_on_first, uint32_t tri_prog, uint32_t proc_prog, void *hits_) {
    const Ray *rays = static_cast<const Ray *>(rays_);
    CommittedHit *hits = static_cast<CommittedHit *>(hits_);
    for (uint64_t i = 0; i < n; i++) {
        St st;
        BuiltinPrograms cb{tri_prog, proc_prog};
        Traversal<MODE_QUERY, false, St, BuiltinPrograms> tr;
        tr.init(static_cast<const BlobHeader *>(tlas), static_cast<const Instance *>(instances), &rays[i], mask_stride ? masks[i] : masks[0], &st, &cb);
        tr.terminate_on_first = terminate_on_first != 0;
        run_ray(tr, i);
        CommittedHit h;
        tr.result(h.inst, h.prim, h.hit_type, h.t, h.u, h.v);
        if (h.hit_type == HIT_MISS) h.t = rays[i].tmax;
        hits[i] = h;
    }
}
void hs_query(const void *tlas, const void *instances, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride,
              int terminate_on_first, uint32_t tri_prog, uint32_t proc_prog, void *hits_) {
    query_impl<LocalStore>(tlas, instances, rays_, n, masks, mask_stride, terminate_on_first, tri_prog, proc_prog, hits_);
}

// The same query with the candidate programs run by the CALLER between two steps (SuspendingCallbacks / Traversal::resolve): the
// shape of the inline RayQuery lowering, `while (proceed) { handler }`.  Must produce what hs_query produces.
void hs_query_suspending(const void *tlas, const void *instances_, const void *rays_, uint64_t n, const uint32_t *masks, int mask_stride,
                         int terminate_on_first, uint32_t tri_prog, uint32_t proc_prog, void *hits_) {
    const Ray *rays = static_cast<const Ray *>(rays_);
    const Instance *instances = static_cast<const Instance *>(instances_);
    CommittedHit *hits = static_cast<CommittedHit *>(hits_);
    const BuiltinPrograms programs{tri_prog, proc_prog};
    for (uint64_t i = 0; i < n; i++) {
        LocalStore st;
        SuspendingCallbacks cb;
        cb.pend.kind = 0;
        Traversal<MODE_QUERY, false, LocalStore, SuspendingCallbacks> tr;
        tr.init(static_cast<const BlobHeader *>(tlas), instances, &rays[i], mask_stride ? masks[i] : masks[0], &st, &cb);
        tr.terminate_on_first = terminate_on_first != 0;
        while (tr.step()) {
            if (!tr.candidate_pending()) continue;
            const PendingCandidate pc = cb.pend;
            const Ray wr = tr.candidate_ray();
            CandidateResult r;
            if (pc.kind == 1) {
                r = programs.on_triangle(pc.inst, pc.prim, pc.u, pc.v, pc.t, wr);
            } else {
                const f4 *br = reinterpret_cast<const f4 *>(tr.prims + (uint64_t)pc.pi * sizeof(BoxRecord));
                r = programs.on_procedural(pc.inst, pc.prim, wr, instances, br[0], br[1]);
            }
            tr.resolve(r.committed, r.terminated, r.t_hit);
        }
        CommittedHit h;
        tr.result(h.inst, h.prim, h.hit_type, h.t, h.u, h.v);
        if (h.hit_type == HIT_MISS) h.t = rays[i].tmax;
        hits[i] = h;
    }
}

// The batch kernel's warp schedule (k_trace, lcrt_kernels.cu) replayed lane by lane: `warps` warps of 32 lanes share one ray
// cursor, reserve 128 rays at a time, refill idle lanes when at least 12 are idle, advance every live lane by one node step per
// iteration and run the primitive phase only when `vote` lanes (or every live lane) wait for it.  Returns the largest number
// of iterations in a row that a lane with a pending primitive was made to wait (the quantity LCRT_TRI_VOTE_PATIENCE bounds
// on the device), or -1 if a warp exceeds `max_iterations` (a lane starves: the schedule is not live).
long long hs_query_warp_schedule(const void *tlas, const void *instances_, const void *rays_, uint64_t n, uint32_t mask, int terminate_on_first,
                                 uint32_t tri_prog, uint32_t proc_prog, void *hits_, uint32_t warps, uint32_t vote, uint32_t refill_min,
                                 uint64_t max_iterations) {
    const Ray *rays = static_cast<const Ray *>(rays_);
    const Instance *instances = static_cast<const Instance *>(instances_);
    CommittedHit *hits = static_cast<CommittedHit *>(hits_);
    struct Lane {
        LocalStore st;
        BuiltinPrograms cb;
        Traversal<MODE_QUERY, false, LocalStore, BuiltinPrograms> tr;
        bool active = false, unsaved = false;
        uint64_t ray = 0;
    };
    struct Warp {
        std::vector<Lane> lane;
        uint64_t next = 0, end = 0, iterations = 0;
        bool exhausted = false, finished = false;
        uint32_t denied = 0;
    };
    std::vector<Warp> w(warps);
    for (auto &x : w) x.lane.resize(32);
    uint64_t cursor = 0;
    long long worst = 0;
    uint32_t running = warps;
    while (running) {
        for (auto &x : w) {
            if (x.finished) continue;
            if (++x.iterations > max_iterations) return -1;
            uint32_t idle = 0;
            for (int l = 0; l < 32; l++) idle += x.lane[l].active ? 0u : 1u;
            if (idle && ((!x.exhausted && idle >= refill_min) || idle == 32u)) {
                for (auto &L : x.lane)
                    if (L.unsaved) {
                        CommittedHit h;
                        L.tr.result(h.inst, h.prim, h.hit_type, h.t, h.u, h.v);
                        if (h.hit_type == HIT_MISS) h.t = rays[L.ray].tmax;
                        hits[L.ray] = h;
                        L.unsaved = false;
                    }
                if (x.exhausted) {
                    x.finished = true, running--;
                    continue;
                }
                if (x.next == x.end) {
                    const uint64_t base = cursor;
                    cursor += 128;
                    x.next = base < n ? base : n, x.end = base + 128 < n ? base + 128 : n;
                    if (x.next == x.end) x.exhausted = true;
                }
                const uint64_t avail = x.end - x.next;
                uint32_t rank = 0;
                for (auto &L : x.lane) {
                    if (L.active) continue;
                    if (rank < avail) {
                        L.ray = x.next + rank;
                        L.cb = BuiltinPrograms{tri_prog, proc_prog};
                        L.tr.init(static_cast<const BlobHeader *>(tlas), instances, &rays[L.ray], mask, &L.st, &L.cb);
                        L.tr.terminate_on_first = terminate_on_first != 0;
                        L.active = true;
                    }
                    rank++;
                }
                x.next += idle < avail ? idle : avail;
            }
            bool alive[32];
            uint32_t am = 0, pm = 0;
            for (int l = 0; l < 32; l++) {
                alive[l] = x.lane[l].active && x.lane[l].tr.step_node();
                if (alive[l]) am++;
                if (alive[l] && x.lane[l].tr.prim_pending()) pm++;
            }
            if (pm >= vote || pm == am) {
                for (int l = 0; l < 32; l++) alive[l] = alive[l] && x.lane[l].tr.step_prim();
                x.denied = 0;
            } else if (pm) {
                x.denied++;
                if ((long long)x.denied > worst) worst = x.denied;
            }
            for (int l = 0; l < 32; l++)
                if (x.lane[l].active && !alive[l]) x.lane[l].active = false, x.lane[l].unsaved = true;
        }
    }
    return worst;
}

}// extern "C"
