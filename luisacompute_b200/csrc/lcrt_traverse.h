// lcrt_traverse.h — two-level BVH8 traversal core (closest / any / RayQuery), host+device portable.
//
// Replaces what the reference delegates to OptiX: `_optix_hitobject_traverse`
// (src/backends/cuda/cuda_builtin/cuda_device_resource.h:1741-1781) and the IS/AH programs of the RayQuery path
// (:2146-2385).  Semantics follow cuda_device_resource.h:1783-1829 (intersect / intersect_any: all geometry opaque,
// `mask & 0xff` against the instance mask, no face culling, procedural primitives invisible) and :1896-1967
// (query_all / query_any).
//
// Algorithm: compressed-wide-BVH traversal (Ylitie et al. 2017): a node group = (child base, hit bits | imask),
// a primitive group = (prim base, hit bits); octant-ordered child visiting via `slot ^ octinv`; per-thread stack.
// Box tests are fp32 and conservative (never stricter than exact arithmetic); the triangle test is a conservative
// fp32 edge-function prefilter followed by a double-precision evaluation that decides the hit and produces
// (t, u, v).  Edge functions are evaluated as non-contracted scalar triple products d.(B x C), which are exactly
// antisymmetric under exchanging the two vertices of a shared edge => watertight.
#pragma once
#include "lcrt_common.h"

namespace lcrt {

#ifndef LCRT_STACK_SIZE
#define LCRT_STACK_SIZE 96
#endif

struct TraceCounters {
    uint32_t nodes, prims, insts;
};

enum TraceMode : int { MODE_CLOSEST = 0, MODE_ANY = 1, MODE_QUERY = 2 };

// Result of a candidate program (LCIntersectionResult, cuda_device_resource.h:2054-2058)
struct CandidateResult {
    float t_hit;
    bool committed;
    bool terminated;
};

// Default callbacks: never invoked in closest / any mode.
struct NoCallbacks {
    LCRT_HD CandidateResult on_triangle(uint32_t, uint32_t, float, float, float, const Ray &) const { return {0.f, true, false}; }
    LCRT_HD CandidateResult on_procedural(uint32_t, uint32_t, const Ray &, const Instance *, const f4, const f4) const { return {0.f, false, false}; }
};

// Simple local-memory stack (inline device-library path and host simulation).
struct LocalStack {
    u2 e[LCRT_STACK_SIZE];
    LCRT_HD void store(int i, u2 v) { e[i] = v; }
    LCRT_HD u2 load(int i) const { return e[i]; }
};

LCRT_HD float safe_dir(float d) {
    // avoid 0 * inf in the slab test; 1e-20 keeps 255 * 2^e * idir finite for any sane scene
    return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
}

// --------------------------------------------------------------------------------------------------------------
// Node test: returns the hit bits (top byte: internal children by traversal priority, low 24: primitives)
// --------------------------------------------------------------------------------------------------------------
LCRT_HD uint32_t intersect_node8(const u4 n0, const u4 n1, const u4 n2, const u4 n3, const u4 n4, float ox, float oy, float oz,
                                 float idx, float idy, float idz, float tmin, float tmax, uint32_t octinv4) {
    const float sx = as_float((n0.w & 0xffu) << 23), sy = as_float(((n0.w >> 8) & 0xffu) << 23),
                sz = as_float(((n0.w >> 16) & 0xffu) << 23);
    // t = q * a + b; (a, b) widened by K so that the near (far) distance is a lower (upper) bound of the exact one
    constexpr float K = 6.0e-7f;
    const float ax = sx * idx, ay = sy * idy, az = sz * idz;
    const float bx = (as_float(n0.x) - ox) * idx, by = (as_float(n0.y) - oy) * idy, bz = (as_float(n0.z) - oz) * idz;
    const float axn = fmaf(-K, fabsf(ax), ax), ayn = fmaf(-K, fabsf(ay), ay), azn = fmaf(-K, fabsf(az), az);
    const float axf = fmaf(K, fabsf(ax), ax), ayf = fmaf(K, fabsf(ay), ay), azf = fmaf(K, fabsf(az), az);
    const float bxn = fmaf(-K, fabsf(bx), bx), byn = fmaf(-K, fabsf(by), by), bzn = fmaf(-K, fabsf(bz), bz);
    const float bxf = fmaf(K, fabsf(bx), bx), byf = fmaf(K, fabsf(by), by), bzf = fmaf(K, fabsf(bz), bz);
    uint32_t hitmask = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 2; i++) {
        const uint32_t meta4 = i ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = i ? n2.y : n2.x, qloy = i ? n2.w : n2.z, qloz = i ? n3.y : n3.x;
        const uint32_t qhix = i ? n3.w : n3.z, qhiy = i ? n4.y : n4.x, qhiz = i ? n4.w : n4.z;
        const uint32_t nx = idx < 0.f ? qhix : qlox, fx = idx < 0.f ? qlox : qhix;
        const uint32_t ny = idy < 0.f ? qhiy : qloy, fy = idy < 0.f ? qloy : qhiy;
        const uint32_t nz = idz < 0.f ? qhiz : qloz, fz = idz < 0.f ? qloz : qhiz;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; j++) {
            const float tnx = fmaf((float)extract_byte(nx, j), axn, bxn);
            const float tny = fmaf((float)extract_byte(ny, j), ayn, byn);
            const float tnz = fmaf((float)extract_byte(nz, j), azn, bzn);
            const float tfx = fmaf((float)extract_byte(fx, j), axf, bxf);
            const float tfy = fmaf((float)extract_byte(fy, j), ayf, byf);
            const float tfz = fmaf((float)extract_byte(fz, j), azf, bzf);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            if (tn <= tf) hitmask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
        }
    }
    return hitmask;
}

// --------------------------------------------------------------------------------------------------------------
// Triangle test
// --------------------------------------------------------------------------------------------------------------
// fp32 conservative prefilter: false => the ray certainly misses (in exact arithmetic).
// `oerr`: absolute error bound of the (fp32, possibly instance-transformed) ray origin.
LCRT_HD bool tri_prefilter(float ox, float oy, float oz, float dx, float dy, float dz, float oerr, const f4 v0, const f4 v1, const f4 v2) {
    const float Ax = v0.x - ox, Ay = v0.y - oy, Az = v0.z - oz;
    const float Bx = v1.x - ox, By = v1.y - oy, Bz = v1.z - oz;
    const float Cx = v2.x - ox, Cy = v2.y - oy, Cz = v2.z - oz;
    const float U = dx * (By * Cz - Bz * Cy) + dy * (Bz * Cx - Bx * Cz) + dz * (Bx * Cy - By * Cx);
    const float V = dx * (Cy * Az - Cz * Ay) + dy * (Cz * Ax - Cx * Az) + dz * (Cx * Ay - Cy * Ax);
    const float W = dx * (Ay * Bz - Az * By) + dy * (Az * Bx - Ax * Bz) + dz * (Ax * By - Ay * Bx);
    const float bm = fmaxf(fmaxf(fmaxf(fabsf(Ax), fabsf(Ay)), fmaxf(fabsf(Az), fabsf(Bx))),
                           fmaxf(fmaxf(fmaxf(fabsf(By), fabsf(Bz)), fmaxf(fabsf(Cx), fabsf(Cy))), fabsf(Cz)));
    const float dm = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
    // |fl(U) - U| <= ~12 eps * dm * bm^2 (three products of three factors, two levels of sums); use 32 eps
    const float E = dm * bm * fmaf(3.9e-6f, bm, 12.f * oerr);
    const bool neg = (U < -E) | (V < -E) | (W < -E);
    const bool pos = (U > E) | (V > E) | (W > E);
    return !(neg & pos);
}

// Double-precision evaluation: decides the hit exactly like the oracle's formula (oracle/oracle.cpp tri_hit).
LCRT_HD bool tri_exact(double ox, double oy, double oz, double dx, double dy, double dz, const f4 v0, const f4 v1, const f4 v2,
                       double &t, double &u, double &v) {
    const double Ax = (double)v0.x - ox, Ay = (double)v0.y - oy, Az = (double)v0.z - oz;
    const double Bx = (double)v1.x - ox, By = (double)v1.y - oy, Bz = (double)v1.z - oz;
    const double Cx = (double)v2.x - ox, Cy = (double)v2.y - oy, Cz = (double)v2.z - oz;
#if defined(__CUDA_ARCH__)
#define LCRT_DM(a, b) __dmul_rn(a, b)
#define LCRT_DS(a, b) __dsub_rn(a, b)
#define LCRT_DA(a, b) __dadd_rn(a, b)
#else
#define LCRT_DM(a, b) ((a) * (b))
#define LCRT_DS(a, b) ((a) - (b))
#define LCRT_DA(a, b) ((a) + (b))
#endif
    const double U = LCRT_DA(LCRT_DA(LCRT_DM(dx, LCRT_DS(LCRT_DM(By, Cz), LCRT_DM(Bz, Cy))), LCRT_DM(dy, LCRT_DS(LCRT_DM(Bz, Cx), LCRT_DM(Bx, Cz)))),
                             LCRT_DM(dz, LCRT_DS(LCRT_DM(Bx, Cy), LCRT_DM(By, Cx))));
    const double V = LCRT_DA(LCRT_DA(LCRT_DM(dx, LCRT_DS(LCRT_DM(Cy, Az), LCRT_DM(Cz, Ay))), LCRT_DM(dy, LCRT_DS(LCRT_DM(Cz, Ax), LCRT_DM(Cx, Az)))),
                             LCRT_DM(dz, LCRT_DS(LCRT_DM(Cx, Ay), LCRT_DM(Cy, Ax))));
    const double W = LCRT_DA(LCRT_DA(LCRT_DM(dx, LCRT_DS(LCRT_DM(Ay, Bz), LCRT_DM(Az, By))), LCRT_DM(dy, LCRT_DS(LCRT_DM(Az, Bx), LCRT_DM(Ax, Bz)))),
                             LCRT_DM(dz, LCRT_DS(LCRT_DM(Ax, By), LCRT_DM(Ay, Bx))));
#undef LCRT_DM
#undef LCRT_DS
#undef LCRT_DA
    const double det = U + V + W;
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    const double b0 = U * id, b1 = V * id, b2 = W * id;
    if (b0 < 0.0 || b1 < 0.0 || b2 < 0.0) return false;
    const double e1x = (double)v1.x - (double)v0.x, e1y = (double)v1.y - (double)v0.y, e1z = (double)v1.z - (double)v0.z;
    const double e2x = (double)v2.x - (double)v0.x, e2y = (double)v2.y - (double)v0.y, e2z = (double)v2.z - (double)v0.z;
    const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const double dn = dx * nx + dy * ny + dz * nz;
    if (dn == 0.0) return false;
    t = (Ax * nx + Ay * ny + Az * nz) / dn;
    u = b1, v = b2;
    return true;
}

// conservative fp32 slab test against an explicit box (procedural primitives)
LCRT_HD bool slab_conservative(float lox, float loy, float loz, float hix, float hiy, float hiz, float ox, float oy, float oz, float idx,
                               float idy, float idz, float tmin, float tmax) {
    constexpr float K = 6.0e-7f;
    float t0x = (lox - ox) * idx, t1x = (hix - ox) * idx;
    float t0y = (loy - oy) * idy, t1y = (hiy - oy) * idy;
    float t0z = (loz - oz) * idz, t1z = (hiz - oz) * idz;
    float nx = fminf(t0x, t1x), fx = fmaxf(t0x, t1x), ny = fminf(t0y, t1y), fy = fmaxf(t0y, t1y), nz = fminf(t0z, t1z),
          fz = fmaxf(t0z, t1z);
    nx = fmaf(-K, fabsf(nx), nx), ny = fmaf(-K, fabsf(ny), ny), nz = fmaf(-K, fabsf(nz), nz);
    fx = fmaf(K, fabsf(fx), fx), fy = fmaf(K, fabsf(fy), fy), fz = fmaf(K, fabsf(fz), fz);
    const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float tf = fminf(fminf(fx, fy), fminf(fz, tmax));
    return tn <= tf;
}

// --------------------------------------------------------------------------------------------------------------
// Traversal state machine.  `step()` performs one outer iteration (one node, then the primitives it exposed);
// the persistent batch kernel interleaves step() with warp-cooperative ray refetch, the inline device-library
// entry points simply loop until done.
// --------------------------------------------------------------------------------------------------------------
template<int MODE, bool COUNT, class Stack, class Callbacks>
struct Traversal {
    // current-space ray
    float ox, oy, oz, dx, dy, dz, idx, idy, idz;
    float tmin;
    float tmax;   // fp32 upper bound of t_best (rounded up), used for box culling
    double t_best;// exact comparison value
    uint32_t octinv4;
    // level state
    const u4 *nodes;
    const char *prims;
    const u4 *tlas_nodes;
    const InstanceRecord *tlas_recs;
    u2 G, Gt;
    int sp, blas_sp;// blas_sp < 0: traversing the TLAS
    uint32_t cur_inst, cur_flags, ray_mask;
    const InstanceRecord *cur_rec;
    float oerr;// error bound of the object-space origin (fp32 transform), widens the triangle prefilter
    // result
    uint32_t hit_inst, hit_prim, hit_type;
    float hit_u, hit_v;
    bool done;
    Ray world;// kept for instance exit and for the RayQuery callbacks
    const Instance *instances;
    Stack *stack;
    Callbacks *cb;
    TraceCounters ctr;

    LCRT_HD void set_ray(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_) {
        ox = ox_, oy = oy_, oz = oz_, dx = dx_, dy = dy_, dz = dz_;
        const float sdx = safe_dir(dx_), sdy = safe_dir(dy_), sdz = safe_dir(dz_);
        idx = 1.0f / sdx, idy = 1.0f / sdy, idz = 1.0f / sdz;
        const uint32_t oct = (sdx < 0.f ? 1u : 0u) | (sdy < 0.f ? 2u : 0u) | (sdz < 0.f ? 4u : 0u);
        octinv4 = (7u - oct) * 0x01010101u;
    }

    LCRT_HD void init(const BlobHeader *tlas, const Instance *instances_, const Ray &ray, uint32_t mask, Stack *stack_, Callbacks *cb_) {
        world = ray;
        instances = instances_;
        stack = stack_, cb = cb_;
        tlas_nodes = reinterpret_cast<const u4 *>(reinterpret_cast<const char *>(tlas) + tlas->nodes_offset);
        tlas_recs = reinterpret_cast<const InstanceRecord *>(reinterpret_cast<const char *>(tlas) + tlas->prims_offset);
        nodes = tlas_nodes, prims = reinterpret_cast<const char *>(tlas_recs);
        set_ray(ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz);
        tmin = ray.tmin, tmax = ray.tmax, t_best = (double)ray.tmax;
        ray_mask = mask & 0xffu;
        G = make_u2(0u, 0x80000000u), Gt = make_u2(0u, 0u);
        sp = 0, blas_sp = -1;
        cur_inst = ~0u, cur_flags = 0u, cur_rec = nullptr, oerr = 0.f;
        hit_inst = ~0u, hit_prim = ~0u, hit_type = HIT_MISS, hit_u = 0.f, hit_v = 0.f;
        ctr.nodes = ctr.prims = ctr.insts = 0u;
        // NaN / empty-interval rays never hit anything (Embree: tnear <= t <= tfar)
        done = !(ray.tmin <= ray.tmax);
    }

    LCRT_HD void accept(double t, float u, float v, uint32_t prim, uint32_t type) {
        t_best = t;
        // round the culling bound up so that the box test never rejects the accepted primitive's neighbours
        float tf = (float)t;
        if ((double)tf < t) tf = as_float(as_uint(tf) + (tf >= 0.f ? 1u : 0xffffffffu));
        tmax = tf;
        hit_inst = cur_inst, hit_prim = prim, hit_u = u, hit_v = v, hit_type = type;
    }

    LCRT_HD void leave_instance() {
        blas_sp = -1;
        nodes = tlas_nodes, prims = reinterpret_cast<const char *>(tlas_recs);
        set_ray(world.ox, world.oy, world.oz, world.dx, world.dy, world.dz);
        cur_inst = ~0u;
    }

    LCRT_HD void enter_instance(const InstanceRecord *rec) {
        // keep what is left of the TLAS groups below the BLAS frames
        if (G.y & 0xff000000u) stack->store(sp++, G);
        if (Gt.y) stack->store(sp++, Gt);
        Gt = make_u2(0u, 0u);
        blas_sp = sp;
        const double *m = rec->inv;
        const float i00 = (float)m[0], i01 = (float)m[1], i02 = (float)m[2], i03 = (float)m[3];
        const float i10 = (float)m[4], i11 = (float)m[5], i12 = (float)m[6], i13 = (float)m[7];
        const float i20 = (float)m[8], i21 = (float)m[9], i22 = (float)m[10], i23 = (float)m[11];
        const float lox = fmaf(i00, world.ox, fmaf(i01, world.oy, fmaf(i02, world.oz, i03)));
        const float loy = fmaf(i10, world.ox, fmaf(i11, world.oy, fmaf(i12, world.oz, i13)));
        const float loz = fmaf(i20, world.ox, fmaf(i21, world.oy, fmaf(i22, world.oz, i23)));
        const float ldx = fmaf(i00, world.dx, fmaf(i01, world.dy, i02 * world.dz));
        const float ldy = fmaf(i10, world.dx, fmaf(i11, world.dy, i12 * world.dz));
        const float ldz = fmaf(i20, world.dx, fmaf(i21, world.dy, i22 * world.dz));
        set_ray(lox, loy, loz, ldx, ldy, ldz);
        // |fl(o') - o'| <= ~4 eps * sum |inv_ij o_j| + |inv_i3|
        const float ax = fabsf(world.ox), ay = fabsf(world.oy), az = fabsf(world.oz);
        const float mx = fmaf(fabsf(i00), ax, fmaf(fabsf(i01), ay, fmaf(fabsf(i02), az, fabsf(i03))));
        const float my = fmaf(fabsf(i10), ax, fmaf(fabsf(i11), ay, fmaf(fabsf(i12), az, fabsf(i13))));
        const float mz = fmaf(fabsf(i20), ax, fmaf(fabsf(i21), ay, fmaf(fabsf(i22), az, fabsf(i23))));
        oerr = 6.0e-7f * fmaxf(mx, fmaxf(my, mz));
        cur_rec = rec;
        nodes = reinterpret_cast<const u4 *>(rec->nodes);
        prims = reinterpret_cast<const char *>(rec->prims);
        cur_inst = rec->inst, cur_flags = rec->flags;
        G = make_u2(0u, 0x80000000u);
        if (COUNT) ctr.insts++;
    }

    // returns true while there is more work
    LCRT_HD bool step() {
        if (done) return false;
        if (G.y & 0xff000000u) {
            // ---- one internal node
            const int child_bit = bfind(G.y);
            const uint32_t base = G.x;
            G.y &= ~(1u << child_bit);
            const uint32_t imask_parent = G.y & 0xffu;
            if (G.y & 0xff000000u) stack->store(sp++, G);
            const uint32_t slot = (uint32_t)(child_bit - 24) ^ (octinv4 & 0xffu);
            const uint32_t rel = (uint32_t)popc(imask_parent & ~(0xffffffffu << slot));
            const u4 *n = nodes + (uint64_t)(base + rel) * 5u;
            const u4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3], n4 = n[4];
            if (COUNT) ctr.nodes++;
            const uint32_t hits = intersect_node8(n0, n1, n2, n3, n4, ox, oy, oz, idx, idy, idz, tmin, tmax, octinv4);
            G = make_u2(n1.x, (hits & 0xff000000u) | (n0.w >> 24));
            Gt = make_u2(n1.y, hits & 0x00ffffffu);
        } else {
            Gt = G;
            G = make_u2(0u, 0u);
        }
        // ---- primitives exposed by that node
        while (Gt.y) {
            const int bit = bfind(Gt.y);
            Gt.y &= ~(1u << bit);
            const uint32_t pi = Gt.x + (uint32_t)bit;
            if (blas_sp < 0) {
                // TLAS leaf: an instance
                const InstanceRecord *rec = tlas_recs + pi;
                const u4 r4 = reinterpret_cast<const u4 *>(rec)[7];// { inst, mask, flags, pad }
                const uint32_t kind = r4.z & IREC_KIND_MASK;
                if ((r4.y & ray_mask) == 0u) continue;
                if (MODE != MODE_QUERY && kind != KIND_TRIANGLES) continue;// intersect() does not see procedural prims
                enter_instance(rec);
                break;
            }
            if (COUNT) ctr.prims++;
            if ((cur_flags & IREC_KIND_MASK) == KIND_TRIANGLES) {
                const f4 *tr = reinterpret_cast<const f4 *>(prims + (uint64_t)pi * sizeof(TriRecord));
                const f4 v0 = tr[0], v1 = tr[1], v2 = tr[2];
                if (!tri_prefilter(ox, oy, oz, dx, dy, dz, oerr, v0, v1, v2)) continue;
                double t, u, v;
                {
                    // exact test on the object-space ray re-derived in double from the world ray
                    const double *m = cur_rec->inv;
                    const double wox = world.ox, woy = world.oy, woz = world.oz, wdx = world.dx, wdy = world.dy, wdz = world.dz;
                    const double eox = m[0] * wox + m[1] * woy + m[2] * woz + m[3], eoy = m[4] * wox + m[5] * woy + m[6] * woz + m[7],
                                 eoz = m[8] * wox + m[9] * woy + m[10] * woz + m[11];
                    const double edx = m[0] * wdx + m[1] * wdy + m[2] * wdz, edy = m[4] * wdx + m[5] * wdy + m[6] * wdz,
                                 edz = m[8] * wdx + m[9] * wdy + m[10] * wdz;
                    if (!tri_exact(eox, eoy, eoz, edx, edy, edz, v0, v1, v2, t, u, v)) continue;
                }
                if (t < (double)tmin || t > t_best) continue;// Embree: tnear <= t <= tfar
                const uint32_t prim = as_uint(v0.w);
                // exact ties are resolved towards the lower (inst, prim) so the result does not depend on traversal order
                if (t == t_best && hit_inst != ~0u && (cur_inst > hit_inst || (cur_inst == hit_inst && prim >= hit_prim))) continue;
                if (MODE == MODE_QUERY && !(cur_flags & IREC_OPAQUE)) {
                    Ray wr = world;
                    wr.tmax = (float)t;// accel.rs:663-672: tmax shrinks to the candidate while the callback runs
                    const CandidateResult r = cb->on_triangle(cur_inst, prim, (float)u, (float)v, (float)t, wr);
                    if (r.committed) accept(t, (float)u, (float)v, prim, HIT_SURFACE);
                    if (r.terminated || (r.committed && MODE_ANY_QUERY())) {
                        done = true;
                        return false;
                    }
                    continue;
                }
                accept(t, (float)u, (float)v, prim, HIT_SURFACE);
                if (MODE == MODE_ANY || MODE_ANY_QUERY()) {
                    done = true;
                    return false;
                }
            } else if (MODE == MODE_QUERY) {
                // procedural primitive: conservative box test, then the candidate program decides
                const f4 *br = reinterpret_cast<const f4 *>(prims + (uint64_t)pi * sizeof(BoxRecord));
                const f4 lo = br[0], hi = br[1];
                if (!slab_conservative(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, ox, oy, oz, idx, idy, idz, tmin, tmax)) continue;
                Ray wr = world;
                wr.tmax = tmax;
                const uint32_t prim = as_uint(lo.w);
                const CandidateResult r = cb->on_procedural(cur_inst, prim, wr, instances, lo, hi);
                if (r.committed && r.t_hit >= tmin && (double)r.t_hit < t_best) {// accel.rs:711-713
                    accept((double)r.t_hit, 0.f, 0.f, prim, HIT_PROCEDURAL);
                    if (MODE_ANY_QUERY()) {
                        done = true;
                        return false;
                    }
                }
                if (r.terminated) {
                    done = true;
                    return false;
                }
            }
        }
        // ---- pop
        if ((G.y & 0xff000000u) == 0u && Gt.y == 0u) {
            if (blas_sp >= 0 && sp == blas_sp) leave_instance();
            if (sp == 0) {
                done = true;
                return false;
            }
            G = stack->load(--sp);
        }
        return true;
    }

    bool terminate_on_first = false;// query_any
    LCRT_HD bool MODE_ANY_QUERY() const { return MODE == MODE_QUERY && terminate_on_first; }
};

}// namespace lcrt
