// lcrt_traverse.h — two-level BVH8 traversal core (closest / any / RayQuery), host+device portable.
//
// Replaces what the reference delegates to OptiX: `_optix_hitobject_traverse`
// (src/backends/cuda/cuda_builtin/cuda_device_resource.h:1741-1781) and the IS/AH programs of the RayQuery path
// (:2146-2385).  Semantics follow cuda_device_resource.h:1783-1829 (intersect / intersect_any: all geometry opaque,
// `mask & 0xff` against the instance mask, no face culling, procedural primitives invisible) and :1896-1967
// (query_all / query_any).
//
// Algorithm: compressed-wide-BVH traversal (Ylitie et al. 2017).  A node group = (child base, hit bits | imask), a
// primitive group = (node index, hit leaf slots); octant-ordered child visiting via `slot ^ octinv`; per-thread stack.
// One step() = at most one node test followed by at most one primitive test, so lanes of a warp that are in
// different phases of their rays share both code blocks.
//
// Arithmetic, tuned to the measured sm_100a issue costs (tools/ubench_alu.cu: FFMA/FMNMX/FSETP 1 cycle per warp
// instruction, LOP3/PRMT/SHF/SEL 2, I2F.U8/FLO/POPC 8):
//   * box test: the quantised plane byte is planted into the mantissa of 2^23 by ONE PRMT (val = 2^23 + 256 q) and
//     the plane distance is ONE FFMA val * a' + b' with the 2^23 offset folded into b'; no integer-to-float
//     conversion, no per-child bit arithmetic (the node test returns an 8-bit slot mask).  Conservative: the
//     (a', b') pairs are widened so that the near (far) distance is a lower (upper) bound of the exact one.
//   * triangle test: fp32 edge functions in a ray-aligned sheared frame (after Woop, Benthin, Wald 2013) with an
//     explicit forward error bound.  A decision is taken in fp32 only when every edge function clears its bound;
//     the fp32 distance comes with its own error bound and candidates are ordered by interval comparison.  Whatever
//     fp32 can not decide (ray within rounding of an edge, two surfaces within rounding of each other, t within
//     rounding of the ray's interval ends) is re-evaluated by tri_exact_world() in double precision with the
//     oracle's formula (non-contracted scalar triple products, exactly antisymmetric under exchanging the two
//     vertices of a shared edge => watertight).  Decisions therefore equal those of an all-double evaluation.
//   * the reported (t, u, v) of the winning triangle are always recomputed in double (refine_hit): fp32 can not
//     deliver 1e-4 barycentrics for a 4 mm triangle seen from 2 m.  The batch kernels do that in a separate,
//     fully convergent pass over the hit records (k_refine_hits); the inline entry points do it at the end.
#pragma once
#include "lcrt_common.h"

namespace lcrt {

#ifndef LCRT_STACK_SIZE
#define LCRT_STACK_SIZE 96
#endif
#ifndef LCRT_RUN_VOTE
#define LCRT_RUN_VOTE 1// Traversal::run_voting(): the lanes of a warp vote on the node / primitive phase
#endif
#ifndef LCRT_RUN_NODE_VOTE
#define LCRT_RUN_NODE_VOTE 12
#endif
#ifndef LCRT_RUN_PRIM_VOTE
#define LCRT_RUN_PRIM_VOTE 1// 1: the primitive phase runs whenever a lane waits for it (C5: 1 -> 420, 4 -> 409, 8 -> 407 spp/s)
#endif

struct TraceCounters {
    uint32_t nodes, prims, insts, exact;
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

// A RayQuery whose candidate handlers run in the CALLER's scope (the inline lowering of RayQueryStmt, SURVEY.md 7.7: `while
// (proceed) { handler }` as in src/backends/common/hlsl/hlsl_codegen.cpp:555-567): a callbacks type that declares
// `suspends` is never called — the traversal records the candidate in its `pend` member and returns from the step; the caller
// runs the handler and hands the verdict back through Traversal::resolve().
struct PendingCandidate {
    int kind;// 0 none, 1 triangle (non-opaque instance), 2 procedural primitive
    uint32_t inst, prim, pi;
    float t, terr, u, v;// triangle: the candidate hit
    float ht;           // what the handler sees as the ray's t_max: the candidate's t (triangle), the committed distance (procedural)
    bool exact;
    double td;
};
struct SuspendingCallbacks {
    static constexpr bool suspends = true;
    PendingCandidate pend;
    LCRT_HD CandidateResult on_triangle(uint32_t, uint32_t, float, float, float, const Ray &) const { return {0.f, false, false}; }
    LCRT_HD CandidateResult on_procedural(uint32_t, uint32_t, const Ray &, const Instance *, const f4, const f4) const { return {0.f, false, false}; }
};
template<class C>
struct callbacks_suspend {
    template<class T> static char test(decltype(T::suspends) *);
    template<class T> static long test(...);
    static constexpr bool value = sizeof(test<C>(nullptr)) == sizeof(char);
};

// Per-ray storage outside the register file: the traversal stack plus the "cold" ray state that is touched once per
// primitive test or once per accepted hit.  The batch kernel keeps the first stack entries and all cold state in
// shared memory (lcrt_kernels.cu HybridStore); this plain version serves the host simulation and the inline
// device-library entry points.
enum ColdWord : int {
    CW_SZ = 0,    // 1 / d_kz of the current-space direction
    CW_KZ,        // dominant axis of the current-space direction
    CW_REC,       // TLAS record index of the instance being traversed
    CW_HIT_T,     // best accepted t (fp32)
    CW_HIT_TERR,  // its absolute error bound; < 0: no hit yet (CW_HIT_T is then the ray's tmax, exact)
    CW_HIT_REC,   // TLAS record index of the best hit
    CW_HIT_PI,    // primitive record index of the best hit inside its BLAS
    CW_HIT_INST,  // instance / primitive ids of the best hit (tie-breaks)
    CW_HIT_PRIM,
    CW_HIT_TYPE,
    CW_HIT_TD_LO, // exact t of the best hit as a double, valid iff CW_HIT_EXACT
    CW_HIT_TD_HI,
    CW_HIT_EXACT,
    CW_COUNT
};
struct LocalStore {
    static constexpr bool world_in_global = false;
    u2 e[LCRT_STACK_SIZE];
    f4 rows[2];
    uint32_t words[CW_COUNT];
    LCRT_HD void store(int i, u2 v) { e[i] = v; }
    LCRT_HD u2 load(int i) const { return e[i]; }
    LCRT_HD void st4(int k, f4 v) { rows[k] = v; }
    LCRT_HD f4 ld4(int k) const { return rows[k]; }
    LCRT_HD void stw(int k, uint32_t v) { words[k] = v; }
    LCRT_HD uint32_t ldw(int k) const { return words[k]; }
};

LCRT_HD float safe_dir(float d) {
    // avoid 0 * inf in the slab test; 1e-20 keeps 2^15 * 255 * 2^e * idir finite for any sane scene
    return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
}

// reciprocal within 1 ulp (MUFU.RCP on the device); used only where the error analysis budgets 2 u for it
LCRT_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// 16-byte read-only loads of BVH data (never written while a trace is in flight)
template<class T>
LCRT_HD T ldg16(const T *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// rays are read exactly once per batch: streaming loads keep them from pushing node / triangle lines out of L2
template<class T>
LCRT_HD T ld16_stream(const T *p) {
#if defined(__CUDA_ARCH__) && !defined(LCRT_NO_STREAM_RAYS)
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    T r;
    memcpy(&r, &v, 16);
    return r;
#else
    return ldg16(p);
#endif
}

// float whose mantissa carries byte `j` of `w` in bits 8..15: 2^23 + 256 * byte
// `magic` must hold PLANT_MAGIC; it is passed in (from a kernel parameter in the batch kernels, from the TLAS header in the
// inline entry points) so that it lives in a register and the selector can be the instruction's immediate — as a literal, ptxas makes it the immediate and re-materialises
// the selector into a register for every PRMT.
constexpr uint32_t PLANT_MAGIC = 0x4B000000u;
template<int J>
LCRT_HD float plant_byte(uint32_t w, uint32_t magic) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7604 | (J << 4)));
    return __uint_as_float(r);
#else
    return as_float(magic | (((w >> (8 * J)) & 0xffu) << 8));
#endif
}

// --------------------------------------------------------------------------------------------------------------
// Node test: returns the 8-bit mask of child slots whose (conservatively widened) box the ray interval overlaps.
// Empty slots are stored inverted (lo 255, hi 0) and fail unless the node is degenerate; callers mask the result
// with imask (internal children) or check meta != 0 (leaf slots).
// --------------------------------------------------------------------------------------------------------------
struct AxisCoef {
    float an, bn, af, bf;// t_near(q) >= val(q) * an + bn ; t_far(q) <= val(q) * af + bf
};
LCRT_HD AxisCoef axis_coef(uint32_t ebyte, float p, float o, float id) {
    // exact: t(q) = q * a + b, a = 2^(e-127) / d, b = (p - o) / d.  Rounding of (1/d, p - o, the products, the final
    // FFMA) is covered by widening a and b by K = 10 ulp relative (as before); planting q as 2^23 + 256 q turns the
    // FFMA into val * (a / 256) + (b - 2^15 a), whose extra cancellation error (<= 2^15 |a| * 2^-24 = |a| / 512 per
    // rounding) is covered by a slack of |a| / 128, i.e. less than 1 % of one quantisation step.
    constexpr float K = 6.0e-7f;
    const float s = as_float(ebyte << 23);
    const float a256 = (s * id) * (1.0f / 256.0f);
    const float b = (p - o) * id;
    AxisCoef c;
    c.an = fmaf(-K, fabsf(a256), a256);
    c.af = fmaf(K, fabsf(a256), a256);
    c.bn = fmaf(-2.0f, fabsf(a256), fmaf(-8388608.0f, c.an, fmaf(-K, fabsf(b), b)));
    c.bf = fmaf(2.0f, fabsf(a256), fmaf(-8388608.0f, c.af, fmaf(K, fabsf(b), b)));
    return c;
}

LCRT_HD uint32_t intersect_node8(const u4 n0, const u4 n2, const u4 n3, const u4 n4, float ox, float oy, float oz, float idx, float idy,
                                 float idz, float tmin, float tmax, uint32_t magic) {
    const AxisCoef cx = axis_coef(n0.w & 0xffu, as_float(n0.x), ox, idx);
    const AxisCoef cy = axis_coef((n0.w >> 8) & 0xffu, as_float(n0.y), oy, idy);
    const AxisCoef cz = axis_coef((n0.w >> 16) & 0xffu, as_float(n0.z), oz, idz);
    uint32_t hm = 0u;
#define LCRT_CHILD(J, SLOT)                                                                                \
    {                                                                                                      \
        const float tnx = fmaf(plant_byte<J>(nx, magic), cx.an, cx.bn), tfx = fmaf(plant_byte<J>(fx, magic), cx.af, cx.bf); \
        const float tny = fmaf(plant_byte<J>(ny, magic), cy.an, cy.bn), tfy = fmaf(plant_byte<J>(fy, magic), cy.af, cy.bf); \
        const float tnz = fmaf(plant_byte<J>(nz, magic), cz.an, cz.bn), tfz = fmaf(plant_byte<J>(fz, magic), cz.af, cz.bf); \
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                         \
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));                                         \
        if (tn <= tf) hm += 1u << (SLOT);                                                                  \
    }
    {
        const uint32_t nx = idx < 0.f ? n3.z : n2.x, fx = idx < 0.f ? n2.x : n3.z;
        const uint32_t ny = idy < 0.f ? n4.x : n2.z, fy = idy < 0.f ? n2.z : n4.x;
        const uint32_t nz = idz < 0.f ? n4.z : n3.x, fz = idz < 0.f ? n3.x : n4.z;
        LCRT_CHILD(0, 0) LCRT_CHILD(1, 1) LCRT_CHILD(2, 2) LCRT_CHILD(3, 3)
    }
    {
        const uint32_t nx = idx < 0.f ? n3.w : n2.y, fx = idx < 0.f ? n2.y : n3.w;
        const uint32_t ny = idy < 0.f ? n4.y : n2.w, fy = idy < 0.f ? n2.w : n4.y;
        const uint32_t nz = idz < 0.f ? n4.w : n3.y, fz = idz < 0.f ? n3.y : n4.w;
        LCRT_CHILD(0, 4) LCRT_CHILD(1, 5) LCRT_CHILD(2, 6) LCRT_CHILD(3, 7)
    }
#undef LCRT_CHILD
    return hm;
}

// permute the 8 slot bits by slot -> slot ^ k (k = 7 - octant): the traversal visits the highest set bit first
LCRT_HD uint32_t permute_slots(uint32_t m, uint32_t k) {
    if (k & 1u) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
    if (k & 2u) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
    if (k & 4u) m = ((m & 0x0fu) << 4) | ((m >> 4) & 0x0fu);
    return m;
}

// --------------------------------------------------------------------------------------------------------------
// Triangle test
// --------------------------------------------------------------------------------------------------------------
constexpr float LCRT_U = 5.9604645e-8f;// 2^-24: unit roundoff of fp32 (round to nearest)

enum TriVerdict : int { TRI_MISS = 0, TRI_HIT = 1, TRI_UNSURE = 2 };

// fp32 test in the sheared frame.  r1 / r2 are the rows of the shear (entries 1, 0 and -d_kx/d_kz resp. -d_ky/d_kz
// with kz the dominant axis of d), so (A.r1, A.r2) are the coordinates of A in the plane z' = 0 of a frame whose z'
// axis is the ray; the 2D edge functions equal the exact scalar triple products d.(B x C) / d_kz.
//   r1.w = 2 * (absolute error bound of the fp32 ray origin, 0 for untransformed rays)
//   r2.w = 4 u + (error bound of the shear entries beyond their own 2 u, 0 for untransformed rays)
// Error model (DESIGN.md "triangle test"): with bm the largest |component| of the ray-relative vertices and m2 the
// largest |2D coordinate|, every 2D coordinate is within e = bm * r2.w + r1.w + u * m2 of its exact value
// (subtraction u, shear product u, shear entry 2 u, FFMA rounding), an edge function within 4 e m2 + 2 e^2 + 4 u m2^2.
// The bound E carries 25 % head room.  TRI_HIT: all three edge functions exceed E with the same sign; TRI_MISS: two
// of them provably differ in sign; everything else (including NaN) is TRI_UNSURE.  For TRI_HIT, t is the fp32
// distance and terr a bound of |t - t_exact|.
LCRT_HD int tri_test_f32(float ox, float oy, float oz, const f4 r1, const f4 r2, float sz, uint32_t kz, const f4 v0, const f4 v1,
                         const f4 v2, float &t, float &terr) {
    const float oerr2 = r1.w, cerr = r2.w;
    const float Ax = v0.x - ox, Ay = v0.y - oy, Az = v0.z - oz;
    const float Bx = v1.x - ox, By = v1.y - oy, Bz = v1.z - oz;
    const float Cx = v2.x - ox, Cy = v2.y - oy, Cz = v2.z - oz;
    const float ax = fmaf(Az, r1.z, fmaf(Ay, r1.y, mul_rn(Ax, r1.x))), ay = fmaf(Az, r2.z, fmaf(Ay, r2.y, mul_rn(Ax, r2.x)));
    const float bx = fmaf(Bz, r1.z, fmaf(By, r1.y, mul_rn(Bx, r1.x))), by = fmaf(Bz, r2.z, fmaf(By, r2.y, mul_rn(Bx, r2.x)));
    const float cx = fmaf(Cz, r1.z, fmaf(Cy, r1.y, mul_rn(Cx, r1.x))), cy = fmaf(Cz, r2.z, fmaf(Cy, r2.y, mul_rn(Cx, r2.x)));
    const float U = fmaf(cx, by, -mul_rn(cy, bx));
    const float V = fmaf(ax, cy, -mul_rn(ay, cx));
    const float W = fmaf(bx, ay, -mul_rn(by, ax));
    const float bm = fmaxf(fmaxf(fmaxf(fabsf(Ax), fabsf(Ay)), fmaxf(fabsf(Az), fabsf(Bx))),
                           fmaxf(fmaxf(fmaxf(fabsf(By), fabsf(Bz)), fmaxf(fabsf(Cx), fabsf(Cy))), fabsf(Cz)));
    const float m2 = fmaxf(fmaxf(fmaxf(fabsf(ax), fabsf(ay)), fmaxf(fabsf(bx), fabsf(by))), fmaxf(fabsf(cx), fabsf(cy)));
    const float e = fmaf(bm, cerr, fmaf(LCRT_U, m2, oerr2));
    const float E = fmaf(m2, fmaf(5.0f * LCRT_U, m2, 5.0f * e), 2.5f * e * e);
    const bool pu = U > E, pv = V > E, pw = W > E, nu = U < -E, nv = V < -E, nw = W < -E;
    if ((pu | pv | pw) & (nu | nv | nw)) return TRI_MISS;
    if (!((pu & pv & pw) | (nu & nv & nw))) return TRI_UNSURE;
    // ---- certainly inside: t = sum b_i z_i with z_i the vertex distances along the ray
    const float det = U + V + W;
    const float za = (kz == 0u ? Ax : kz == 1u ? Ay : Az) * sz, zb = (kz == 0u ? Bx : kz == 1u ? By : Bz) * sz,
                zc = (kz == 0u ? Cx : kz == 1u ? Cy : Cz) * sz;
    const float T = fmaf(U, za, fmaf(V, zb, W * zc));
    const float rdet = rcp_fast(det);// 2 u, budgeted below
    t = T * rdet;
    // |b_i - b_i_exact| <= 4 E / |det| and the errors sum to zero, so they move t by at most 6 (E / |det|) zspread;
    // the z_i are within (r2.w + 3 u) zmax + |sz| oerr of exact; the accumulation and the division add 8 u zmax.
    const float zmax = fmaxf(fabsf(za), fmaxf(fabsf(zb), fabsf(zc)));
    const float zspread = fmaxf(za, fmaxf(zb, zc)) - fminf(za, fminf(zb, zc));
    terr = 1.25f * fmaf(6.0f * E * fabsf(rdet), zspread, fmaf(cerr + 11.0f * LCRT_U, zmax, fabsf(sz) * oerr2));
    return terr == terr ? TRI_HIT : TRI_UNSURE;
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

// The fallback: object-space ray re-derived in double from the world ray and the instance's double inverse, then
// tri_exact.  Out of line on the device so that its registers do not weigh on the traversal loop.
LCRT_HD_COLD bool tri_exact_world(const Ray *world, const InstanceRecord *rec, const f4 *tri, double &t, float &u, float &v) {
    const f4 v0 = tri[0], v1 = tri[1], v2 = tri[2];
    const double wox = world->ox, woy = world->oy, woz = world->oz, wdx = world->dx, wdy = world->dy, wdz = world->dz;
    double eox = wox, eoy = woy, eoz = woz, edx = wdx, edy = wdy, edz = wdz;
    if (!(rec->flags & IREC_IDENTITY)) {
        const double *m = rec->inv_exact;
        eox = m[0] * wox + m[1] * woy + m[2] * woz + m[3], eoy = m[4] * wox + m[5] * woy + m[6] * woz + m[7],
        eoz = m[8] * wox + m[9] * woy + m[10] * woz + m[11];
        edx = m[0] * wdx + m[1] * wdy + m[2] * wdz, edy = m[4] * wdx + m[5] * wdy + m[6] * wdz, edz = m[8] * wdx + m[9] * wdy + m[10] * wdz;
    }
    double ud, vd;
    if (!tri_exact(eox, eoy, eoz, edx, edy, edz, v0, v1, v2, t, ud, vd)) return false;
    u = (float)ud, v = (float)vd;
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

// Exact (t, u, v) and primitive id of an accepted surface hit, given where it was found: TLAS record `rec`,
// primitive record `pi` of that instance's BLAS.  Returns false if the double evaluation does not confirm the hit
// (can not happen while the fp32 error bounds hold; callers then keep the fp32 distance).
LCRT_HD bool refine_surface_hit(const Ray *world, const InstanceRecord *tlas_recs, uint32_t rec, uint32_t pi, uint32_t &prim, float &t,
                                float &u, float &v) {
    const InstanceRecord *r = tlas_recs + rec;
    const f4 *tr = reinterpret_cast<const f4 *>(reinterpret_cast<const char *>(r->prims) + (uint64_t)pi * sizeof(TriRecord));
    prim = as_uint(tr[0].w);
    double td;
    if (!tri_exact_world(world, r, tr, td, u, v)) return false;
    t = (float)td;
    return true;
}

// --------------------------------------------------------------------------------------------------------------
// Traversal state machine.  `step()` performs one outer iteration (at most one node, then at most one primitive);
// the persistent batch kernel interleaves step() with warp-cooperative ray refetch, the inline device-library
// entry points simply loop until done.
// --------------------------------------------------------------------------------------------------------------
template<int MODE, bool COUNT, class Store, class Callbacks>
struct Traversal {
    // ---- hot state (registers)
    float ox, oy, oz, idx, idy, idz;// current-space ray (box tests)
    float tmin;
    float tmax;        // culling bound: upper end of the best hit's t interval (CW_HIT_T / CW_HIT_TERR)
    uint32_t octinv;   // 7 - octant of the current-space direction
    const u4 *nodes;
    const char *prims;
    u2 G;              // node group: x = child base, y = hit internal children (by priority) << 24 | imask
    u2 Gt;             // primitive group: x = node index, y = pending leaf slots | remaining prims of the current slot << 8
    uint32_t tp;       // next primitive record of the current slot
    int sp, blas_sp;   // blas_sp < 0: traversing the TLAS
    uint32_t cur_inst, cur_flags, ray_mask;
    bool done;
    // ---- level-independent
    const Ray *world;  // the ray as given (re-read on instance exit, by the fallback and by the RayQuery callbacks)
    const u4 *tlas_nodes;
    const InstanceRecord *tlas_recs;
    const Instance *instances;
    Store *store;
    Callbacks *cb;
    TraceCounters ctr;
    bool terminate_on_first = false;// query_any
    uint32_t magic = PLANT_MAGIC;   // see plant_byte(); overwritten with the same value from a kernel parameter / the TLAS header
    static constexpr uint32_t RAY_PENDING = 0xffffffffu;// octinv value: the current-space ray must be (re)derived
    LCRT_HD bool MODE_ANY_QUERY() const { return MODE == MODE_QUERY && terminate_on_first; }

    // the world ray lives in global memory for the batch kernels (read-only path) and in the caller's local memory for
    // the inline entry points
    LCRT_HD f4 load_world(int k) const {
        const f4 *p = reinterpret_cast<const f4 *>(world) + k;
        if (Store::world_in_global) return ld16_stream(p);
        return *p;
    }

    // oerr: bound of |o - o_exact| per component; derr: likewise for d (both 0 for the untransformed ray)
    LCRT_HD void set_ray(float ox_, float oy_, float oz_, float dx, float dy, float dz, float oerr, float derr) {
        ox = ox_, oy = oy_, oz = oz_;
        const float sdx = safe_dir(dx), sdy = safe_dir(dy), sdz = safe_dir(dz);
        // 1 ulp reciprocals: the box test budgets 10 u per coefficient and uses at most 7 with these (axis_coef)
        idx = rcp_fast(sdx), idy = rcp_fast(sdy), idz = rcp_fast(sdz);
        const uint32_t oct = (sdx < 0.f ? 1u : 0u) | (sdy < 0.f ? 2u : 0u) | (sdz < 0.f ? 4u : 0u);
        octinv = 7u - oct;
        // shear rows for the triangle test (only primitives are tested with them, and those hang below BLAS nodes: a ray
        // that returns to the TLAS skips this part): kz = dominant axis, (kx, ky, kz) cyclic.  Branch-free on purpose.
        if (blas_sp < 0) return;
        const float adx = fabsf(dx), ady = fabsf(dy), adz = fabsf(dz);
        const bool k0 = adx >= ady && adx >= adz, k1 = !k0 && ady >= adz, k2 = !k0 && !k1;
        const float dk = k0 ? dx : k1 ? dy : dz;
        const float dkx = k0 ? dy : k1 ? dz : dx;
        const float dky = k0 ? dz : k1 ? dx : dy;
        const float rk = 1.0f / dk;// correctly rounded; NaN / inf for a null direction: every test becomes TRI_UNSURE
        const float nsx = -(dkx * rk), nsy = -(dky * rk);// within 2 u of exact
        f4 r1, r2;
        r1.x = k2 ? 1.f : k0 ? nsx : 0.f, r1.y = k0 ? 1.f : k1 ? nsx : 0.f, r1.z = k1 ? 1.f : k2 ? nsx : 0.f;
        r2.x = k1 ? 1.f : k0 ? nsy : 0.f, r2.y = k2 ? 1.f : k1 ? nsy : 0.f, r2.z = k0 ? 1.f : k2 ? nsy : 0.f;
        r1.w = 2.0f * oerr;
        r2.w = 4.0f * LCRT_U + 2.5f * derr * fabsf(rk);
        store->st4(0, r1), store->st4(1, r2);
        store->stw(CW_SZ, as_uint(rk)), store->stw(CW_KZ, k0 ? 0u : k1 ? 1u : 2u);
    }

    // `single`: the TLAS holds exactly one instance (record 0): its root node is skipped
    LCRT_HD void init(const u4 *tlas_nodes_, const InstanceRecord *tlas_recs_, bool single, const Instance *instances_, const Ray *ray,
                      uint32_t mask, Store *store_, Callbacks *cb_) {
        world = ray;
        instances = instances_;
        store = store_, cb = cb_;
        tlas_nodes = tlas_nodes_, tlas_recs = tlas_recs_;
        nodes = tlas_nodes, prims = reinterpret_cast<const char *>(tlas_recs);
        const float tmin_ = load_world(0).w, tmax_ = load_world(1).w;
        tmin = tmin_, tmax = tmax_;
        // the current-space ray is set up by the first step() that needs it (a single-instance TLAS goes straight to
        // its instance and never needs the world-space setup)
        ox = oy = oz = idx = idy = idz = 0.f;
        octinv = single ? 0u : RAY_PENDING;
        ray_mask = mask & 0xffu;
        G = single ? make_u2(0u, 0u) : make_u2(0u, 0x80000000u);
        Gt = single ? make_u2(0u, 1u << 8) : make_u2(0u, 0u);
        tp = 0u;
        sp = 0, blas_sp = -1;
        cur_inst = ~0u, cur_flags = 0u;
        store->stw(CW_HIT_T, as_uint(tmax_)), store->stw(CW_HIT_TERR, as_uint(-1.0f));
        store->stw(CW_HIT_INST, ~0u), store->stw(CW_HIT_PRIM, ~0u), store->stw(CW_HIT_TYPE, HIT_MISS);
        store->stw(CW_HIT_EXACT, 0u), store->stw(CW_REC, 0u);
        ctr.nodes = ctr.prims = ctr.insts = ctr.exact = 0u;
        // NaN / empty-interval rays never hit anything (Embree: tnear <= t <= tfar)
        done = !(tmin_ <= tmax_);
    }
    LCRT_HD void init(const BlobHeader *tlas, const Instance *instances_, const Ray *ray, uint32_t mask, Store *store_, Callbacks *cb_) {
        init(reinterpret_cast<const u4 *>(reinterpret_cast<const char *>(tlas) + tlas->nodes_offset),
             reinterpret_cast<const InstanceRecord *>(reinterpret_cast<const char *>(tlas) + tlas->prims_offset), tlas->prim_count == 1u, instances_,
             ray, mask, store_, cb_);
        // PRMT has one immediate operand.  With a compile-time magic the compiler spends it on the magic and keeps the four
        // byte selectors in registers, which it re-materialises about 40 times per node test (301 instructions); a magic
        // it can not see through goes into a register once and the selectors become immediates (as in the batch kernel,
        // where it is a kernel parameter).
        magic = tlas->plant_magic;
        check_stack_bound(tlas->max_stack);
    }

    // ---- result (valid once step() has returned false).  hit_rec / hit_pi locate a surface hit for refine_surface_hit.
    LCRT_HD uint32_t hit_inst() const { return store->ldw(CW_HIT_INST); }
    LCRT_HD uint32_t hit_prim() const { return store->ldw(CW_HIT_PRIM); }
    LCRT_HD uint32_t hit_type() const { return store->ldw(CW_HIT_TYPE); }
    LCRT_HD uint32_t hit_rec() const { return store->ldw(CW_HIT_REC); }
    LCRT_HD uint32_t hit_pi() const { return store->ldw(CW_HIT_PI); }
    LCRT_HD float hit_t() const { return as_float(store->ldw(CW_HIT_T)); }
    LCRT_HD bool has_hit() const { return store->ldw(CW_HIT_INST) != ~0u; }

    // final (inst, prim, t, u, v) with the double-precision refinement done in place (inline entry points, hostsim)
    LCRT_HD void result(uint32_t &inst, uint32_t &prim, uint32_t &type, float &t, float &u, float &v) const {
        inst = hit_inst(), prim = hit_prim(), type = MODE == MODE_QUERY ? hit_type() : (inst != ~0u ? (uint32_t)HIT_SURFACE : (uint32_t)HIT_MISS);
        t = hit_t(), u = 0.f, v = 0.f;
        if (type == HIT_SURFACE) refine_surface_hit(world, tlas_recs, hit_rec(), hit_pi(), prim, t, u, v);
    }

    LCRT_HD void accept(float t, float terr, uint32_t prim, uint32_t pi, uint32_t type, bool exact, double td) {
        store->stw(CW_HIT_T, as_uint(t)), store->stw(CW_HIT_TERR, as_uint(terr));
        store->stw(CW_HIT_INST, cur_inst), store->stw(CW_HIT_PRIM, prim);
        store->stw(CW_HIT_REC, store->ldw(CW_REC)), store->stw(CW_HIT_PI, pi);
        if (MODE == MODE_QUERY) store->stw(CW_HIT_TYPE, type);
        store->stw(CW_HIT_EXACT, exact ? 1u : 0u);
        if (exact) {
            uint64_t bits;
            memcpy(&bits, &td, 8);
            store->stw(CW_HIT_TD_LO, (uint32_t)bits), store->stw(CW_HIT_TD_HI, (uint32_t)(bits >> 32));
        }
        // cull beyond the upper end of the accepted interval (a hair more: coincident primitives must still be
        // visited so that the lowest (inst, prim) wins an exact tie)
        tmax = fmaf(fabsf(t), 2.4e-7f, t + terr);
    }

    // (inst, prim) order for exact ties: the result must not depend on traversal order
    LCRT_HD bool loses_tie(uint32_t prim) const {
        const uint32_t hi = store->ldw(CW_HIT_INST);
        return hi != ~0u && (cur_inst > hi || (cur_inst == hi && prim >= store->ldw(CW_HIT_PRIM)));
    }

    // exact t of the current best hit (evaluated on demand, then cached)
    LCRT_HD double best_exact_t() {
        if (store->ldw(CW_HIT_EXACT)) {
            const uint64_t bits = ((uint64_t)store->ldw(CW_HIT_TD_HI) << 32) | store->ldw(CW_HIT_TD_LO);
            double td;
            memcpy(&td, &bits, 8);
            return td;
        }
        double td = (double)hit_t();
        if (MODE != MODE_QUERY || hit_type() == HIT_SURFACE) {
            const InstanceRecord *r = tlas_recs + store->ldw(CW_HIT_REC);
            const f4 *tr = reinterpret_cast<const f4 *>(reinterpret_cast<const char *>(r->prims) + (uint64_t)store->ldw(CW_HIT_PI) * sizeof(TriRecord));
            float u, v;
            double te;
            if (COUNT) ctr.exact++;
            if (tri_exact_world(world, r, tr, te, u, v)) td = te;
        }
        uint64_t bits;
        memcpy(&bits, &td, 8);
        store->stw(CW_HIT_TD_LO, (uint32_t)bits), store->stw(CW_HIT_TD_HI, (uint32_t)(bits >> 32)), store->stw(CW_HIT_EXACT, 1u);
        return td;
    }

    LCRT_HD void leave_instance() {
        blas_sp = -1;
        nodes = tlas_nodes, prims = reinterpret_cast<const char *>(tlas_recs);
        cur_inst = ~0u, cur_flags = 0u;
        octinv = RAY_PENDING;
    }

    // The stack can not overflow on a structure this library built: the builder records the depth of every hierarchy
    // (BlobHeader::depth), a TLAS build adds up the most entries a ray can need (max_stack = TLAS depth + 2 + deepest
    // instanced BLAS) and refuses to mark itself built beyond LCRT_STACK_SIZE (k_build_finalize, error 5).  The traversal
    // checks that bound ONCE per ray, in init(), instead of guarding every push (a compare and a branch per node cost the
    // batch kernel 5 % — profiles/r2_trace_guard.md): a blob that claims more ends the ray at once and raises
    // STACK_OVERFLOW, which the callers report through BlobHeader::trace_error.
    static constexpr uint32_t STACK_OVERFLOW = 0x80000000u;// flag bit in ray_mask (the mask itself is 8 bits)
    LCRT_HD bool overflowed() const { return (ray_mask & STACK_OVERFLOW) != 0u; }
    LCRT_HD void push(u2 v) { store->store(sp++, v); }
    LCRT_HD void check_stack_bound(uint32_t max_stack) {
        if (max_stack > (uint32_t)LCRT_STACK_SIZE) ray_mask |= STACK_OVERFLOW, done = true;
    }

    LCRT_HD void enter_instance(uint32_t rec_index, const u4 r3, const u4 r4) {
        // keep what is left of the TLAS groups below the BLAS frames
        if (G.y & 0xff000000u) push(G);
        if (Gt.y & 0xffu) push(make_u2(Gt.x, Gt.y & 0xffu));
        Gt = make_u2(0u, 0u);
        blas_sp = sp;
        nodes = reinterpret_cast<const u4 *>(((uint64_t)r3.y << 32) | r3.x);
        prims = reinterpret_cast<const char *>(((uint64_t)r3.w << 32) | r3.z);
        cur_inst = r4.x, cur_flags = r4.z;
        store->stw(CW_REC, rec_index);
        G = make_u2(0u, 0x80000000u);
        octinv = RAY_PENDING;
        if (COUNT) ctr.insts++;
    }

    // (re)derive the current-space ray after a level change: the world ray, moved to object space by the fp32
    // inverse when inside a non-identity instance.  One call site, so that lanes which start a ray, enter an
    // instance or return from one share the same instructions.
    LCRT_HD void setup_ray() {
        const f4 w0 = load_world(0), w1 = load_world(1);
        float sox = w0.x, soy = w0.y, soz = w0.z, sdx = w1.x, sdy = w1.y, sdz = w1.z, oerr = 0.f, derr = 0.f;
        if (blas_sp >= 0 && !(cur_flags & IREC_IDENTITY)) {
            const f4 *rec = reinterpret_cast<const f4 *>(tlas_recs + store->ldw(CW_REC));
            const f4 m0 = ldg16(rec), m1 = ldg16(rec + 1), m2 = ldg16(rec + 2);
#if LCRT_RECENTER
            // o' = inv_linear * (o - T): the subtraction is exact for an origin near the instance, the rest works on small numbers
            const f4 tw = ldg16(rec + 5);// { T.x, T.y, T.z, - } as fp32 bits (InstanceRecord::spare)
            const float rx = w0.x - tw.x, ry = w0.y - tw.y, rz = w0.z - tw.z;
            sox = fmaf(m0.x, rx, fmaf(m0.y, ry, m0.z * rz));
            soy = fmaf(m1.x, rx, fmaf(m1.y, ry, m1.z * rz));
            soz = fmaf(m2.x, rx, fmaf(m2.y, ry, m2.z * rz));
#else
            sox = fmaf(m0.x, w0.x, fmaf(m0.y, w0.y, fmaf(m0.z, w0.z, m0.w)));
            soy = fmaf(m1.x, w0.x, fmaf(m1.y, w0.y, fmaf(m1.z, w0.z, m1.w)));
            soz = fmaf(m2.x, w0.x, fmaf(m2.y, w0.y, fmaf(m2.z, w0.z, m2.w)));
#endif
            sdx = fmaf(m0.x, w1.x, fmaf(m0.y, w1.y, m0.z * w1.z));
            sdy = fmaf(m1.x, w1.x, fmaf(m1.y, w1.y, m1.z * w1.z));
            sdz = fmaf(m2.x, w1.x, fmaf(m2.y, w1.y, m2.z * w1.z));
            // |fl(o') - o'_exact| <= 5 u * max_i (sum_j |inv_ij o_j| + |inv_i3|): entries rounded from double (1 u),
            // three FFMA roundings; likewise for the direction
#if LCRT_RECENTER
            // |o - T| is within u of exact (exact when the two are within a factor of two); the entries and three roundings as below
            const float ax = fabsf(rx), ay = fabsf(ry), az = fabsf(rz);
            const float mx = 1.25f * fmaf(fabsf(m0.x), ax, fmaf(fabsf(m0.y), ay, fabsf(m0.z) * az));
            const float my = 1.25f * fmaf(fabsf(m1.x), ax, fmaf(fabsf(m1.y), ay, fabsf(m1.z) * az));
            const float mz = 1.25f * fmaf(fabsf(m2.x), ax, fmaf(fabsf(m2.y), ay, fabsf(m2.z) * az));
#else
            const float ax = fabsf(w0.x), ay = fabsf(w0.y), az = fabsf(w0.z);
            const float mx = fmaf(fabsf(m0.x), ax, fmaf(fabsf(m0.y), ay, fmaf(fabsf(m0.z), az, fabsf(m0.w))));
            const float my = fmaf(fabsf(m1.x), ax, fmaf(fabsf(m1.y), ay, fmaf(fabsf(m1.z), az, fabsf(m1.w))));
            const float mz = fmaf(fabsf(m2.x), ax, fmaf(fabsf(m2.y), ay, fmaf(fabsf(m2.z), az, fabsf(m2.w))));
#endif
            const float bx = fabsf(w1.x), by = fabsf(w1.y), bz = fabsf(w1.z);
            const float nx = fmaf(fabsf(m0.x), bx, fmaf(fabsf(m0.y), by, fabsf(m0.z) * bz));
            const float ny = fmaf(fabsf(m1.x), bx, fmaf(fabsf(m1.y), by, fabsf(m1.z) * bz));
            const float nz = fmaf(fabsf(m2.x), bx, fmaf(fabsf(m2.y), by, fabsf(m2.z) * bz));
            oerr = 5.0f * LCRT_U * fmaxf(mx, fmaxf(my, mz)), derr = 5.0f * LCRT_U * fmaxf(nx, fmaxf(ny, nz));
        }
        set_ray(sox, soy, soz, sdx, sdy, sdz, oerr, derr);
    }

    // one candidate triangle; returns true if the traversal terminates
    LCRT_HD bool test_triangle(uint32_t pi) {
        const f4 *tr = reinterpret_cast<const f4 *>(prims + (uint64_t)pi * sizeof(TriRecord));
        const f4 v0 = ldg16(tr), v1 = ldg16(tr + 1), v2 = ldg16(tr + 2);
        float t, terr;
        const int verdict = tri_test_f32(ox, oy, oz, store->ld4(0), store->ld4(1), as_float(store->ldw(CW_SZ)), store->ldw(CW_KZ), v0, v1, v2, t, terr);
        if (verdict == TRI_MISS) return false;
        const uint32_t prim = as_uint(v0.w);
        const bool candidate_program = MODE == MODE_QUERY && !(cur_flags & IREC_OPAQUE);
        bool take = false, exact = false, need_exact = verdict == TRI_UNSURE || candidate_program;
        double td = 0.0;
        float u = 0.f, v = 0.f;
        const float ht = hit_t(), he = as_float(store->ldw(CW_HIT_TERR));// he < 0: no hit yet, ht is the ray's tmax
        if (!need_exact) {
            // interval comparison: take / drop only what is certain (Embree: tnear <= t <= tfar, nearest wins)
            const float hw = fmaxf(he, 0.f);
            if (t - terr > ht + hw || t + terr < tmin) return false;
            if (t - terr >= tmin && t + terr < ht - hw) take = true;
            else need_exact = true;
        }
        if (need_exact) {
            if (COUNT) ctr.exact++;
            if (!tri_exact_world(world, tlas_recs + store->ldw(CW_REC), tr, td, u, v) || !(td >= (double)tmin)) return false;
            if (he < 0.f) {
                take = td <= (double)ht;
            } else {
                const double bd = best_exact_t();
                take = td < bd || (td == bd && !loses_tie(prim));// exact ties go to the lower (inst, prim)
            }
            if (!take) return false;
            t = (float)td, terr = 1.2e-7f * fabsf(t), exact = true;
        }
        if constexpr (callbacks_suspend<Callbacks>::value) {
            if (candidate_program) {// the caller runs the handler (resolve())
                PendingCandidate &pc = cb->pend;
                pc.kind = 1, pc.inst = cur_inst, pc.prim = prim, pc.pi = pi, pc.t = t, pc.terr = terr, pc.u = u, pc.v = v, pc.ht = t;
                pc.exact = exact, pc.td = td;
                return false;
            }
        }
        if (candidate_program) {
            Ray wr = *world;
            wr.tmax = t;// accel.rs:663-672: tmax shrinks to the candidate while the callback runs
            const CandidateResult r = cb->on_triangle(cur_inst, prim, u, v, t, wr);
            if (r.committed) accept(t, terr, prim, pi, HIT_SURFACE, exact, td);
            return r.terminated || (r.committed && MODE_ANY_QUERY());
        }
        accept(t, terr, prim, pi, HIT_SURFACE, exact, td);
        return MODE == MODE_ANY || MODE_ANY_QUERY();
    }

    // ---- suspending callbacks: the candidate the last step left for the caller, and its verdict
    LCRT_HD bool candidate_pending() const {
        if constexpr (callbacks_suspend<Callbacks>::value) return cb->pend.kind != 0;
        else return false;
    }
    LCRT_HD Ray candidate_ray() const {// the world ray as the handler sees it (accel.rs:663-672, 705-716)
        Ray wr = *world;
        if constexpr (callbacks_suspend<Callbacks>::value) wr.tmax = cb->pend.ht;
        return wr;
    }
    LCRT_HD void resolve(bool committed, bool terminated, float t_hit) {
        if constexpr (callbacks_suspend<Callbacks>::value) {
            const PendingCandidate pc = cb->pend;
            cb->pend.kind = 0;
            if (pc.kind == 1) {
                if (committed) accept(pc.t, pc.terr, pc.prim, pc.pi, HIT_SURFACE, pc.exact, pc.td);
                if (terminated || (committed && MODE_ANY_QUERY())) done = true;
            } else if (pc.kind == 2) {
                if (committed && t_hit >= tmin && t_hit < pc.ht) {// accel.rs:711-713
                    accept(t_hit, 0.f, pc.prim, pc.pi, HIT_PROCEDURAL, true, (double)t_hit);
                    if (MODE_ANY_QUERY()) done = true;
                }
                if (terminated) done = true;
            }
        }
    }

    // step() = step_node() followed by step_prim(); the batch kernel calls the halves separately so that the warp can
    // vote on when the (poorly converged) primitive phase is worth running.  Both return true while there is more work.
    LCRT_HD bool prim_pending() const { return Gt.y != 0u; }
    LCRT_HD bool step() { return step_node() && step_prim(); }
    // whole ray, for the inline entry points (one thread, one ray, no refill).  One node, then one primitive per
    // iteration; the "while-while" alternative (nodes until a primitive group turns up, then the whole group) measured
    // 27 % slower on coherent and 18 % slower on incoherent rays through lc-backend-b200 (C2, B200).
    LCRT_HD void run() {
        while (step()) {}
    }
    // run() for traversals with expensive primitive phases (RayQuery candidates, many instance entries per ray): on the
    // device the lanes that entered together vote on the two phases of an iteration.  The primitive phase runs when
    // LCRT_RUN_PRIM_VOTE lanes wait for it (default 1: always); the node phase runs when LCRT_RUN_NODE_VOTE lanes, or
    // at least as many lanes as wait for primitives, want it.  A path tracer over a small scene (36 triangles behind 9
    // instances, BASELINE config C5) ran the 289-instruction node test with 7 of 32 lanes under run(); through the drop-in
    // backend the vote takes it from 287 to 410 spp/s.  A lane that loses a vote waits, so a latency-bound traversal of
    // one large mesh without refill loses (C2 through the drop-in backend: -16 %); closest / any-hit use run().
    // Lanes that are done keep voting until the group is, so the loop exit is uniform.
    // Measured and rejected on top of this: entering a BLAS of <= 4 primitives without its root node test (record count in
    // the instance record): C5 went from 411 to 364-372 spp/s, the extra set-up-only iterations and the lost box culling
    // cost more than the node tests saved.
    LCRT_HD void run_voting() {
#if defined(__CUDA_ARCH__) && LCRT_RUN_VOTE
        const uint32_t group = __activemask();
        bool alive = !done;
        for (;;) {
            const bool wants_prim = alive && prim_pending(), wants_node = alive && !wants_prim;
            const uint32_t pm = __ballot_sync(group, wants_prim), nm = __ballot_sync(group, wants_node);
            if ((pm | nm) == 0u) break;
            const int np = popc(pm), nn = popc(nm);
            const bool run_node = nn != 0 && (nn >= LCRT_RUN_NODE_VOTE || nn >= np);
            if (run_node && wants_node) alive = step_node();
            if ((np >= LCRT_RUN_PRIM_VOTE || !run_node) && alive && prim_pending()) alive = step_prim();
        }
#else
        run();
#endif
    }

    LCRT_HD bool step_node() {
        if (done) return false;
        // ---- pop when both groups are exhausted
        if ((G.y & 0xff000000u) == 0u && Gt.y == 0u) {
            if (blas_sp >= 0 && sp == blas_sp) leave_instance();
            if (sp == 0) {
                done = true;
                return false;
            }
            const u2 e = store->load(--sp);
            if (e.y & 0xff000000u) G = e;
            else Gt = e;
        }
        if (octinv == RAY_PENDING && (blas_sp >= 0 || (G.y & 0xff000000u) != 0u)) setup_ray();
        // ---- one internal node (not while primitives of the previous node are pending: they may shrink tmax)
        if ((G.y & 0xff000000u) != 0u && Gt.y == 0u) {
            const int child_bit = bfind(G.y);
            const uint32_t base = G.x;
            G.y &= ~(1u << child_bit);
            const uint32_t imask_parent = G.y & 0xffu;
            if (G.y & 0xff000000u) push(G);
            const uint32_t slot = (uint32_t)(child_bit - 24) ^ octinv;
            const uint32_t rel = (uint32_t)popc(imask_parent & ~(0xffffffffu << slot));
            const uint32_t ni = base + rel;
            const u4 *n = nodes + (uint64_t)ni * 5u;
            const u4 n0 = ldg16(n), n2 = ldg16(n + 2), n3 = ldg16(n + 3), n4 = ldg16(n + 4);
            const uint32_t child_base = ldg16(n + 1).x;
            if (COUNT) ctr.nodes++;
            const uint32_t hm = intersect_node8(n0, n2, n3, n4, ox, oy, oz, idx, idy, idz, tmin, tmax, magic);
            const uint32_t imask = n0.w >> 24;
            G = make_u2(child_base, (permute_slots(hm & imask, octinv) << 24) | imask);
            Gt = make_u2(ni, hm & ~imask);
        }
        return true;
    }

    LCRT_HD bool step_prim() {
        // ---- one primitive
        if (Gt.y != 0u) {
            if ((Gt.y >> 8) == 0u) {
                // open the next hit leaf slot: its records are prim_base + offset .. + count
                const uint32_t s = (uint32_t)bfind(Gt.y);
                Gt.y &= ~(1u << s);
                const u4 n1 = ldg16(nodes + (uint64_t)Gt.x * 5u + 1);
                const uint32_t meta = (((s & 4u) ? n1.w : n1.z) >> (8u * (s & 3u))) & 0xffu;
                const uint32_t unary = meta >> 5;// 0 (empty slot), 1, 3 or 7
                tp = n1.y + (meta & 31u);
                Gt.y |= ((unary & 1u) + ((unary >> 1) & 1u) + (unary >> 2)) << 8;
            }
            if ((Gt.y >> 8) != 0u) {
                Gt.y -= 1u << 8;
                const uint32_t pi = tp++;
                if (blas_sp < 0) {
                    // TLAS leaf: an instance
                    const u4 *rec = reinterpret_cast<const u4 *>(tlas_recs + pi);
                    const u4 r4 = ldg16(rec + 4);// { inst, mask, flags, pad }
                    if ((r4.y & ray_mask) != 0u && (MODE == MODE_QUERY || (r4.z & IREC_KIND_MASK) == KIND_TRIANGLES))
                        enter_instance(pi, ldg16(rec + 3), r4);// intersect() does not see procedural prims
                } else if ((cur_flags & IREC_KIND_MASK) == KIND_TRIANGLES) {
                    if (COUNT) ctr.prims++;
                    if (test_triangle(pi)) {
                        done = true;
                        return false;
                    }
                } else if (MODE == MODE_QUERY) {
                    if (COUNT) ctr.prims++;
                    // procedural primitive: conservative box test, then the candidate program decides
                    const f4 *br = reinterpret_cast<const f4 *>(prims + (uint64_t)pi * sizeof(BoxRecord));
                    const f4 lo = ldg16(br), hi = ldg16(br + 1);
                    if (slab_conservative(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, ox, oy, oz, idx, idy, idz, tmin, tmax)) {
                        Ray wr = *world;
                        // the program sees the committed distance so far; a surface hit reports its exact value
                        const float ht = store->ldw(CW_HIT_INST) != ~0u ? (float)best_exact_t() : hit_t();
                        wr.tmax = ht;
                        const uint32_t prim = as_uint(lo.w);
                        if constexpr (callbacks_suspend<Callbacks>::value) {
                            PendingCandidate &pc = cb->pend;
                            pc.kind = 2, pc.inst = cur_inst, pc.prim = prim, pc.pi = pi, pc.ht = ht;
                            return true;// the caller runs the handler (resolve())
                        }
                        const CandidateResult r = cb->on_procedural(cur_inst, prim, wr, instances, lo, hi);
                        if (r.committed && r.t_hit >= tmin && r.t_hit < ht) {// accel.rs:711-713
                            accept(r.t_hit, 0.f, prim, pi, HIT_PROCEDURAL, true, (double)r.t_hit);
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
            }
        }
        return true;
    }
};

}// namespace lcrt
