// lcrt_query_programs.h — built-in RayQuery candidate programs for the batch C-ABI (lcrt_query_*).
//
// In LuisaCompute the candidate handlers are user DSL code that the codegen outlines
// (src/backends/cuda/cuda_codegen_ast.cpp:255-543) and OptiX calls from __intersection__ray_query /
// __anyhit__ray_query (cuda_device_resource.h:2146-2385).  Arbitrary handlers are supported through the device
// library (lcrt_device_library.cuh: template callbacks inlined into the traversal loop); the fixed programs below
// exist so that RayQuery can be exercised and parity-checked through the plain C-ABI.  The oracle implements the
// same rules (oracle/oracle.cpp TriProgram / ProcProgram).
#pragma once
#include "lcrt_build.h"
#include "lcrt_traverse.h"

namespace lcrt {

enum TriProgram : uint32_t { TRI_COMMIT_ALL = 0, TRI_CUTOUT_MOD3 = 1, TRI_IGNORE_ALL = 2, TRI_TERMINATE_FIRST = 3 };
enum ProcProgram : uint32_t { PROC_SPHERE = 0, PROC_BOX = 1, PROC_IGNORE_ALL = 2 };

struct BuiltinPrograms {
    uint32_t tri_prog, proc_prog;
    LCRT_HD CandidateResult on_triangle(uint32_t, uint32_t prim, float, float, float t, const Ray &) const {
        CandidateResult r{t, true, false};
        switch (tri_prog) {
            case TRI_CUTOUT_MOD3: r.committed = (prim % 3u) != 0u; break;
            case TRI_IGNORE_ALL: r.committed = false; break;
            case TRI_TERMINATE_FIRST: r.terminated = true; break;
            default: break;
        }
        return r;
    }
    // `lo`/`hi`: the candidate's object-space AABB; the world ray carries the current tmax
    LCRT_HD CandidateResult on_procedural(uint32_t inst, uint32_t, const Ray &wr, const Instance *instances, const f4 lo, const f4 hi) const {
        CandidateResult r{0.f, false, false};
        if (proc_prog == PROC_SPHERE) {
            // sphere inscribed in the AABB, moved to world space by the instance transform (uniform scale), as the
            // handler of src/tests/test_procedural.cpp:101-125 does; evaluated in double so that it agrees with the oracle
            const float *m = instances[inst].affine;
            const double cxo = 0.5 * ((double)lo.x + (double)hi.x), cyo = 0.5 * ((double)lo.y + (double)hi.y), czo = 0.5 * ((double)lo.z + (double)hi.z);
            const double ro = 0.5 * ((double)hi.x - (double)lo.x);
            const double cx = m[0] * cxo + m[1] * cyo + m[2] * czo + m[3], cy = m[4] * cxo + m[5] * cyo + m[6] * czo + m[7],
                         cz = m[8] * cxo + m[9] * cyo + m[10] * czo + m[11];
            const double rad = ro * sqrt((double)m[0] * m[0] + (double)m[4] * m[4] + (double)m[8] * m[8]);
            const double Lx = wr.ox - cx, Ly = wr.oy - cy, Lz = wr.oz - cz;
            const double dx = wr.dx, dy = wr.dy, dz = wr.dz;
            const double a = dx * dx + dy * dy + dz * dz, hb = Lx * dx + Ly * dy + Lz * dz, cc = Lx * Lx + Ly * Ly + Lz * Lz - rad * rad;
            const double disc = hb * hb - a * cc;
            if (disc >= 0.0) {
                const double sq = sqrt(disc);
                const double t0 = (-hb - sq) / a, t1 = (-hb + sq) / a;
                if (t0 >= (double)wr.tmin && t0 < (double)wr.tmax) r.committed = true, r.t_hit = (float)t0;
                else if (t1 >= (double)wr.tmin && t1 < (double)wr.tmax) r.committed = true, r.t_hit = (float)t1;
            }
        } else if (proc_prog == PROC_BOX) {
            // the AABB itself: entry distance of the object-space box, the world ray moved to object space in double
            // (oracle/oracle.cpp slab() + PROC_BOX: t = max(tmin, entry), committed iff the slabs overlap and tmin <= t < tmax)
            double inv[12];
            if (!invert_affine(instances[inst].affine, inv)) return r;
            const double wo[3] = {wr.ox, wr.oy, wr.oz}, wd[3] = {wr.dx, wr.dy, wr.dz};
            const double blo[3] = {lo.x, lo.y, lo.z}, bhi[3] = {hi.x, hi.y, hi.z};
            double t0 = wr.tmin, t1 = wr.tmax;
            bool ok = true;
            for (int a = 0; a < 3; a++) {
                const double o = inv[4 * a] * wo[0] + inv[4 * a + 1] * wo[1] + inv[4 * a + 2] * wo[2] + inv[4 * a + 3];
                const double d = inv[4 * a] * wd[0] + inv[4 * a + 1] * wd[1] + inv[4 * a + 2] * wd[2];
                const double id = 1.0 / d;
                double ta = (blo[a] - o) * id, tb = (bhi[a] - o) * id;
                if (ta != ta || tb != tb) {// 0 * inf: origin on the slab plane, direction parallel to it
                    if (o < blo[a] || o > bhi[a]) ok = false;
                    continue;
                }
                if (ta > tb) {
                    const double s = ta;
                    ta = tb, tb = s;
                }
                ta -= fabs(ta) * 1e-12, tb += fabs(tb) * 1e-12;
                t0 = ta > t0 ? ta : t0, t1 = tb < t1 ? tb : t1;
            }
            if (ok && t0 <= t1 && t0 >= (double)wr.tmin && t0 < (double)wr.tmax) r.committed = true, r.t_hit = (float)t0;
        }
        return r;
    }
};

}// namespace lcrt
