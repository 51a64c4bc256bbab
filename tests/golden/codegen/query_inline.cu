// Output of THIS repository's code generator (lc-backend-b200: the reference's CUDA generator with plugin/hybrid/b200_ray_query_lowering.inl
// spliced in, RayQuery handlers lowered inline; LUISA_DUMP_SOURCE=1 LUISA_B200_RAY_QUERY=inline) for the `query` kernel of baseline/ref_optix_bench.cpp
// (scene mode), with the built-in device library text cut out at the marker below.  Fixture of tests/test_codegen_compat.py.
#define LUISA_ENABLE_OPTIX
#define LUISA_ENABLE_OPTIX_RAY_QUERY
#define LUISA_RAY_QUERY_IMPL_COUNT 0
#define LCRT_RAY_QUERY_INLINE_LOWERING 1
#define LCRT_RAY_QUERY_SITE_COUNT 1
#define LC_BLOCK_SIZE lc_make_uint3(256, 1, 1)

/* built-in device library */

template<> __device__ inline auto lc_zero<LCRay>() noexcept {
  return LCRay{
    lc_zero<lc_array<lc_float, 3>>(),
    lc_zero<lc_float>(),
    lc_zero<lc_array<lc_float, 3>>(),
    lc_zero<lc_float>(),
  };
}

template<> __device__ inline auto lc_one<LCRay>() noexcept {
  return LCRay{
    lc_one<lc_array<lc_float, 3>>(),
    lc_one<lc_float>(),
    lc_one<lc_array<lc_float, 3>>(),
    lc_one<lc_float>(),
  };
}

__device__ inline void lc_accumulate_grad(LCRay *dst, LCRay grad) noexcept {
  lc_accumulate_grad(&dst->m0, grad.m0);
  lc_accumulate_grad(&dst->m1, grad.m1);
  lc_accumulate_grad(&dst->m2, grad.m2);
  lc_accumulate_grad(&dst->m3, grad.m3);
}

struct alignas(4) SE5A58585AF30682A {
  lc_array<lc_float, 3> m0{};
  lc_array<lc_float, 3> m1{};
};

struct alignas(4) SE5A58585AF30682A_int {
  lc_array<lc_uint, 3> m0{};
  lc_array<lc_uint, 3> m1{};
};

template<> __device__ inline auto lc_zero<SE5A58585AF30682A>() noexcept {
  return SE5A58585AF30682A{
    lc_zero<lc_array<lc_float, 3>>(),
    lc_zero<lc_array<lc_float, 3>>(),
  };
}

template<> __device__ inline auto lc_one<SE5A58585AF30682A>() noexcept {
  return SE5A58585AF30682A{
    lc_one<lc_array<lc_float, 3>>(),
    lc_one<lc_array<lc_float, 3>>(),
  };
}

__device__ inline void lc_accumulate_grad(SE5A58585AF30682A *dst, SE5A58585AF30682A grad) noexcept {
  lc_accumulate_grad(&dst->m0, grad.m0);
  lc_accumulate_grad(&dst->m1, grad.m1);
}

template<> __device__ inline auto lc_zero<LCCommittedHit>() noexcept {
  return LCCommittedHit{
    lc_zero<lc_uint>(),
    lc_zero<lc_uint>(),
    lc_zero<lc_float2>(),
    lc_zero<lc_uint>(),
    lc_zero<lc_float>(),
  };
}

template<> __device__ inline auto lc_one<LCCommittedHit>() noexcept {
  return LCCommittedHit{
    lc_one<lc_uint>(),
    lc_one<lc_uint>(),
    lc_one<lc_float2>(),
    lc_one<lc_uint>(),
    lc_one<lc_float>(),
  };
}

__device__ inline void lc_accumulate_grad(LCCommittedHit *dst, LCCommittedHit grad) noexcept {
  lc_accumulate_grad(&dst->m0, grad.m0);
  lc_accumulate_grad(&dst->m1, grad.m1);
  lc_accumulate_grad(&dst->m2, grad.m2);
  lc_accumulate_grad(&dst->m3, grad.m3);
  lc_accumulate_grad(&dst->m4, grad.m4);
}

template<> __device__ inline auto lc_zero<LCTriangleHit>() noexcept {
  return LCTriangleHit{
    lc_zero<lc_uint>(),
    lc_zero<lc_uint>(),
    lc_zero<lc_float2>(),
    lc_zero<lc_float>(),
  };
}

template<> __device__ inline auto lc_one<LCTriangleHit>() noexcept {
  return LCTriangleHit{
    lc_one<lc_uint>(),
    lc_one<lc_uint>(),
    lc_one<lc_float2>(),
    lc_one<lc_float>(),
  };
}

__device__ inline void lc_accumulate_grad(LCTriangleHit *dst, LCTriangleHit grad) noexcept {
  lc_accumulate_grad(&dst->m0, grad.m0);
  lc_accumulate_grad(&dst->m1, grad.m1);
  lc_accumulate_grad(&dst->m2, grad.m2);
  lc_accumulate_grad(&dst->m3, grad.m3);
}

template<> __device__ inline auto lc_zero<LCProceduralHit>() noexcept {
  return LCProceduralHit{
    lc_zero<lc_uint>(),
    lc_zero<lc_uint>(),
  };
}

template<> __device__ inline auto lc_one<LCProceduralHit>() noexcept {
  return LCProceduralHit{
    lc_one<lc_uint>(),
    lc_one<lc_uint>(),
  };
}

__device__ inline void lc_accumulate_grad(LCProceduralHit *dst, LCProceduralHit grad) noexcept {
  lc_accumulate_grad(&dst->m0, grad.m0);
  lc_accumulate_grad(&dst->m1, grad.m1);
}

struct alignas(16) Params {
  alignas(16) const LCBuffer<const LCRay> b0{};
  alignas(16) const LCBuffer<LCCommittedHit> b1{};
  alignas(16) const LCAccel a2{};
  alignas(16) const LCBuffer<const SE5A58585AF30682A> b3{};
  alignas(16) lc_uint v4{};
  alignas(16) lc_uint4 ls_kid;
};

extern "C" { __constant__ Params params; }


/* inline RayQuery API (luisacompute_b200) begin */
struct LCRayQueryInline {
lcrt::LocalStore store;
lcrt::SuspendingCallbacks cb;
lcrt::Traversal<lcrt::MODE_QUERY, false, lcrt::LocalStore, lcrt::SuspendingCallbacks> tr;
LCRTCandidate candidate;
lc_bool triangle_candidate;
};
#ifndef LCRT_RAY_QUERY_SITE_COUNT
#define LCRT_RAY_QUERY_SITE_COUNT 1
#endif
__device__ inline void lc_ray_query_inline_begin(LCRayQueryInline &s, LCRayQuery &q) noexcept {
s.cb.pend.kind = 0;
s.tr.init(lcrt::devlib_tlas(q.accel), reinterpret_cast<const lcrt::Instance *>(q.accel.instances), reinterpret_cast<const lcrt::Ray *>(&q.ray), q.mask,
&s.store, &s.cb);
s.tr.terminate_on_first = (q.flags & LC_RAY_FLAG_TERMINATE_ON_FIRST_HIT) != 0u;
}
#if LCRT_RAY_QUERY_SITE_COUNT <= 1
[[nodiscard]] __device__ inline
#else
[[nodiscard]] __device__ __noinline__
#endif
bool lc_ray_query_inline_proceed(LCRayQueryInline &s) noexcept {
while (s.tr.step()) {
if (!s.tr.candidate_pending()) continue;
const lcrt::PendingCandidate &pc = s.cb.pend;
s.triangle_candidate = pc.kind == 1;
s.candidate = LCRTCandidate{};
if (pc.kind == 1) s.candidate.triangle = LCTriangleHit{pc.inst, pc.prim, lc_make_float2(pc.u, pc.v), pc.t};
else s.candidate.procedural = LCProceduralHit{pc.inst, pc.prim};
s.candidate.world_ray = lcrt::DevlibQueryCallbacks::to_lc_ray(s.tr.candidate_ray());
return true;
}
return false;
}
__device__ inline void lc_ray_query_inline_resolve(LCRayQueryInline &s, const LCIntersectionResult &r) noexcept {
s.tr.resolve(r.committed, r.terminated, r.t_hit);
}
__device__ inline void lc_ray_query_inline_end(LCRayQueryInline &s, LCRayQuery &q) noexcept {
lcrt::devlib_report(q.accel, s.tr.overflowed());
s.tr.result(q.hit.m0, q.hit.m1, q.hit.m3, q.hit.m4, q.hit.m2.x, q.hit.m2.y);
if (q.hit.m3 == static_cast<lc_uint>(LCHitType::MISS)) q.hit.m0 = ~0u;
}
/* inline RayQuery API end */

extern "C" __global__ void __raygen__main() {
  const auto &b0 = params.b0;
  auto b1 = params.b1;
  const auto &a2 = params.a2;
  const auto &b3 = params.b3;
  const auto &v4 = params.v4;
  constexpr auto bs = lc_block_size();
  const auto ls = lc_dispatch_size();
  const auto did = lc_dispatch_id();
  const auto tid = lc_thread_id();
  const auto bid = lc_block_id();
  const auto kid = lc_kernel_id();
  const auto ws = lc_warp_size();
  const auto lid = lc_warp_lane_id();
  lc_uint3 v6{};
  lc_uint v7{};
  lc_uint v8{};
  lc_uint v9{};
  lc_uint v10{};
  LCRay v11{};
  lc_array<lc_float, 3> v12{};
  lc_uint v13{};
  lc_uint v14{};
  lc_uint v15{};
  lc_float v16{};
  lc_array<lc_float, 3> v17{};
  lc_uint v18{};
  lc_uint v19{};
  lc_uint v20{};
  lc_float v21{};
  lc_uint v22{};
  lc_uint v23{};
  lc_uint v24{};
  lc_uint v25{};
  lc_uint v26{};
  lc_uint v27{};
  LCRayQueryAll v28{};
  LCTriangleHit v29{};
  lc_uint v30{};
  lc_uint v31{};
  lc_float2 v32{};
  lc_float v33{};
  lc_float v34{};
  lc_float v35{};
  lc_uint v36{};
  lc_bool v37{};
  LCProceduralHit v38{};
  lc_uint v39{};
  lc_uint v40{};
  LCRay v41{};
  lc_array<lc_float, 3> v42{};
  lc_uint v43{};
  lc_uint v44{};
  lc_uint v45{};
  lc_float v46{};
  lc_array<lc_float, 3> v47{};
  lc_uint v48{};
  lc_uint v49{};
  lc_uint v50{};
  lc_float v51{};
  lc_uint v52{};
  lc_uint v53{};
  lc_uint v54{};
  lc_uint v55{};
  lc_uint v56{};
  lc_uint v57{};
  SE5A58585AF30682A v58{};
  lc_array<lc_float, 3> v59{};
  lc_uint v60{};
  lc_uint v61{};
  lc_uint v62{};
  lc_array<lc_float, 3> v63{};
  lc_uint v64{};
  lc_uint v65{};
  lc_uint v66{};
  lc_uint v67{};
  lc_uint v68{};
  lc_uint v69{};
  lc_uint v70{};
  lc_uint v71{};
  lc_uint v72{};
  lc_float4x4 v73{};
  lc_array<lc_float, 3> v74{};
  lc_uint v75{};
  lc_uint v76{};
  lc_uint v77{};
  lc_float v78{};
  lc_array<lc_float, 3> v79{};
  lc_uint v80{};
  lc_uint v81{};
  lc_uint v82{};
  lc_float v83{};
  lc_array<lc_float, 3> v84{};
  lc_uint v85{};
  lc_uint v86{};
  lc_uint v87{};
  lc_float v88{};
  lc_float3 v89{};
  lc_float v90{};
  lc_float v91{};
  lc_float v92{};
  lc_array<lc_float, 3> v93{};
  lc_uint v94{};
  lc_uint v95{};
  lc_uint v96{};
  lc_float v97{};
  lc_array<lc_float, 3> v98{};
  lc_uint v99{};
  lc_uint v100{};
  lc_uint v101{};
  lc_float v102{};
  lc_array<lc_float, 3> v103{};
  lc_uint v104{};
  lc_uint v105{};
  lc_uint v106{};
  lc_float v107{};
  lc_float3 v108{};
  lc_float v109{};
  lc_float v110{};
  lc_float v111{};
  lc_float3 v112{};
  lc_float v113{};
  lc_float v114{};
  lc_float v115{};
  lc_float3 v116{};
  lc_float v117{};
  lc_float v118{};
  lc_float v119{};
  lc_array<lc_float, 3> v120{};
  lc_uint v121{};
  lc_uint v122{};
  lc_uint v123{};
  lc_float v124{};
  lc_array<lc_float, 3> v125{};
  lc_uint v126{};
  lc_uint v127{};
  lc_uint v128{};
  lc_float v129{};
  lc_array<lc_float, 3> v130{};
  lc_uint v131{};
  lc_uint v132{};
  lc_uint v133{};
  lc_float v134{};
  lc_float3 v135{};
  lc_float v136{};
  lc_float v137{};
  lc_float v138{};
  lc_array<lc_float, 3> v139{};
  lc_uint v140{};
  lc_uint v141{};
  lc_uint v142{};
  lc_float v143{};
  lc_array<lc_float, 3> v144{};
  lc_uint v145{};
  lc_uint v146{};
  lc_uint v147{};
  lc_float v148{};
  lc_array<lc_float, 3> v149{};
  lc_uint v150{};
  lc_uint v151{};
  lc_uint v152{};
  lc_float v153{};
  lc_float3 v154{};
  lc_float v155{};
  lc_float v156{};
  lc_float v157{};
  lc_float v158{};
  lc_float v159{};
  lc_float4 v160{};
  lc_float v161{};
  lc_float v162{};
  lc_float v163{};
  lc_float v164{};
  lc_float4 v165{};
  lc_float v166{};
  lc_float v167{};
  lc_float v168{};
  lc_float v169{};
  lc_float3 v170{};
  lc_float v171{};
  lc_float v172{};
  lc_float v173{};
  lc_int v174{};
  lc_float3 v175{};
  lc_float v176{};
  lc_float v177{};
  lc_float v178{};
  lc_float v179{};
  lc_float v180{};
  lc_float3 v181{};
  lc_float v182{};
  lc_float v183{};
  lc_float v184{};
  lc_float v185{};
  lc_float v186{};
  lc_float v187{};
  lc_float v188{};
  lc_float v189{};
  lc_float v190{};
  lc_array<lc_float, 3> v191{};
  lc_uint v192{};
  lc_uint v193{};
  lc_uint v194{};
  lc_float v195{};
  lc_uint v196{};
  lc_array<lc_float, 3> v197{};
  lc_uint v198{};
  lc_uint v199{};
  lc_uint v200{};
  lc_float v201{};
  lc_uint v202{};
  lc_array<lc_float, 3> v203{};
  lc_uint v204{};
  lc_uint v205{};
  lc_uint v206{};
  lc_float v207{};
  lc_uint v208{};
  lc_float3 v209{};
  lc_float v210{};
  lc_float v211{};
  lc_float v212{};
  lc_float3 v213{};
  lc_float v214{};
  lc_float v215{};
  lc_float v216{};
  lc_float v217{};
  lc_float v218{};
  lc_float v219{};
  lc_float v220{};
  lc_float v221{};
  lc_float v222{};
  lc_array<lc_float, 3> v223{};
  lc_uint v224{};
  lc_uint v225{};
  lc_uint v226{};
  lc_float v227{};
  lc_uint v228{};
  lc_array<lc_float, 3> v229{};
  lc_uint v230{};
  lc_uint v231{};
  lc_uint v232{};
  lc_float v233{};
  lc_uint v234{};
  lc_array<lc_float, 3> v235{};
  lc_uint v236{};
  lc_uint v237{};
  lc_uint v238{};
  lc_float v239{};
  lc_uint v240{};
  lc_float v241{};
  lc_float v242{};
  lc_float v243{};
  lc_float v244{};
  lc_float v245{};
  lc_float v246{};
  lc_float v247{};
  lc_float v248{};
  lc_bool v249{};
  lc_float v250{};
  lc_float v251{};
  lc_float v252{};
  lc_float v253{};
  lc_float v254{};
  lc_float v255{};
  lc_float v256{};
  lc_float v257{};
  lc_bool v258{};
  lc_float v259{};
  lc_bool v260{};
  lc_bool v261{};
  lc_float v262{};
  lc_bool v263{};
  lc_float v264{};
  lc_bool v265{};
  lc_bool v266{};
  LCCommittedHit v267{};
  lc_uint v268{};
  lc_uint v269{};
  lc_float2 v270{};
  lc_float v271{};
  lc_float v272{};
  lc_uint v273{};
  lc_float v274{};
  v7 = 0u;
  v7 = v6.x;
  v8 = 0u;
  v8 = v6.y;
  v9 = 0u;
  v9 = v6.z;
  v6 = lc_make_uint3(0u, 0u, 0u);
  v6 = did;
  v10 = 0u;
  v10 = 255u;
  v10 = v4;
  v13 = 0u;
  v13 = 0u;
  v12[v13] = 0x0p-1022f;
  v14 = 0u;
  v14 = 1u;
  v12[v14] = 0x0p-1022f;
  v15 = 0u;
  v15 = 2u;
  v12[v15] = 0x0p-1022f;
  v12 = v11.m0;
  v16 = 0x0p-1022f;
  v16 = v11.m1;
  v18 = 0u;
  v18 = 0u;
  v17[v18] = 0x0p-1022f;
  v19 = 0u;
  v19 = 1u;
  v17[v19] = 0x0p-1022f;
  v20 = 0u;
  v20 = 2u;
  v17[v20] = 0x0p-1022f;
  v17 = v11.m2;
  v21 = 0x0p-1022f;
  v21 = v11.m3;
  v22 = 0u;
  v22 = 0u;
  v11.m0[v22] = 0x0p-1022f;
  v23 = 0u;
  v23 = 1u;
  v11.m0[v23] = 0x0p-1022f;
  v24 = 0u;
  v24 = 2u;
  v11.m0[v24] = 0x0p-1022f;
  v11.m1 = 0x0p-1022f;
  v25 = 0u;
  v25 = 0u;
  v11.m2[v25] = 0x0p-1022f;
  v26 = 0u;
  v26 = 1u;
  v11.m2[v26] = 0x0p-1022f;
  v27 = 0u;
  v27 = 2u;
  v11.m2[v27] = 0x0p-1022f;
  v11.m3 = 0x0p-1022f;
  v11 = lc_buffer_read(b0, v6.x);
  v28 = lc_accel_query_all(a2, v11, v10);
  
  { // ray query #0, lowered inline
    LCRayQueryInline lc_rq;
    lc_ray_query_inline_begin(lc_rq, v28);
    while (lc_ray_query_inline_proceed(lc_rq)) {
      LCIntersectionResult result{};
      [[maybe_unused]] const LCRTCandidate &lc_rt_candidate = lc_rq.candidate;
      if (lc_rq.triangle_candidate) {
        v30 = 0u;
        v30 = v29.m0;
        v31 = 0u;
        v31 = v29.m1;
        v33 = 0x0p-1022f;
        v33 = v32.x;
        v34 = 0x0p-1022f;
        v34 = v32.y;
        v32 = lc_make_float2(0x0p-1022f, 0x0p-1022f);
        v32 = v29.m2;
        v35 = 0x0p-1022f;
        v35 = v29.m3;
        v29.m0 = 0u;
        v29.m1 = 0u;
        v29.m2 = lc_make_float2(0x0p-1022f, 0x0p-1022f);
        v29.m3 = 0x0p-1022f;
        v29 = LC_RAY_QUERY_TRIANGLE_CANDIDATE_HIT(v28);
        v36 = 0u;
        v36 = (v29.m1 % 3u);
        v37 = false;
        v37 = (v36 != 0u);
        // ../baseline/ref_optix_bench.cpp:148
        if (v37) {
          LC_RAY_QUERY_COMMIT_TRIANGLE(v28);
        }
      } else {
        v39 = 0u;
        v39 = v38.m0;
        v40 = 0u;
        v40 = v38.m1;
        v38.m0 = 0u;
        v38.m1 = 0u;
        v38 = LC_RAY_QUERY_PROCEDURAL_CANDIDATE_HIT(v28);
        v43 = 0u;
        v43 = 0u;
        v42[v43] = 0x0p-1022f;
        v44 = 0u;
        v44 = 1u;
        v42[v44] = 0x0p-1022f;
        v45 = 0u;
        v45 = 2u;
        v42[v45] = 0x0p-1022f;
        v42 = v41.m0;
        v46 = 0x0p-1022f;
        v46 = v41.m1;
        v48 = 0u;
        v48 = 0u;
        v47[v48] = 0x0p-1022f;
        v49 = 0u;
        v49 = 1u;
        v47[v49] = 0x0p-1022f;
        v50 = 0u;
        v50 = 2u;
        v47[v50] = 0x0p-1022f;
        v47 = v41.m2;
        v51 = 0x0p-1022f;
        v51 = v41.m3;
        v52 = 0u;
        v52 = 0u;
        v41.m0[v52] = 0x0p-1022f;
        v53 = 0u;
        v53 = 1u;
        v41.m0[v53] = 0x0p-1022f;
        v54 = 0u;
        v54 = 2u;
        v41.m0[v54] = 0x0p-1022f;
        v41.m1 = 0x0p-1022f;
        v55 = 0u;
        v55 = 0u;
        v41.m2[v55] = 0x0p-1022f;
        v56 = 0u;
        v56 = 1u;
        v41.m2[v56] = 0x0p-1022f;
        v57 = 0u;
        v57 = 2u;
        v41.m2[v57] = 0x0p-1022f;
        v41.m3 = 0x0p-1022f;
        v41 = LC_RAY_QUERY_WORLD_RAY(v28);
        v60 = 0u;
        v60 = 0u;
        v59[v60] = 0x0p-1022f;
        v61 = 0u;
        v61 = 1u;
        v59[v61] = 0x0p-1022f;
        v62 = 0u;
        v62 = 2u;
        v59[v62] = 0x0p-1022f;
        v59 = v58.m0;
        v64 = 0u;
        v64 = 0u;
        v63[v64] = 0x0p-1022f;
        v65 = 0u;
        v65 = 1u;
        v63[v65] = 0x0p-1022f;
        v66 = 0u;
        v66 = 2u;
        v63[v66] = 0x0p-1022f;
        v63 = v58.m1;
        v67 = 0u;
        v67 = 0u;
        v58.m0[v67] = 0x0p-1022f;
        v68 = 0u;
        v68 = 1u;
        v58.m0[v68] = 0x0p-1022f;
        v69 = 0u;
        v69 = 2u;
        v58.m0[v69] = 0x0p-1022f;
        v70 = 0u;
        v70 = 0u;
        v58.m1[v70] = 0x0p-1022f;
        v71 = 0u;
        v71 = 1u;
        v58.m1[v71] = 0x0p-1022f;
        v72 = 0u;
        v72 = 2u;
        v58.m1[v72] = 0x0p-1022f;
        v58 = lc_buffer_read(b3, v38.m1);
        v73 = lc_make_float4x4(0x1p+0f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x1p+0f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x1p+0f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x1p+0f);
        v73 = lc_accel_instance_transform(a2, v38.m0);
        v75 = 0u;
        v75 = 0u;
        v74[v75] = 0x0p-1022f;
        v76 = 0u;
        v76 = 1u;
        v74[v76] = 0x0p-1022f;
        v77 = 0u;
        v77 = 2u;
        v74[v77] = 0x0p-1022f;
        v74 = v58.m1;
        v78 = 0x0p-1022f;
        v78 = v74[2];
        v80 = 0u;
        v80 = 0u;
        v79[v80] = 0x0p-1022f;
        v81 = 0u;
        v81 = 1u;
        v79[v81] = 0x0p-1022f;
        v82 = 0u;
        v82 = 2u;
        v79[v82] = 0x0p-1022f;
        v79 = v58.m1;
        v83 = 0x0p-1022f;
        v83 = v79[1];
        v85 = 0u;
        v85 = 0u;
        v84[v85] = 0x0p-1022f;
        v86 = 0u;
        v86 = 1u;
        v84[v86] = 0x0p-1022f;
        v87 = 0u;
        v87 = 2u;
        v84[v87] = 0x0p-1022f;
        v84 = v58.m1;
        v88 = 0x0p-1022f;
        v88 = v84[0];
        v90 = 0x0p-1022f;
        v90 = v89.x;
        v91 = 0x0p-1022f;
        v91 = v89.y;
        v92 = 0x0p-1022f;
        v92 = v89.z;
        v89 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v89 = lc_make_float3(v88, v83, v78);
        v94 = 0u;
        v94 = 0u;
        v93[v94] = 0x0p-1022f;
        v95 = 0u;
        v95 = 1u;
        v93[v95] = 0x0p-1022f;
        v96 = 0u;
        v96 = 2u;
        v93[v96] = 0x0p-1022f;
        v93 = v58.m0;
        v97 = 0x0p-1022f;
        v97 = v93[2];
        v99 = 0u;
        v99 = 0u;
        v98[v99] = 0x0p-1022f;
        v100 = 0u;
        v100 = 1u;
        v98[v100] = 0x0p-1022f;
        v101 = 0u;
        v101 = 2u;
        v98[v101] = 0x0p-1022f;
        v98 = v58.m0;
        v102 = 0x0p-1022f;
        v102 = v98[1];
        v104 = 0u;
        v104 = 0u;
        v103[v104] = 0x0p-1022f;
        v105 = 0u;
        v105 = 1u;
        v103[v105] = 0x0p-1022f;
        v106 = 0u;
        v106 = 2u;
        v103[v106] = 0x0p-1022f;
        v103 = v58.m0;
        v107 = 0x0p-1022f;
        v107 = v103[0];
        v109 = 0x0p-1022f;
        v109 = v108.x;
        v110 = 0x0p-1022f;
        v110 = v108.y;
        v111 = 0x0p-1022f;
        v111 = v108.z;
        v108 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v108 = lc_make_float3(v107, v102, v97);
        v113 = 0x0p-1022f;
        v113 = v112.x;
        v114 = 0x0p-1022f;
        v114 = v112.y;
        v115 = 0x0p-1022f;
        v115 = v112.z;
        v112 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v112 = (v108 + v89);
        v117 = 0x0p-1022f;
        v117 = v116.x;
        v118 = 0x0p-1022f;
        v118 = v116.y;
        v119 = 0x0p-1022f;
        v119 = v116.z;
        v116 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v116 = (v112 * 0x1p-1f);
        v121 = 0u;
        v121 = 0u;
        v120[v121] = 0x0p-1022f;
        v122 = 0u;
        v122 = 1u;
        v120[v122] = 0x0p-1022f;
        v123 = 0u;
        v123 = 2u;
        v120[v123] = 0x0p-1022f;
        v120 = v58.m0;
        v124 = 0x0p-1022f;
        v124 = v120[2];
        v126 = 0u;
        v126 = 0u;
        v125[v126] = 0x0p-1022f;
        v127 = 0u;
        v127 = 1u;
        v125[v127] = 0x0p-1022f;
        v128 = 0u;
        v128 = 2u;
        v125[v128] = 0x0p-1022f;
        v125 = v58.m0;
        v129 = 0x0p-1022f;
        v129 = v125[1];
        v131 = 0u;
        v131 = 0u;
        v130[v131] = 0x0p-1022f;
        v132 = 0u;
        v132 = 1u;
        v130[v132] = 0x0p-1022f;
        v133 = 0u;
        v133 = 2u;
        v130[v133] = 0x0p-1022f;
        v130 = v58.m0;
        v134 = 0x0p-1022f;
        v134 = v130[0];
        v136 = 0x0p-1022f;
        v136 = v135.x;
        v137 = 0x0p-1022f;
        v137 = v135.y;
        v138 = 0x0p-1022f;
        v138 = v135.z;
        v135 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v135 = lc_make_float3(v134, v129, v124);
        v140 = 0u;
        v140 = 0u;
        v139[v140] = 0x0p-1022f;
        v141 = 0u;
        v141 = 1u;
        v139[v141] = 0x0p-1022f;
        v142 = 0u;
        v142 = 2u;
        v139[v142] = 0x0p-1022f;
        v139 = v58.m1;
        v143 = 0x0p-1022f;
        v143 = v139[2];
        v145 = 0u;
        v145 = 0u;
        v144[v145] = 0x0p-1022f;
        v146 = 0u;
        v146 = 1u;
        v144[v146] = 0x0p-1022f;
        v147 = 0u;
        v147 = 2u;
        v144[v147] = 0x0p-1022f;
        v144 = v58.m1;
        v148 = 0x0p-1022f;
        v148 = v144[1];
        v150 = 0u;
        v150 = 0u;
        v149[v150] = 0x0p-1022f;
        v151 = 0u;
        v151 = 1u;
        v149[v151] = 0x0p-1022f;
        v152 = 0u;
        v152 = 2u;
        v149[v152] = 0x0p-1022f;
        v149 = v58.m1;
        v153 = 0x0p-1022f;
        v153 = v149[0];
        v155 = 0x0p-1022f;
        v155 = v154.x;
        v156 = 0x0p-1022f;
        v156 = v154.y;
        v157 = 0x0p-1022f;
        v157 = v154.z;
        v154 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v154 = lc_make_float3(v153, v148, v143);
        v158 = 0x0p-1022f;
        v158 = (v154.x - v135.x);
        v159 = 0x0p-1022f;
        v159 = (v158 * 0x1p-1f);
        v161 = 0x0p-1022f;
        v161 = v160.x;
        v162 = 0x0p-1022f;
        v162 = v160.y;
        v163 = 0x0p-1022f;
        v163 = v160.z;
        v164 = 0x0p-1022f;
        v164 = v160.w;
        v160 = lc_make_float4(0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v160 = lc_make_float4(v116, 0x1p+0f);
        v166 = 0x0p-1022f;
        v166 = v165.x;
        v167 = 0x0p-1022f;
        v167 = v165.y;
        v168 = 0x0p-1022f;
        v168 = v165.z;
        v169 = 0x0p-1022f;
        v169 = v165.w;
        v165 = lc_make_float4(0x0p-1022f, 0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v165 = (v73 * v160);
        v171 = 0x0p-1022f;
        v171 = v170.x;
        v172 = 0x0p-1022f;
        v172 = v170.y;
        v173 = 0x0p-1022f;
        v173 = v170.z;
        v170 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v170 = lc_make_float3(v165.x, v165.y, v165.z);
        v174 = 0;
        v174 = 0;
        v176 = 0x0p-1022f;
        v176 = v175.x;
        v177 = 0x0p-1022f;
        v177 = v175.y;
        v178 = 0x0p-1022f;
        v178 = v175.z;
        v175 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v175 = lc_make_float3(v73[v174].x, v73[v174].y, v73[v174].z);
        v179 = 0x0p-1022f;
        v179 = lc_length(v175);
        v180 = 0x0p-1022f;
        v180 = (v159 * v179);
        v182 = 0x0p-1022f;
        v182 = v181.x;
        v183 = 0x0p-1022f;
        v183 = v181.y;
        v184 = 0x0p-1022f;
        v184 = v181.z;
        v181 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v185 = 0x0p-1022f;
        v185 = v181.x;
        v186 = 0x0p-1022f;
        v186 = v181.y;
        v187 = 0x0p-1022f;
        v187 = v181.z;
        v188 = 0x0p-1022f;
        v188 = v185;
        v189 = 0x0p-1022f;
        v189 = v186;
        v190 = 0x0p-1022f;
        v190 = v187;
        v192 = 0u;
        v192 = 0u;
        v191[v192] = 0x0p-1022f;
        v193 = 0u;
        v193 = 1u;
        v191[v193] = 0x0p-1022f;
        v194 = 0u;
        v194 = 2u;
        v191[v194] = 0x0p-1022f;
        v191 = v41.m0;
        v195 = 0x0p-1022f;
        v195 = v191[0u];
        v196 = 0u;
        v196 = 0u;
        v181[v196] = v195;
        v198 = 0u;
        v198 = 0u;
        v197[v198] = 0x0p-1022f;
        v199 = 0u;
        v199 = 1u;
        v197[v199] = 0x0p-1022f;
        v200 = 0u;
        v200 = 2u;
        v197[v200] = 0x0p-1022f;
        v197 = v41.m0;
        v201 = 0x0p-1022f;
        v201 = v197[1u];
        v202 = 0u;
        v202 = 1u;
        v181[v202] = v201;
        v204 = 0u;
        v204 = 0u;
        v203[v204] = 0x0p-1022f;
        v205 = 0u;
        v205 = 1u;
        v203[v205] = 0x0p-1022f;
        v206 = 0u;
        v206 = 2u;
        v203[v206] = 0x0p-1022f;
        v203 = v41.m0;
        v207 = 0x0p-1022f;
        v207 = v203[2u];
        v208 = 0u;
        v208 = 2u;
        v181[v208] = v207;
        v210 = 0x0p-1022f;
        v210 = v209.x;
        v211 = 0x0p-1022f;
        v211 = v209.y;
        v212 = 0x0p-1022f;
        v212 = v209.z;
        v209 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v209 = (v181 - v170);
        v214 = 0x0p-1022f;
        v214 = v213.x;
        v215 = 0x0p-1022f;
        v215 = v213.y;
        v216 = 0x0p-1022f;
        v216 = v213.z;
        v213 = lc_make_float3(0x0p-1022f, 0x0p-1022f, 0x0p-1022f);
        v217 = 0x0p-1022f;
        v217 = v213.x;
        v218 = 0x0p-1022f;
        v218 = v213.y;
        v219 = 0x0p-1022f;
        v219 = v213.z;
        v220 = 0x0p-1022f;
        v220 = v217;
        v221 = 0x0p-1022f;
        v221 = v218;
        v222 = 0x0p-1022f;
        v222 = v219;
        v224 = 0u;
        v224 = 0u;
        v223[v224] = 0x0p-1022f;
        v225 = 0u;
        v225 = 1u;
        v223[v225] = 0x0p-1022f;
        v226 = 0u;
        v226 = 2u;
        v223[v226] = 0x0p-1022f;
        v223 = v41.m2;
        v227 = 0x0p-1022f;
        v227 = v223[0u];
        v228 = 0u;
        v228 = 0u;
        v213[v228] = v227;
        v230 = 0u;
        v230 = 0u;
        v229[v230] = 0x0p-1022f;
        v231 = 0u;
        v231 = 1u;
        v229[v231] = 0x0p-1022f;
        v232 = 0u;
        v232 = 2u;
        v229[v232] = 0x0p-1022f;
        v229 = v41.m2;
        v233 = 0x0p-1022f;
        v233 = v229[1u];
        v234 = 0u;
        v234 = 1u;
        v213[v234] = v233;
        v236 = 0u;
        v236 = 0u;
        v235[v236] = 0x0p-1022f;
        v237 = 0u;
        v237 = 1u;
        v235[v237] = 0x0p-1022f;
        v238 = 0u;
        v238 = 2u;
        v235[v238] = 0x0p-1022f;
        v235 = v41.m2;
        v239 = 0x0p-1022f;
        v239 = v235[2u];
        v240 = 0u;
        v240 = 2u;
        v213[v240] = v239;
        v241 = 0x0p-1022f;
        v241 = lc_dot(v213, v213);
        v242 = 0x0p-1022f;
        v242 = lc_dot(v209, v213);
        v243 = 0x0p-1022f;
        v243 = (v180 * v180);
        v244 = 0x0p-1022f;
        v244 = lc_dot(v209, v209);
        v245 = 0x0p-1022f;
        v245 = (v244 - v243);
        v246 = 0x0p-1022f;
        v246 = (v241 * v245);
        v247 = 0x0p-1022f;
        v247 = (v242 * v242);
        v248 = 0x0p-1022f;
        v248 = (v247 - v246);
        v249 = false;
        v249 = (v248 >= 0x0p-1022f);
        // ../baseline/ref_optix_bench.cpp:163
        if (v249) {
          v250 = 0x0p-1022f;
          v250 = lc_sqrt(v248);
          v251 = 0x0p-1022f;
          v251 = -(v242);
          v252 = 0x0p-1022f;
          v252 = (v251 - v250);
          v253 = 0x0p-1022f;
          v253 = (v252 / v241);
          v254 = 0x0p-1022f;
          v254 = -(v242);
          v255 = 0x0p-1022f;
          v255 = (v254 + v250);
          v256 = 0x0p-1022f;
          v256 = (v255 / v241);
          v257 = 0x0p-1022f;
          v257 = v41.m3;
          v258 = false;
          v258 = (v253 < v257);
          v259 = 0x0p-1022f;
          v259 = v41.m1;
          v260 = false;
          v260 = (v253 >= v259);
          v261 = false;
          v261 = (v260 & v258);
          // ../baseline/ref_optix_bench.cpp:166
          if (v261) {
            LC_RAY_QUERY_COMMIT_PROCEDURAL(v28, v253);
          } else {
            v262 = 0x0p-1022f;
            v262 = v41.m3;
            v263 = false;
            v263 = (v256 < v262);
            v264 = 0x0p-1022f;
            v264 = v41.m1;
            v265 = false;
            v265 = (v256 >= v264);
            v266 = false;
            v266 = (v265 & v263);
            if (v266) {
              LC_RAY_QUERY_COMMIT_PROCEDURAL(v28, v256);
            }
          }
        }
      }
      lc_ray_query_inline_resolve(lc_rq, result);
    }
    lc_ray_query_inline_end(lc_rq, v28);
  } // ray query #0

  v268 = 0u;
  v268 = v267.m0;
  v269 = 0u;
  v269 = v267.m1;
  v271 = 0x0p-1022f;
  v271 = v270.x;
  v272 = 0x0p-1022f;
  v272 = v270.y;
  v270 = lc_make_float2(0x0p-1022f, 0x0p-1022f);
  v270 = v267.m2;
  v273 = 0u;
  v273 = v267.m3;
  v274 = 0x0p-1022f;
  v274 = v267.m4;
  v267.m0 = 0u;
  v267.m1 = 0u;
  v267.m2 = lc_make_float2(0x0p-1022f, 0x0p-1022f);
  v267.m3 = 0u;
  v267.m4 = 0x0p-1022f;
  v267 = lc_ray_query_committed_hit(v28);
  lc_buffer_write(b1, v6.x, v267);
}

