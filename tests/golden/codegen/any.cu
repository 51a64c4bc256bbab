// Output of the reference's CUDA code generator (LuisaCompute cuda backend, LUISA_DUMP_SOURCE=1) for the `any` kernel of
// baseline/ref_optix_bench.cpp (scene mode), with the built-in device library text cut out at the marker below.
// Test fixture for tests/test_codegen_compat.py; regenerate with the commands in that file.
#define LUISA_ENABLE_OPTIX
#define LUISA_ENABLE_OPTIX_TRACE_ANY
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

struct alignas(16) Params {
  alignas(16) const LCBuffer<const LCRay> b0{};
  alignas(16) const LCBuffer<lc_uint> b1{};
  alignas(16) const LCAccel a2{};
  alignas(16) lc_uint v3{};
  alignas(16) lc_uint4 ls_kid;
};

extern "C" { __constant__ Params params; }

extern "C" __global__ void __raygen__main() {
  const auto &b0 = params.b0;
  auto b1 = params.b1;
  const auto &a2 = params.a2;
  const auto &v3 = params.v3;
  constexpr auto bs = lc_block_size();
  const auto ls = lc_dispatch_size();
  const auto did = lc_dispatch_id();
  const auto tid = lc_thread_id();
  const auto bid = lc_block_id();
  const auto kid = lc_kernel_id();
  const auto ws = lc_warp_size();
  const auto lid = lc_warp_lane_id();
  lc_uint3 v5{};
  lc_uint v6{};
  lc_uint v7{};
  lc_uint v8{};
  lc_uint v9{};
  LCRay v10{};
  lc_array<lc_float, 3> v11{};
  lc_uint v12{};
  lc_uint v13{};
  lc_uint v14{};
  lc_float v15{};
  lc_array<lc_float, 3> v16{};
  lc_uint v17{};
  lc_uint v18{};
  lc_uint v19{};
  lc_float v20{};
  lc_uint v21{};
  lc_uint v22{};
  lc_uint v23{};
  lc_uint v24{};
  lc_uint v25{};
  lc_uint v26{};
  lc_bool v27{};
  lc_uint v28{};
  v6 = 0u;
  v6 = v5.x;
  v7 = 0u;
  v7 = v5.y;
  v8 = 0u;
  v8 = v5.z;
  v5 = lc_make_uint3(0u, 0u, 0u);
  v5 = did;
  v9 = 0u;
  v9 = 255u;
  v9 = v3;
  v12 = 0u;
  v12 = 0u;
  v11[v12] = 0x0p-1022f;
  v13 = 0u;
  v13 = 1u;
  v11[v13] = 0x0p-1022f;
  v14 = 0u;
  v14 = 2u;
  v11[v14] = 0x0p-1022f;
  v11 = v10.m0;
  v15 = 0x0p-1022f;
  v15 = v10.m1;
  v17 = 0u;
  v17 = 0u;
  v16[v17] = 0x0p-1022f;
  v18 = 0u;
  v18 = 1u;
  v16[v18] = 0x0p-1022f;
  v19 = 0u;
  v19 = 2u;
  v16[v19] = 0x0p-1022f;
  v16 = v10.m2;
  v20 = 0x0p-1022f;
  v20 = v10.m3;
  v21 = 0u;
  v21 = 0u;
  v10.m0[v21] = 0x0p-1022f;
  v22 = 0u;
  v22 = 1u;
  v10.m0[v22] = 0x0p-1022f;
  v23 = 0u;
  v23 = 2u;
  v10.m0[v23] = 0x0p-1022f;
  v10.m1 = 0x0p-1022f;
  v24 = 0u;
  v24 = 0u;
  v10.m2[v24] = 0x0p-1022f;
  v25 = 0u;
  v25 = 1u;
  v10.m2[v25] = 0x0p-1022f;
  v26 = 0u;
  v26 = 2u;
  v10.m2[v26] = 0x0p-1022f;
  v10.m3 = 0x0p-1022f;
  v10 = lc_buffer_read(b0, v5.x);
  v27 = false;
  v27 = lc_accel_trace_any(a2, v10, v9);
  v28 = 0u;
  v28 = lc_select(0u, 1u, v27);
  lc_buffer_write(b1, v5.x, v28);
}

