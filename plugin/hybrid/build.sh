#!/bin/bash
# Builds liblc-backend-b200.so: the reference's CUDA backend (its own objects, compiled by its own build system in
# $BUILD/build_cuda by baseline/build_reference_cuda_backend.sh) with
#   * cuda_primitive / cuda_mesh / cuda_procedural_primitive / cuda_accel  ->  plugin/hybrid/b200_cuda_rt.cpp (over liblcb200_rt.so)
#   * the embedded device-library text `luisa_cuda_builtin_cuda_device_resource`  ->  the same text with its RT section replaced by
#     this repository's device library (plugin/hybrid/make_device_resource.py), same array size
#   * cuda_shader_optix  ->  plugin/hybrid/b200_shader_rt.cpp: kernels that trace rays are launched with cuLaunchKernel
#   * optix_api          ->  plugin/hybrid/b200_optix_stub.cpp: libnvoptix is never opened
#   * cuda_codegen_ast   ->  the same translation unit with class RayQueryLowering replaced by plugin/hybrid/b200_ray_query_lowering.inl
#     (make_codegen.py splices it into a build-time copy): RayQuery handlers are emitted inline, `while (proceed) { handler }`
# swapped in at link time.  Reference sources are neither modified nor copied.  usage: build.sh [REF] [BUILD] [OUT]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${1:-/root/reference}
B=${2:-/tmp/lcref}
OUT=${3:-$ROOT/baseline/_ref/bin}
BD=$B/build_cuda
OBJ=$BD/src/backends/cuda/CMakeFiles/luisa-compute-backend-cuda.dir
[ -f "$OBJ/ref_cuda_vulkan_stubs.o" ] || bash "$ROOT/baseline/build_reference_cuda_backend.sh" "$REF" "$B" "$OUT"
W=$B/hybrid
mkdir -p "$W"
python "$HERE/make_device_resource.py" "$REF/src/backends/cuda/cuda_builtin/cuda_device_resource.h" "$W/b200_device_resource.cpp"
g++ -O1 -fPIC -c -o "$W/b200_device_resource.o" "$W/b200_device_resource.cpp"
# the other three embedded arrays stay the reference's; its definition of the device-resource array becomes weak
objcopy --weaken-symbol=luisa_cuda_builtin_cuda_device_resource "$OBJ/cuda_builtin_embedded.cpp.o" "$W/cuda_builtin_embedded_weak.o"
cmd=$(ninja -C "$BD" -t commands luisa-compute-backend-cuda | grep -E "/cuda_accel\.cpp\.o " | head -1)
rt=$(echo "$cmd" | sed -E "s# -o [^ ]*cuda_accel\.cpp\.o# -o $W/b200_cuda_rt.o#; s# -c [^ ]*cuda_accel\.cpp# -I$REF/src/backends/cuda -I$ROOT/include -c $HERE/b200_cuda_rt.cpp#; s# -MD -MT [^ ]* -MF [^ ]*##")
(cd "$BD" && eval "$rt")
# the code generator with this repository's RayQuery lowering (inline `while (proceed) { handler }` loops; the reference's outlined
# form stays available behind LUISA_B200_RAY_QUERY=outline): a build-time copy of cuda_codegen_ast.cpp, see make_codegen.py
python "$HERE/make_codegen.py" "$REF/src/backends/cuda/cuda_codegen_ast.cpp" "$HERE/b200_ray_query_lowering.inl" "$W/b200_codegen_ast.cpp"
python "$HERE/make_device_resource.py" --inline-api "$W/b200_inline_api.inc"   # lc_ray_query_inline_* as text, emitted by the lowering
cg=$(echo "$cmd" | sed -E "s# -o [^ ]*cuda_accel\.cpp\.o# -o $W/b200_codegen_ast.o#; s# -c [^ ]*cuda_accel\.cpp# -I$REF/src/backends/cuda -I$W -c $W/b200_codegen_ast.cpp#; s# -MD -MT [^ ]* -MF [^ ]*##")
(cd "$BD" && eval "$cg")
for unit in b200_shader_rt b200_optix_stub; do
  one=$(echo "$cmd" | sed -E "s# -o [^ ]*cuda_accel\.cpp\.o# -o $W/$unit.o#; s# -c [^ ]*cuda_accel\.cpp# -I$REF/src/backends/cuda -I$ROOT/include -c $HERE/$unit.cpp#; s# -MD -MT [^ ]* -MF [^ ]*##")
  (cd "$BD" && eval "$one")
done
link=$(ninja -C "$BD" -t commands luisa-compute-backend-cuda | grep -E "^: && .* -o bin/liblc-backend-cuda.so" | sed -E 's#^: && ##; s# && cd .*$##')
link=$(echo "$link" | sed -E "s# [^ ]*/cuda_accel\.cpp\.o##; s# [^ ]*/cuda_mesh\.cpp\.o##; s# [^ ]*/cuda_primitive\.cpp\.o##; s# [^ ]*/cuda_procedural_primitive\.cpp\.o##; s# [^ ]*/cuda_builtin_embedded\.cpp\.o##; s# [^ ]*/cuda_shader_optix\.cpp\.o##; s# [^ ]*/optix_api\.cpp\.o##; s# [^ ]*/cuda_codegen_ast\.cpp\.o##")
link=${link/ -o bin\/liblc-backend-cuda.so / -o bin/liblc-backend-b200.so $OBJ/ref_cuda_vulkan_stubs.o $W/b200_cuda_rt.o $W/b200_shader_rt.o $W/b200_optix_stub.o $W/b200_codegen_ast.o $W/b200_device_resource.o $W/cuda_builtin_embedded_weak.o }
cp -f "$ROOT/luisacompute_b200/liblcb200_rt.so" "$BD/bin/"
(cd "$BD" && eval "$link -L$BD/bin -llcb200_rt")
cp -f "$BD/bin/liblc-backend-b200.so" "$ROOT/luisacompute_b200/liblcb200_rt.so" "$OUT"/
echo "hybrid backend -> $OUT/liblc-backend-b200.so"

# developer variant: the same backend with the reference's OptiX launch path kept (lc-backend-b200o), to separate launch-path
# effects from traversal effects.  HYBRID_OPTIX_LAUNCH=1 bash plugin/hybrid/build.sh
if [ "${HYBRID_OPTIX_LAUNCH:-0}" = "1" ]; then
  python "$HERE/make_device_resource.py" "$REF/src/backends/cuda/cuda_builtin/cuda_device_resource.h" "$W/b200o_device_resource.cpp" optix-launch
  g++ -O1 -fPIC -c -o "$W/b200o_device_resource.o" "$W/b200o_device_resource.cpp"
  link=$(ninja -C "$BD" -t commands luisa-compute-backend-cuda | grep -E "^: && .* -o bin/liblc-backend-cuda.so" | sed -E 's#^: && ##; s# && cd .*$##')
  link=$(echo "$link" | sed -E "s# [^ ]*/cuda_accel\.cpp\.o##; s# [^ ]*/cuda_mesh\.cpp\.o##; s# [^ ]*/cuda_primitive\.cpp\.o##; s# [^ ]*/cuda_procedural_primitive\.cpp\.o##; s# [^ ]*/cuda_builtin_embedded\.cpp\.o##")
  link=${link/ -o bin\/liblc-backend-cuda.so / -o bin/liblc-backend-b200o.so $OBJ/ref_cuda_vulkan_stubs.o $W/b200_cuda_rt.o $W/b200o_device_resource.o $W/cuda_builtin_embedded_weak.o }
  (cd "$BD" && eval "$link -L$BD/bin -llcb200_rt")
  cp -f "$BD/bin/liblc-backend-b200o.so" "$OUT"/
  echo "hybrid backend with OptiX launch -> $OUT/liblc-backend-b200o.so"
fi
