#!/bin/bash
# Builds the REFERENCE's own CUDA/OptiX backend (liblc-backend-cuda.so + its luisa_nvrtc helper) out of tree, for the
# reference-side measurement (baseline/ref_optix_bench.cpp).  The reference's sources and build system are used as they
# are; the only deviation: its cuda_event.h refuses to compile without the Vulkan SDK (timeline events are Vulkan
# semaphores), which this image does not have, so the two translation units that include it are compiled against the
# type-only shim in baseline/shim/ and the event / swapchain classes are linked from baseline/ref_cuda_vulkan_stubs.cpp
# (abort if used; the ray-tracing benchmark uses neither).  usage: build_reference_cuda_backend.sh [REF] [BUILD] [OUT]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${1:-/root/reference}
B=${2:-/tmp/lcref}
OUT=${3:-$HERE/_ref/bin}
BD=$B/build_cuda
mkdir -p "$B/shim/uuid" "$OUT"
printf '%s\n' '#pragma once' 'extern "C" {' 'typedef unsigned char uuid_t[16];' 'void uuid_generate(uuid_t);' \
    'void uuid_unparse(const uuid_t,char*);' 'int uuid_parse(const char*,uuid_t);' '}' > "$B/shim/uuid/uuid.h"
if [ ! -f "$BD/build.ninja" ]; then
    cmake -S "$REF" -B "$BD" -G Ninja -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
        -DLibUUID_LIBRARY=/usr/lib/x86_64-linux-gnu/libuuid.so.1 -DLibUUID_INCLUDE_DIR="$B/shim" \
        -DLUISA_COMPUTE_ENABLE_RUST=OFF -DLUISA_COMPUTE_ENABLE_CPU=OFF -DLUISA_COMPUTE_ENABLE_REMOTE=OFF -DLUISA_COMPUTE_ENABLE_GUI=OFF \
        -DLUISA_COMPUTE_ENABLE_DX=OFF -DLUISA_COMPUTE_ENABLE_METAL=OFF -DLUISA_COMPUTE_ENABLE_VULKAN=OFF -DLUISA_COMPUTE_ENABLE_CUDA=ON \
        -DLUISA_COMPUTE_BUILD_TESTS=OFF -DLUISA_COMPUTE_ENABLE_DSL=ON -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc -DCMAKE_CUDA_ARCHITECTURES=100 > "$B/cmake_cuda.log" 2>&1
fi
# everything except cuda_device.cpp / cuda_stream.cpp (which stop at cuda_event.h:60) builds as is
ninja -C "$BD" -k 0 luisa-compute-backend-cuda > "$B/ninja_cuda.log" 2>&1 || true
OBJ=$BD/src/backends/cuda/CMakeFiles/luisa-compute-backend-cuda.dir
for f in cuda_device cuda_stream; do
    cmd=$(ninja -C "$BD" -t commands luisa-compute-backend-cuda | grep -E "/$f\.cpp\.o " | head -1)
    (cd "$BD" && eval "${cmd/ -c / -DLUISA_BACKEND_ENABLE_VULKAN_SWAPCHAIN=1 -I$HERE/shim -c }")
done
cmd=$(ninja -C "$BD" -t commands luisa-compute-backend-cuda | grep -E "/cuda_stream\.cpp\.o " | head -1)
stub=$(echo "$cmd" | sed -E "s# -o [^ ]*cuda_stream\.cpp\.o# -o $OBJ/ref_cuda_vulkan_stubs.o#; s# -c [^ ]*cuda_stream\.cpp# -DLUISA_BACKEND_ENABLE_VULKAN_SWAPCHAIN=1 -I$HERE/shim -I$REF/src/backends/cuda -c $HERE/ref_cuda_vulkan_stubs.cpp#; s# -MD -MT [^ ]* -MF [^ ]*##")
(cd "$BD" && eval "$stub")
link=$(ninja -C "$BD" -t commands luisa-compute-backend-cuda | grep -E "^: && .* -o bin/liblc-backend-cuda.so" | sed -E 's#^: && ##; s# && cd .*$##')
(cd "$BD" && eval "${link/ -o bin\/liblc-backend-cuda.so / -o bin/liblc-backend-cuda.so $OBJ/ref_cuda_vulkan_stubs.o }")
cp -f "$BD"/bin/liblc-*.so "$BD"/bin/luisa_nvrtc "$OUT"/
[ -f /usr/local/cuda/targets/x86_64-linux/lib/libcudadevrt.a ] && cp -f /usr/local/cuda/targets/x86_64-linux/lib/libcudadevrt.a "$OUT"/ || true
echo "reference CUDA backend -> $OUT/liblc-backend-cuda.so"
