#!/bin/bash
# Builds the reference's UNMODIFIED host runtime (liblc-core / -ast / -runtime / -dsl ...) out of tree, for the plugin
# integration test and the reference-side measurement (SURVEY.md Appendix B recipe, CUDA backend enabled so that
# baseline/build_reference_cuda_backend.sh can continue in the same build directory).  Nothing is written under
# /root/reference; nothing here is part of the product.  usage: build_reference_runtime.sh [REF=/root/reference] [BUILD=/tmp/lcref]
set -e
REF=${1:-/root/reference}
B=${2:-/tmp/lcref}
BD=$B/build_cuda
[ -f "$BD/bin/liblc-runtime.so" ] && [ -f "$BD/bin/liblc-dsl.so" ] && exit 0
command -v cmake > /dev/null && command -v ninja > /dev/null
mkdir -p "$B/shim/uuid"
printf '%s\n' '#pragma once' 'extern "C" {' 'typedef unsigned char uuid_t[16];' 'void uuid_generate(uuid_t);' \
    'void uuid_unparse(const uuid_t,char*);' 'int uuid_parse(const char*,uuid_t);' '}' > "$B/shim/uuid/uuid.h"
if [ ! -f "$BD/build.ninja" ]; then
    cmake -S "$REF" -B "$BD" -G Ninja -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
        -DLibUUID_LIBRARY=/usr/lib/x86_64-linux-gnu/libuuid.so.1 -DLibUUID_INCLUDE_DIR="$B/shim" \
        -DLUISA_COMPUTE_ENABLE_RUST=OFF -DLUISA_COMPUTE_ENABLE_CPU=OFF -DLUISA_COMPUTE_ENABLE_REMOTE=OFF -DLUISA_COMPUTE_ENABLE_GUI=OFF \
        -DLUISA_COMPUTE_ENABLE_DX=OFF -DLUISA_COMPUTE_ENABLE_METAL=OFF -DLUISA_COMPUTE_ENABLE_VULKAN=OFF -DLUISA_COMPUTE_ENABLE_CUDA=ON \
        -DLUISA_COMPUTE_BUILD_TESTS=OFF -DLUISA_COMPUTE_ENABLE_DSL=ON -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc -DCMAKE_CUDA_ARCHITECTURES=100 > "$B/cmake_cuda.log" 2>&1
fi
ninja -C "$BD" luisa-compute-runtime luisa-compute-dsl > "$B/ninja.log" 2>&1
