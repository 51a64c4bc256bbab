#!/bin/bash
# Builds the reference's UNMODIFIED host runtime (liblc-core / -ast / -runtime / -dsl ...) out of tree, for the plugin
# integration test only (SURVEY.md Appendix B recipe).  Nothing is written under /root/reference; nothing here is part
# of the product.  usage: build_reference_runtime.sh [REF=/root/reference] [BUILD=/tmp/lcref]
set -e
REF=${1:-/root/reference}
B=${2:-/tmp/lcref}
[ -f "$B/build/bin/liblc-runtime.so" ] && [ -f "$B/build/bin/liblc-dsl.so" ] && exit 0
command -v cmake > /dev/null && command -v ninja > /dev/null
mkdir -p "$B/shim/uuid"
printf '%s\n' '#pragma once' 'extern "C" {' 'typedef unsigned char uuid_t[16];' 'void uuid_generate(uuid_t);' \
    'void uuid_unparse(const uuid_t,char*);' 'int uuid_parse(const char*,uuid_t);' '}' > "$B/shim/uuid/uuid.h"
cmake -S "$REF" -B "$B/build" -G Ninja -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
    -DLibUUID_LIBRARY=/usr/lib/x86_64-linux-gnu/libuuid.so.1 -DLibUUID_INCLUDE_DIR="$B/shim" \
    -DLUISA_COMPUTE_ENABLE_RUST=OFF -DLUISA_COMPUTE_ENABLE_CPU=OFF -DLUISA_COMPUTE_ENABLE_REMOTE=OFF -DLUISA_COMPUTE_ENABLE_GUI=OFF \
    -DLUISA_COMPUTE_ENABLE_DX=OFF -DLUISA_COMPUTE_ENABLE_METAL=OFF -DLUISA_COMPUTE_ENABLE_VULKAN=OFF -DLUISA_COMPUTE_ENABLE_CUDA=OFF \
    -DLUISA_COMPUTE_BUILD_TESTS=OFF -DLUISA_COMPUTE_ENABLE_DSL=ON > "$B/cmake.log" 2>&1
ninja -C "$B/build" luisa-compute-runtime luisa-compute-dsl > "$B/ninja.log" 2>&1
