#!/bin/bash
# Developer tool: build a tuned variant of the library.  usage: tools/build_variant.sh NAME -DFOO=1 ...
# -> gpurun_variants/liblcb200_rt_NAME.so (travels to the GPU box; not part of the product)
set -e
cd "$(dirname "$0")/../luisacompute_b200/csrc"
name=$1; shift
out=../../gpurun_variants
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr "$@" -c -o /tmp/lcrt_kernels_$name.o lcrt_kernels.cu 2> /tmp/ptxas_$name.log
[ -f lcrt_devlib_embedded.o ] || make -s lcrt_devlib_embedded.o
[ -f lcb_device.o ] || make -s lcb_device.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/liblcb200_rt_$name.so /tmp/lcrt_kernels_$name.o lcrt_devlib_embedded.o lcb_device.o -cudart static 2>/dev/null
grep -A2 "k_traceILi0ELb0" /tmp/ptxas_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " -> $name"
