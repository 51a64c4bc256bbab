#!/bin/bash
# Developer tool: variants of the generated-code-style test kernels -> gpurun_variants/libdevlibtest_NAME.so
# usage: tools/_build_devlib_variants.sh "name:-DFOO=1 -DBAR=2" ...
R=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $R/gpurun_variants
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -shared -cudart static $f \
    -o $R/gpurun_variants/libdevlibtest_$n.so $R/tests/devlib/devlib_kernels.cu 2>&1 | grep -v deprecated
done
