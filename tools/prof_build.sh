#!/bin/bash
# Developer tool (run under gpurun): ncu --set full of the build kernels of the LAST build in tools/build_perf.py (5 M soup)
name=${1:-build}
python tools/build_perf.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:"k_collapse|k_climb|k_sort_scatter|k_sort_hist|k_emit_records|k_morton|k_centroid" -s 70 -c 11 -o gpurun_out/${name} -f python tools/build_perf.py > gpurun_out/${name}.log 2>&1
tail -1 gpurun_out/${name}.log
