#!/bin/bash
# Developer tool (run under gpurun): launch list of the builds in tools/build_perf.py + ncu --set full of k_collapse (5 M soup) and k_sort_scatter
name=${1:-build}
python tools/build_perf.py 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${name}_launches.csv python tools/build_perf.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_collapse -s 8 -c 1 -o gpurun_out/${name}_collapse -f python tools/build_perf.py > gpurun_out/${name}_collapse.log 2>&1
tail -1 gpurun_out/${name}_collapse.log
