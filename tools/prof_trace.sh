#!/bin/bash
# Developer tool (run under gpurun): ncu --set full capture of ONE incoherent closest-hit launch of the C2 workload.
# usage: tools/prof_trace.sh NAME [LIB] -> gpurun_out/NAME.ncu-rep
name=${1:-prof}
[ -n "$2" ] && export LCB200_RT_LIB=$2
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 13 -c 1 -o gpurun_out/$name -f \
    python tools/trace_perf.py > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
