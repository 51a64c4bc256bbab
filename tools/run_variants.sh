for v in base rf4 rf8 rf16 rf20 mb6 mb4; do LCB200_RT_LIB=$PWD/gpurun_variants/liblcb200_rt_$v.so python tools/trace_perf.py --tag $v 2>&1 | tail -1; done
