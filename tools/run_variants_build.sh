python tools/build_perf.py 2>&1 | tail -1
for v in $(ls gpurun_variants | sed 's/liblcb200_rt_//; s/.so//'); do LCB200_RT_LIB=$PWD/gpurun_variants/liblcb200_rt_$v.so python tools/build_perf.py 2>&1 | tail -1; done
