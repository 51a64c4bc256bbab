mkdir -p gpurun_out
for b in 7 5 4 3 2; do LCRT_COLLAPSE_BLOCKS_PER_SM=$b timeout 200 python tools/build_perf.py 2>&1 | tail -1 | sed "s/^/bps=$b /"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_build4.csv python tools/build_perf.py > /dev/null 2>&1
(timeout 1200 python -m pytest tests/test_gpu_holes.py tests/test_gpu_parity.py tests/test_gpu_hybrid.py tests/test_gpu_plugin.py tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | tail -40) > gpurun_out/r2_pytest7.log; tail -30 gpurun_out/r2_pytest7.log
echo done
