mkdir -p gpurun_out
timeout 200 python tools/build_perf.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_build3.csv python tools/build_perf.py > /dev/null 2>&1
(timeout 1200 python -m pytest tests/test_gpu_holes.py tests/test_gpu_parity.py tests/test_gpu_hybrid.py tests/test_gpu_plugin.py tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | tail -40) > gpurun_out/r2_pytest5.log; tail -40 gpurun_out/r2_pytest5.log
echo done
