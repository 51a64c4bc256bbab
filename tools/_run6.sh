mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collapse --launch-skip 3 -c 1 -o gpurun_out/r2_collapse_v3 python tools/build_perf.py > gpurun_out/ncu_collapse.log 2>&1
tail -2 gpurun_out/ncu_collapse.log
(timeout 1200 python -m pytest tests/test_gpu_holes.py tests/test_gpu_parity.py tests/test_gpu_hybrid.py tests/test_gpu_plugin.py tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | tail -40) > gpurun_out/r2_pytest6.log; tail -40 gpurun_out/r2_pytest6.log
echo done
