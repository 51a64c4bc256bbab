mkdir -p gpurun_out
timeout 200 python tools/build_perf.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_build2.csv python tools/build_perf.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collapse --launch-skip 3 -c 1 -o gpurun_out/r2_collapse python tools/build_perf.py > gpurun_out/ncu_collapse.log 2>&1
tail -3 gpurun_out/ncu_collapse.log
timeout 600 python -m pytest tests/test_gpu_holes.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
echo done
