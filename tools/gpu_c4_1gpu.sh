#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pvgo.py -m gpu -x -q -k "one_gpu or dense_root or config4" > gpurun_out/j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/j_tests.log
timeout 300 python tools/c4_bench.py --tries 2 > gpurun_out/j_c4_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/j_c4_1gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_root --csv --log-file gpurun_out/j_c4_root_launches.csv python tools/c4_bench.py --tries 0 > gpurun_out/j_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/j_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_root_syrk -s 3 -c 1 -f -o gpurun_out/j_syrk_full3 python tools/c4_bench.py --tries 0 > gpurun_out/j_ncu2.log 2>&1; echo "rc=$?" >> gpurun_out/j_ncu2.log
tail -8 gpurun_out/j_tests.log; tail -3 gpurun_out/j_c4_1gpu.log; tail -3 gpurun_out/j_ncu.log; tail -3 gpurun_out/j_ncu2.log
