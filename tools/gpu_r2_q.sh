#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pvgo.py tests/test_gpu_reproj.py -m gpu -x -q -k "one_gpu or dense_root or reproj or config4" > gpurun_out/q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q_tests.log
tail -12 gpurun_out/q_tests.log
