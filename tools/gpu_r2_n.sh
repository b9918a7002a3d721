#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small.py tests/test_gpu_dropin.py tests/test_gpu_reference_files.py -m gpu -x -q > gpurun_out/n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/n_tests.log
timeout 300 python tools/small_bench.py > gpurun_out/n_small.log 2>&1; echo "rc=$?" >> gpurun_out/n_small.log
ISLAM_SMALL_SPEC=0 timeout 300 python tools/small_bench.py > gpurun_out/n_small_nospec.log 2>&1
tail -15 gpurun_out/n_tests.log; cat gpurun_out/n_small.log; grep batch gpurun_out/n_small_nospec.log
