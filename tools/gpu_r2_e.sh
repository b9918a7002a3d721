#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small.py tests/test_gpu_dropin.py tests/test_gpu_bilevel.py tests/test_gpu_reference_files.py -m gpu -q > gpurun_out/e_small.log 2>&1; echo "rc=$?" >> gpurun_out/e_small.log
timeout 300 python tools/small_bench.py > gpurun_out/e_small_bench.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_small.py -m gpu -q -x -k "window_matches or largest" > gpurun_out/e_sanitizer.log 2>&1
tail -25 gpurun_out/e_small.log; cat gpurun_out/e_small_bench.log; tail -5 gpurun_out/e_sanitizer.log
