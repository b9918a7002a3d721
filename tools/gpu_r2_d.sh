#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reproj.py tests/test_gpu_dropin.py tests/test_gpu_reference_files.py -m gpu -q > gpurun_out/d_reproj.log 2>&1; echo "rc=$?" >> gpurun_out/d_reproj.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/d_all.log 2>&1; echo "rc=$?" >> gpurun_out/d_all.log
tail -30 gpurun_out/d_reproj.log; tail -4 gpurun_out/d_all.log
