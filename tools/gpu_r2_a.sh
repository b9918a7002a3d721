#!/bin/bash
# round 2, call A: the new parity tests + whole GPU suite + smoke + bench (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_reference_files.py tests/test_gpu_dropin.py tests/test_gpu_imu.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/a_new_tests.log 2>&1; echo "rc=$?" >> gpurun_out/a_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/a_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/a_smoke.log
timeout 600 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
tail -15 gpurun_out/a_new_tests.log; tail -5 gpurun_out/a_gpu_all.log; tail -2 gpurun_out/a_smoke.log; cut -c1-600 gpurun_out/a_bench_n1.json
