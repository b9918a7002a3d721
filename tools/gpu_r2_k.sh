#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pvgo.py -m gpu -x -q -k "one_gpu or dense_root or config4" > gpurun_out/k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/k_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/c4_bench.py --one-gpu --N 5000 --n-lc 150 --tries 2 > gpurun_out/k_c4_2r.log 2>&1; echo "rc=$?" >> gpurun_out/k_c4_2r.log
tail -8 gpurun_out/k_tests.log; grep -v "^\*\|OMP" gpurun_out/k_c4_2r.log | tail -6
