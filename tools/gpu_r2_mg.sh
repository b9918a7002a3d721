#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pvgo.py -m gpu -q -k "sharded" > gpurun_out/mg_tests.log 2>&1; echo "rc=$?" >> gpurun_out/mg_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/mg_bench2.json 2> gpurun_out/mg_bench2.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/mg_bench1.json 2> gpurun_out/mg_bench1.err
tail -5 gpurun_out/mg_tests.log; cut -c1-250 gpurun_out/mg_bench2.json; cut -c1-200 gpurun_out/mg_bench1.json; tail -3 gpurun_out/mg_bench2.err
