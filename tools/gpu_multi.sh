#!/bin/bash
# N GPUs (gpurun --gpus N): sharded tests over NCCL, C4 per-try breakdown, the bench line at N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pvgo.py -m gpu -x -q -k "sharded" > gpurun_out/h_tests_$N.log 2>&1; echo "rc=$?" >> gpurun_out/h_tests_$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/c4_bench.py --tries 3 > gpurun_out/h_c4_$N.log 2>&1; echo "rc=$?" >> gpurun_out/h_c4_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/h_bench_$N.json 2> gpurun_out/h_bench_$N.err; echo "rc=$?" >> gpurun_out/h_bench_$N.err
tail -6 gpurun_out/h_tests_$N.log; grep -v "^\*\|OMP" gpurun_out/h_c4_$N.log | tail -8; cut -c1-300 gpurun_out/h_bench_$N.json; tail -3 gpurun_out/h_bench_$N.err
