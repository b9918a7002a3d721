#!/bin/bash
# N GPUs: C4 per-try breakdown with and without look-ahead, then the C2 bench line at N
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/c4_bench.py --tries 3 > gpurun_out/l_c4_$N.log 2>&1; echo "rc=$?" >> gpurun_out/l_c4_$N.log
ISLAM_ROOT_LOOKAHEAD=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/c4_bench.py --tries 2 > gpurun_out/l_c4_nola_$N.log 2>&1; echo "rc=$?" >> gpurun_out/l_c4_nola_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/l_bench_$N.json 2> gpurun_out/l_bench_$N.err; echo "rc=$?" >> gpurun_out/l_bench_$N.err
grep -v "^\*\|OMP" gpurun_out/l_c4_$N.log | tail -7; grep -v "^\*\|OMP" gpurun_out/l_c4_nola_$N.log | tail -5; cut -c1-300 gpurun_out/l_bench_$N.json; tail -2 gpurun_out/l_bench_$N.err
