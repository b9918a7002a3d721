#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/c4_bench.py --tries 3 > gpurun_out/m_c4_$N.log 2>&1; echo "rc=$?" >> gpurun_out/m_c4_$N.log
grep -v "^\*\|OMP" gpurun_out/m_c4_$N.log | tail -7
