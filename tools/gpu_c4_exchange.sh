#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
ISLAM_ROOT_EXCHANGE=allreduce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/c4_bench.py --tries 3 > gpurun_out/r_c4_ar_$N.log 2>&1; echo "rc=$?" >> gpurun_out/r_c4_ar_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/c4_bench.py --tries 3 > gpurun_out/r_c4_bc_$N.log 2>&1; echo "rc=$?" >> gpurun_out/r_c4_bc_$N.log
echo "== allreduce exchange"; grep -v "^\*\|OMP" gpurun_out/r_c4_ar_$N.log | tail -6; echo "== broadcast exchange"; grep -v "^\*\|OMP" gpurun_out/r_c4_bc_$N.log | tail -6
