#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/o_bench_1.json 2> gpurun_out/o_bench_1.err; echo "rc=$?" >> gpurun_out/o_bench_1.err
timeout 600 python -m pytest tests/test_gpu_pvgo.py -m gpu -x -q -k "sharded" > gpurun_out/o_tests_$N.log 2>&1; echo "rc=$?" >> gpurun_out/o_tests_$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/o_bench_$N.json 2> gpurun_out/o_bench_$N.err; echo "rc=$?" >> gpurun_out/o_bench_$N.err
python -c "
import json
for f in ('gpurun_out/o_bench_1.json','gpurun_out/o_bench_$N.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d.get('other_configs'))
    except Exception as e: print(f, 'ERR', e)
"
tail -4 gpurun_out/o_tests_$N.log; tail -3 gpurun_out/o_bench_1.err; tail -3 gpurun_out/o_bench_$N.err
