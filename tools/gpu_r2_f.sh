#!/bin/bash
# linearise from Jacobian blocks + persisting L2: parity, A/B bench, DRAM traffic of one try
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pvgo.py tests/test_gpu_reproj.py -m gpu -q -x > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/f_bench_on.json 2> gpurun_out/f_bench_on.err
ISLAM_L2_PERSIST=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/f_bench_off.json 2> gpurun_out/f_bench_off.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 400 -c 44 --csv \
    --log-file gpurun_out/f_launches_on.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_on.log 2>&1
ISLAM_L2_PERSIST=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 400 -c 44 --csv \
    --log-file gpurun_out/f_launches_off.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_off.log 2>&1
tail -3 gpurun_out/f_tests.log; python - <<'PY'
import json
for k in ('on','off'):
    d=json.loads(open(f'gpurun_out/f_bench_{k}.json').read().strip().splitlines()[-1])
    print(k, d['value'], d['e2e']['value'], d['roofline']['phases_ms'], d['lm']['rel_pose_error_vs_oracle']['rel'])
PY
