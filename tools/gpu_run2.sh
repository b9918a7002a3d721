#!/bin/bash
# developer GPU session: pvgo parity tests, bench line, per-level timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pvgo.py -q -k "not sharded" > gpurun_out/t_pvgo.log 2>&1
echo "pvgo rc=$?" >> gpurun_out/t_pvgo.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2>&1
timeout 200 python tools/level_timeline.py > gpurun_out/timeline.log 2>&1
tail -4 gpurun_out/t_pvgo.log; cat gpurun_out/timeline.log
