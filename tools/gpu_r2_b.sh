#!/bin/bash
# round 2, call B: the pipelined front kernel (front4) — parity, timeline, phase clocks, bench; A/B against k_factor3
mkdir -p gpurun_out
./tools/lat_bench > gpurun_out/b_lat.log 2>&1
timeout 900 python -m pytest tests/test_gpu_pvgo.py -m gpu -q -x > gpurun_out/b_pvgo.log 2>&1; echo "rc=$?" >> gpurun_out/b_pvgo.log
timeout 200 python tools/level_timeline.py > gpurun_out/b_timeline.log 2>&1
timeout 200 python tools/phase_clocks.py 1 4 64 256 512 > gpurun_out/b_phase.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err

tail -15 gpurun_out/b_pvgo.log; tail -25 gpurun_out/b_timeline.log; cut -c1-200 gpurun_out/b_bench.json
