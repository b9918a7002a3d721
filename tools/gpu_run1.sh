#!/bin/bash
# developer GPU session: sanitizer on a small solve, pvgo parity tests, bench line, phase clocks
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_pvgo.py -q -x -k "test_solve_matches_dense and (C1 or band8_300)" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
timeout 900 python -m pytest tests/test_gpu_pvgo.py -q -k "not sharded" > gpurun_out/t_pvgo.log 2>&1
echo "pvgo rc=$?" >> gpurun_out/t_pvgo.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2>&1
timeout 200 python tools/phase_clocks.py 1 2 4 128 256 512 > gpurun_out/phase.log 2>&1
tail -5 gpurun_out/sanitizer.log; tail -15 gpurun_out/t_pvgo.log; tail -3 gpurun_out/bench1.log; cat gpurun_out/phase.log | head -80
