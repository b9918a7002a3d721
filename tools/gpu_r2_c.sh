#!/bin/bash
mkdir -p gpurun_out
ISLAM_FRONT4=1 timeout 900 python -m pytest tests/test_gpu_pvgo.py -m gpu -q -x > gpurun_out/c_pvgo.log 2>&1; echo "rc=$?" >> gpurun_out/c_pvgo.log
ISLAM_FRONT4=1 timeout 200 python tools/level_timeline.py > gpurun_out/c_timeline.log 2>&1
ISLAM_FRONT4=1 timeout 200 python tools/phase_clocks.py 1 64 > gpurun_out/c_phase.log 2>&1
ISLAM_FRONT4=1 ISLAM_DBG_LIB=libislam_dbg2.so timeout 200 python tools/phase_clocks.py 1 64 > gpurun_out/c_phase_chain_only.log 2>&1
ISLAM_FRONT4=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -3 gpurun_out/c_pvgo.log; head -12 gpurun_out/c_timeline.log | cut -c1-130; cut -c1-160 gpurun_out/c_bench.json
