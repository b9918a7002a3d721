#!/bin/bash
# round-2 rehearsal on one B200: full GPU test-suite, smoke, both bench arms, ncu launch list + full capture, timeline probes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1000 -c 54 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_factor3 -s 23 -c 1 -o gpurun_out/f3_l3 -f python tools/solve_loop.py 3 > gpurun_out/ncu_f3.log 2>&1
timeout 200 python tools/level_timeline.py > gpurun_out/timeline.log 2>&1
timeout 200 python tools/phase_clocks.py 1 4 64 256 512 > gpurun_out/phase.log 2>&1
ISLAM_FRONT4=1 timeout 200 python tools/level_timeline.py > gpurun_out/timeline_front4.log 2>&1
ISLAM_FRONT4=1 timeout 200 python tools/phase_clocks.py 64 > gpurun_out/phase_front4.log 2>&1
ISLAM_FRONT4=1 timeout 200 python -m pytest tests/test_gpu_pvgo.py -m gpu -q -k "linearize or solve_matches or lm_steps or full_size" > gpurun_out/t_front4.log 2>&1; echo "rc=$?" >> gpurun_out/t_front4.log
timeout 200 python tools/scale_bench.py > gpurun_out/scale_bench.log 2>&1
timeout 600 python tools/configs_bench.py > gpurun_out/configs.log 2>&1
timeout 200 python tools/c4_bench.py --tries 2 > gpurun_out/c4_1gpu.log 2>&1
timeout 300 python tools/small_bench.py > gpurun_out/small_bench.log 2>&1
./tools/lat_bench > gpurun_out/lat_bench.log 2>&1
./tools/chain_bench > gpurun_out/chain_bench.log 2>&1
tail -3 gpurun_out/t_gpu_all.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_n1.json | cut -c1-300; cat gpurun_out/bench_ref.json | cut -c1-300; tail -12 gpurun_out/timeline.log; tail -3 gpurun_out/t_front4.log
