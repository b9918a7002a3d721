#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for what in root small; do
    timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py $what > gpurun_out/p_${tool}_${what}.log 2>&1; echo "rc=$?" >> gpurun_out/p_${tool}_${what}.log
    echo "== $tool $what"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|probe done|dense root|small window|rc=" gpurun_out/p_${tool}_${what}.log | tail -5
  done
done
