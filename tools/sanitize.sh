#!/bin/bash
# compute-sanitizer passes (one GPU) over the kernels with hand-offs between threads / blocks; logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in mppi cem net rollout fleet; do
    log=gpurun_out/sanitizer_${tool}_${part}.log
    timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > $log 2>&1
    echo "$tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver finished' $log | tr '\n' ' ')"
  done
done
