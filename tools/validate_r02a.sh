#!/bin/bash
# Round-2 validation pass A (one GPU): all GPU parity tests (no -x: every failure is wanted), measured errors, baseline
# timings of the kernels round 2 works on.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_mppi.py > gpurun_out/bench_mppi.txt 2>&1; tail -20 gpurun_out/bench_mppi.txt
timeout 300 python tools/bench_net.py > gpurun_out/bench_net.txt 2>&1; tail -20 gpurun_out/bench_net.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
ls -la gpurun_out
